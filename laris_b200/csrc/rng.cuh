// Counter-based RNG shared by the null-graph and p-value kernels.
// CPU twin: oracle/rng.py (bit-exact; tests/test_rng_* compare them).
//
// Philox4x32-10 (Salmon et al. 2011, Random123 constants).
//   counter = (c0, c1, c2, stream)   key = (seed & 0xffffffff, seed >> 32)
//   STREAM_COL  = 1 : c0,c1 = edge id lo/hi, c2 = repeat; col = mulhi64(o0|o1<<32, n)
//   STREAM_PERM = 2 : c0 = round, c1 = 0, c2 = repeat; Feistel round key = o0
//   STREAM_PVAL = 3 : c0 = t / 4, c1,c2 = row id lo/hi; draw[t%4] = mulhi32(o[t%4], nb)
#pragma once
#include <stdint.h>

namespace laris {

constexpr uint32_t kStreamCol = 1, kStreamPerm = 2, kStreamPval = 3;
constexpr int kFeistelRounds = 6;

struct u32x4 { uint32_t x, y, z, w; };

__host__ __device__ __forceinline__ u32x4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                                        uint64_t seed) {
    uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        uint64_t p0 = (uint64_t)0xD2511F53u * c0;
        uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
        c1 = (uint32_t)p1;
        c3 = (uint32_t)p0;
        c0 = n0;
        c2 = n2;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    return {c0, c1, c2, c3};
}

__host__ __device__ __forceinline__ uint32_t fmix32(uint32_t x) {
    x ^= x >> 16; x *= 0x85EBCA6Bu; x ^= x >> 13; x *= 0xC2B2AE35u; x ^= x >> 16;
    return x;
}

constexpr uint32_t kLabelPermBase = 0x40000000u;      // permutation ids of the label-permutation null (K4b repeats count from 0)

struct FeistelKeys { uint32_t k[kFeistelRounds]; uint32_t half_bits; };

inline FeistelKeys make_feistel_keys(uint64_t m, uint32_t repeat, uint64_t seed) {
    FeistelKeys fk;
    for (int r = 0; r < kFeistelRounds; ++r) fk.k[r] = philox4x32_10((uint32_t)r, 0u, repeat, kStreamPerm, seed).x;
    uint32_t b = 2;
    while (b < 64 && ((uint64_t)1 << b) < m) ++b;      // bit_length(m-1), at least 2
    fk.half_bits = (b + 1) / 2;
    return fk;
}

// bijection of [0, m): balanced Feistel network over 2*half_bits bits + cycle walking
__device__ __forceinline__ uint64_t feistel_perm(uint64_t i, uint64_t m, const FeistelKeys& fk) {
    const uint32_t h = fk.half_bits;
    const uint64_t mask = ((uint64_t)1 << h) - 1;
    uint64_t x = i;
    do {
        uint64_t L = x >> h, R = x & mask;
#pragma unroll
        for (int r = 0; r < kFeistelRounds; ++r) {
            uint32_t f = fmix32((uint32_t)R * 0x9E3779B1u + fk.k[r]);
            uint64_t nL = R;
            R = L ^ ((uint64_t)f & mask);
            L = nL;
        }
        x = (L << h) | R;
    } while (x >= m);
    return x;
}

}  // namespace laris
