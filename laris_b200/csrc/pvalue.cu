// K6: exceedance counts of the background-resampling null, K7: per-group
// Benjamini-Hochberg.  See include/laris_b200.h for the reference call sites.
//
// K6: one warp per (group, pair) row.  The nb candidate null values of the row
// (scores of the pair's background pairs inside the same sender-receiver group)
// are compared with the observed score ONCE: what a draw contributes is two bits
// per candidate (null >= obs, null > 0), held as register bitmasks (nb <= 32) or
// flag bytes in shared memory.  Each lane then turns Philox outputs into draws -
// for nb <= 64 three per 32-bit word (the first three base-nb digits of w / 2^32:
// idx = umulhi(w, nb); w *= nb; independent and uniform to nb^3 / 2^32 < 6.2e-5),
// i.e. 12 per Philox call - and adds the bits.  The score table (C^2 x P fp64) is
// read once per row and stays L2-resident; the kernel is integer-ALU bound.
#include <math.h>

#include "common.cuh"
#include "rng.cuh"

namespace laris {

constexpr int kPvWarps = 8;
constexpr int kPvMaxNb = 256;

// SMALL: nb <= 32 (register bitmasks); DIGITS: draws per 32-bit word; ALL3: also count nz / ge_nz (only the
// conditional p-value needs them - the default path adds one bit per draw)
template <bool SMALL, int DIGITS, bool ALL3>
__global__ void __launch_bounds__(kPvWarps * 32)
pvalue_counts_kernel(const double* __restrict__ table, int64_t ldt, int n_groups, int n_slots,
                     const int32_t* __restrict__ bg, int nb, const uint8_t* __restrict__ active,
                     const int64_t* __restrict__ slot_gid, int64_t P_global, int g0, int t0, int n_perm, int rng_mode,
                     uint64_t seed, const uint8_t* __restrict__ draws, int32_t* __restrict__ ge, int32_t* __restrict__ nz,
                     int32_t* __restrict__ ge_nz) {
    __shared__ uint8_t flagv[kPvWarps][kPvMaxNb];                   // bit 0: null >= obs, bit 1: null > 0
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    constexpr int per_call = 4 * DIGITS;
    const int64_t nrows = (int64_t)n_groups * n_slots;
    for (int64_t r = (int64_t)blockIdx.x * kPvWarps + warp; r < nrows; r += (int64_t)gridDim.x * kPvWarps) {
        const int g = (int)(r / n_slots), s = (int)(r - (int64_t)g * n_slots);
        if (active && !active[s]) continue;                         // warp-uniform
        const double* trow = table + (int64_t)g * ldt;
        const double obs = trow[s];
        uint32_t mask_ge = 0u, mask_nz = 0u;                        // nb <= 32: the two bits of candidate m, bit m
        __syncwarp();
        for (int m0 = 0; m0 < nb; m0 += 32) {
            const int m = m0 + lane;
            bool a = false, b = false;
            if (m < nb) {
                const double v = trow[bg[(int64_t)s * nb + m]];
                a = v >= obs;
                b = v > 0.0;
                flagv[warp][m] = (uint8_t)((a ? 1 : 0) | (b ? 2 : 0));
            }
            if (m0 == 0) { mask_ge = __ballot_sync(0xffffffffu, a); mask_nz = __ballot_sync(0xffffffffu, b); }
        }
        __syncwarp();
        const uint32_t mask_both = mask_ge & mask_nz;
        int c_ge = 0, c_nz = 0, c_both = 0;
        auto count = [&](uint32_t idx) {
            if (SMALL) {
                c_ge += (mask_ge >> idx) & 1u;
                if (ALL3) { c_nz += (mask_nz >> idx) & 1u; c_both += (mask_both >> idx) & 1u; }
            } else {
                const uint32_t f = flagv[warp][idx];
                c_ge += f & 1u;
                if (ALL3) { c_nz += (f >> 1) & 1u; c_both += f == 3u; }
            }
        };
        if (rng_mode == 0) {
            const uint64_t grow = (uint64_t)(g0 + g) * (uint64_t)P_global + (uint64_t)(slot_gid ? slot_gid[s] : s);
            const int q_beg = t0 / per_call, q_end = (t0 + n_perm + per_call - 1) / per_call;
            for (int q = q_beg + lane; q < q_end; q += 32) {
                const u32x4 o = philox4x32_10((uint32_t)q, (uint32_t)grow, (uint32_t)(grow >> 32), kStreamPval, seed);
                const uint32_t ov[4] = {o.x, o.y, o.z, o.w};
                const int tq = q * per_call;
                if (tq >= t0 && tq + per_call <= t0 + n_perm) {      // every draw of this call is in range: straight-line code
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        uint32_t wv = ov[u];
#pragma unroll
                        for (int j = 0; j < DIGITS; ++j) {
                            count(__umulhi(wv, (uint32_t)nb));
                            wv *= (uint32_t)nb;
                        }
                    }
                } else {                                             // first / last call of a permutation shard
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        uint32_t wv = ov[u];
#pragma unroll
                        for (int j = 0; j < DIGITS; ++j) {
                            const uint32_t idx = __umulhi(wv, (uint32_t)nb);
                            wv *= (uint32_t)nb;
                            const int t = tq + u * DIGITS + j;
                            if (t >= t0 && t < t0 + n_perm) count(idx);
                        }
                    }
                }
            }
        } else {
            const uint8_t* d = draws + r * (int64_t)n_perm;
            for (int t = lane; t < n_perm; t += 32) count(d[t]);
        }
        c_ge = __reduce_add_sync(0xffffffffu, c_ge);
        if (ALL3) { c_nz = __reduce_add_sync(0xffffffffu, c_nz); c_both = __reduce_add_sync(0xffffffffu, c_both); }
        if (lane == 0) {
            ge[r] += c_ge;                                          // one warp per row: plain RMW
            if (ALL3) { nz[r] += c_nz; ge_nz[r] += c_both; }
        }
    }
}

// ---- BH per group: bitonic sort of (p, index) in shared memory, rank scaling, suffix minimum
__global__ void bh_fdr_kernel(const double* __restrict__ p, const uint8_t* __restrict__ sel, int len, int cap,
                              double* __restrict__ fdr) {
    extern __shared__ unsigned char raw[];
    double* key = reinterpret_cast<double*>(raw);                   // [cap]
    int32_t* id = reinterpret_cast<int32_t*>(key + cap);            // [cap]
    __shared__ int m_sel;
    const int g = blockIdx.x, tid = threadIdx.x;
    const double* pg = p + (int64_t)g * len;
    const uint8_t* sg = sel + (int64_t)g * len;
    double* fg = fdr + (int64_t)g * len;
    if (tid == 0) m_sel = 0;
    __syncthreads();
    int local = 0;
    for (int i = tid; i < cap; i += blockDim.x) {
        const bool on = i < len && sg[i];
        key[i] = on ? pg[i] : INFINITY;
        id[i] = i;
        local += on;
        if (i < len && !on) fg[i] = 1.0;
    }
    atomicAdd(&m_sel, local);
    __syncthreads();
    const int m = m_sel;
    for (int kk = 2; kk <= cap; kk <<= 1)
        for (int j = kk >> 1; j > 0; j >>= 1) {
            for (int i = tid; i < cap; i += blockDim.x) {
                const int l = i ^ j;
                if (l > i) {
                    const bool up = (i & kk) == 0;
                    const double a = key[i], b = key[l];
                    const int ia = id[i], ib = id[l];
                    const bool gt = a > b || (a == b && ia > ib);
                    if (gt == up) { key[i] = b; key[l] = a; id[i] = ib; id[l] = ia; }
                }
            }
            __syncthreads();
        }
    // scaled values p_(i) * m / i for the m selected (sorted first), +inf elsewhere
    for (int i = tid; i < cap; i += blockDim.x) key[i] = i < m ? key[i] * (double)m / (double)(i + 1) : INFINITY;
    __syncthreads();
    // suffix minimum (Hillis-Steele)
    for (int off = 1; off < cap; off <<= 1) {
        double v[16];
        int cnt = 0;
        for (int i = tid; i < cap; i += blockDim.x) { v[cnt++] = i + off < cap ? fmin(key[i], key[i + off]) : key[i]; }
        __syncthreads();
        cnt = 0;
        for (int i = tid; i < cap; i += blockDim.x) key[i] = v[cnt++];
        __syncthreads();
    }
    for (int i = tid; i < m; i += blockDim.x) fg[id[i]] = fmin(key[i], 1.0);
}

}  // namespace laris

using namespace laris;

extern "C" int laris_pvalue_counts(const double* table, int64_t ldt, int32_t n_groups, int32_t n_slots, const int32_t* bg,
                                   int32_t nb, const uint8_t* active, const int64_t* slot_gid, int64_t P_global,
                                   int32_t g0, int32_t t0, int32_t n_perm, int rng_mode, uint64_t seed,
                                   const uint8_t* draws, int32_t* ge, int32_t* nz, int32_t* ge_nz, void* stream) {
    LARIS_CHECK_ARG(table && bg && ge && ((nz != nullptr) == (ge_nz != nullptr)), "laris_pvalue_counts: null pointer");
    LARIS_CHECK_ARG(n_groups >= 1 && n_slots >= 1 && ldt >= n_slots && n_perm >= 0, "laris_pvalue_counts: bad sizes");
    LARIS_CHECK_ARG(nb >= 1 && nb <= kPvMaxNb, "laris_pvalue_counts: nb must be in [1,%d] (got %d)", kPvMaxNb, nb);
    LARIS_CHECK_ARG(rng_mode == 0 || (rng_mode == 1 && draws), "laris_pvalue_counts: rng_mode 1 needs host draws");
    if (n_perm == 0) return LARIS_OK;
    const int64_t nrows = (int64_t)n_groups * n_slots;
    int64_t grid = ceil_div(nrows, kPvWarps);
    if (grid > (int64_t)kNumSMs * 16) grid = (int64_t)kNumSMs * 16;
    const bool all3 = nz != nullptr;
#define LARIS_PV_LAUNCH(SMALL, DIGITS, ALL3)                                                                        \
    pvalue_counts_kernel<SMALL, DIGITS, ALL3><<<(unsigned)grid, kPvWarps * 32, 0, (cudaStream_t)stream>>>(          \
        table, ldt, n_groups, n_slots, bg, nb, active, slot_gid, P_global, g0, t0, n_perm, rng_mode, seed, draws, ge, nz, ge_nz)
    if (nb <= 32) { if (all3) LARIS_PV_LAUNCH(true, 3, true); else LARIS_PV_LAUNCH(true, 3, false); }
    else if (nb <= 64) { if (all3) LARIS_PV_LAUNCH(false, 3, true); else LARIS_PV_LAUNCH(false, 3, false); }
    else { if (all3) LARIS_PV_LAUNCH(false, 1, true); else LARIS_PV_LAUNCH(false, 1, false); }
#undef LARIS_PV_LAUNCH
    LARIS_LAUNCH_CHECK();
    return LARIS_OK;
}

extern "C" int laris_bh_fdr(const double* p, const uint8_t* sel, int32_t n_groups, int32_t len, double* fdr, void* stream) {
    LARIS_CHECK_ARG(p && sel && fdr && n_groups >= 1 && len >= 1, "laris_bh_fdr: bad arguments");
    int cap = 32;
    while (cap < len) cap <<= 1;
    if (cap > 16384) { set_error("laris_bh_fdr: group length %d exceeds the shared-memory sort (limit 16384 LR pairs per group)", len); return LARIS_E_LIMIT; }
    const size_t smem = (size_t)cap * 12;
    LARIS_CUDA(cudaFuncSetAttribute(bh_fdr_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int threads = cap >= 1024 ? 1024 : cap;
    bh_fdr_kernel<<<(unsigned)n_groups, threads, smem, (cudaStream_t)stream>>>(p, sel, len, cap, fdr);
    LARIS_LAUNCH_CHECK();
    return LARIS_OK;
}
