// K5, key-gather form ("sparse B"): num[p,(a,b)] = sum_i S[i,p] N[i,a] N[i,b] when a cell's neighbourhood holds only a
// few cell types.  Reference: the P x C^2 cosine numerators of sklearn cosine_similarity(S^T, pairwise products of the
// neighbourhood composition) at laris/tools/_utils.py:859-866.
//
// With k <= ~10 neighbours a cell has 1-3 non-zero N[i,a]; the B operand of the dense contraction (n x C^2) is then
// > 99.8 % zeros.  Here it is stored as what it is - a sparse matrix Q with the key (a <= b) as column:
//   1. pair_key_count_kernel   per cell: bit mask of its non-zero types, histogram of its keys (shared-memory
//                              histogram per CTA, one global atomic per (CTA, key))
//   2. scan                    key_ptr = exclusive scan of the counts (single CTA, C^2 <= 4096 keys)
//   3. pair_key_fill_kernel    entries {cell, q = N_a N_b} into the key's segment (a CTA reserves its range per key
//                              with one atomic, ranks inside it with shared-memory atomics)
//   4. pair_units_kernel       work units of <= kUnit entries of ONE key
//   5. pair_gather_kernel      CTA = (unit, 512-column slab): gathers the unit's S rows (each a contiguous 2 KB
//                              piece, 128-bit streaming loads, 8 rows in flight per thread) and accumulates
//                              q * S in fp64 registers; writes one partial row per unit
//   6. pair_reduce_kernel      sums the partials of a key in unit order into num[:, (a,b)] and its mirror (b,a)
// HBM traffic: 4*P bytes per entry, i.e. (entries / n) * 4nP - the S matrix is read once per key a cell takes part in
// (~1.2-2x at spatially coherent labels) instead of the dense kernel's split pre-pass (S read + 1.5x S written as
// bf16 + re-read).  fp64 accumulation of fp32 products: no operand split, no TMEM truncation - the numerators are
// exact to ~1e-7.  Entry order inside a key follows the atomics, so fp64 sums may differ in the last bits between
// runs (like the COSG gene sums); everything else is deterministic.
#include "common.cuh"
#include "scan.cuh"

namespace laris {
namespace ksp {

constexpr int kMaxC = 64;
constexpr int kUnit = 256;          // entries per work unit
constexpr int kCellsPerCta = 256;

__device__ __forceinline__ uint64_t type_mask(const float* __restrict__ row, int C) {
    uint64_t m = 0ull;
    for (int a = 0; a < C; ++a)
        if (row[a] != 0.f) m |= 1ull << a;
    return m;
}

// key_count[a*C+b] (a <= b) += #cells with N[i,a] != 0 and N[i,b] != 0
__global__ void __launch_bounds__(kCellsPerCta)
pair_key_count_kernel(const float* __restrict__ N, int64_t n, int C, int32_t* __restrict__ key_count) {
    extern __shared__ int32_t hist[];                               // [C*C]
    const int CC = C * C;
    for (int t = threadIdx.x; t < CC; t += kCellsPerCta) hist[t] = 0;
    __syncthreads();
    const int64_t i = (int64_t)blockIdx.x * kCellsPerCta + threadIdx.x;
    if (i < n) {
        const uint64_t m = type_mask(N + i * C, C);
        for (uint64_t ma = m; ma; ma &= ma - 1) {
            const int a = __ffsll((long long)ma) - 1;
            for (uint64_t mb = m & ~((1ull << a) - 1ull); mb; mb &= mb - 1)
                atomicAdd(&hist[a * C + __ffsll((long long)mb) - 1], 1);
        }
    }
    __syncthreads();
    for (int t = threadIdx.x; t < CC; t += kCellsPerCta)
        if (hist[t]) atomicAdd(&key_count[t], hist[t]);
}

// single CTA: key_ptr[0..CC] = exclusive scan of key_count; key_count[CC] = total; cursor = key_ptr (fill positions)
__global__ void __launch_bounds__(1024)
pair_key_scan_kernel(int32_t* __restrict__ key_count, int CC, int32_t* __restrict__ key_ptr, int32_t* __restrict__ cursor) {
    __shared__ int32_t warp_tot[32];
    __shared__ int32_t carry;
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    if (tid == 0) carry = 0;
    __syncthreads();
    for (int base = 0; base < CC; base += 1024) {
        const int t = base + tid;
        const int32_t v = t < CC ? key_count[t] : 0;
        int32_t inc = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int32_t x = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= o) inc += x;
        }
        if (lane == 31) warp_tot[w] = inc;
        __syncthreads();
        int32_t before = carry;
        for (int q = 0; q < w; ++q) before += warp_tot[q];
        if (t < CC) { key_ptr[t] = before + inc - v; cursor[t] = before + inc - v; }
        __syncthreads();
        if (tid == 0) { int32_t s = 0; for (int q = 0; q < 32; ++q) s += warp_tot[q]; carry += s; }
        __syncthreads();
    }
    if (tid == 0) { key_ptr[CC] = carry; key_count[CC] = carry; }
}

__global__ void __launch_bounds__(kCellsPerCta)
pair_key_fill_kernel(const float* __restrict__ N, int64_t n, int C, int32_t* __restrict__ cursor, int2* __restrict__ entries) {
    extern __shared__ int32_t sm[];                                 // hist [C*C], then base [C*C]
    const int CC = C * C;
    int32_t* hist = sm;
    int32_t* base = sm + CC;
    for (int t = threadIdx.x; t < CC; t += kCellsPerCta) hist[t] = 0;
    __syncthreads();
    const int64_t i = (int64_t)blockIdx.x * kCellsPerCta + threadIdx.x;
    uint64_t m = 0ull;
    const float* row = N + i * C;
    if (i < n) {
        m = type_mask(row, C);
        for (uint64_t ma = m; ma; ma &= ma - 1) {
            const int a = __ffsll((long long)ma) - 1;
            for (uint64_t mb = m & ~((1ull << a) - 1ull); mb; mb &= mb - 1)
                atomicAdd(&hist[a * C + __ffsll((long long)mb) - 1], 1);
        }
    }
    __syncthreads();
    for (int t = threadIdx.x; t < CC; t += kCellsPerCta) {
        const int32_t c = hist[t];
        base[t] = c ? atomicAdd(&cursor[t], c) : 0;
        hist[t] = 0;                                                // becomes the rank counter inside the CTA's range
    }
    __syncthreads();
    if (i < n) {
        for (uint64_t ma = m; ma; ma &= ma - 1) {
            const int a = __ffsll((long long)ma) - 1;
            const float va = row[a];
            for (uint64_t mb = m & ~((1ull << a) - 1ull); mb; mb &= mb - 1) {
                const int b = __ffsll((long long)mb) - 1;
                const int key = a * C + b;
                entries[base[key] + atomicAdd(&hist[key], 1)] = make_int2((int)i, __float_as_int(va * row[b]));
            }
        }
    }
}

// units of <= kUnit entries of one key: unit_key / unit_first, *nunits; one thread per key, positions by a serial scan
__global__ void __launch_bounds__(1024)
pair_units_kernel(const int32_t* __restrict__ key_ptr, int CC, int32_t* __restrict__ unit_key, int32_t* __restrict__ unit_first,
                  int32_t* __restrict__ key_unit_ptr /*[CC+1]*/) {
    __shared__ int32_t warp_tot[32];
    __shared__ int32_t carry;
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    if (tid == 0) carry = 0;
    __syncthreads();
    for (int base = 0; base < CC; base += 1024) {
        const int t = base + tid;
        const int32_t cnt = t < CC ? key_ptr[t + 1] - key_ptr[t] : 0;
        const int32_t v = (cnt + kUnit - 1) / kUnit;
        int32_t inc = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int32_t x = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= o) inc += x;
        }
        if (lane == 31) warp_tot[w] = inc;
        __syncthreads();
        int32_t before = carry;
        for (int q = 0; q < w; ++q) before += warp_tot[q];
        const int32_t first_unit = before + inc - v;
        if (t < CC) {
            key_unit_ptr[t] = first_unit;
            for (int32_t u = 0; u < v; ++u) { unit_key[first_unit + u] = t; unit_first[first_unit + u] = key_ptr[t] + u * kUnit; }
        }
        __syncthreads();
        if (tid == 0) { int32_t s = 0; for (int q = 0; q < 32; ++q) s += warp_tot[q]; carry += s; }
        __syncthreads();
    }
    if (tid == 0) key_unit_ptr[CC] = carry;
}

__global__ void __launch_bounds__(128)
pair_gather_kernel(const float4* __restrict__ S4, int64_t ld4, const int2* __restrict__ entries,
                   const int32_t* __restrict__ key_ptr, const int32_t* __restrict__ unit_key,
                   const int32_t* __restrict__ unit_first, const int32_t* __restrict__ key_unit_ptr, int CC,
                   double* __restrict__ partial /*[units][ld4*4]*/) {
    __shared__ int2 ent[kUnit];
    const int u = blockIdx.x;
    if (u >= key_unit_ptr[CC]) return;
    const int key = unit_key[u], e0 = unit_first[u];
    const int cnt = min(kUnit, key_ptr[key + 1] - e0);
    for (int t = threadIdx.x; t < cnt; t += 128) ent[t] = entries[e0 + t];
    __syncthreads();
    const int64_t col4 = (int64_t)blockIdx.y * 128 + threadIdx.x;
    if (col4 >= ld4) return;
    const float4* base = S4 + col4;
    double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
    int t = 0;
    for (; t + 8 <= cnt; t += 8) {
        float4 v[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] = ld_stream_f4(base + (int64_t)ent[t + j].x * ld4);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const double q = (double)__int_as_float(ent[t + j].y);
            a0 = fma(q, (double)v[j].x, a0); a1 = fma(q, (double)v[j].y, a1);
            a2 = fma(q, (double)v[j].z, a2); a3 = fma(q, (double)v[j].w, a3);
        }
    }
    for (; t < cnt; ++t) {
        const float4 v = ld_stream_f4(base + (int64_t)ent[t].x * ld4);
        const double q = (double)__int_as_float(ent[t].y);
        a0 = fma(q, (double)v.x, a0); a1 = fma(q, (double)v.y, a1); a2 = fma(q, (double)v.z, a2); a3 = fma(q, (double)v.w, a3);
    }
    double* o = partial + (int64_t)u * ld4 * 4 + col4 * 4;
    o[0] = a0; o[1] = a1; o[2] = a2; o[3] = a3;
}

// num[p, a*C+b] = num[p, b*C+a] = sum over the key's units, in unit order
__global__ void __launch_bounds__(128)
pair_reduce_kernel(const double* __restrict__ partial, int64_t ld, const int32_t* __restrict__ key_unit_ptr, int C, int P,
                   double* __restrict__ num) {
    const int key = blockIdx.x, a = key / C, b = key - a * C;
    if (a > b) return;
    const int u0 = key_unit_ptr[key], u1 = key_unit_ptr[key + 1];
    if (u0 == u1) return;
    const int CC = C * C;
    for (int p = blockIdx.y * 128 + threadIdx.x; p < P; p += gridDim.y * 128) {
        double t = 0.0;
        for (int u = u0; u < u1; ++u) t += partial[(int64_t)u * ld + p];
        num[(int64_t)p * CC + key] = t;
        if (a != b) num[(int64_t)p * CC + b * C + a] = t;
    }
}

}  // namespace ksp
}  // namespace laris

using namespace laris;

extern "C" int laris_celltype_pair_keys(const float* Ncomp, int64_t n, int32_t C, int32_t* key_count, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    LARIS_CHECK_ARG(Ncomp && key_count && n > 0 && C >= 1 && C <= ksp::kMaxC && n < (int64_t)INT32_MAX,
                    "laris_celltype_pair_keys: needs 1 <= C <= %d and n < 2^31", ksp::kMaxC);
    const int CC = C * C;
    LARIS_CUDA(cudaMemsetAsync(key_count, 0, sizeof(int32_t) * (size_t)(CC + 1), st));
    ksp::pair_key_count_kernel<<<(unsigned)ceil_div(n, ksp::kCellsPerCta), ksp::kCellsPerCta, sizeof(int32_t) * CC, st>>>(
        Ncomp, n, C, key_count);
    LARIS_LAUNCH_CHECK();
    return LARIS_OK;
}

extern "C" int laris_celltype_pair_contract_sparse(const float* S, int64_t ldS, int64_t n, int32_t P, const float* Ncomp,
                                                   int32_t C, int32_t* key_count, int64_t total_entries, double* num,
                                                   void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    LARIS_CHECK_ARG(S && Ncomp && key_count && num && n > 0 && P >= 1 && ldS >= P && C >= 1 && C <= ksp::kMaxC,
                    "laris_celltype_pair_contract_sparse: bad arguments");
    LARIS_CHECK_ARG(ldS % 4 == 0 && ((uintptr_t)S & 15) == 0, "laris_celltype_pair_contract_sparse: S needs ldS %% 4 == 0 and 16-byte alignment");
    LARIS_CHECK_ARG(total_entries >= 0 && total_entries < (int64_t)INT32_MAX, "laris_celltype_pair_contract_sparse: too many entries");
    const int CC = C * C;
    LARIS_CUDA(cudaMemsetAsync(num, 0, sizeof(double) * (size_t)P * CC, st));
    if (total_entries == 0) return LARIS_OK;
    const int64_t unit_bound = total_entries / ksp::kUnit + (int64_t)C * (C + 1) / 2 + 1;
    Scratch<int32_t> key_ptr(st), cursor(st), unit_key(st), unit_first(st), key_unit_ptr(st);
    Scratch<int2> entries(st);
    Scratch<double> partial(st);
    LARIS_CUDA(key_ptr.alloc((size_t)CC + 1));
    LARIS_CUDA(cursor.alloc((size_t)CC + 1));
    LARIS_CUDA(unit_key.alloc((size_t)unit_bound));
    LARIS_CUDA(unit_first.alloc((size_t)unit_bound));
    LARIS_CUDA(key_unit_ptr.alloc((size_t)CC + 1));
    LARIS_CUDA(entries.alloc((size_t)total_entries));
    LARIS_CUDA(partial.alloc((size_t)unit_bound * ldS));
    ksp::pair_key_scan_kernel<<<1, 1024, 0, st>>>(key_count, CC, key_ptr.p, cursor.p);
    LARIS_LAUNCH_CHECK();
    ksp::pair_key_fill_kernel<<<(unsigned)ceil_div(n, ksp::kCellsPerCta), ksp::kCellsPerCta, 2 * sizeof(int32_t) * CC, st>>>(
        Ncomp, n, C, cursor.p, entries.p);
    LARIS_LAUNCH_CHECK();
    ksp::pair_units_kernel<<<1, 1024, 0, st>>>(key_ptr.p, CC, unit_key.p, unit_first.p, key_unit_ptr.p);
    LARIS_LAUNCH_CHECK();
    const int64_t ld4 = ldS / 4;
    dim3 grid((unsigned)unit_bound, (unsigned)ceil_div(ld4, 128));
    ksp::pair_gather_kernel<<<grid, 128, 0, st>>>((const float4*)S, ld4, entries.p, key_ptr.p, unit_key.p, unit_first.p,
                                                 key_unit_ptr.p, CC, partial.p);
    LARIS_LAUNCH_CHECK();
    ksp::pair_reduce_kernel<<<dim3((unsigned)CC, (unsigned)(P >= 1024 ? 4 : 1)), 128, 0, st>>>(partial.p, ldS, key_unit_ptr.p, C,
                                                                                             P, num);
    LARIS_LAUNCH_CHECK();
    return LARIS_OK;
}
