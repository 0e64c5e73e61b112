// K4: per-pair spatial statistics in one pass over the dense score matrix
// (observed graph and each random-graph null repeat), plus the device
// generator of the random graphs.
//
// spatial_stats_kernel: a thread owns 4 adjacent pair columns (one float4 of
// every row it visits) and keeps their five running sums in registers, so the
// only cross-thread traffic is the final partial write.  A CTA covers 512
// columns; rows are dealt round-robin to the row-chunk CTAs in the spatial
// processing order, so all CTAs sweep the tissue together and a neighbour row
// fetched by one CTA is an L2 hit for the others (real graph).  For the random
// graph every neighbour row is a fresh 4*P-byte HBM read - that pass is the
// HBM-bound one: (1+k)*4*n*P + 8*n*k bytes.
#include <cuda.h>
#include <stdlib.h>

#include "async.cuh"
#include "common.cuh"
#include "rng.cuh"
#include "scan.cuh"

namespace laris {

constexpr int kStThreads = 128;
constexpr int kStRowBatch = 8;
constexpr int kStMaxK = 64;

__device__ __forceinline__ float4 ld_cached_f4(const float4* p) {
    float4 r;
    asm volatile("ld.global.nc.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
    return r;
}

// LOCAL: the graph is spatial and `order` walks the tissue - a CTA then takes one CONTIGUOUS range of the order
// and loads through L1, so the neighbour rows of consecutive cells (largely the same rows) are often L1 hits
// instead of L2 round trips (measured 1.1x at 1e5 cells x 500 pairs, 1.7x at 3e5 x 2000).  Otherwise (random
// graph / no order) rows are dealt round-robin and streamed past L1.
template <bool L1N, bool LOCAL>
__global__ void __launch_bounds__(kStThreads)
spatial_stats_kernel(const float4* __restrict__ S4, int64_t ld4, int64_t n, const int32_t* __restrict__ idx,
                     const float* __restrict__ w, int k, const int32_t* __restrict__ order, int nchunks,
                     double* __restrict__ partial) {
    __shared__ int32_t s_idx[kStRowBatch * kStMaxK];
    __shared__ float s_w[kStRowBatch * kStMaxK];
    __shared__ int32_t s_row[kStRowBatch];
    const int tid = threadIdx.x;
    const int64_t col4 = (int64_t)blockIdx.x * kStThreads + tid;
    const bool active = col4 < ld4;
    const int chunk = blockIdx.y;
    double a_dot[4] = {0, 0, 0, 0}, a_ss[4] = {0, 0, 0, 0}, a_tt[4] = {0, 0, 0, 0}, a_s[4] = {0, 0, 0, 0};
    int a_nz[4] = {0, 0, 0, 0};

    const int64_t per_chunk = (n + nchunks - 1) / nchunks;
    const int64_t r_begin = LOCAL ? chunk * per_chunk : chunk, r_end = LOCAL ? min(n, r_begin + per_chunk) : n;
    const int64_t r_step = LOCAL ? 1 : nchunks;
    auto ldrow = [&](const float4* p) { return LOCAL ? ld_cached_f4(p) : ld_stream_f4(p); };
    for (int64_t m0 = 0; r_begin + m0 * r_step < r_end; m0 += kStRowBatch) {
        __syncthreads();
        if (tid < kStRowBatch) {
            const int64_t r = r_begin + (m0 + tid) * r_step;
            s_row[tid] = r < r_end ? (order ? order[r] : (int32_t)r) : -1;
        }
        __syncthreads();
        for (int e = tid; e < kStRowBatch * k; e += kStThreads) {
            const int rb = e / k, j = e - rb * k;
            const int64_t i = s_row[rb];
            if (i >= 0) { s_idx[e] = idx[i * k + j]; s_w[e] = w[i * k + j]; }
        }
        __syncthreads();
        if (!active) continue;
#pragma unroll 1
        for (int rb = 0; rb < kStRowBatch; ++rb) {
            const int64_t i = s_row[rb];
            if (i < 0) break;
            const float4 s = ldrow(S4 + i * ld4 + col4);
            float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
            float wsum = 0.f;
            const int32_t* ni = s_idx + rb * k;
            const float* nw = s_w + rb * k;
            int j = 0;
            for (; j + 4 <= k; j += 4) {
                const float4 v0 = ldrow(S4 + (int64_t)ni[j] * ld4 + col4);
                const float4 v1 = ldrow(S4 + (int64_t)ni[j + 1] * ld4 + col4);
                const float4 v2 = ldrow(S4 + (int64_t)ni[j + 2] * ld4 + col4);
                const float4 v3 = ldrow(S4 + (int64_t)ni[j + 3] * ld4 + col4);
                const float w0 = nw[j], w1 = nw[j + 1], w2 = nw[j + 2], w3 = nw[j + 3];
                t.x = fmaf(w0, v0.x, t.x); t.y = fmaf(w0, v0.y, t.y); t.z = fmaf(w0, v0.z, t.z); t.w = fmaf(w0, v0.w, t.w);
                t.x = fmaf(w1, v1.x, t.x); t.y = fmaf(w1, v1.y, t.y); t.z = fmaf(w1, v1.z, t.z); t.w = fmaf(w1, v1.w, t.w);
                t.x = fmaf(w2, v2.x, t.x); t.y = fmaf(w2, v2.y, t.y); t.z = fmaf(w2, v2.z, t.z); t.w = fmaf(w2, v2.w, t.w);
                t.x = fmaf(w3, v3.x, t.x); t.y = fmaf(w3, v3.y, t.y); t.z = fmaf(w3, v3.z, t.z); t.w = fmaf(w3, v3.w, t.w);
                wsum += (w0 + w1) + (w2 + w3);
            }
            for (; j < k; ++j) {
                const float4 v0 = ldrow(S4 + (int64_t)ni[j] * ld4 + col4);
                const float w0 = nw[j];
                t.x = fmaf(w0, v0.x, t.x); t.y = fmaf(w0, v0.y, t.y); t.z = fmaf(w0, v0.z, t.z); t.w = fmaf(w0, v0.w, t.w);
                wsum += w0;
            }
            if (L1N && wsum != 0.f) {
                const float inv = 1.0f / wsum;
                t.x *= inv; t.y *= inv; t.z *= inv; t.w *= inv;
            }
            const float sv[4] = {s.x, s.y, s.z, s.w}, tv[4] = {t.x, t.y, t.z, t.w};
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                a_dot[c] += (double)(sv[c] * tv[c]);
                a_ss[c] += (double)(sv[c] * sv[c]);
                a_tt[c] += (double)(tv[c] * tv[c]);
                a_s[c] += (double)sv[c];
                a_nz[c] += sv[c] != 0.f ? 1 : 0;
            }
        }
    }
    if (active) {
        const int64_t Ppad = ld4 * 4;
        double* o = partial + (int64_t)chunk * 5 * Ppad + col4 * 4;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            o[c] = a_dot[c];
            o[Ppad + c] = a_ss[c];
            o[2 * Ppad + c] = a_tt[c];
            o[3 * Ppad + c] = a_s[c];
            o[4 * Ppad + c] = (double)a_nz[c];
        }
    }
}

// partial [nchunks][5][Ppad] -> out [5][P].  CTA (32 columns x 8 chunk lanes): coalesced reads, the 8 chunk lanes
// are combined in a fixed order, so the result is deterministic.
__global__ void __launch_bounds__(256)
stats_reduce_kernel(const double* __restrict__ partial, int nchunks, int64_t Ppad, int P, double* __restrict__ out) {
    __shared__ double red[8][33];
    const int cx = threadIdx.x & 31, cy = threadIdx.x >> 5;
    const int p = blockIdx.x * 32 + cx, q = blockIdx.y;
    double s = 0.0;
    if (p < P)
        for (int c = cy; c < nchunks; c += 8) s += partial[((int64_t)c * 5 + q) * Ppad + p];
    red[cy][cx] = s;
    __syncthreads();
    if (cy == 0 && p < P) {
        double t = 0.0;
#pragma unroll
        for (int r = 0; r < 8; ++r) t += red[r][cx];
        out[(int64_t)q * P + p] = t;
    }
}


// ---------------------------------------------------------------------------------------------------------
// K4b batched: RB random-graph repeats in one pass.  S[i,:] is read once for all repeats, sum S^2 / sum S / nnz
// are formed once; every neighbour row of every repeat is a fresh HBM read (k * RB independent 128-bit streaming
// loads per thread and row).  partial layout [chunk][2*RB+3][Ppad]: (dot_r, tt_r) per repeat, then ss, sum, nnz.
// ---------------------------------------------------------------------------------------------------------
template <int RB>
__global__ void __launch_bounds__(kStThreads)
spatial_null_stats_kernel(const float4* __restrict__ S4, int64_t ld4, int64_t n, const int32_t* __restrict__ nidx,
                          const float* __restrict__ nw, int k, int nchunks, double* __restrict__ partial) {
    constexpr int RBATCH = 4;                                        // rows staged per barrier
    __shared__ int32_t s_idx[RBATCH * RB * kStMaxK];
    __shared__ float s_w[RBATCH * RB * kStMaxK];
    const int tid = threadIdx.x;
    const int64_t col4 = (int64_t)blockIdx.x * kStThreads + tid;
    const bool active = col4 < ld4;
    const int chunk = blockIdx.y;
    double a_dot[RB][4], a_tt[RB][4], a_ss[4] = {0, 0, 0, 0}, a_s[4] = {0, 0, 0, 0};
    int a_nz[4] = {0, 0, 0, 0};
#pragma unroll
    for (int r = 0; r < RB; ++r)
#pragma unroll
        for (int c = 0; c < 4; ++c) { a_dot[r][c] = 0.0; a_tt[r][c] = 0.0; }
    const int64_t nk = n * (int64_t)k;
    for (int64_t m0 = chunk; m0 < n; m0 += (int64_t)RBATCH * nchunks) {
        __syncthreads();
        for (int e = tid; e < RBATCH * RB * k; e += kStThreads) {
            const int rb = e / (RB * k), rem = e - rb * (RB * k), r = rem / k, j = rem - r * k;
            const int64_t i = m0 + (int64_t)rb * nchunks;
            if (i < n) { s_idx[e] = nidx[(int64_t)r * nk + i * k + j]; s_w[e] = nw[(int64_t)r * nk + i * k + j]; }
        }
        __syncthreads();
        if (!active) continue;
#pragma unroll 1
        for (int rb = 0; rb < RBATCH; ++rb) {
            const int64_t i = m0 + (int64_t)rb * nchunks;
            if (i >= n) break;
            const float4 s = ld_stream_f4(S4 + i * ld4 + col4);
            const float sv[4] = {s.x, s.y, s.z, s.w};
#pragma unroll
            for (int r = 0; r < RB; ++r) {
                const int32_t* ni = s_idx + (rb * RB + r) * k;
                const float* wv = s_w + (rb * RB + r) * k;
                float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
                float wsum = 0.f;
                int j = 0;
                for (; j + 4 <= k; j += 4) {
                    const float4 v0 = ld_stream_f4(S4 + (int64_t)ni[j] * ld4 + col4);
                    const float4 v1 = ld_stream_f4(S4 + (int64_t)ni[j + 1] * ld4 + col4);
                    const float4 v2 = ld_stream_f4(S4 + (int64_t)ni[j + 2] * ld4 + col4);
                    const float4 v3 = ld_stream_f4(S4 + (int64_t)ni[j + 3] * ld4 + col4);
                    const float w0 = wv[j], w1 = wv[j + 1], w2 = wv[j + 2], w3 = wv[j + 3];
                    t.x = fmaf(w0, v0.x, t.x); t.y = fmaf(w0, v0.y, t.y); t.z = fmaf(w0, v0.z, t.z); t.w = fmaf(w0, v0.w, t.w);
                    t.x = fmaf(w1, v1.x, t.x); t.y = fmaf(w1, v1.y, t.y); t.z = fmaf(w1, v1.z, t.z); t.w = fmaf(w1, v1.w, t.w);
                    t.x = fmaf(w2, v2.x, t.x); t.y = fmaf(w2, v2.y, t.y); t.z = fmaf(w2, v2.z, t.z); t.w = fmaf(w2, v2.w, t.w);
                    t.x = fmaf(w3, v3.x, t.x); t.y = fmaf(w3, v3.y, t.y); t.z = fmaf(w3, v3.z, t.z); t.w = fmaf(w3, v3.w, t.w);
                    wsum += (w0 + w1) + (w2 + w3);
                }
                for (; j < k; ++j) {
                    const float4 v0 = ld_stream_f4(S4 + (int64_t)ni[j] * ld4 + col4);
                    const float w0 = wv[j];
                    t.x = fmaf(w0, v0.x, t.x); t.y = fmaf(w0, v0.y, t.y); t.z = fmaf(w0, v0.z, t.z); t.w = fmaf(w0, v0.w, t.w);
                    wsum += w0;
                }
                if (wsum != 0.f) {                                   // L1 row normalisation of the random graph
                    const float inv = 1.0f / wsum;
                    t.x *= inv; t.y *= inv; t.z *= inv; t.w *= inv;
                }
                const float tv[4] = {t.x, t.y, t.z, t.w};
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    a_dot[r][c] += (double)(sv[c] * tv[c]);
                    a_tt[r][c] += (double)(tv[c] * tv[c]);
                }
            }
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                a_ss[c] += (double)(sv[c] * sv[c]);
                a_s[c] += (double)sv[c];
                a_nz[c] += sv[c] != 0.f ? 1 : 0;
            }
        }
    }
    if (active) {
        const int64_t Ppad = ld4 * 4;
        constexpr int Q = 2 * RB + 3;
        double* o = partial + (int64_t)chunk * Q * Ppad + col4 * 4;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
#pragma unroll
            for (int r = 0; r < RB; ++r) {
                o[(2 * r) * Ppad + c] = a_dot[r][c];
                o[(2 * r + 1) * Ppad + c] = a_tt[r][c];
            }
            o[(2 * RB) * Ppad + c] = a_ss[c];
            o[(2 * RB + 1) * Ppad + c] = a_s[c];
            o[(2 * RB + 2) * Ppad + c] = (double)a_nz[c];
        }
    }
}

// partial [nchunks][2*RB+3][Ppad] -> out rows [RB][5][P] (dot, ss, tt, sum, nnz per repeat; ss / sum / nnz repeated)
__global__ void __launch_bounds__(256)
null_stats_reduce_kernel(const double* __restrict__ partial, int nchunks, int64_t Ppad, int P, int RB,
                         double* __restrict__ out) {
    __shared__ double red[8][33];
    const int cx = threadIdx.x & 31, cy = threadIdx.x >> 5;
    const int p = blockIdx.x * 32 + cx, q = blockIdx.y, Q = 2 * RB + 3;
    double s = 0.0;
    if (p < P)
        for (int c = cy; c < nchunks; c += 8) s += partial[((int64_t)c * Q + q) * Ppad + p];
    red[cy][cx] = s;
    __syncthreads();
    if (cy == 0 && p < P) {
        double t = 0.0;
#pragma unroll
        for (int r = 0; r < 8; ++r) t += red[r][cx];
        if (q < 2 * RB) {
            out[((int64_t)(q >> 1) * 5 + ((q & 1) ? 2 : 0)) * P + p] = t;
        } else {
            const int stat = q == 2 * RB ? 1 : (q == 2 * RB + 1 ? 3 : 4);
            for (int r = 0; r < RB; ++r) out[((int64_t)r * 5 + stat) * P + p] = t;
        }
    }
}

// ---------------------------------------------------------------------------------------------------------
// K4b from L2.  The random neighbour rows of the null cannot be cached when every gather drags a whole 4*P-byte row
// through the SM - but a BLOCK of 16 or 32 columns of S (64 / 128 bytes per row, n rows) fits the 126 MB L2 for n up to
// ~1.6e6 cells.  Measured on B200 (tools/probes/l2_gather_probe.cu): random 128-byte row gathers run at 14.7-17 TB/s from a
// working set of <= 128 MB, against 5.4-6.8 TB/s from HBM.  So the pass is run once per column block (one launch per
// block: the stream order is the grid-wide barrier that keeps all SMs on the same block): the block is read from HBM
// once (compulsory misses) and the R*k gathers per cell are L2 hits.  Thread mapping: a warp covers 4 (8) rows x 8 (4)
// column quads, a thread owns one quad of the rows of its lane; the R*k (index, weight) pairs of a batch of rows are
// staged in shared memory.  HBM bytes: 4nP (S once) + 8nkR per column block (the edge lists are re-read per block).
// ---------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, uint32_t src_bytes) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async4(uint32_t dst, const void* src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit_group() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait_group() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// L2 residency control for the column-blocked pass: the compact block (written once, gathered R*k times per row) is
// marked evict_last, everything that merely streams past it (S on its way into the block, the edge lists) evict_first, so
// that 80 MB of edge lists per launch do not push block lines out (ncu before: 38 % of the gathers missed L2).
__device__ __forceinline__ uint64_t l2_policy_keep() {
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ uint64_t l2_policy_stream() {
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ float4 ldg_f4_policy(const float4* p, uint64_t pol) {
    float4 r;
    asm("ld.global.nc.L2::cache_hint.v4.f32 {%0,%1,%2,%3}, [%4], %5;" : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p), "l"(pol));
    return r;
}
__device__ __forceinline__ float4 ldg_f4_stream_policy(const float4* p, uint64_t pol) {
    float4 r;
    asm("ld.global.nc.L1::no_allocate.L2::cache_hint.v4.f32 {%0,%1,%2,%3}, [%4], %5;"
        : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p), "l"(pol));
    return r;
}
__device__ __forceinline__ void st_f4_policy(float4* p, float4 v, uint64_t pol) {
    asm volatile("st.global.L2::cache_hint.v4.f32 [%0], {%1,%2,%3,%4}, %5;" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w), "l"(pol) : "memory");
}
__device__ __forceinline__ void cp_async16_policy(uint32_t dst, const void* src, uint64_t pol) {
    asm volatile("cp.async.cg.shared.global.L2::cache_hint [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "l"(pol) : "memory");
}
__device__ __forceinline__ void cp_async4_policy(uint32_t dst, const void* src, uint64_t pol) {
    asm volatile("cp.async.ca.shared.global.L2::cache_hint [%0], [%1], 4, %2;" ::"r"(dst), "l"(src), "l"(pol) : "memory");
}

// One (column block, repeat) per launch over the COMPACT block (see compact_block_kernel).  Warps are independent: a warp
// covers 32/QUADS rows x QUADS column quads per step, brings the edge lists of its NEXT step into its own slice of
// shared memory with cp.async while the current step gathers (no CTA barrier in the loop), keeps ten 128-bit L2
// gathers in flight per thread, and its sums stay in fp32 (a thread sees <= 64 rows per launch; the CTA reduction and
// everything after it is fp64): ~76 registers, three CTAs per SM.  FIRST: also sum S^2, sum S and nnz of the block
// (needed once, not per repeat).
template <int QUADS, bool FIRST, bool HINT>
__global__ void __launch_bounds__(256, 3)
spatial_null_stats_l2_kernel(const float4* __restrict__ B4, int64_t n, const int32_t* __restrict__ nidx,
                             const float* __restrict__ nw, int k, double* __restrict__ partial /*[grid][5][QUADS*4]*/) {
    constexpr int ROWS = 256 / QUADS, WB = QUADS * 4, RW = 32 / QUADS;              // rows per CTA step / per warp step
    extern __shared__ __align__(16) unsigned char l2_smem[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, quad = tid % QUADS, rl = tid / QUADS;
    const int per_buf = RW * k;                                      // words of idx (and of w) per warp buffer
    int32_t* w_idx = reinterpret_cast<int32_t*>(l2_smem) + (size_t)warp * 4 * per_buf;     // [2][per_buf] idx, then [2][per_buf] w
    float* w_w = reinterpret_cast<float*>(w_idx + 2 * per_buf);
    float a_dot[4] = {0, 0, 0, 0}, a_tt[4] = {0, 0, 0, 0}, a_ss[4] = {0, 0, 0, 0}, a_s[4] = {0, 0, 0, 0};
    int a_nz[4] = {0, 0, 0, 0};
    const float4* base = B4 + quad;
    const int64_t nwarps = (int64_t)gridDim.x * 8, gw = (int64_t)blockIdx.x * 8 + warp;
    const uint64_t keep = HINT ? l2_policy_keep() : 0, pass = HINT ? l2_policy_stream() : 0;
    auto row_piece = [&](int64_t r) { return HINT ? ldg_f4_policy(base + r * QUADS, keep) : __ldg(base + r * QUADS); };
    auto stage = [&](int64_t b, int buf) {
        const int64_t i0 = b * RW;
        if (i0 < n) {
            const int rows = (int)min((int64_t)RW, n - i0);
            const uint32_t di = aio::smem_addr(w_idx + buf * per_buf), dw = aio::smem_addr(w_w + buf * per_buf);
            const bool vec = rows == RW && ((((uintptr_t)(nidx + i0 * k)) | ((uintptr_t)(nw + i0 * k))) & 15) == 0;
            if (vec) {                                               // RW * k * 4 bytes is a multiple of 16 (RW >= 4)
                for (int c = lane; c < per_buf / 4; c += 32) {
                    if (HINT) {
                        cp_async16_policy(di + 16 * c, nidx + i0 * k + 4 * c, pass);
                        cp_async16_policy(dw + 16 * c, nw + i0 * k + 4 * c, pass);
                    } else {
                        cp_async16(di + 16 * c, nidx + i0 * k + 4 * c, 16);
                        cp_async16(dw + 16 * c, nw + i0 * k + 4 * c, 16);
                    }
                }
            } else {
                for (int e = lane; e < rows * k; e += 32) {
                    if (HINT) {
                        cp_async4_policy(di + 4 * e, nidx + i0 * k + e, pass);
                        cp_async4_policy(dw + 4 * e, nw + i0 * k + e, pass);
                    } else {
                        cp_async4(di + 4 * e, nidx + i0 * k + e);
                        cp_async4(dw + 4 * e, nw + i0 * k + e);
                    }
                }
            }
        }
        cp_async_commit_group();
    };
    stage(gw, 0);
    int buf = 0;
    for (int64_t b = gw; b * RW < n; b += nwarps, buf ^= 1) {
        stage(b + nwarps, buf ^ 1);
        cp_async_wait_group<1>();
        __syncwarp();
        const int64_t i = b * RW + (lane / QUADS);
        if (i < n) {
            const int32_t* ni = w_idx + buf * per_buf + (lane / QUADS) * k;
            const float* wv = w_w + buf * per_buf + (lane / QUADS) * k;
            const float4 s = row_piece(i);
            float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
            float wsum = 0.f;                                        // same association as the streaming kernel
            {
                int j = 0;
                for (; j + 4 <= k; j += 4) wsum += (wv[j] + wv[j + 1]) + (wv[j + 2] + wv[j + 3]);
                for (; j < k; ++j) wsum += wv[j];
            }
            int j = 0;
            for (; j + 10 <= k; j += 10) {                           // ten independent 128-bit L2 gathers in flight
                float4 v[10];
#pragma unroll
                for (int u = 0; u < 10; ++u) v[u] = row_piece(ni[j + u]);
#pragma unroll
                for (int u = 0; u < 10; ++u) {
                    const float w0 = wv[j + u];
                    t.x = fmaf(w0, v[u].x, t.x); t.y = fmaf(w0, v[u].y, t.y); t.z = fmaf(w0, v[u].z, t.z); t.w = fmaf(w0, v[u].w, t.w);
                }
            }
            for (; j < k; ++j) {
                const float4 v0 = row_piece(ni[j]);
                const float w0 = wv[j];
                t.x = fmaf(w0, v0.x, t.x); t.y = fmaf(w0, v0.y, t.y); t.z = fmaf(w0, v0.z, t.z); t.w = fmaf(w0, v0.w, t.w);
            }
            if (wsum != 0.f) {                                       // L1 row normalisation of the random graph
                const float inv = 1.0f / wsum;
                t.x *= inv; t.y *= inv; t.z *= inv; t.w *= inv;
            }
            const float sv[4] = {s.x, s.y, s.z, s.w}, tv[4] = {t.x, t.y, t.z, t.w};
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                a_dot[c] = fmaf(sv[c], tv[c], a_dot[c]);
                a_tt[c] = fmaf(tv[c], tv[c], a_tt[c]);
                if (FIRST) {
                    a_ss[c] = fmaf(sv[c], sv[c], a_ss[c]);
                    a_s[c] += sv[c];
                    a_nz[c] += sv[c] != 0.f ? 1 : 0;
                }
            }
        }
        __syncwarp();                                                // buffer `buf` is refilled two steps from now
    }
    cp_async_wait_group<0>();
    __syncthreads();
    // the ROWS row lanes of every quad, summed in lane order in fp64; one partial row block per CTA
    double* red = reinterpret_cast<double*>(l2_smem);                // [ROWS][WB] doubles = 8 KB
    double* out = partial + (int64_t)blockIdx.x * 5 * WB;
#pragma unroll
    for (int stat = 0; stat < 5; ++stat) {
        if (!FIRST && stat != 0 && stat != 2) continue;
#pragma unroll
        for (int c = 0; c < 4; ++c)
            red[(rl * QUADS + quad) * 4 + c] = stat == 0 ? (double)a_dot[c] : stat == 1 ? (double)a_ss[c] : stat == 2 ? (double)a_tt[c]
                                               : stat == 3 ? (double)a_s[c] : (double)a_nz[c];
        __syncthreads();
        if (tid < WB) {
            double t = 0.0;
            for (int r = 0; r < ROWS; ++r) t += red[r * WB + tid];
            out[stat * WB + tid] = t;
        }
        __syncthreads();
    }
}

// the column block as a COMPACT [n][QUADS] matrix: the gathers then address n*QUADS*16 contiguous bytes - inside the TLB
// reach (256 MB) and the L2 - instead of pieces strewn over the whole 4nP-byte matrix (measured: gathering the strided
// pieces in place is bound by address translation at ~4e10 requests/s, whatever their size)
template <int QUADS, bool HINT>
__global__ void compact_block_kernel(const float4* __restrict__ S4, int64_t ld4, int64_t n, int64_t q0, int nq,
                                     float4* __restrict__ blk) {
    const int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (e >= n * QUADS) return;
    const int64_t i = e / QUADS;
    const int q = (int)(e - i * QUADS);
    if (HINT) {
        const float4 v = q < nq ? ldg_f4_stream_policy(S4 + i * ld4 + q0 + q, l2_policy_stream()) : make_float4(0.f, 0.f, 0.f, 0.f);
        st_f4_policy(blk + e, v, l2_policy_keep());
    } else {
        blk[e] = q < nq ? ld_stream_f4(S4 + i * ld4 + q0 + q) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
}

// hands the block's L2 lines back: without this they would stay marked evict_last after the pass and shrink the L2 the
// following kernels see.  The buffer is scratch, its contents are dead.
__global__ void release_block_kernel(float4* __restrict__ blk, int64_t n128) {
    const int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (e < n128) asm volatile("discard.global.L2 [%0], 128;" ::"l"(reinterpret_cast<char*>(blk) + e * 128) : "memory");
}

// partial [nblk][R][ncta][5][WB] -> out rows [R][5][P]; sum S^2 / sum S / nnz come from repeat 0's launch
__global__ void __launch_bounds__(256)
null_l2_reduce_kernel(const double* __restrict__ partial, int nblk, int R, int ncta, int WB, int P, double* __restrict__ out) {
    const int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;           // (blk, r, stat, col)
    if (e >= (int64_t)nblk * R * 5 * WB) return;
    const int col = (int)(e % WB), stat = (int)((e / WB) % 5), r = (int)((e / (5 * WB)) % R), blk = (int)(e / ((int64_t)5 * WB * R));
    const int p = blk * WB + col;
    if (p >= P) return;
    const int src_r = (stat == 0 || stat == 2) ? r : 0;
    double t = 0.0;
    for (int c = 0; c < ncta; ++c) t += partial[((((int64_t)blk * R + src_r) * ncta + c) * 5 + stat) * WB + col];
    out[((int64_t)r * 5 + stat) * P + p] = t;
}

// ---------------------------------------------------------------------------------------------------------
// K4a tiled: observed-graph statistics with the neighbour rows staged in shared memory by the TMA engine.
//
// The graph is spatial and `order` walks the tissue, so the k neighbours of kTileCells consecutive cells are
// mostly the same few rows.  tile_plan_kernel finds, per tile, the DISTINCT rows (own rows included) and rewrites
// every edge as (slot in that list, weight).  spatial_obs_tile_kernel is persistent (one CTA per SM): a producer
// warp brings the 64-column pieces of a tile's distinct rows into shared memory with cp.async.bulk (one bulk copy
// per row piece, completion counted on an mbarrier), double buffered, while 16 consumer warps form
// T = sum_j w_j * row[slot_j] and the five sums from shared memory.  HBM / L2 then see each row piece about
// u/128 ~ 1.6 times per slab instead of 1+k times.  Work items (slab, tile) are dealt to CTAs as contiguous
// ranges; a CTA writes one partial record per slab segment and the records are reduced in a fixed order.
// ---------------------------------------------------------------------------------------------------------
constexpr int kTileCells = 128;
constexpr int kTileCap = 352;            // distinct rows staged per tile; larger unions read global rows directly
constexpr int kTileIds = 2048;           // kTileCells * (k + 1) <= kTileIds
constexpr int kTileMaxK = 15;
constexpr int kSlabCols = 64;            // columns per work item: 256-byte row pieces
constexpr int kTileConsumers = 512;      // 16 quads x 32 cell lanes
constexpr int kTileThreads = kTileConsumers + 32;
constexpr int kTileSelf = kTileCells + 4;          // self slots + the tile's row count (16-byte granule)

struct TilePlan {
    int32_t* cnt;      // [ntiles]                 distinct rows, or -1: the tile reads global rows directly
    int32_t* rows;     // [ntiles][kTileCap]       the distinct rows
    int2* ent;         // [ntiles][kTileCells][k]  {slot (or row when cnt < 0), weight bits}
    int32_t* self;     // [ntiles][kTileSelf]      slot (or row) of the cell itself, -1 = past the end; [kTileCells] = cnt
};

__global__ void __launch_bounds__(256)
tile_plan_kernel(const int32_t* __restrict__ idx, const float* __restrict__ w, int k, const int32_t* __restrict__ order,
                 int64_t n, TilePlan plan) {
    __shared__ int32_t arr[kTileIds];
    __shared__ int32_t uniq[kTileIds];
    __shared__ int32_t cell[kTileCells];
    __shared__ int32_t scan_s[33];
    const int tid = threadIdx.x;
    const int64_t tile = blockIdx.x;
    for (int c = tid; c < kTileCells; c += 256) {
        const int64_t q = tile * kTileCells + c;
        cell[c] = q < n ? (order ? order[q] : (int32_t)q) : -1;
    }
    __syncthreads();
    for (int e = tid; e < kTileIds; e += 256) {
        int32_t v = INT32_MAX;
        if (e < kTileCells) {
            if (cell[e] >= 0) v = cell[e];
        } else if (e < kTileCells * (k + 1)) {
            const int c = (e - kTileCells) / k, j = (e - kTileCells) - c * k;
            if (cell[c] >= 0) {
                const int32_t id = idx[(int64_t)cell[c] * k + j];
                if (id >= 0) v = id;
            }
        }
        arr[e] = v;
    }
    __syncthreads();
    for (int kk = 2; kk <= kTileIds; kk <<= 1)
        for (int j = kk >> 1; j > 0; j >>= 1) {
            for (int i = tid; i < kTileIds; i += 256) {
                const int l = i ^ j;
                if (l > i) {
                    const bool up = (i & kk) == 0;
                    const int32_t a = arr[i], b = arr[l];
                    if ((a > b) == up) { arr[i] = b; arr[l] = a; }
                }
            }
            __syncthreads();
        }
    // distinct values -> uniq (8 consecutive elements per thread, block scan of the counts)
    int flags = 0, cntl = 0;
    const int base = tid * 8;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const int i = base + j;
        const bool f = arr[i] != INT32_MAX && (i == 0 || arr[i] != arr[i - 1]);
        flags |= (f ? 1 : 0) << j;
        cntl += f;
    }
    int total;
    int pos = block_exclusive_scan<int>(cntl, &total, scan_s);
#pragma unroll
    for (int j = 0; j < 8; ++j)
        if (flags & (1 << j)) uniq[pos++] = arr[base + j];
    __syncthreads();
    const int u = total;
    const bool staged = u <= kTileCap;
    if (tid == 0) { plan.cnt[tile] = staged ? u : -1; plan.self[tile * kTileSelf + kTileCells] = staged ? u : -1; }
    if (staged)
        for (int r = tid; r < kTileCap; r += 256) plan.rows[tile * kTileCap + r] = r < u ? uniq[r] : 0;
    auto slot_of = [&](int32_t id) {
        int lo = 0, hi = u;
        while (lo < hi) {
            const int mid = (lo + hi) >> 1;
            if (uniq[mid] < id) lo = mid + 1; else hi = mid;
        }
        return lo;
    };
    for (int c = tid; c < kTileCells; c += 256)
        plan.self[tile * kTileSelf + c] = cell[c] < 0 ? -1 : (staged ? slot_of(cell[c]) : cell[c]);
    for (int e = tid; e < kTileCells * k; e += 256) {
        const int c = e / k, j = e - c * k;
        int2 out = make_int2(0, 0);
        if (cell[c] >= 0) {
            const int32_t id = idx[(int64_t)cell[c] * k + j];
            const float wv = w[(int64_t)cell[c] * k + j];
            if (id >= 0) out = make_int2(staged ? slot_of(id) : id, __float_as_int(wv));
            else out = make_int2(staged ? slot_of(cell[c]) : cell[c], 0);       // flagged edge: weight 0 on the own row
        }
        plan.ent[(tile * kTileCells + c) * k + j] = out;
    }
}

struct TileSmem {
    static __host__ __device__ size_t stage_bytes(int k) {
        return (((size_t)kTileCap * kSlabCols * 4 + (size_t)kTileCells * k * 8 + (size_t)kTileSelf * 4) + 127) & ~(size_t)127;
    }
    static __host__ __device__ size_t total(int k) { return 2 * stage_bytes(k) + 16 * 16 * 4 * 8 + 64; }
};

// per-thread running sums of a (column quad, cell lane) of the tile kernels
struct TileAcc {
    double dot[4] = {0, 0, 0, 0}, ss[4] = {0, 0, 0, 0}, tt[4] = {0, 0, 0, 0}, s[4] = {0, 0, 0, 0};
    int nz[4] = {0, 0, 0, 0};
};

// one work item (128 cells x 64 columns) from the staged tile: T = sum_j w_j * row[slot_j] and the five sums, fp32 over the
// four cells of this thread, fp64 across items.  `direct`: the tile's rows were not staged (union over the cap) - read global.
// KT > 0: k known at compile time - the edge words and the row quads of a cell are loaded in batches of 5 before the FMAs,
// so the two dependent shared-memory latencies per edge overlap instead of forming a chain (the pass is latency-bound:
// 16 warps per SM, ncu short-scoreboard stall 2.9 of 5.6 cycles per issue).
template <int KT>
__device__ __forceinline__ void tile_accumulate(TileAcc& A, const float4* __restrict__ tile_s, const int2* __restrict__ ent_s,
                                                const int32_t* __restrict__ self_s, const float* __restrict__ S, int64_t ldS,
                                                int64_t col0, int k, int quad, int cl) {
    if (!(col0 + quad * 4 < ldS)) return;
    const bool direct = self_s[kTileCells] < 0;
    float p_dot[4] = {0, 0, 0, 0}, p_ss[4] = {0, 0, 0, 0}, p_tt[4] = {0, 0, 0, 0}, p_s[4] = {0, 0, 0, 0};
#pragma unroll
    for (int m = 0; m < kTileCells / 32; ++m) {
        const int c = cl + 32 * m;
        const int sf = self_s[c];
        if (sf < 0) continue;
        const int2* e = ent_s + c * k;
        float4 s, t = make_float4(0.f, 0.f, 0.f, 0.f);
        if (!direct) {
            s = tile_s[sf * (kSlabCols / 4) + quad];
            if (KT > 0) {
                constexpr int B = 5;
#pragma unroll
                for (int j0 = 0; j0 < KT; j0 += B) {
                    int2 ej[B];
                    float4 v[B];
#pragma unroll
                    for (int j = 0; j < B; ++j) if (j0 + j < KT) ej[j] = e[j0 + j];
#pragma unroll
                    for (int j = 0; j < B; ++j) if (j0 + j < KT) v[j] = tile_s[ej[j].x * (kSlabCols / 4) + quad];
#pragma unroll
                    for (int j = 0; j < B; ++j)
                        if (j0 + j < KT) {
                            const float wj = __int_as_float(ej[j].y);
                            t.x = fmaf(wj, v[j].x, t.x); t.y = fmaf(wj, v[j].y, t.y); t.z = fmaf(wj, v[j].z, t.z); t.w = fmaf(wj, v[j].w, t.w);
                        }
                }
            } else {
                for (int j = 0; j < k; ++j) {
                    const int2 ej = e[j];
                    const float wj = __int_as_float(ej.y);
                    const float4 v = tile_s[ej.x * (kSlabCols / 4) + quad];
                    t.x = fmaf(wj, v.x, t.x); t.y = fmaf(wj, v.y, t.y); t.z = fmaf(wj, v.z, t.z); t.w = fmaf(wj, v.w, t.w);
                }
            }
        } else {
            const float4* S4 = reinterpret_cast<const float4*>(S + col0) + quad;
            s = ld_cached_f4(S4 + (int64_t)sf * (ldS / 4));
            for (int j = 0; j < k; ++j) {
                const int2 ej = e[j];
                const float wj = __int_as_float(ej.y);
                const float4 v = ld_cached_f4(S4 + (int64_t)ej.x * (ldS / 4));
                t.x = fmaf(wj, v.x, t.x); t.y = fmaf(wj, v.y, t.y); t.z = fmaf(wj, v.z, t.z); t.w = fmaf(wj, v.w, t.w);
            }
        }
        const float sv[4] = {s.x, s.y, s.z, s.w}, tv[4] = {t.x, t.y, t.z, t.w};
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            p_dot[q] = fmaf(sv[q], tv[q], p_dot[q]);
            p_ss[q] = fmaf(sv[q], sv[q], p_ss[q]);
            p_tt[q] = fmaf(tv[q], tv[q], p_tt[q]);
            p_s[q] += sv[q];
            A.nz[q] += sv[q] != 0.f ? 1 : 0;
        }
    }
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        A.dot[q] += (double)p_dot[q]; A.ss[q] += (double)p_ss[q]; A.tt[q] += (double)p_tt[q]; A.s[q] += (double)p_s[q];
    }
}

// slab segment finished: reduce the 32 cell lanes of every quad in a fixed order, write one record, clear the sums.
// Called by the kTileConsumers consumer threads together (named barrier 1).
__device__ __forceinline__ void tile_flush_record(TileAcc& A, double* red, double* out, int tid) {
    const int lane = tid & 31, warp = tid >> 5, quad = tid & 15;
    for (int stat = 0; stat < 5; ++stat) {
        double v[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            v[q] = stat == 0 ? A.dot[q] : stat == 1 ? A.ss[q] : stat == 2 ? A.tt[q] : stat == 3 ? A.s[q] : (double)A.nz[q];
            v[q] += __shfl_xor_sync(0xffffffffu, v[q], 16);                     // the warp's two cell lanes
        }
        asm volatile("bar.sync 1, %0;" ::"n"(kTileConsumers) : "memory");
        if (lane < 16) {
#pragma unroll
            for (int q = 0; q < 4; ++q) red[(warp * 16 + quad) * 4 + q] = v[q];
        }
        asm volatile("bar.sync 1, %0;" ::"n"(kTileConsumers) : "memory");
        if (tid < kSlabCols) {
            const int qd = tid >> 2, q = tid & 3;
            double t = 0.0;
            for (int wq = 0; wq < 16; ++wq) t += red[(wq * 16 + qd) * 4 + q];
            out[stat * kSlabCols + tid] = t;
        }
    }
    A = TileAcc();
}

__global__ void __launch_bounds__(kTileThreads, 1)
spatial_obs_tile_kernel(const float* __restrict__ S, int64_t ldS, int k, TilePlan plan, int64_t ntiles, int nslabs,
                        int maxseg, double* __restrict__ partial, int32_t* __restrict__ rec_slab,
                        const __grid_constant__ CUtensorMap tmap, int gather4) {
    extern __shared__ __align__(128) unsigned char smem[];
    const size_t stage_b = TileSmem::stage_bytes(k);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    double* red = reinterpret_cast<double*>(smem + 2 * stage_b);                 // [16 warps][16 quads][4]
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 2 * stage_b + 16 * 16 * 4 * 8);
    const uint32_t full0 = aio::smem_addr(bars), empty0 = aio::smem_addr(bars + 2);
    if (tid == 0) {
        aio::mbar_init(full0, 1); aio::mbar_init(full0 + 8, 1);
        aio::mbar_init(empty0, kTileConsumers / 32); aio::mbar_init(empty0 + 8, kTileConsumers / 32);
        aio::mbar_init_fence();
    }
    __syncthreads();
    const int64_t W = (int64_t)nslabs * ntiles;
    const int64_t w0 = W * blockIdx.x / gridDim.x, w1 = W * (blockIdx.x + 1) / gridDim.x;
    const uint32_t ent_bytes = (uint32_t)kTileCells * k * 8, self_bytes = kTileSelf * 4;

    if (warp == kTileConsumers / 32) {
        // ------------------------------ producer warp ------------------------------
        // The row list of the NEXT item is fetched (one batch of independent loads per lane) while this item's copies
        // are issued and the stage is awaited, so no global-load latency sits between two bulk copies.
        // gather4 != 0: rows travel four at a time through a 2-D tensor map of S (cp.async.bulk.tensor ... tile::gather4:
        // one TMA operation gathers the 64-column pieces of four arbitrary rows); else one 1-D bulk copy per row.
        constexpr int GPL = (kTileCap / 4 + 31) / 32;                           // row groups per lane
        int4 rows_nx[GPL];
        int cnt_nx = 0;
        auto fetch = [&](int64_t wi) {
            const int64_t tile = wi % ntiles;
            cnt_nx = plan.cnt[tile];
            const int4* rp = reinterpret_cast<const int4*>(plan.rows + tile * kTileCap);
#pragma unroll
            for (int j = 0; j < GPL; ++j) rows_nx[j] = lane + 32 * j < kTileCap / 4 ? rp[lane + 32 * j] : make_int4(0, 0, 0, 0);
        };
        if (w0 < w1) fetch(w0);
        for (int64_t wi = w0, it = 0; wi < w1; ++wi, ++it) {
            const int st = (int)(it & 1);
            const uint32_t ph = (uint32_t)((it >> 1) & 1);
            const int slab = (int)(wi / ntiles);
            const int64_t tile = wi - (int64_t)slab * ntiles;
            int4 rows_cur[GPL];
            const int cnt = cnt_nx;
#pragma unroll
            for (int j = 0; j < GPL; ++j) rows_cur[j] = rows_nx[j];
            if (wi + 1 < w1) fetch(wi + 1);
            aio::mbar_wait_sleep(empty0 + 8 * st, ph ^ 1u);
            unsigned char* stage = smem + st * stage_b;
            const uint32_t tile_s = aio::smem_addr(stage);
            const uint32_t ent_s = tile_s + kTileCap * kSlabCols * 4, self_s = ent_s + ent_bytes;
            const int64_t col0 = (int64_t)slab * kSlabCols;
            const uint32_t colbytes = (uint32_t)min((int64_t)kSlabCols, ldS - col0) * 4u;
            const uint32_t bar = full0 + 8 * st;
            const int groups = cnt > 0 ? (cnt + 3) >> 2 : 0;
            if (lane == 0) {
                const uint32_t row_bytes = gather4 ? (uint32_t)groups * (4u * kSlabCols * 4u) : (cnt > 0 ? (uint32_t)cnt * colbytes : 0u);
                aio::mbar_expect_tx(bar, ent_bytes + self_bytes + row_bytes);
                aio::bulk_g2s(ent_s, plan.ent + tile * kTileCells * k, ent_bytes, bar);
                aio::bulk_g2s(self_s, plan.self + tile * kTileSelf, self_bytes, bar);
            }
            __syncwarp();
#pragma unroll
            for (int j = 0; j < GPL; ++j) {
                const int g = lane + 32 * j;
                if (g >= groups) continue;
                const int4 r4 = rows_cur[j];
                const uint32_t dst = tile_s + (uint32_t)g * (4u * kSlabCols * 4u);
                if (gather4) {
                    aio::tma_gather4(dst, &tmap, (int)col0, r4.x, r4.y, r4.z, r4.w, bar);
                } else {
                    const int rr[4] = {r4.x, r4.y, r4.z, r4.w};
#pragma unroll
                    for (int q = 0; q < 4; ++q)
                        if (4 * g + q < cnt) aio::bulk_g2s(dst + (uint32_t)q * (kSlabCols * 4), S + (int64_t)rr[q] * ldS + col0, colbytes, bar);
                }
            }
        }
        return;
    }
    // --------------------------------- consumers ---------------------------------
    const int quad = tid & 15, cl = tid >> 4;                                   // 16 column quads x 32 cell lanes
    TileAcc A;
    int seg = 0;
    for (int64_t wi = w0, it = 0; wi < w1; ++wi, ++it) {
        const int st = (int)(it & 1);
        const uint32_t ph = (uint32_t)((it >> 1) & 1);
        const int slab = (int)(wi / ntiles);
        const int64_t col0 = (int64_t)slab * kSlabCols;
        aio::mbar_wait(full0 + 8 * st, ph);
        const unsigned char* stage = smem + st * stage_b;
        tile_accumulate<0>(A, reinterpret_cast<const float4*>(stage),
                        reinterpret_cast<const int2*>(stage + (size_t)kTileCap * kSlabCols * 4),
                        reinterpret_cast<const int32_t*>(stage + (size_t)kTileCap * kSlabCols * 4 + ent_bytes), S, ldS, col0, k,
                        quad, cl);
        __syncwarp();
        if (lane == 0) aio::mbar_arrive(empty0 + 8 * st);                       // this warp is done with the stage
        if (wi + 1 == w1 || (wi + 1) / ntiles != slab) {
            const int64_t rec = (int64_t)blockIdx.x * maxseg + seg;
            tile_flush_record(A, red, partial + rec * 5 * kSlabCols, tid);
            if (tid == 0) rec_slab[rec] = slab;
            ++seg;
        }
    }
}

// The same pass with the rows staged by cp.async (LDGSTS): no producer warp, no TMA operations - all 512 threads issue
// 16-byte global->shared copies (16 per 256-byte row piece).  A 1-D bulk copy costs the TMA engine ~71 cycles and a
// gather4 ~146 cycles per operation per SM (measured); LDGSTS sustains ~64 B/clk/SM.  Three plan buffers
// (row list, edges, own slots) and two tile buffers are software-pipelined with cp.async groups:
//   iteration `it`:  issue plan(it+2) | wait plan(it+1), barrier | issue rows(it+1) | wait rows(it), barrier | compute(it)
struct TileSmemCp {
    static __host__ __device__ size_t plan_bytes(int k) {
        return (((size_t)kTileCap * 4 + (size_t)kTileCells * k * 8 + (size_t)kTileSelf * 4) + 127) & ~(size_t)127;
    }
    static __host__ __device__ size_t tile_bytes() { return (size_t)kTileCap * kSlabCols * 4; }
    static __host__ __device__ size_t total(int k) { return 2 * tile_bytes() + 3 * plan_bytes(k) + 16 * 16 * 4 * 8; }
};

__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

template <int KT>
__global__ void __launch_bounds__(kTileConsumers, 1)
spatial_obs_tile_cpasync_kernel(const float* __restrict__ S, int64_t ldS, int k, TilePlan plan, int64_t ntiles, int nslabs,
                                int maxseg, double* __restrict__ partial, int32_t* __restrict__ rec_slab) {
    extern __shared__ __align__(128) unsigned char smem[];
    const size_t tile_b = TileSmemCp::tile_bytes(), plan_b = TileSmemCp::plan_bytes(k);
    const int tid = threadIdx.x;
    unsigned char* tile_base = smem;
    unsigned char* plan_base = smem + 2 * tile_b;
    double* red = reinterpret_cast<double*>(smem + 2 * tile_b + 3 * plan_b);
    const uint32_t rows_bytes = kTileCap * 4, ent_bytes = (uint32_t)kTileCells * k * 8, self_bytes = kTileSelf * 4;
    const int64_t W = (int64_t)nslabs * ntiles;
    const int64_t w0 = W * blockIdx.x / gridDim.x, w1 = W * (blockIdx.x + 1) / gridDim.x;

    auto issue_plan = [&](int64_t wi, int buf) {           // rows | ent | self of the item's tile -> plan buffer `buf`
        if (wi < w1) {
            const int64_t tile = wi % ntiles;
            const uint32_t dst = aio::smem_addr(plan_base + (size_t)buf * plan_b);
            const unsigned char* src_rows = reinterpret_cast<const unsigned char*>(plan.rows + tile * kTileCap);
            const unsigned char* src_ent = reinterpret_cast<const unsigned char*>(plan.ent + tile * kTileCells * k);
            const unsigned char* src_self = reinterpret_cast<const unsigned char*>(plan.self + tile * kTileSelf);
            const int n_rows = rows_bytes / 16, n_ent = ent_bytes / 16, n_self = self_bytes / 16;
            for (int c = tid; c < n_rows + n_ent + n_self; c += kTileConsumers) {
                if (c < n_rows) cp_async16(dst + 16 * c, src_rows + 16 * c, 16);
                else if (c < n_rows + n_ent) cp_async16(dst + 16 * c, src_ent + 16 * (c - n_rows), 16);
                else cp_async16(dst + 16 * c, src_self + 16 * (c - n_rows - n_ent), 16);
            }
        }
        cp_async_commit();
    };
    auto issue_rows = [&](int64_t wi, int pbuf, int tbuf) { // the tile's distinct row pieces, 16 copies of 16 bytes per row
        if (wi < w1) {
            const unsigned char* pl = plan_base + (size_t)pbuf * plan_b;
            const int32_t* rows_s = reinterpret_cast<const int32_t*>(pl);
            const int cnt = reinterpret_cast<const int32_t*>(pl + rows_bytes + ent_bytes)[kTileCells];
            const int64_t col0 = (wi / ntiles) * kSlabCols;
            const uint32_t dst = aio::smem_addr(tile_base + (size_t)tbuf * tile_b);
            for (int c = tid; c < cnt * 16; c += kTileConsumers) {
                const int r = c >> 4, piece = c & 15;
                const bool ok = col0 + piece * 4 < ldS;
                const float* src = S + (int64_t)rows_s[r] * ldS + (ok ? col0 + piece * 4 : 0);
                cp_async16(dst + (uint32_t)r * (kSlabCols * 4) + piece * 16, src, ok ? 16u : 0u);
            }
        }
        cp_async_commit();
    };

    const int quad = tid & 15, cl = tid >> 4;
    TileAcc A;
    int seg = 0;
    issue_plan(w0, 0);
    issue_plan(w0 + 1, 1);
    cp_async_wait<1>();                                     // plan(0)
    __syncthreads();
    issue_rows(w0, 0, 0);
    for (int64_t wi = w0, it = 0; wi < w1; ++wi, ++it) {
        issue_plan(wi + 2, (int)((it + 2) % 3));
        cp_async_wait<2>();                                 // plan(it+1): everything but the two youngest groups
        __syncthreads();
        issue_rows(wi + 1, (int)((it + 1) % 3), (int)((it + 1) & 1));
        cp_async_wait<2>();                                 // rows(it)
        __syncthreads();
        const int slab = (int)(wi / ntiles);
        const unsigned char* pl = plan_base + (size_t)(it % 3) * plan_b;
        tile_accumulate<KT>(A, reinterpret_cast<const float4*>(tile_base + (size_t)(it & 1) * tile_b),
                        reinterpret_cast<const int2*>(pl + rows_bytes), reinterpret_cast<const int32_t*>(pl + rows_bytes + ent_bytes),
                        S, ldS, (int64_t)slab * kSlabCols, k, quad, cl);
        if (wi + 1 == w1 || (wi + 1) / ntiles != slab) {
            const int64_t rec = (int64_t)blockIdx.x * maxseg + seg;
            tile_flush_record(A, red, partial + rec * 5 * kSlabCols, tid);
            if (tid == 0) rec_slab[rec] = slab;
            ++seg;
        }
        __syncthreads();                                    // the buffers of item `it` are free for items it+2 / it+3
    }
    cp_async_wait<0>();
}

// records [nrec][5][kSlabCols] (+ their slab) -> out [5][P]; one CTA per slab, records summed in index order
__global__ void __launch_bounds__(5 * kSlabCols)
tile_records_reduce_kernel(const double* __restrict__ partial, const int32_t* __restrict__ rec_slab, int nrec, int P,
                           double* __restrict__ out) {
    const int slab = blockIdx.x, stat = threadIdx.x / kSlabCols, c = threadIdx.x % kSlabCols;
    const int p = slab * kSlabCols + c;
    double t = 0.0;
    for (int r = 0; r < nrec; ++r)
        if (rec_slab[r] == slab) t += partial[((int64_t)r * 5 + stat) * kSlabCols + c];
    if (p < P) out[(int64_t)stat * P + p] = t;
}

__global__ void null_graph_kernel(int64_t n, int64_t m, const float* __restrict__ w, uint32_t repeat, uint64_t seed,
                                  FeistelKeys fk, int32_t* __restrict__ cols, float* __restrict__ wout) {
    const int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (e >= m) return;
    const u32x4 o = philox4x32_10((uint32_t)e, (uint32_t)((uint64_t)e >> 32), repeat, kStreamCol, seed);
    const uint64_t u = (uint64_t)o.x | ((uint64_t)o.y << 32);
    cols[e] = (int32_t)__umul64hi(u, (uint64_t)n);
    wout[e] = w[feistel_perm((uint64_t)e, (uint64_t)m, fk)];
}

}  // namespace laris

using namespace laris;

extern "C" int laris_spatial_stats(const float* S, int64_t ldS, int64_t n, int32_t P, const int32_t* idx, const float* w,
                                   int k, int l1_normalise, const int32_t* order, double* out, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    LARIS_CHECK_ARG(S && idx && w && out, "laris_spatial_stats: null pointer");
    LARIS_CHECK_ARG(n > 0 && P >= 1 && ldS >= P && ldS % 4 == 0 && ((uintptr_t)S & 15) == 0,
                    "laris_spatial_stats: S needs ldS >= P, ldS %% 4 == 0 and 16-byte alignment");
    LARIS_CHECK_ARG(k >= 1 && k <= kStMaxK, "laris_spatial_stats: k must be in [1,%d] (got %d)", kStMaxK, k);
    const int64_t ld4 = ldS / 4;
    const int slabs = (int)ceil_div(ld4, kStThreads);
    int dev = 0, sms = kNumSMs;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const bool local = order != nullptr && !l1_normalise;
    int64_t nchunks = ceil_div((int64_t)sms * (local ? 6 : 12), slabs);   // resident 128-thread CTAs per SM
    nchunks = nchunks < 1 ? 1 : nchunks;
    if (nchunks > ceil_div(n, kStRowBatch)) nchunks = ceil_div(n, kStRowBatch);
    if (nchunks > 65535) nchunks = 65535;
    Scratch<double> partial(st);
    LARIS_CUDA(partial.alloc((size_t)nchunks * 5 * ld4 * 4));
    dim3 grid((unsigned)slabs, (unsigned)nchunks);
    if (l1_normalise)
        spatial_stats_kernel<true, false><<<grid, kStThreads, 0, st>>>((const float4*)S, ld4, n, idx, w, k, order, (int)nchunks, partial.p);
    else if (local)
        spatial_stats_kernel<false, true><<<grid, kStThreads, 0, st>>>((const float4*)S, ld4, n, idx, w, k, order, (int)nchunks, partial.p);
    else
        spatial_stats_kernel<false, false><<<grid, kStThreads, 0, st>>>((const float4*)S, ld4, n, idx, w, k, order, (int)nchunks, partial.p);
    LARIS_LAUNCH_CHECK();
    stats_reduce_kernel<<<dim3((unsigned)ceil_div(P, 32), 5), 256, 0, st>>>(partial.p, (int)nchunks, ld4 * 4, P, out);
    LARIS_LAUNCH_CHECK();
    return LARIS_OK;
}

extern "C" int laris_null_graph_philox(int64_t n, int k, const float* w, int32_t repeat, uint64_t seed, int32_t* cols,
                                       float* wout, void* stream) {
    LARIS_CHECK_ARG(w && cols && wout && n > 0 && k >= 1 && n < (int64_t)INT32_MAX, "laris_null_graph_philox: bad arguments");
    const int64_t m = n * k;
    const FeistelKeys fk = make_feistel_keys((uint64_t)m, (uint32_t)repeat, seed);
    null_graph_kernel<<<(unsigned)ceil_div(m, 256), 256, 0, (cudaStream_t)stream>>>(n, m, w, (uint32_t)repeat, seed, fk, cols, wout);
    LARIS_LAUNCH_CHECK();
    return LARIS_OK;
}

// staging of the tile rows: 2 = cp.async 16-byte copies by all threads (default for k <= 12), 1 = TMA tile::gather4 through
// a tensor map (k 13..15, or on request), 0 = one 1-D bulk copy per row.  LARIS_B200_K4A_STAGING=bulk|gather4|cpasync
// overrides it (used to measure the three against each other, tools/time_k4.py).
static int tile_staging_mode() {
    static int mode = -1;
    if (mode < 0) {
        const char* e = getenv("LARIS_B200_K4A_STAGING");
        mode = !e ? 2 : strcmp(e, "bulk") == 0 ? 0 : strcmp(e, "gather4") == 0 ? 1 : 2;
    }
    return mode;
}

static int encode_row_gather_map(CUtensorMap* tmap, const float* S, int64_t ldS, int64_t n) {
    typedef CUresult (*EncodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                    const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    static EncodeTiled encode = nullptr;
    if (!encode) {
        void* fn = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q) != cudaSuccess || !fn) return -1;
        encode = (EncodeTiled)fn;
    }
    const cuuint64_t dims[2] = {(cuuint64_t)ldS, (cuuint64_t)n};
    const cuuint64_t strides[1] = {(cuuint64_t)ldS * sizeof(float)};
    const cuuint32_t box[2] = {(cuuint32_t)kSlabCols, 1u};
    const cuuint32_t estr[2] = {1u, 1u};
    const CUresult r = encode(tmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void*)S, dims, strides, box, estr,
                              CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? 0 : -1;
}

// observed statistics through the staged-tile kernel -> out [5][P]
static int observed_stats_tiled(const float* S, int64_t ldS, int64_t n, int32_t P, const int32_t* idx, const float* w, int k,
                                const int32_t* order, double* out, cudaStream_t st) {
    const int64_t ntiles = ceil_div(n, kTileCells);
    Scratch<int32_t> cnt(st), rows(st), self(st), rec_slab(st);
    Scratch<int2> ent(st);
    Scratch<double> partial(st);
    LARIS_CUDA(cnt.alloc((size_t)ntiles));
    LARIS_CUDA(rows.alloc((size_t)ntiles * kTileCap));
    LARIS_CUDA(self.alloc((size_t)ntiles * kTileSelf));
    LARIS_CUDA(ent.alloc((size_t)ntiles * kTileCells * k));
    TilePlan plan{cnt.p, rows.p, ent.p, self.p};
    tile_plan_kernel<<<(unsigned)ntiles, 256, 0, st>>>(idx, w, k, order, n, plan);
    LARIS_LAUNCH_CHECK();
    int dev = 0, sms = kNumSMs;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int nslabs = (int)ceil_div(ldS, kSlabCols);
    const int64_t W = (int64_t)nslabs * ntiles;
    const int grid = (int)(W < sms ? W : sms);
    const int maxseg = (int)ceil_div(nslabs, grid) + 2;
    const int nrec = grid * maxseg;
    LARIS_CUDA(partial.alloc((size_t)nrec * 5 * kSlabCols));
    LARIS_CUDA(rec_slab.alloc((size_t)nrec));
    LARIS_CUDA(cudaMemsetAsync(rec_slab.p, 0xff, sizeof(int32_t) * (size_t)nrec, st));
    int mode = tile_staging_mode();
    if (mode == 2 && TileSmemCp::total(k) <= (size_t)227 * 1024) {
        const size_t smem_cp = TileSmemCp::total(k);
        auto kern = k == 10 ? spatial_obs_tile_cpasync_kernel<10> : k == 8 ? spatial_obs_tile_cpasync_kernel<8>
                  : k == 6 ? spatial_obs_tile_cpasync_kernel<6> : spatial_obs_tile_cpasync_kernel<0>;
        LARIS_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_cp));
        kern<<<grid, kTileConsumers, smem_cp, st>>>(S, ldS, k, plan, ntiles, nslabs, maxseg, partial.p, rec_slab.p);
        LARIS_LAUNCH_CHECK();
        tile_records_reduce_kernel<<<nslabs, 5 * kSlabCols, 0, st>>>(partial.p, rec_slab.p, nrec, P, out);
        LARIS_LAUNCH_CHECK();
        return LARIS_OK;
    }
    if (mode == 2) mode = 1;                                      // k 13..15: the three plan buffers do not fit
    const size_t smem = TileSmem::total(k);
    LARIS_CUDA(cudaFuncSetAttribute(spatial_obs_tile_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    // 2-D tensor map of S for the gather4 staging: [n rows][ldS columns] fp32, box = 64 columns x 1 row (four rows per
    // operation come from the instruction's four row coordinates); columns past ldS are zero-filled by the TMA engine
    CUtensorMap tmap;
    memset(&tmap, 0, sizeof(tmap));
    int gather4 = mode;
    if (gather4) {
        const int rc = encode_row_gather_map(&tmap, S, ldS, n);
        if (rc) gather4 = 0;                                      // driver without tensor-map support: 1-D bulk copies
    }
    spatial_obs_tile_kernel<<<grid, kTileThreads, smem, st>>>(S, ldS, k, plan, ntiles, nslabs, maxseg, partial.p, rec_slab.p,
                                                              tmap, gather4);
    LARIS_LAUNCH_CHECK();
    tile_records_reduce_kernel<<<nslabs, 5 * kSlabCols, 0, st>>>(partial.p, rec_slab.p, nrec, P, out);
    LARIS_LAUNCH_CHECK();
    return LARIS_OK;
}

// bytes of one column block (n rows x WB columns) that still gather from L2 at full rate (probe: 128-byte rows hold
// 14.7 TB/s at 128 MB, 64-byte rows 16.8 TB/s at 64 MB and 9.8 TB/s at 96 MB)
static int null_l2_quads(int64_t n) {
    static int mode = -1;                 // LARIS_B200_K4B=hbm forces the streaming kernel, =q2|q4|q8 a block width (A/B runs)
    if (mode < 0) {
        const char* e = getenv("LARIS_B200_K4B");
        mode = !e ? 1 : strcmp(e, "hbm") == 0 ? 0 : strcmp(e, "q2") == 0 ? 2 : strcmp(e, "q4") == 0 ? 4 : strcmp(e, "q8") == 0 ? 8 : 1;
    }
    if (mode != 1) return mode;
    // ncu: a 128 MB block (1e6 rows x 128 B) misses L2 72 % of the time - data shared by the SMs of both dies is cached in
    // both L2 partitions, so the budget is ~half of the 126 MB (probe: 64-byte rows gather at 16.8 TB/s from 64 MB, 14.5 from
    // 80 MB, 9.8 from 96 MB).  Blocks are sized to <= 72 MB.
    if (n * 128 <= (int64_t)48 << 20) return 8;             // 32-column blocks (<= 393 k cells)
    if (n * 64 <= (int64_t)72 << 20) return 4;              // 16-column blocks (<= 1.18 M cells)
    return 0;   // larger n: 8-column blocks measured no faster than streaming (2e6 cells: 23.7 vs ~22 ms) - stream from HBM
}

static bool null_l2_hints() {             // LARIS_B200_K4B_HINTS=0 switches the L2 eviction hints off (A/B runs)
    static int on = -1;
    if (on < 0) {
        const char* e = getenv("LARIS_B200_K4B_HINTS");
        on = !(e && e[0] == '0');
    }
    return on != 0;
}

template <int QUADS, bool HINT>
static int null_stats_l2(const float* S, int64_t ldS, int64_t n, int32_t P, const int32_t* nidx, const float* nw, int k, int R,
                         double* out, cudaStream_t st) {
    constexpr int WB = QUADS * 4, ROWS = 256 / QUADS;
    const int64_t ld4 = ldS / 4, nk = n * (int64_t)k;
    int dev = 0, sms = kNumSMs;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    int64_t ncta = (int64_t)sms * 3;
    if (ncta < ceil_div(n, (int64_t)ROWS * 64)) ncta = ceil_div(n, (int64_t)ROWS * 64);   // <= 64 rows per thread: fp32 partials
    if (ncta > ceil_div(n, ROWS)) ncta = ceil_div(n, ROWS);
    const int nblk = (int)ceil_div(ld4, QUADS);
    Scratch<double> partial(st);
    Scratch<float4> blkbuf(st);
    LARIS_CUDA(partial.alloc((size_t)nblk * R * ncta * 5 * WB));
    LARIS_CUDA(blkbuf.alloc((size_t)n * QUADS));
    size_t smem = (size_t)4 * ROWS * k * 4;                                              // per warp: two buffers of (idx, w)
    if (smem < (size_t)ROWS * WB * 8) smem = (size_t)ROWS * WB * 8;
    LARIS_CUDA(cudaFuncSetAttribute(spatial_null_stats_l2_kernel<QUADS, true, HINT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    LARIS_CUDA(cudaFuncSetAttribute(spatial_null_stats_l2_kernel<QUADS, false, HINT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    for (int blk = 0; blk < nblk; ++blk) {
        const int64_t q0 = (int64_t)blk * QUADS;
        const int nq = (int)(ld4 - q0 < QUADS ? ld4 - q0 : QUADS);
        compact_block_kernel<QUADS, HINT><<<(unsigned)ceil_div(n * QUADS, 256), 256, 0, st>>>((const float4*)S, ld4, n, q0, nq, blkbuf.p);
        LARIS_LAUNCH_CHECK();
        for (int r = 0; r < R; ++r) {                      // the block's R*k gathers per cell, one repeat per launch, all L2 hits
            double* dst = partial.p + ((size_t)blk * R + r) * ncta * 5 * WB;
            if (r == 0)
                spatial_null_stats_l2_kernel<QUADS, true, HINT><<<(unsigned)ncta, 256, smem, st>>>(blkbuf.p, n, nidx, nw, k, dst);
            else
                spatial_null_stats_l2_kernel<QUADS, false, HINT><<<(unsigned)ncta, 256, smem, st>>>(blkbuf.p, n, nidx + (int64_t)r * nk,
                                                                                             nw + (int64_t)r * nk, k, dst);
            LARIS_LAUNCH_CHECK();
        }
    }
    null_l2_reduce_kernel<<<(unsigned)ceil_div((int64_t)nblk * R * 5 * WB, 256), 256, 0, st>>>(partial.p, nblk, R, (int)ncta, WB, P, out);
    LARIS_LAUNCH_CHECK();
    if (HINT) {
        const int64_t n128 = n * QUADS * 16 / 128;          // whole 128-byte lines of the block buffer
        if (n128 > 0) release_block_kernel<<<(unsigned)ceil_div(n128, 256), 256, 0, st>>>(blkbuf.p, n128);
        LARIS_LAUNCH_CHECK();
    }
    return LARIS_OK;
}

template <int RB>
static int null_stats_batch(const float* S, int64_t ldS, int64_t n, int32_t P, const int32_t* nidx, const float* nw, int k,
                            double* out, cudaStream_t st) {
    const int64_t ld4 = ldS / 4;
    const int slabs = (int)ceil_div(ld4, kStThreads);
    int dev = 0, sms = kNumSMs;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    int64_t nchunks = ceil_div((int64_t)sms * 8, slabs);
    if (nchunks > n) nchunks = n;
    if (nchunks > 65535) nchunks = 65535;
    constexpr int Q = 2 * RB + 3;
    Scratch<double> partial(st);
    LARIS_CUDA(partial.alloc((size_t)nchunks * Q * ld4 * 4));
    spatial_null_stats_kernel<RB><<<dim3((unsigned)slabs, (unsigned)nchunks), kStThreads, 0, st>>>(
        (const float4*)S, ld4, n, nidx, nw, k, (int)nchunks, partial.p);
    LARIS_LAUNCH_CHECK();
    null_stats_reduce_kernel<<<dim3((unsigned)ceil_div(P, 32), Q), 256, 0, st>>>(partial.p, (int)nchunks, ld4 * 4, P, RB, out);
    LARIS_LAUNCH_CHECK();
    return LARIS_OK;
}

extern "C" int laris_spatial_stats_batched(const float* S, int64_t ldS, int64_t n, int32_t P, const int32_t* idx,
                                           const float* w, int k, const int32_t* order, const int32_t* null_idx,
                                           const float* null_w, int32_t n_repeats, double* out, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    LARIS_CHECK_ARG(S && out && (idx || n_repeats > 0), "laris_spatial_stats_batched: null pointer");
    LARIS_CHECK_ARG(n > 0 && P >= 1 && ldS >= P && ldS % 4 == 0 && ((uintptr_t)S & 15) == 0,
                    "laris_spatial_stats_batched: S needs ldS >= P, ldS %% 4 == 0 and 16-byte alignment");
    LARIS_CHECK_ARG(k >= 1 && k <= kStMaxK, "laris_spatial_stats_batched: k must be in [1,%d] (got %d)", kStMaxK, k);
    LARIS_CHECK_ARG(n_repeats >= 0 && (n_repeats == 0 || (null_idx && null_w)), "laris_spatial_stats_batched: bad null graphs");
    int row0 = 0;
    if (idx) {
        LARIS_CHECK_ARG(w, "laris_spatial_stats_batched: null weights");
        int rc;
        if (k <= kTileMaxK) rc = observed_stats_tiled(S, ldS, n, P, idx, w, k, order, out, st);
        else rc = laris_spatial_stats(S, ldS, n, P, idx, w, k, 0, order, out, stream);
        if (rc) return rc;
        row0 = 1;
    }
    const int64_t nk = n * (int64_t)k;
    const int quads = n_repeats > 0 ? null_l2_quads(n) : 0;
    if (quads) {
        double* o = out + (int64_t)row0 * 5 * P;
        if (null_l2_hints())
            return quads == 8   ? null_stats_l2<8, true>(S, ldS, n, P, null_idx, null_w, k, n_repeats, o, st)
                   : quads == 4 ? null_stats_l2<4, true>(S, ldS, n, P, null_idx, null_w, k, n_repeats, o, st)
                                : null_stats_l2<2, true>(S, ldS, n, P, null_idx, null_w, k, n_repeats, o, st);
        return quads == 8   ? null_stats_l2<8, false>(S, ldS, n, P, null_idx, null_w, k, n_repeats, o, st)
               : quads == 4 ? null_stats_l2<4, false>(S, ldS, n, P, null_idx, null_w, k, n_repeats, o, st)
                            : null_stats_l2<2, false>(S, ldS, n, P, null_idx, null_w, k, n_repeats, o, st);
    }
    for (int r = 0; r < n_repeats;) {
        const int rb = n_repeats - r >= 3 ? 3 : n_repeats - r;
        double* o = out + (int64_t)(row0 + r) * 5 * P;
        const int32_t* ni = null_idx + (int64_t)r * nk;
        const float* nw = null_w + (int64_t)r * nk;
        int rc = rb == 3 ? null_stats_batch<3>(S, ldS, n, P, ni, nw, k, o, st)
               : rb == 2 ? null_stats_batch<2>(S, ldS, n, P, ni, nw, k, o, st)
                         : null_stats_batch<1>(S, ldS, n, P, ni, nw, k, o, st);
        if (rc) return rc;
        r += rb;
    }
    return LARIS_OK;
}
