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
#include "common.cuh"
#include "rng.cuh"

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
