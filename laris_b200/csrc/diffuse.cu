// K2+K3: fused neighbour-aggregated expression + per-cell LR score, plus the
// gene-selection and CSR (de)compaction kernels either side of it.
//
// Data layout in HBM
//   Xu      packed CSR over the G_u genes the LR table uses: indptr int64[n+1],
//           entries int64 = {lo: int32 column in [0,G_u), hi: fp32 bits}
//   graph   idx int32[n*k], w fp32[n*k]  (K1 output, fixed degree)
//   S       dense fp32 [n][ldS], cell-major (a row = all pair scores of a cell,
//           which is the unit every downstream kernel gathers)
//
// Kernel shape (diffuse_score_kernel): a CTA owns a tile of TC spatially
// adjacent cells (consecutive entries of the K1 processing order).
//   phase 1  each warp owns some of the tile's cells and streams the packed
//            sparse rows of that cell's k neighbours (64-bit coalesced loads,
//            software-prefetched one neighbour ahead) into a shared-memory
//            tile D[g][cell] (row stride TC+1 so both phases are conflict-light)
//   phase 2  lane = cell: for each LR pair both operands D[lig][cell],
//            D[rec][cell] are conflict-free shared loads, the expression mask is
//            a broadcast word per gene (bit per cell); each lane emits 8
//            consecutive pairs of its cell as two 128-bit streaming stores.
// D never touches HBM.  Algorithmic bytes: 8*nnz(Xu) + 8*n*k + 4*n*P.
#include "common.cuh"

namespace laris {

// ------------------------------------------------------- gene selection --
__global__ void csr_select_count_kernel(const int64_t* __restrict__ indptr, const int32_t* __restrict__ indices,
                                        const float* __restrict__ data, int64_t n, const int32_t* __restrict__ gene_map,
                                        int64_t* __restrict__ counts) {
    const int lane = threadIdx.x & 31;
    const int64_t row = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    if (row >= n) return;
    const int64_t s = indptr[row], e = indptr[row + 1];
    int c = 0;
    for (int64_t t = s + lane; t < e; t += 32) c += (gene_map[indices[t]] >= 0 && data[t] != 0.0f) ? 1 : 0;
    c = warp_sum(c);
    if (lane == 0) counts[row] = c;
}

__global__ void csr_select_fill_kernel(const int64_t* __restrict__ indptr, const int32_t* __restrict__ indices,
                                       const float* __restrict__ data, int64_t n, const int32_t* __restrict__ gene_map,
                                       const int64_t* __restrict__ out_indptr, int64_t* __restrict__ out) {
    const int lane = threadIdx.x & 31;
    const int64_t row = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    if (row >= n) return;
    const int64_t s = indptr[row], e = indptr[row + 1];
    int64_t o = out_indptr[row];
    for (int64_t t0 = s; t0 < e; t0 += 32) {
        const int64_t t = t0 + lane;
        int g = -1;
        float v = 0.0f;
        if (t < e) { g = gene_map[indices[t]]; v = data[t]; }
        const bool keep = g >= 0 && v != 0.0f;
        const unsigned m = __ballot_sync(0xffffffffu, keep);
        if (keep) {
            const int pos = __popc(m & ((1u << lane) - 1u));
            out[o + pos] = (int64_t)(((uint64_t)__float_as_uint(v) << 32) | (uint32_t)g);
        }
        o += __popc(m);
    }
}

// ---------------------------------------------------------------- K2+K3 --
constexpr int kDsThreads = 256;
constexpr int kDsWarps = kDsThreads / 32;
constexpr int kPrefetch = 4;                       // packed entries per lane held in flight per neighbour

template <int TC>
__global__ void __launch_bounds__(kDsThreads)
diffuse_score_kernel(const int64_t* __restrict__ xu_indptr, const int64_t* __restrict__ xu_packed, int64_t n, int G_u,
                     const int32_t* __restrict__ idx, const float* __restrict__ w, int k,
                     const int32_t* __restrict__ order, const int32_t* __restrict__ lig, const int32_t* __restrict__ rec,
                     int P, float* __restrict__ S, int64_t ldS, int64_t* __restrict__ row_nnz) {
    constexpr int STRIDE = TC + 1;
    constexpr int SUB = 32 / TC;                   // pair groups handled side by side by one warp
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int P8 = (int)((ldS + 7) / 8) * 8;       // pair slots rounded to the 8-wide store groups
    const int d_words = (G_u * STRIDE + 3) & ~3, f_words = (G_u + 3) & ~3;   // keep every region 16-byte aligned
    float* D = reinterpret_cast<float*>(smem_raw);                         // [G_u][STRIDE]
    uint32_t* flags = reinterpret_cast<uint32_t*>(D + d_words);             // [G_u]  bit c = cell c expresses g
    uint32_t* pairs = flags + f_words;                                      // [P8]   lig | rec << 16, 0xffffffff = pad
    __shared__ int32_t cell_row[TC];
    __shared__ int32_t cell_cnt[TC];

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int p = tid; p < P8; p += kDsThreads)
        pairs[p] = p < P ? ((uint32_t)lig[p] | ((uint32_t)rec[p] << 16)) : 0xffffffffu;

    const int64_t ntiles = (n + TC - 1) / TC;
    for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        __syncthreads();                                            // previous tile fully consumed
        {   // zero the D tile and the flag words (contiguous, 128-bit stores)
            float4* z = reinterpret_cast<float4*>(D);
            const int nz4 = (d_words + f_words) / 4;                // D and flags are adjacent
            for (int q = tid; q < nz4; q += kDsThreads) z[q] = make_float4(0.f, 0.f, 0.f, 0.f);
        }
        if (tid < TC) {
            const int64_t q = tile * TC + tid;
            cell_row[tid] = q < n ? (order ? order[q] : (int32_t)q) : -1;
            cell_cnt[tid] = 0;
        }
        __syncthreads();

        // ---------------- phase 1: D[g][c] = sum_j w[c,j] * Xu[nb(c,j), g] ----------------
        for (int c = warp; c < TC; c += kDsWarps) {
            const int64_t i = cell_row[c];
            if (i < 0) continue;
            // expression flags of the cell itself
            {
                const int64_t s = xu_indptr[i], e = xu_indptr[i + 1];
                for (int64_t t = s + lane; t < e; t += 32) {
                    const uint32_t g = (uint32_t)(xu_packed[t] & 0xffffffffll);
                    atomicOr(&flags[g], 1u << c);
                }
            }
            for (int j0 = 0; j0 < k; j0 += 32) {
                const int kk = min(32, k - j0);
                int64_t ns = 0, ne = 0;
                float ws = 0.f;
                if (lane < kk) {
                    const int64_t nb = idx[i * k + j0 + lane];
                    ws = w[i * k + j0 + lane];
                    ns = xu_indptr[nb];
                    ne = xu_indptr[nb + 1];
                }
                int64_t buf[kPrefetch];
                {   // prefetch neighbour 0
                    const int64_t s0 = __shfl_sync(0xffffffffu, ns, 0), e0 = __shfl_sync(0xffffffffu, ne, 0);
#pragma unroll
                    for (int u = 0; u < kPrefetch; ++u) {
                        const int64_t t = s0 + lane + 32 * u;
                        buf[u] = t < e0 ? xu_packed[t] : -1ll;
                    }
                }
                for (int j = 0; j < kk; ++j) {
                    const int64_t sj = __shfl_sync(0xffffffffu, ns, j), ej = __shfl_sync(0xffffffffu, ne, j);
                    const float wj = __shfl_sync(0xffffffffu, ws, j);
                    int64_t nxt[kPrefetch];
                    if (j + 1 < kk) {
                        const int64_t s1 = __shfl_sync(0xffffffffu, ns, j + 1), e1 = __shfl_sync(0xffffffffu, ne, j + 1);
#pragma unroll
                        for (int u = 0; u < kPrefetch; ++u) {
                            const int64_t t = s1 + lane + 32 * u;
                            nxt[u] = t < e1 ? xu_packed[t] : -1ll;
                        }
                    } else {
#pragma unroll
                        for (int u = 0; u < kPrefetch; ++u) nxt[u] = -1ll;
                    }
#pragma unroll
                    for (int u = 0; u < kPrefetch; ++u) {
                        if (buf[u] != -1ll) {
                            const uint32_t g = (uint32_t)(buf[u] & 0xffffffffll);
                            const float v = __uint_as_float((uint32_t)((uint64_t)buf[u] >> 32));
                            float* d = D + (size_t)g * STRIDE + c;
                            *d = fmaf(wj, v, *d);
                        }
                    }
                    for (int64_t t = sj + 32 * kPrefetch + lane; t < ej; t += 32) {     // rows longer than the buffer
                        const int64_t pk = xu_packed[t];
                        const uint32_t g = (uint32_t)(pk & 0xffffffffll);
                        const float v = __uint_as_float((uint32_t)((uint64_t)pk >> 32));
                        float* d = D + (size_t)g * STRIDE + c;
                        *d = fmaf(wj, v, *d);
                    }
                    __syncwarp();                                   // next neighbour may hit the same genes
#pragma unroll
                    for (int u = 0; u < kPrefetch; ++u) buf[u] = nxt[u];
                }
            }
        }
        __syncthreads();

        // ---------------- phase 2: S[c][p] = D[l][c] * D[r][c] * (flag_l | flag_r)[c] ----------------
        {
            const int c = lane % TC, sub = lane / TC;
            const int64_t i = cell_row[c];
            int cnt = 0;
            for (int p0 = (warp * SUB + sub) * 8; p0 < P8; p0 += kDsWarps * SUB * 8) {
                const uint4 pa = *reinterpret_cast<const uint4*>(pairs + p0);
                const uint4 pb = *reinterpret_cast<const uint4*>(pairs + p0 + 4);
                const uint32_t pr[8] = {pa.x, pa.y, pa.z, pa.w, pb.x, pb.y, pb.z, pb.w};
                float v[8];
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    float val = 0.f;
                    if (pr[u] != 0xffffffffu) {
                        const uint32_t l = pr[u] & 0xffffu, r = pr[u] >> 16;
                        const float a = D[(size_t)l * STRIDE + c], b = D[(size_t)r * STRIDE + c];
                        const uint32_t m = (flags[l] | flags[r]) >> c;
                        val = (m & 1u) ? a * b : 0.f;
                    }
                    v[u] = val;
                    cnt += val != 0.f ? 1 : 0;
                }
                if (i >= 0) {
                    float* out = S + i * ldS + p0;
                    if (p0 + 4 <= ldS) st_stream_f4(reinterpret_cast<float4*>(out), make_float4(v[0], v[1], v[2], v[3]));
                    if (p0 + 8 <= ldS) st_stream_f4(reinterpret_cast<float4*>(out + 4), make_float4(v[4], v[5], v[6], v[7]));
                }
            }
            if (row_nnz) atomicAdd(&cell_cnt[c], cnt);
        }
        if (row_nnz) {
            __syncthreads();
            if (tid < TC && cell_row[tid] >= 0) row_nnz[cell_row[tid]] = cell_cnt[tid];
        }
    }
}

// --------------------------------------------------------- CSR <-> dense --
__global__ void csr_compact_kernel(const float* __restrict__ S, int64_t ldS, int64_t n, int P,
                                   const int64_t* __restrict__ indptr, int32_t* __restrict__ out_idx,
                                   float* __restrict__ out_val) {
    const int lane = threadIdx.x & 31;
    const int64_t row = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    if (row >= n) return;
    const float* s = S + row * ldS;
    int64_t o = indptr[row];
    for (int p0 = 0; p0 < P; p0 += 32) {
        const int p = p0 + lane;
        const float v = p < P ? s[p] : 0.f;
        const unsigned m = __ballot_sync(0xffffffffu, v != 0.f);
        if (v != 0.f) {
            const int64_t pos = o + __popc(m & ((1u << lane) - 1u));
            out_idx[pos] = p;
            out_val[pos] = v;
        }
        o += __popc(m);
    }
}

__global__ void csr_scatter_kernel(const int64_t* __restrict__ indptr, const int32_t* __restrict__ indices,
                                   const float* __restrict__ data, int64_t n, float* __restrict__ S, int64_t ldS) {
    const int lane = threadIdx.x & 31;
    const int64_t row = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    if (row >= n) return;
    for (int64_t t = indptr[row] + lane; t < indptr[row + 1]; t += 32) S[row * ldS + indices[t]] = data[t];
}

template <int TC>
static int launch_diffuse(const int64_t* xu_indptr, const int64_t* xu_packed, int64_t n, int G_u, const int32_t* idx,
                          const float* w, int k, const int32_t* order, const int32_t* lig, const int32_t* rec, int P,
                          float* S, int64_t ldS, int64_t* row_nnz, size_t smem, cudaStream_t st) {
    auto kern = diffuse_score_kernel<TC>;
    LARIS_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int per_sm = 1;
    LARIS_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, kDsThreads, smem));
    if (per_sm < 1) per_sm = 1;
    int dev = 0, sms = kNumSMs;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int64_t ntiles = (n + TC - 1) / TC;
    const int64_t grid = ntiles < (int64_t)sms * per_sm ? ntiles : (int64_t)sms * per_sm;
    kern<<<(unsigned)grid, kDsThreads, smem, st>>>(xu_indptr, xu_packed, n, G_u, idx, w, k, order, lig, rec, P, S, ldS,
                                                   row_nnz);
    LARIS_LAUNCH_CHECK();
    return LARIS_OK;
}

}  // namespace laris

using namespace laris;

extern "C" int laris_csr_select_count(const int64_t* indptr, const int32_t* indices, const float* data, int64_t n,
                                      const int32_t* gene_map, int64_t* out_counts, void* stream) {
    LARIS_CHECK_ARG(indptr && gene_map && out_counts && n >= 0, "laris_csr_select_count: bad arguments");
    if (n == 0) return LARIS_OK;
    csr_select_count_kernel<<<(unsigned)ceil_div(n * 32, 256), 256, 0, (cudaStream_t)stream>>>(indptr, indices, data, n,
                                                                                             gene_map, out_counts);
    LARIS_LAUNCH_CHECK();
    return LARIS_OK;
}

extern "C" int laris_csr_select_fill(const int64_t* indptr, const int32_t* indices, const float* data, int64_t n,
                                     const int32_t* gene_map, const int64_t* out_indptr, int64_t* out_packed,
                                     void* stream) {
    LARIS_CHECK_ARG(indptr && gene_map && out_indptr && n >= 0, "laris_csr_select_fill: bad arguments");
    if (n == 0) return LARIS_OK;
    csr_select_fill_kernel<<<(unsigned)ceil_div(n * 32, 256), 256, 0, (cudaStream_t)stream>>>(
        indptr, indices, data, n, gene_map, out_indptr, out_packed);
    LARIS_LAUNCH_CHECK();
    return LARIS_OK;
}

extern "C" int laris_diffuse_lr_score(const int64_t* xu_indptr, const int64_t* xu_packed, int64_t n, int32_t G_u,
                                      const int32_t* idx, const float* w, int k, const int32_t* order,
                                      const int32_t* lig, const int32_t* rec, int32_t P, float* S, int64_t ldS,
                                      int64_t* row_nnz, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    LARIS_CHECK_ARG(xu_indptr && idx && w && lig && rec && S, "laris_diffuse_lr_score: null pointer");
    LARIS_CHECK_ARG(n > 0 && k >= 1 && P >= 1, "laris_diffuse_lr_score: empty problem");
    LARIS_CHECK_ARG(G_u >= 1 && G_u <= 65535, "laris_diffuse_lr_score: G_u must be in [1, 65535] (got %d)", G_u);
    LARIS_CHECK_ARG(ldS >= P && ldS % 4 == 0 && ((uintptr_t)S & 15) == 0,
                    "laris_diffuse_lr_score: S needs ldS >= P, ldS %% 4 == 0 and 16-byte alignment");
    const size_t P8 = (size_t)((ldS + 7) / 8) * 8;
    auto need = [&](int tc) {
        return ((((size_t)G_u * (tc + 1) + 3) & ~(size_t)3) + (((size_t)G_u + 3) & ~(size_t)3) + P8) * 4;
    };
    const size_t two_per_sm = 113 * 1024 - 1024, one_per_sm = 227 * 1024 - 1024;
    int tc = 0;
    for (int cand : {32, 16, 8}) if (need(cand) <= two_per_sm) { tc = cand; break; }
    if (!tc) for (int cand : {32, 16, 8}) if (need(cand) <= one_per_sm) { tc = cand; break; }
    if (!tc) {
        set_error("laris_diffuse_lr_score: G_u=%d, P=%d do not fit the shared-memory tile (limit ~6000 genes)", G_u, P);
        return LARIS_E_LIMIT;
    }
#define ARGS xu_indptr, xu_packed, n, G_u, idx, w, k, order, lig, rec, P, S, ldS, row_nnz, need(tc), st
    if (tc == 32) return launch_diffuse<32>(ARGS);
    if (tc == 16) return launch_diffuse<16>(ARGS);
    return launch_diffuse<8>(ARGS);
#undef ARGS
}

extern "C" int laris_csr_compact(const float* S, int64_t ldS, int64_t n, int32_t P, const int64_t* indptr,
                                 int32_t* out_indices, float* out_data, void* stream) {
    LARIS_CHECK_ARG(S && indptr && n >= 0 && P >= 1 && ldS >= P, "laris_csr_compact: bad arguments");
    if (n == 0) return LARIS_OK;
    csr_compact_kernel<<<(unsigned)ceil_div(n * 32, 256), 256, 0, (cudaStream_t)stream>>>(S, ldS, n, P, indptr,
                                                                                        out_indices, out_data);
    LARIS_LAUNCH_CHECK();
    return LARIS_OK;
}

extern "C" int laris_csr_to_dense(const int64_t* indptr, const int32_t* indices, const float* data, int64_t n, int32_t P,
                                  float* S, int64_t ldS, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    LARIS_CHECK_ARG(indptr && S && n >= 0 && ldS >= P, "laris_csr_to_dense: bad arguments");
    if (n == 0) return LARIS_OK;
    LARIS_CUDA(cudaMemsetAsync(S, 0, (size_t)n * ldS * sizeof(float), st));
    csr_scatter_kernel<<<(unsigned)ceil_div(n * 32, 256), 256, 0, st>>>(indptr, indices, data, n, S, ldS);
    LARIS_LAUNCH_CHECK();
    return LARIS_OK;
}
