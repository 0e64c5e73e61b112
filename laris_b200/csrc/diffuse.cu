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
// Shared-memory layout of one CTA (all regions 16-byte aligned):
//   D      [(G_u+1)][TC+1] fp32   neighbour-aggregated expression of the tile, gene-major;
//                                 row G_u is a zero "dummy gene" that padding pairs point at
//   flags  [G_u+1]         u32    bit c set <=> cell c of the tile expresses the gene
//   pairs  [P8]            u32    lig | rec << 16  (padding: dummy | dummy << 16)
//   nb_ptr [TC][KC] i64, nb_len [TC][KC] i32, nb_w [TC][KC] f32   neighbour rows of the tile
//   self_ptr [TC] i64, self_len [TC] i32, cell_row [TC] i32, cell_cnt [TC] i32
constexpr int kDsKC = 16;                          // neighbours staged per pass (k > 16 takes several passes)

struct DsLayout {
    int d_words, f_words, p_words;
    size_t off_flags, off_pairs, off_nbptr, off_nblen, off_nbw, off_selfptr, off_selflen, off_row, off_cnt, total;
    __host__ __device__ DsLayout(int G_u, int64_t ldS, int TC) {
        d_words = ((G_u + 1) * (TC + 1) + 3) & ~3;
        f_words = (G_u + 1 + 3) & ~3;
        p_words = (int)((ldS + 7) / 8) * 8;
        size_t o = (size_t)d_words * 4;
        off_flags = o; o += (size_t)f_words * 4;
        off_pairs = o; o += (size_t)p_words * 4;
        o = (o + 7) & ~(size_t)7;
        off_nbptr = o; o += (size_t)TC * kDsKC * 8;
        off_selfptr = o; o += (size_t)TC * 8;
        off_nblen = o; o += (size_t)TC * kDsKC * 4;
        off_nbw = o; o += (size_t)TC * kDsKC * 4;
        off_selflen = o; o += (size_t)TC * 4;
        off_row = o; o += (size_t)TC * 4;
        off_cnt = o; o += (size_t)TC * 4;
        total = (o + 15) & ~(size_t)15;
    }
};

__device__ __forceinline__ int2 ld_entry(const int2* p) {
    int2 r;
    asm volatile("ld.global.nc.v2.s32 {%0,%1}, [%2];" : "=r"(r.x), "=r"(r.y) : "l"(p));
    return r;
}

template <int TC, int NT>
__global__ void __launch_bounds__(NT)
diffuse_score_kernel(const int64_t* __restrict__ xu_indptr, const int64_t* __restrict__ xu_packed, int64_t n, int G_u,
                     const int32_t* __restrict__ idx, const float* __restrict__ w, int k,
                     const int32_t* __restrict__ order, const int32_t* __restrict__ lig, const int32_t* __restrict__ rec,
                     int P, float* __restrict__ S, int64_t ldS, int64_t* __restrict__ row_nnz) {
    constexpr int STRIDE = TC + 1;
    constexpr int SUB = 32 / TC;                   // pair groups handled side by side by one warp
    constexpr int NW = NT / 32;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const DsLayout L(G_u, ldS, TC);
    float* D = reinterpret_cast<float*>(smem_raw);
    uint32_t* flags = reinterpret_cast<uint32_t*>(smem_raw + L.off_flags);
    uint32_t* pairs = reinterpret_cast<uint32_t*>(smem_raw + L.off_pairs);
    int64_t* nb_ptr = reinterpret_cast<int64_t*>(smem_raw + L.off_nbptr);
    int32_t* nb_len = reinterpret_cast<int32_t*>(smem_raw + L.off_nblen);
    float* nb_w = reinterpret_cast<float*>(smem_raw + L.off_nbw);
    int64_t* self_ptr = reinterpret_cast<int64_t*>(smem_raw + L.off_selfptr);
    int32_t* self_len = reinterpret_cast<int32_t*>(smem_raw + L.off_selflen);
    int32_t* cell_row = reinterpret_cast<int32_t*>(smem_raw + L.off_row);
    int32_t* cell_cnt = reinterpret_cast<int32_t*>(smem_raw + L.off_cnt);

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t pad_pair = (uint32_t)G_u | ((uint32_t)G_u << 16);
    for (int p = tid; p < L.p_words; p += NT)
        pairs[p] = p < P ? ((uint32_t)lig[p] | ((uint32_t)rec[p] << 16)) : pad_pair;
    const int2* __restrict__ entries = reinterpret_cast<const int2*>(xu_packed);

    const int64_t ntiles = (n + TC - 1) / TC;
    for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        __syncthreads();                                            // previous tile fully consumed
        {   // zero the D tile and the flag words (adjacent regions, 128-bit stores)
            float4* z = reinterpret_cast<float4*>(D);
            const int nz4 = (L.d_words + L.f_words) / 4;
            for (int q = tid; q < nz4; q += NT) z[q] = make_float4(0.f, 0.f, 0.f, 0.f);
        }
        if (tid < TC) {
            const int64_t q = tile * TC + tid;
            const int32_t i = q < n ? (order ? order[q] : (int32_t)q) : -1;
            cell_row[tid] = i;
            cell_cnt[tid] = 0;
            int64_t s = 0, e = 0;
            if (i >= 0) { s = xu_indptr[i]; e = xu_indptr[i + 1]; }
            self_ptr[tid] = s;
            self_len[tid] = (int)(e - s);
        }
        __syncthreads();

        // ---------------- phase 1: D[g][c] = sum_j w[c,j] * Xu[nb(c,j), g] ----------------
        for (int j0 = 0; j0 < k; j0 += kDsKC) {
            const int kc = min(kDsKC, k - j0);
            if (j0) __syncthreads();
            for (int e = tid; e < TC * kc; e += NT) {                // neighbour table of this pass
                const int c = e / kc, j = e - c * kc;
                const int64_t i = cell_row[c];
                int64_t s = 0;
                int len = 0;
                float wj = 0.f;
                if (i >= 0) {
                    const int64_t nb = idx[i * k + j0 + j];
                    wj = w[i * k + j0 + j];
                    s = xu_indptr[nb];
                    len = (int)(xu_indptr[nb + 1] - s);
                }
                nb_ptr[c * kDsKC + j] = s;
                nb_len[c * kDsKC + j] = len;
                nb_w[c * kDsKC + j] = wj;
            }
            __syncthreads();
            for (int c = warp; c < TC; c += NW) {
                float* Dc = D + c;
                if (j0 == 0) {                                       // expression flags of the cell itself
                    const int2* row = entries + self_ptr[c];
                    const int len = self_len[c];
                    for (int t = lane; t < len; t += 32) atomicOr(&flags[ld_entry(row + t).x], 1u << c);
                }
                for (int j = 0; j < kc; ++j) {
                    const int2* row = entries + nb_ptr[c * kDsKC + j];
                    const int len = nb_len[c * kDsKC + j];
                    const float wj = nb_w[c * kDsKC + j];
                    for (int t = lane; t < len; t += 128) {          // 4 independent 64-bit loads in flight per lane
                        const bool v1 = t + 32 < len, v2 = t + 64 < len, v3 = t + 96 < len;
                        const int2 e0 = ld_entry(row + t);
                        int2 e1 = make_int2(0, 0), e2 = make_int2(0, 0), e3 = make_int2(0, 0);
                        if (v1) e1 = ld_entry(row + t + 32);
                        if (v2) e2 = ld_entry(row + t + 64);
                        if (v3) e3 = ld_entry(row + t + 96);
                        { float* d = Dc + e0.x * STRIDE; *d = fmaf(wj, __int_as_float(e0.y), *d); }
                        if (v1) { float* d = Dc + e1.x * STRIDE; *d = fmaf(wj, __int_as_float(e1.y), *d); }
                        if (v2) { float* d = Dc + e2.x * STRIDE; *d = fmaf(wj, __int_as_float(e2.y), *d); }
                        if (v3) { float* d = Dc + e3.x * STRIDE; *d = fmaf(wj, __int_as_float(e3.y), *d); }
                    }
                    __syncwarp();                                    // the next neighbour may hit the same genes
                }
            }
        }
        __syncthreads();

        // ---------------- phase 2: S[c][p] = D[l][c] * D[r][c] * (flag_l | flag_r)[c] ----------------
        {
            const int c = lane % TC, sub = lane / TC;
            const int64_t i = cell_row[c];
            const float* Dc = D + c;
            float* srow = S + (i >= 0 ? i : 0) * ldS;
            int cnt = 0;
            for (int p0 = (warp * SUB + sub) * 8; p0 < L.p_words; p0 += NW * SUB * 8) {
                const uint4 pa = *reinterpret_cast<const uint4*>(pairs + p0);
                const uint4 pb = *reinterpret_cast<const uint4*>(pairs + p0 + 4);
                const uint32_t pr[8] = {pa.x, pa.y, pa.z, pa.w, pb.x, pb.y, pb.z, pb.w};
                float v[8];
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    const int l = (int)(pr[u] & 0xffffu), r = (int)(pr[u] >> 16);
                    const float prod = Dc[l * STRIDE] * Dc[r * STRIDE];
                    const uint32_t m = ((flags[l] | flags[r]) >> c) & 1u;
                    v[u] = m ? prod : 0.f;
                    cnt += v[u] != 0.f;
                }
                if (i >= 0) {
                    if (p0 + 4 <= ldS) st_stream_f4(reinterpret_cast<float4*>(srow + p0), make_float4(v[0], v[1], v[2], v[3]));
                    if (p0 + 8 <= ldS) st_stream_f4(reinterpret_cast<float4*>(srow + p0 + 4), make_float4(v[4], v[5], v[6], v[7]));
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

template <int TC, int NT>
static int launch_diffuse(const int64_t* xu_indptr, const int64_t* xu_packed, int64_t n, int G_u, const int32_t* idx,
                          const float* w, int k, const int32_t* order, const int32_t* lig, const int32_t* rec, int P,
                          float* S, int64_t ldS, int64_t* row_nnz, size_t smem, int per_sm, cudaStream_t st) {
    auto kern = diffuse_score_kernel<TC, NT>;
    LARIS_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int occ = 1;
    LARIS_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, NT, smem));
    if (occ < 1) occ = 1;
    if (occ < per_sm) per_sm = occ;
    int dev = 0, sms = kNumSMs;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int64_t ntiles = (n + TC - 1) / TC;
    const int64_t grid = ntiles < (int64_t)sms * per_sm ? ntiles : (int64_t)sms * per_sm;
    kern<<<(unsigned)grid, NT, smem, st>>>(xu_indptr, xu_packed, n, G_u, idx, w, k, order, lig, rec, P, S, ldS, row_nnz);
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
    // tile shape: the largest tile that still leaves >= 6 CTAs (48 warps) per SM; when the gene
    // set is too large for that, the largest tile with >= 2 CTAs and 512-thread CTAs to keep warps up
    const size_t smem_sm = 227 * 1024;
    auto need = [&](int tc) { return DsLayout(G_u, ldS, tc).total; };
    int tc = 0, nt = 256, per_sm = 1;
    for (int cand : {32, 16, 8}) if ((need(cand) + 1024) * 6 <= smem_sm) { tc = cand; per_sm = (int)(smem_sm / (need(cand) + 1024)); break; }
    if (!tc) for (int cand : {32, 16, 8}) if ((need(cand) + 1024) * 2 <= smem_sm) { tc = cand; nt = 512; per_sm = (int)(smem_sm / (need(cand) + 1024)); break; }
    if (!tc) for (int cand : {16, 8}) if (need(cand) + 1024 <= smem_sm) { tc = cand; nt = 512; per_sm = 1; break; }
    if (!tc) {
        set_error("laris_diffuse_lr_score: G_u=%d, P=%d do not fit the shared-memory tile (limit ~6000 genes)", G_u, P);
        return LARIS_E_LIMIT;
    }
    if (per_sm > 8) per_sm = 8;
#define ARGS xu_indptr, xu_packed, n, G_u, idx, w, k, order, lig, rec, P, S, ldS, row_nnz, need(tc), per_sm, st
    if (nt == 256) {
        if (tc == 32) return launch_diffuse<32, 256>(ARGS);
        if (tc == 16) return launch_diffuse<16, 256>(ARGS);
        return launch_diffuse<8, 256>(ARGS);
    }
    if (tc == 32) return launch_diffuse<32, 512>(ARGS);
    if (tc == 16) return launch_diffuse<16, 512>(ARGS);
    return launch_diffuse<8, 512>(ARGS);
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
