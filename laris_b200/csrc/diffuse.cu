// K2+K3: fused neighbour-aggregated expression + per-cell LR score, plus the
// gene-selection and CSR (de)compaction kernels either side of it.
//
// Data layout in HBM
//   Xu      packed CSR over the columns the LR table uses: indptr int64[n+1],
//           entries int64 = {lo: int32 column in [0,G_u), hi: fp32 bits}
//   graph   idx int32[n*k], w fp32[n*k]  (K1 output, fixed degree)
//   S       dense fp32 [n][ldS], cell-major (a row = all pair scores of a cell,
//           which is the unit every downstream kernel gathers)
//
// Kernel shape (diffuse_score_kernel): ONE WARP OWNS ONE CELL, no CTA barrier in
// the loop.  The warp keeps the cell's neighbour-aggregated expression row
// D[0..G_u) in a private slab of shared memory.
//   phase 1  the k neighbour rows are streamed as a flat list of 32-entry
//            chunks (64-bit coalesced ld.global.nc, 4 chunks in flight per
//            lane); each lane adds w*x into D[column] (columns of one row are
//            distinct, so the read-modify-write needs no atomics)
//   flags    the cell's own row marks "expressed" in the SIGN BIT of D (all
//            contributions non-negative, the normal case), else in a bitmap
//   phase 2  lane owns 4 consecutive pairs: two shared loads per score, and
//            the warp writes 512 contiguous bytes of the cell's S row per step
// D never touches HBM.  Algorithmic bytes: 8*nnz(Xu) + 8*n*k + 4*n*P.
// Shared-memory bank conflicts: phase 2 reads D[lig[p]] for 32 different pairs
// at once; laris_lr_column_layout() numbers the columns so that those land in
// 32 different banks (graph colouring of the LR table).
#include "common.cuh"
#include "complex.cuh"
#include <stdlib.h>

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

// single pass: the kept entries of row r go to out[indptr[r] ...) - the row keeps its offset in the INPUT matrix, so no
// count / scan / second read is needed; row_end[r] marks where the kept entries stop.  Costs nnz(X) slots of scratch.
__global__ void csr_select_rows_kernel(const int64_t* __restrict__ indptr, const int32_t* __restrict__ indices,
                                       const float* __restrict__ data, int64_t n, const int32_t* __restrict__ gene_map,
                                       int64_t* __restrict__ out, int64_t* __restrict__ row_end) {
    const int lane = threadIdx.x & 31;
    const int64_t row = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    if (row >= n) return;
    const int64_t s = indptr[row], e = indptr[row + 1];
    int64_t o = s;
    for (int64_t t0 = s; t0 < e; t0 += 32) {
        const int64_t t = t0 + lane;
        int g = -1;
        float v = 0.0f;
        if (t < e) { g = gene_map[indices[t]]; v = data[t]; }
        const bool keep = g >= 0 && v != 0.0f;
        const unsigned m = __ballot_sync(0xffffffffu, keep);
        if (keep) out[o + __popc(m & ((1u << lane) - 1u))] = (int64_t)(((uint64_t)__float_as_uint(v) << 32) | (uint32_t)g);
        o += __popc(m);
    }
    if (lane == 0) row_end[row] = o;
}

// ---------------------------------------------------------------- K2+K3 --
// Shared memory of one CTA:  [pair table p_words u32, only when the table does not fit registers]
// then per warp  D [dw_words] fp32 (slot G_u = always-zero column of the padding pairs)  and
// B [bw_words] u32 (expression bitmap, only touched for signed data).
__device__ __forceinline__ int2 ld_entry(const int2* p) {
    int2 r;
    asm volatile("ld.global.nc.v2.s32 {%0,%1}, [%2];" : "=r"(r.x), "=r"(r.y) : "l"(p));
    return r;
}

struct DsLayout {
    int dw_words, bw_words;
    size_t total;
    __host__ __device__ DsLayout(int G_u, int table_words, int nw) {
        dw_words = (G_u + 1 + 3) & ~3;
        bw_words = ((G_u + 1 + 31) / 32 + 3) & ~3;
        total = ((size_t)table_words + (size_t)nw * (dw_words + bw_words)) * 4;
    }
};

// One neighbour row in flight: CH chunks of 32 packed entries (lane-strided), its weight, and the
// number of entries left counted from this lane (chunk c holds a valid entry iff 32c < rem).
template <int CH>
struct RowBuf {
    int2 e[CH];
    float w;
    int rem;
};

template <int CH>
__device__ __forceinline__ void row_load(RowBuf<CH>& b, const int2* rowp, int len, float wv, int j, int lane) {
    constexpr unsigned FULL = 0xffffffffu;
    const int2* p = reinterpret_cast<const int2*>(__shfl_sync(FULL, (unsigned long long)rowp, j)) + lane;
    b.rem = __shfl_sync(FULL, len, j) - lane;
    b.w = __shfl_sync(FULL, wv, j);
#pragma unroll
    for (int c = 0; c < CH; ++c)
        if (32 * c < b.rem) b.e[c] = ld_entry(p + 32 * c);
}

// D[col] += w * x for the row's entries.  Columns of one row are distinct, so the chunks of a row need
// no ordering among themselves; the trailing __syncwarp orders this row against the next one.
template <int CH>
__device__ __forceinline__ void row_accumulate(const RowBuf<CH>& b, float* Dw, uint32_t& signs) {
#pragma unroll
    for (int c = 0; c < CH; ++c)
        if (32 * c < b.rem) {
            float* dp = Dw + b.e[c].x;
            *dp = fmaf(b.w, __int_as_float(b.e[c].y), *dp);
            signs |= (uint32_t)b.e[c].y;
        }
    __syncwarp();
}

// one score: D[l]*D[r] gated by "ligand or receptor expressed in this cell"
template <bool SIGNED>
__device__ __forceinline__ float lr_score1(const float* Dw, const uint32_t* Bw, uint32_t pr, int& cnt) {
    const uint32_t l = pr & 0xffffu, r = pr >> 16;
    const float a = Dw[l], b = Dw[r];
    float v;
    if (SIGNED) {
        const uint32_t m = ((Bw[l >> 5] >> (l & 31)) | (Bw[r >> 5] >> (r & 31))) & 1u;
        v = m ? a * b : 0.f;
    } else {                                                         // flag = sign bit of D (D >= 0)
        const bool m = (int)(__float_as_uint(a) | __float_as_uint(b)) < 0;
        v = m ? fabsf(a * b) : 0.f;
    }
    cnt += v != 0.f;
    return v;
}

// NT threads, CH entry chunks per row pass, PPL LR pairs per lane held in registers (0: table in smem)
template <int NT, int CH, int PPL>
__global__ void __launch_bounds__(NT)
diffuse_score_kernel(const int64_t* __restrict__ row_start, const int64_t* __restrict__ row_end,
                     const int64_t* __restrict__ xu_packed, int64_t n, int G_u,
                     const int32_t* __restrict__ idx, const float* __restrict__ w, int k,
                     const int32_t* __restrict__ order, const uint32_t* __restrict__ pairs, int p_words,
                     float* __restrict__ S, int64_t ldS, int64_t* __restrict__ row_nnz, ComplexTable cx) {
    constexpr int NW = NT / 32;
    constexpr unsigned FULL = 0xffffffffu;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int table_words = PPL ? 0 : p_words;
    const DsLayout L(G_u, table_words, NW);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    uint32_t* pairs_s = reinterpret_cast<uint32_t*>(smem_raw);
    float* Dw = reinterpret_cast<float*>(pairs_s + table_words) + (size_t)warp * (L.dw_words + L.bw_words);
    uint32_t* Bw = reinterpret_cast<uint32_t*>(Dw + L.dw_words);
    uint32_t pr[PPL ? PPL : 1];
    if (PPL) {                                                       // this lane's pairs never change: registers
#pragma unroll
        for (int g = 0; g < PPL / 4; ++g) {
            const uint4 t = *reinterpret_cast<const uint4*>(pairs + 128 * g + 4 * lane);   // table padded to 32*PPL
            pr[4 * g] = t.x; pr[4 * g + 1] = t.y; pr[4 * g + 2] = t.z; pr[4 * g + 3] = t.w;
        }
    } else {
        for (int p = tid; p < p_words; p += NT) pairs_s[p] = pairs[p];
        __syncthreads();
    }
    const int2* __restrict__ entries = reinterpret_cast<const int2*>(xu_packed);

    const int64_t nwarps = (int64_t)gridDim.x * NW;
    for (int64_t q = (int64_t)blockIdx.x * NW + warp; q < n; q += nwarps) {
        const int64_t i = order ? (int64_t)order[q] : q;
        for (int t = lane * 4; t < L.dw_words; t += 128) *reinterpret_cast<float4*>(Dw + t) = make_float4(0.f, 0.f, 0.f, 0.f);
        // the cell's own row: its first 32*CH columns are fetched now and consumed after phase 1
        const int64_t self_s = row_start[i];
        const int self_len = (int)(row_end[i] - self_s);
        int selfcol[CH];
#pragma unroll
        for (int c = 0; c < CH; ++c)
            if (lane + 32 * c < self_len) selfcol[c] = ld_entry(entries + self_s + lane + 32 * c).x;
        __syncwarp();

        // ---------------- phase 1: D[g] = sum_j w[j] * Xu[nb(j), g] ----------------
        // Rows are consumed in neighbour order; the loads of row j+1 are in flight while row j is added.
        uint32_t signs = 0;                                          // sign bit set <=> a negative value or weight was seen
        for (int j0 = 0; j0 < k; j0 += 32) {
            const int kc = min(32, k - j0);
            const int2* rowp = entries;
            int len = 0;
            float wv = 0.f;
            if (lane < kc) {
                const int64_t nb = idx[i * k + j0 + lane];
                wv = w[i * k + j0 + lane];
                const int64_t rs = row_start[nb];
                len = (int)(row_end[nb] - rs);
                rowp = entries + rs;
            }
            signs |= __float_as_uint(wv);
            RowBuf<CH> A, B;
            row_load(A, rowp, len, wv, 0, lane);
            for (int j = 0; j < kc; j += 2) {
                if (j + 1 < kc) row_load(B, rowp, len, wv, j + 1, lane);
                row_accumulate(A, Dw, signs);
                if (j + 2 < kc) row_load(A, rowp, len, wv, j + 2, lane);
                if (j + 1 < kc) row_accumulate(B, Dw, signs);
            }
            if (__any_sync(FULL, len > 32 * CH)) {                   // rare: rows longer than one pass
                for (int j = 0; j < kc; ++j) {
                    const int lj = __shfl_sync(FULL, len, j);
                    if (lj <= 32 * CH) continue;
                    const int2* p = reinterpret_cast<const int2*>(__shfl_sync(FULL, (unsigned long long)rowp, j));
                    const float wj = __shfl_sync(FULL, wv, j);
                    for (int t = 32 * CH + lane; t < lj; t += 32) {
                        const int2 e = ld_entry(p + t);
                        float* dp = Dw + e.x;
                        *dp = fmaf(wj, __int_as_float(e.y), *dp);
                        signs |= (uint32_t)e.y;
                    }
                    __syncwarp();
                }
            }
        }
        const bool signed_data = __any_sync(FULL, (signs >> 31) != 0u);

        // ---------------- expression flags of the cell itself ----------------
        if (!signed_data) {                                          // D >= 0: the sign bit is free to carry the flag
#pragma unroll
            for (int c = 0; c < CH; ++c)
                if (lane + 32 * c < self_len) Dw[selfcol[c]] = -fabsf(Dw[selfcol[c]]);
            for (int t = 32 * CH + lane; t < self_len; t += 32) { float* dp = Dw + ld_entry(entries + self_s + t).x; *dp = -fabsf(*dp); }
        } else {
            for (int t = lane * 4; t < L.bw_words; t += 128) *reinterpret_cast<uint4*>(Bw + t) = make_uint4(0u, 0u, 0u, 0u);
            __syncwarp();
            for (int t = lane; t < self_len; t += 32) {
                const int g = ld_entry(entries + self_s + t).x;
                atomicOr(&Bw[g >> 5], 1u << (g & 31));
            }
        }
        __syncwarp();
        if (cx.nc) {                                                 // multi-subunit complexes -> their virtual columns
            if (signed_data) complex_pass<true>(cx, Dw, Bw, lane); else complex_pass<false>(cx, Dw, Bw, lane);
            __syncwarp();
        }

        // ---------------- phase 2: S[p] = D[l] * D[r] * (flag_l | flag_r) ----------------
        // lane owns 4 consecutive pairs of every 128: the warp stores 512 contiguous bytes per step
        float* srow = S + i * ldS;
        int cnt = 0;
        if (PPL) {
#pragma unroll
            for (int g = 0; g < PPL / 4; ++g) {
                float v[4];
                if (!signed_data) {
#pragma unroll
                    for (int u = 0; u < 4; ++u) v[u] = lr_score1<false>(Dw, Bw, pr[4 * g + u], cnt);
                } else {
#pragma unroll
                    for (int u = 0; u < 4; ++u) v[u] = lr_score1<true>(Dw, Bw, pr[4 * g + u], cnt);
                }
                const int p0 = 128 * g + 4 * lane;
                if (p0 < (int)ldS) st_stream_f4(reinterpret_cast<float4*>(srow + p0), make_float4(v[0], v[1], v[2], v[3]));
            }
        } else {
            for (int p0 = lane * 4; p0 < (int)ldS; p0 += 128) {
                const uint4 pr4 = *reinterpret_cast<const uint4*>(pairs_s + p0);
                const uint32_t q4[4] = {pr4.x, pr4.y, pr4.z, pr4.w};
                float v[4];
                if (!signed_data) {
#pragma unroll
                    for (int u = 0; u < 4; ++u) v[u] = lr_score1<false>(Dw, Bw, q4[u], cnt);
                } else {
#pragma unroll
                    for (int u = 0; u < 4; ++u) v[u] = lr_score1<true>(Dw, Bw, q4[u], cnt);
                }
                st_stream_f4(reinterpret_cast<float4*>(srow + p0), make_float4(v[0], v[1], v[2], v[3]));
            }
        }
        if (row_nnz) {
            cnt = __reduce_add_sync(FULL, cnt);
            if (lane == 0) row_nnz[i] = cnt;
        }
        __syncwarp();                                                // D is re-zeroed for the next cell
    }
}

__global__ void pack_pairs_kernel(const int32_t* __restrict__ lig, const int32_t* __restrict__ rec, int P, int G_u,
                                  int p_words, uint32_t* __restrict__ pairs) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p < p_words) pairs[p] = p < P ? ((uint32_t)lig[p] | ((uint32_t)rec[p] << 16)) : ((uint32_t)G_u | ((uint32_t)G_u << 16));
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

static int diffuse_lr_score_impl(const int64_t* row_start, const int64_t* row_end, const int64_t* xu_packed, int64_t n,
                                 int32_t G_u,
                                 const int32_t* idx, const float* w, int k, const int32_t* order,
                                 const int32_t* lig, const int32_t* rec, int32_t P, float* S, int64_t ldS,
                                 int64_t* row_nnz, ComplexTable cx, cudaStream_t st) {
    LARIS_CHECK_ARG(row_start && row_end && idx && w && lig && rec && S, "laris_diffuse_lr_score: null pointer");
    LARIS_CHECK_ARG(n > 0 && k >= 1 && P >= 1, "laris_diffuse_lr_score: empty problem");
    LARIS_CHECK_ARG(G_u >= 1 && G_u <= 65533, "laris_diffuse_lr_score: G_u must be in [1, 65533] (got %d)", G_u);
    LARIS_CHECK_ARG(ldS >= P && ldS % 4 == 0 && ((uintptr_t)S & 15) == 0,
                    "laris_diffuse_lr_score: S needs ldS >= P, ldS %% 4 == 0 and 16-byte alignment");
    // pack the pair table once: lig | rec << 16, padded to whole 128-pair groups with the zero "dummy gene" G_u
    const int p_words = (int)((ldS + 127) / 128) * 128;
    Scratch<uint32_t> pairs(st);
    LARIS_CUDA(pairs.alloc((size_t)p_words));
    pack_pairs_kernel<<<(unsigned)ceil_div(p_words, 256), 256, 0, st>>>(lig, rec, P, G_u, p_words, pairs.p);
    LARIS_LAUNCH_CHECK();
    // 8 warps per CTA, each with a private D slab; persistent grid = SMs x resident CTAs.
    // CH: entry chunks per row pass; large gene panels have longer rows.
    constexpr int NT = 256;
    const int ppl = 0;   // register-resident pair tables cost too many registers next to the row pipeline
    const int ch = G_u > 1024 ? 8 : 4;
    const DsLayout L(G_u, ppl ? 0 : p_words, NT / 32);
    if (L.total + 1024 > (size_t)227 * 1024) {
        set_error("laris_diffuse_lr_score: G_u=%d, P=%d do not fit shared memory (limit ~6500 columns)", G_u, P);
        return LARIS_E_LIMIT;
    }
    using Kern = void (*)(const int64_t*, const int64_t*, const int64_t*, int64_t, int, const int32_t*, const float*, int,
                          const int32_t*, const uint32_t*, int, float*, int64_t, int64_t*, ComplexTable);
    Kern kern = nullptr;
#define DS_PICK(CH_, PPL_) if (ch == CH_ && ppl == PPL_) kern = diffuse_score_kernel<NT, CH_, PPL_>;
    DS_PICK(4, 0) DS_PICK(8, 0)
#undef DS_PICK
    LARIS_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L.total));
    int occ = 1;
    LARIS_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, NT, L.total));
    if (occ < 1) occ = 1;
    int dev = 0, sms = kNumSMs;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int64_t want = ceil_div(n, NT / 32);
    const int64_t grid = want < (int64_t)sms * occ ? want : (int64_t)sms * occ;
    kern<<<(unsigned)grid, NT, L.total, st>>>(row_start, row_end, xu_packed, n, G_u, idx, w, k, order, pairs.p, p_words, S,
                                              ldS, row_nnz, cx);
    LARIS_LAUNCH_CHECK();
    return LARIS_OK;
}

extern "C" int laris_diffuse_lr_score(const int64_t* xu_indptr, const int64_t* xu_packed, int64_t n, int32_t G_u,
                                      const int32_t* idx, const float* w, int k, const int32_t* order,
                                      const int32_t* lig, const int32_t* rec, int32_t P, float* S, int64_t ldS,
                                      int64_t* row_nnz, void* stream) {
    LARIS_CHECK_ARG(xu_indptr, "laris_diffuse_lr_score: null pointer");
    return diffuse_lr_score_impl(xu_indptr, xu_indptr + 1, xu_packed, n, G_u, idx, w, k, order, lig, rec, P, S, ldS, row_nnz,
                                 ComplexTable(), (cudaStream_t)stream);
}

extern "C" int laris_diffuse_lr_score_complex(const int64_t* xu_indptr, const int64_t* xu_packed, int64_t n, int32_t G_u,
                                              const int32_t* idx, const float* w, int k, const int32_t* order,
                                              const int32_t* lig, const int32_t* rec, int32_t P, float* S, int64_t ldS,
                                              int64_t* row_nnz, const int32_t* cx_ptr, const int32_t* cx_cols,
                                              const int32_t* cx_vcol, int32_t n_complex, int32_t rule, void* stream) {
    LARIS_CHECK_ARG(n_complex >= 0 && (n_complex == 0 || (cx_ptr && cx_cols && cx_vcol)) && (rule == 0 || rule == 1),
                    "laris_diffuse_lr_score_complex: bad complex table (n_complex=%d, rule=%d)", n_complex, rule);
    ComplexTable cx;
    cx.ptr = cx_ptr; cx.cols = cx_cols; cx.vcol = cx_vcol; cx.nc = n_complex; cx.rule = rule;
    LARIS_CHECK_ARG(xu_indptr, "laris_diffuse_lr_score_complex: null pointer");
    return diffuse_lr_score_impl(xu_indptr, xu_indptr + 1, xu_packed, n, G_u, idx, w, k, order, lig, rec, P, S, ldS, row_nnz,
                                 cx, (cudaStream_t)stream);
}

extern "C" int laris_diffuse_lr_score_rows(const int64_t* row_start, const int64_t* row_end, const int64_t* xu_packed,
                                           int64_t n, int32_t G_u, const int32_t* idx, const float* w, int k,
                                           const int32_t* order, const int32_t* lig, const int32_t* rec, int32_t P, float* S,
                                           int64_t ldS, int64_t* row_nnz, const int32_t* cx_ptr, const int32_t* cx_cols,
                                           const int32_t* cx_vcol, int32_t n_complex, int32_t rule, void* stream) {
    LARIS_CHECK_ARG(n_complex >= 0 && (n_complex == 0 || (cx_ptr && cx_cols && cx_vcol)) && (rule == 0 || rule == 1),
                    "laris_diffuse_lr_score_rows: bad complex table (n_complex=%d, rule=%d)", n_complex, rule);
    ComplexTable cx;
    cx.ptr = cx_ptr; cx.cols = cx_cols; cx.vcol = cx_vcol; cx.nc = n_complex; cx.rule = rule;
    return diffuse_lr_score_impl(row_start, row_end, xu_packed, n, G_u, idx, w, k, order, lig, rec, P, S, ldS, row_nnz, cx,
                                 (cudaStream_t)stream);
}

extern "C" int laris_csr_select_rows(const int64_t* indptr, const int32_t* indices, const float* data, int64_t n,
                                     const int32_t* gene_map, int64_t* out_packed, int64_t* row_end, void* stream) {
    LARIS_CHECK_ARG(indptr && gene_map && out_packed && row_end && n >= 0, "laris_csr_select_rows: bad arguments");
    if (n == 0) return LARIS_OK;
    csr_select_rows_kernel<<<(unsigned)ceil_div(n * 32, 256), 256, 0, (cudaStream_t)stream>>>(indptr, indices, data, n,
                                                                                            gene_map, out_packed, row_end);
    LARIS_LAUNCH_CHECK();
    return LARIS_OK;
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
