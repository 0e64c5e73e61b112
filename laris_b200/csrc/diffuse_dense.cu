// K2+K3, dense-panel variant: fused neighbour-aggregated expression + per-cell LR score for
// expression panels that are dense enough (>= ~8 % non-zeros over the LR genes) that a dense fp32
// row costs fewer L1 wavefronts than a packed sparse row plus its shared-memory scatter.
//
// Data layout in HBM
//   Xd      dense fp32 [n][ldX] over the packed LR columns, ldX % 128 == 0, ldX > G_u; slot G_u and
//           every slot above it are zero (slot G_u is the "dummy gene" of the padding pairs)
//   graph   idx int32[n*k], w fp32[n*k]  (K1 output, fixed degree)
//   S       dense fp32 [n][ldS], cell-major
//
// Kernel shape (diffuse_dense_kernel): ONE WARP OWNS R = min(3, 32/k) CONSECUTIVE CELLS of the spatial
// order.  Their R*k graph edges sit one per lane; __match_any_sync finds the distinct neighbour rows
// (adjacent cells share most of their neighbours), so each distinct row is loaded ONCE, as 128-bit
// coalesced loads (lane owns 4 consecutive columns of every 128), and fused-multiply-added into the
// R cells' accumulators in REGISTERS with that row's R weights.  Phase 1 therefore costs no shared
// memory traffic at all and ~1/R of the L1 traffic of a per-cell gather.
//   flags    the cell's own row marks "expressed"; it rides in the SIGN BIT of D when the panel and
//            the weights are non-negative (the normal case), else in a per-cell bitmap
//   phase 2  D goes to the warp's shared slab once (128-bit stores); lane owns 4 consecutive pairs of
//            every 128: two shared gathers per score, 512 contiguous bytes of the S row per store
// D never touches HBM.  Algorithmic bytes: 4*n*G_u + 8*n*k + 4*n*P.
#include <type_traits>

#include "common.cuh"
#include "complex.cuh"

namespace laris {

// ------------------------------------------------------------ densify ----
// warp per row: the row is assembled in shared memory and written once with 128-bit stores
__global__ void __launch_bounds__(256)
csr_select_dense_kernel(const int64_t* __restrict__ indptr, const int32_t* __restrict__ indices,
                        const float* __restrict__ data, int64_t n, const int32_t* __restrict__ gene_map,
                        float* __restrict__ Xd, int64_t ldX, int32_t* __restrict__ neg_flag) {
    extern __shared__ __align__(16) float rows_s[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float* rs = rows_s + (size_t)warp * ldX;
    const int64_t nwarps = (int64_t)gridDim.x * (blockDim.x >> 5);
    bool neg = false;
    for (int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + warp; row < n; row += nwarps) {
        for (int t = lane * 4; t < ldX; t += 128) *reinterpret_cast<float4*>(rs + t) = make_float4(0.f, 0.f, 0.f, 0.f);
        __syncwarp();
        const int64_t s = indptr[row], e = indptr[row + 1];
        for (int64_t t = s + lane; t < e; t += 32) {
            const int g = gene_map[indices[t]];
            const float v = data[t];
            if (g >= 0) { rs[g] = v; neg |= v < 0.f; }
        }
        __syncwarp();
        float* out = Xd + row * ldX;
        for (int t = lane * 4; t < ldX; t += 128) *reinterpret_cast<float4*>(out + t) = *reinterpret_cast<const float4*>(rs + t);
        __syncwarp();
    }
    if (neg_flag && __any_sync(0xffffffffu, neg) && lane == 0) atomicOr(neg_flag, 1);
}

// ---------------------------------------------------------------- K2+K3 --
constexpr int kDenseRMax = 3;

__device__ __forceinline__ float4 ldg_f4(const float* p) {
    float4 r;
    asm volatile("ld.global.nc.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
    return r;
}

__device__ __forceinline__ uint32_t nz4(const float4 x) {
    return (x.x != 0.f ? 1u : 0u) | (x.y != 0.f ? 2u : 0u) | (x.z != 0.f ? 4u : 0u) | (x.w != 0.f ? 8u : 0u);
}

// D with the "expressed" flag in the sign bit (D >= 0)
__device__ __forceinline__ float flag_sign(float d, uint32_t bit) {
    return __uint_as_float((__float_as_uint(d) & 0x7fffffffu) | (bit << 31));
}

template <bool SIGNED>
__device__ __forceinline__ float dense_score1(float a, float b, uint32_t m, int& cnt) {
    const float t = a * b;
    bool nz;
    if (SIGNED) {
        nz = m != 0u && t != 0.f;
        if (nz) ++cnt;
        return nz ? t : 0.f;
    }
    const int fm = (int)(__float_as_uint(a) | __float_as_uint(b)) >> 31;   // flag = sign bit of D (D >= 0): all ones / zero
    const uint32_t vb = __float_as_uint(t) & 0x7fffffffu & (uint32_t)fm;
    cnt += vb != 0u;
    return __uint_as_float(vb);
}

#ifndef LARIS_DENSE_DEPTH
#define LARIS_DENSE_DEPTH 2
#endif
#ifndef LARIS_DENSE_MINB
#define LARIS_DENSE_MINB 5
#endif
constexpr int kDenseDepth = LARIS_DENSE_DEPTH;                                       // distinct neighbour rows in flight per warp

// Per-warp shared memory: kDenseRMax cell slabs [slab_words] fp32 and their flag bitmaps [slab_words/32]
// (bitmaps are only touched for signed data).
struct DenseLayout {
    int slab_words, bw_words, warp_words;
    size_t total;
    __host__ __device__ DenseLayout(int ldX, int table_words, int nw) {
        slab_words = ldX;
        bw_words = ldX / 32;
        warp_words = kDenseRMax * (slab_words + bw_words);
        total = ((size_t)table_words + (size_t)nw * warp_words) * 4;
    }
};

// one distinct neighbour row in flight: NBS 128-bit loads per lane, its weight for each of the
// warp's cells, and which of those cells it is the own row of
template <int NBS>
struct DenseRow {
    float4 x[NBS];
    float w[kDenseRMax];
    uint32_t self;
};

template <int NBS>
__device__ __forceinline__ void dense_row_load(DenseRow<NBS>& row, unsigned& lm, int eid, const float (&wr)[kDenseRMax],
                                               uint32_t selfbits, const float* __restrict__ colp, int64_t ldX, int b0, int NB) {
    constexpr unsigned FULL = 0xffffffffu;
    const int l = __ffs(lm) - 1;
    lm &= lm - 1;
    const int id = __shfl_sync(FULL, eid, l);
#pragma unroll
    for (int r = 0; r < kDenseRMax; ++r) row.w[r] = __shfl_sync(FULL, wr[r], l);
    row.self = __shfl_sync(FULL, selfbits, l);
    const float* p = colp + (int64_t)id * ldX;
#pragma unroll
    for (int b = 0; b < NBS; ++b)
        if (b0 + b < NB) row.x[b] = ldg_f4(p + 128 * b);
}

template <int NBS>
__device__ __forceinline__ void dense_row_accumulate(const DenseRow<NBS>& row, float4 (&D)[kDenseRMax][NBS],
                                                     uint32_t (&flags)[kDenseRMax], int b0, int NB) {
#pragma unroll
    for (int b = 0; b < NBS; ++b)
        if (b0 + b < NB) {
#pragma unroll
            for (int r = 0; r < kDenseRMax; ++r) {
                // packed fp32 FMA (FFMA2, sm_100): two columns per instruction, same rounding as two fmaf
                const float2 w2 = make_float2(row.w[r], row.w[r]);
                const float2 lo = __ffma2_rn(w2, make_float2(row.x[b].x, row.x[b].y), make_float2(D[r][b].x, D[r][b].y));
                const float2 hi = __ffma2_rn(w2, make_float2(row.x[b].z, row.x[b].w), make_float2(D[r][b].z, D[r][b].w));
                D[r][b] = make_float4(lo.x, lo.y, hi.x, hi.y);
            }
        }
    if (row.self) {                                                  // the own row of a cell carries its expression flags
        uint32_t f = 0u;
#pragma unroll
        for (int b = 0; b < NBS; ++b)
            if (b0 + b < NB) f |= nz4(row.x[b]) << (4 * b);
#pragma unroll
        for (int r = 0; r < kDenseRMax; ++r)
            if ((row.self >> r) & 1u) flags[r] |= f;
    }
}

// NT threads; NBT = number of 128-column blocks when it is 1..4 (single pass, compile-time slab stride) or
// 0 for wider panels (passes of 4 blocks).  The packed pair table sits in shared memory, read once per
// 4 pairs and group (all cells of the group are scored together).
// Work distribution: the grid is nreg regions x nsub CTAs; region = blockIdx % nreg (CTAs of one region
// are co-resident on one SM under the round-robin block placement), and a region walks ONE contiguous
// range of the spatial order with all its warps interleaved, so the neighbour rows the next groups
// need are the ones the previous groups left in that SM's L1.
template <int NT, int NBT>
__global__ void __launch_bounds__(NT, NBT >= 1 && NBT <= 3 ? LARIS_DENSE_MINB : 3)
diffuse_dense_kernel(const float* __restrict__ Xd, int64_t ldX, const int32_t* __restrict__ neg_flag, int64_t n,
                     const int32_t* __restrict__ idx, const float* __restrict__ w, int k, int R, int nreg,
                     const int32_t* __restrict__ order, const uint32_t* __restrict__ pairs, int p_words,
                     float* __restrict__ S, int64_t ldS, int64_t* __restrict__ row_nnz, ComplexTable cx) {
    constexpr int NW = NT / 32;
    constexpr int RM = kDenseRMax;
    constexpr int NBS = NBT ? NBT : 4;
    constexpr int DEPTH = kDenseDepth;
    constexpr unsigned FULL = 0xffffffffu;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int table_words = p_words;
    const int NB = NBT ? NBT : (int)(ldX >> 7);
    const int slab_words = NB * 128, bw_words = NB * 4;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    uint32_t* pairs_s = reinterpret_cast<uint32_t*>(smem_raw);
    float* slab = reinterpret_cast<float*>(pairs_s + table_words) + (size_t)warp * RM * (slab_words + bw_words);
    uint32_t* Bw = reinterpret_cast<uint32_t*>(slab + RM * slab_words);
    for (int p = tid; p < p_words; p += NT) pairs_s[p] = pairs[p];
    __syncthreads();
    const bool panel_signed = neg_flag ? (*neg_flag != 0) : true;
    const int nedge = R * k;                                         // R == 1 when k > 32

    const int64_t ngroups = (n + R - 1) / R;
    const int region = blockIdx.x % nreg, sub = blockIdx.x / nreg, nsub = gridDim.x / nreg;
    const int64_t per_region = (ngroups + nreg - 1) / nreg;
    const int64_t g_begin = (int64_t)region * per_region, g_end = min(ngroups, g_begin + per_region);
    // Software pipeline over the warp's groups: the cell ids of group g+2 and the first 32 graph edges of group g+1
    // are loaded while group g is processed, so the order -> idx/w -> row dependent-load chain of a group is off
    // its critical path.
    const int64_t gstride = (int64_t)nsub * NW;
    auto load_cells = [&](int64_t grp) -> int {                      // the group's cells live in lanes 0..Rg-1
        const int64_t q0 = grp * R;
        if (grp >= g_end || lane >= (int)min((int64_t)R, n - q0)) return -1;
        return order ? order[q0 + lane] : (int)(q0 + lane);
    };
    auto load_edges = [&](int cq_, int64_t grp, int& eid, float& ew) {   // one graph edge per lane: (cell e/k, edge e%k)
        const int Rg_ = grp < g_end ? (int)min((int64_t)R, n - grp * R) : 0;
        const int er = lane / k;
        const bool valid = lane < nedge && er < Rg_;
        const int c = __shfl_sync(FULL, cq_, valid ? er : 0);
        eid = -1 - lane;                                             // idle lanes never match anything
        ew = 0.f;
        if (valid) {
            eid = idx[(int64_t)c * k + (lane - er * k)];
            ew = w[(int64_t)c * k + (lane - er * k)];
        }
    };
    const int64_t grp0 = g_begin + sub * NW + warp;
    int cq_next = load_cells(grp0), eid_next;
    float ew_next;
    load_edges(cq_next, grp0, eid_next, ew_next);
    int cq_next2 = load_cells(grp0 + gstride);
    for (int64_t grp = grp0; grp < g_end; grp += gstride) {
        const int64_t q0 = grp * R;
        const int Rg = (int)min((int64_t)R, n - q0);
        const int cq = cq_next, eid_first = eid_next;
        const float ew_first = ew_next;
        cq_next = cq_next2;
        load_edges(cq_next, grp + gstride, eid_next, ew_next);
        cq_next2 = load_cells(grp + 2 * gstride);
        bool wneg = false;

        for (int b0 = 0; b0 < NB; b0 += NBS) {                       // passes over the column blocks (1 pass when NB <= 4)
            float4 D[RM][NBS];
            uint32_t flags[RM];
#pragma unroll
            for (int r = 0; r < RM; ++r) {
                flags[r] = 0u;
#pragma unroll
                for (int b = 0; b < NBS; ++b) D[r][b] = make_float4(0.f, 0.f, 0.f, 0.f);
            }
            uint32_t found = 0u;                                     // bit r: cell r's own row was among the neighbour rows
            const float* colp = Xd + ((int64_t)b0 << 7) + 4 * lane;
            for (int e0 = 0; e0 < nedge; e0 += 32) {
                // ---- one graph edge per lane: (cell r, neighbour id, weight) ----
                const int e = e0 + lane;
                const int er = e / k;
                const bool valid = e < nedge && er < Rg;
                int eid = -1 - lane;                                 // idle lanes never match anything
                float ew = 0.f;
                const int c = __shfl_sync(FULL, cq, valid ? er : 0);
                if (e0 == 0) {                                       // prefetched with the previous group
                    eid = eid_first;
                    ew = ew_first;
                } else if (valid) {
                    eid = idx[(int64_t)c * k + (e - er * k)];
                    ew = w[(int64_t)c * k + (e - er * k)];
                }
                wneg |= __any_sync(FULL, ew < 0.f);
                const unsigned same = __match_any_sync(FULL, eid);
                const bool leader = valid && (lane == __ffs(same) - 1);
                const unsigned selfedge = __ballot_sync(FULL, valid && eid == c);
                float wr[RM];
                uint32_t selfbits = 0u;
                bool dup = false;
#pragma unroll
                for (int r = 0; r < RM; ++r) {
                    const unsigned sub_r = same & __ballot_sync(FULL, valid && er == r);
                    const float wv = __shfl_sync(FULL, ew, sub_r ? __ffs(sub_r) - 1 : 0);
                    wr[r] = sub_r ? wv : 0.f;
                    dup |= __popc(sub_r) > 1;
                    if (sub_r & selfedge) selfbits |= 1u << r;
                }
                if (__any_sync(FULL, dup)) {                         // a neighbour listed twice for one cell: weights add
#pragma unroll
                    for (int r = 0; r < RM; ++r) {
                        unsigned sub_r = same & __ballot_sync(FULL, valid && er == r);
                        float acc = 0.f;
                        for (int it = 0; it < 32; ++it) {
                            if (!__ballot_sync(FULL, sub_r != 0u)) break;
                            const float wv = __shfl_sync(FULL, ew, sub_r ? __ffs(sub_r) - 1 : 0);
                            if (sub_r) { acc += wv; sub_r &= sub_r - 1; }
                        }
                        wr[r] = acc;
                    }
                }
                found |= __reduce_or_sync(FULL, leader ? selfbits : 0u);
                // ---- distinct rows, in lane order, DEPTH of them in flight: a ring of register buffers ----
                unsigned lm = __ballot_sync(FULL, leader);
                DenseRow<NBS> buf[DEPTH];
                unsigned live = 0u;
#pragma unroll
                for (int d = 0; d < DEPTH; ++d)
                    if (lm) { dense_row_load(buf[d], lm, eid, wr, selfbits, colp, ldX, b0, NB); live |= 1u << d; }
                while (live) {
#pragma unroll
                    for (int d = 0; d < DEPTH; ++d) {
                        if (!((live >> d) & 1u)) { live = 0u; break; }   // rows are issued in ring order: the rest are empty too
                        dense_row_accumulate(buf[d], D, flags, b0, NB);
                        if (lm) dense_row_load(buf[d], lm, eid, wr, selfbits, colp, ldX, b0, NB);
                        else live &= ~(1u << d);
                    }
                }
            }
#pragma unroll
            for (int r = 0; r < RM; ++r)                             // graph without self edges: fetch the own row
                if (r < Rg && !((found >> r) & 1u)) {
                    const float* p = colp + (int64_t)__shfl_sync(FULL, cq, r) * ldX;
#pragma unroll
                    for (int b = 0; b < NBS; ++b)
                        if (b0 + b < NB) flags[r] |= nz4(ldg_f4(p + 128 * b)) << (4 * b);
                }
            // ---- D (+ flags) -> the warp's shared slabs ----
            const bool signed_now = panel_signed || wneg;
#pragma unroll
            for (int r = 0; r < RM; ++r) {
                if (r >= Rg) continue;
                float* sl = slab + r * slab_words + (b0 << 7) + 4 * lane;
#pragma unroll
                for (int b = 0; b < NBS; ++b) {
                    if (b0 + b >= NB) continue;
                    float4 v = D[r][b];
                    const uint32_t f = flags[r] >> (4 * b);
                    if (!signed_now) {
                        v.x = flag_sign(v.x, f & 1u);
                        v.y = flag_sign(v.y, (f >> 1) & 1u);
                        v.z = flag_sign(v.z, (f >> 2) & 1u);
                        v.w = flag_sign(v.w, (f >> 3) & 1u);
                    } else {                                         // 8 lanes share one bitmap word
                        uint32_t word = (f & 15u) << ((lane & 7) * 4);
                        word |= __shfl_xor_sync(FULL, word, 1);
                        word |= __shfl_xor_sync(FULL, word, 2);
                        word |= __shfl_xor_sync(FULL, word, 4);
                        if ((lane & 7) == 0) Bw[r * bw_words + 4 * (b0 + b) + (lane >> 3)] = word;
                    }
                    *reinterpret_cast<float4*>(sl + 128 * b) = v;
                }
            }
        }
        const bool signed_data = panel_signed || wneg;
        __syncwarp();
        if (cx.nc) {                                                 // multi-subunit complexes -> their virtual columns
            for (int r = 0; r < Rg; ++r) {
                if (signed_data) complex_pass<true>(cx, slab + r * slab_words, Bw + r * bw_words, lane);
                else complex_pass<false>(cx, slab + r * slab_words, Bw + r * bw_words, lane);
            }
            __syncwarp();
        }
        // ---------------- phase 2: S[p] = D[l] * D[r] * (flag_l | flag_r) ----------------
        // the lane's pairs are scored for all cells of the group at once: one address per operand, RM gathers
        int cnt[RM];
        int ci[RM];
        float* srow[RM];
#pragma unroll
        for (int r = 0; r < RM; ++r) {
            cnt[r] = 0;
            ci[r] = __shfl_sync(FULL, cq, r);
            srow[r] = S + (int64_t)(r < Rg ? ci[r] : 0) * ldS + 4 * lane;
        }
        auto phase2 = [&](auto signed_tag) {
            constexpr bool SIGNED = decltype(signed_tag)::value;
            for (int p0 = lane * 4; p0 < (int)ldS; p0 += 128) {
                const uint4 pr4 = *reinterpret_cast<const uint4*>(pairs_s + p0);
                const uint32_t q4[4] = {pr4.x, pr4.y, pr4.z, pr4.w};
                float v[RM][4];
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const uint32_t l = q4[u] & 0xffffu, rr = q4[u] >> 16;
                    const float* pa = slab + l;
                    const float* pb = slab + rr;
#pragma unroll
                    for (int r = 0; r < RM; ++r) {
                        const float a = pa[r * slab_words], b = pb[r * slab_words];
                        uint32_t m = 0u;
                        if (SIGNED) {
                            const uint32_t* Bc = Bw + r * bw_words;
                            m = ((Bc[l >> 5] >> (l & 31)) | (Bc[rr >> 5] >> (rr & 31))) & 1u;
                        }
                        v[r][u] = dense_score1<SIGNED>(a, b, m, cnt[r]);
                    }
                }
#pragma unroll
                for (int r = 0; r < RM; ++r)
                    if (r < Rg) st_stream_f4(reinterpret_cast<float4*>(srow[r] + (p0 - 4 * lane)),
                                             make_float4(v[r][0], v[r][1], v[r][2], v[r][3]));
            }
        };
        if (signed_data) phase2(std::true_type{}); else phase2(std::false_type{});
        if (row_nnz) {
#pragma unroll
            for (int r = 0; r < RM; ++r) {
                const int c = __reduce_add_sync(FULL, cnt[r]);
                if (lane == 0 && r < Rg) row_nnz[ci[r]] = c;
            }
        }
        __syncwarp();                                                // slabs are rewritten by the next group
    }
}


__global__ void pack_pairs_dense_kernel(const int32_t* __restrict__ lig, const int32_t* __restrict__ rec, int P, int G_u,
                                        int p_words, uint32_t* __restrict__ pairs) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p < p_words) pairs[p] = p < P ? ((uint32_t)lig[p] | ((uint32_t)rec[p] << 16)) : ((uint32_t)G_u | ((uint32_t)G_u << 16));
}

}  // namespace laris

using namespace laris;

extern "C" int laris_csr_select_dense(const int64_t* indptr, const int32_t* indices, const float* data, int64_t n,
                                      const int32_t* gene_map, float* Xd, int64_t ldX, int32_t* neg_flag, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    LARIS_CHECK_ARG(indptr && gene_map && Xd && n >= 0, "laris_csr_select_dense: bad arguments");
    LARIS_CHECK_ARG(ldX >= 128 && ldX % 128 == 0 && ldX <= 6144 && ((uintptr_t)Xd & 15) == 0,
                    "laris_csr_select_dense: ldX must be a multiple of 128 in [128, 6144] and Xd 16-byte aligned");
    if (neg_flag) LARIS_CUDA(cudaMemsetAsync(neg_flag, 0, sizeof(int32_t), st));
    if (n == 0) return LARIS_OK;
    const size_t smem = (size_t)8 * ldX * sizeof(float);
    LARIS_CUDA(cudaFuncSetAttribute(csr_select_dense_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int occ = 1;
    LARIS_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, csr_select_dense_kernel, 256, smem));
    if (occ < 1) occ = 1;
    int dev = 0, sms = kNumSMs;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int64_t want = ceil_div(n, 8);
    const int64_t grid = want < (int64_t)sms * occ ? want : (int64_t)sms * occ;
    csr_select_dense_kernel<<<(unsigned)grid, 256, smem, st>>>(indptr, indices, data, n, gene_map, Xd, ldX, neg_flag);
    LARIS_LAUNCH_CHECK();
    return LARIS_OK;
}

static int diffuse_lr_score_dense_impl(const float* Xd, int64_t ldX, const int32_t* neg_flag, int64_t n, int32_t G_u,
                                       const int32_t* idx, const float* w, int k, const int32_t* order,
                                       const int32_t* lig, const int32_t* rec, int32_t P, float* S, int64_t ldS,
                                       int64_t* row_nnz, ComplexTable cx, cudaStream_t st) {
    LARIS_CHECK_ARG(Xd && idx && w && lig && rec && S, "laris_diffuse_lr_score_dense: null pointer");
    LARIS_CHECK_ARG(n > 0 && k >= 1 && P >= 1, "laris_diffuse_lr_score_dense: empty problem");
    LARIS_CHECK_ARG(G_u >= 1 && ldX > G_u && ldX % 128 == 0 && ldX <= 65536 && ((uintptr_t)Xd & 15) == 0,
                    "laris_diffuse_lr_score_dense: need G_u < ldX, ldX %% 128 == 0 (got G_u=%d ldX=%lld)", G_u,
                    (long long)ldX);
    LARIS_CHECK_ARG(ldS >= P && ldS % 4 == 0 && ((uintptr_t)S & 15) == 0,
                    "laris_diffuse_lr_score_dense: S needs ldS >= P, ldS %% 4 == 0 and 16-byte alignment");
    const int p_words = (int)((ldS + 127) / 128) * 128;
    constexpr int NT = 128;
    Scratch<uint32_t> pairs(st);
    LARIS_CUDA(pairs.alloc((size_t)p_words));
    pack_pairs_dense_kernel<<<(unsigned)ceil_div(p_words, 256), 256, 0, st>>>(lig, rec, P, G_u, p_words, pairs.p);
    LARIS_LAUNCH_CHECK();
    const int R = k > 32 ? 1 : (32 / k < kDenseRMax ? 32 / k : kDenseRMax);
    const DenseLayout L((int)ldX, p_words, NT / 32);
    if (L.total + 1024 > (size_t)227 * 1024) {
        set_error("laris_diffuse_lr_score_dense: ldX=%lld, P=%d do not fit shared memory", (long long)ldX, P);
        return LARIS_E_LIMIT;
    }
    using Kern = void (*)(const float*, int64_t, const int32_t*, int64_t, const int32_t*, const float*, int, int, int,
                          const int32_t*, const uint32_t*, int, float*, int64_t, int64_t*, ComplexTable);
    const int NB = (int)(ldX / 128);
    const int nbt = NB <= 4 ? NB : 0;
    Kern kern = nbt == 1 ? diffuse_dense_kernel<NT, 1> : nbt == 2 ? diffuse_dense_kernel<NT, 2>
              : nbt == 3 ? diffuse_dense_kernel<NT, 3> : nbt == 4 ? diffuse_dense_kernel<NT, 4> : diffuse_dense_kernel<NT, 0>;
    LARIS_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L.total));
    int occ = 1;
    LARIS_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, NT, L.total));
    if (occ < 1) occ = 1;
    int dev = 0, sms = kNumSMs;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    // regions x co-resident CTAs: one region per SM, each walking a contiguous range of the spatial order
    const int64_t ngroups = ceil_div(n, R);
    const int nreg = (int)(ngroups < sms ? ngroups : sms);
    int nsub = (int)ceil_div(ceil_div(ngroups, nreg), NT / 32);
    if (nsub > occ) nsub = occ;
    if (nsub < 1) nsub = 1;
    kern<<<(unsigned)(nreg * nsub), NT, L.total, st>>>(Xd, ldX, neg_flag, n, idx, w, k, R, nreg, order, pairs.p, p_words, S,
                                                       ldS, row_nnz, cx);
    LARIS_LAUNCH_CHECK();
    return LARIS_OK;
}

extern "C" int laris_diffuse_lr_score_dense(const float* Xd, int64_t ldX, const int32_t* neg_flag, int64_t n, int32_t G_u,
                                            const int32_t* idx, const float* w, int k, const int32_t* order,
                                            const int32_t* lig, const int32_t* rec, int32_t P, float* S, int64_t ldS,
                                            int64_t* row_nnz, void* stream) {
    return diffuse_lr_score_dense_impl(Xd, ldX, neg_flag, n, G_u, idx, w, k, order, lig, rec, P, S, ldS, row_nnz,
                                       ComplexTable(), (cudaStream_t)stream);
}

extern "C" int laris_diffuse_lr_score_dense_complex(const float* Xd, int64_t ldX, const int32_t* neg_flag, int64_t n,
                                                    int32_t G_u, const int32_t* idx, const float* w, int k,
                                                    const int32_t* order, const int32_t* lig, const int32_t* rec,
                                                    int32_t P, float* S, int64_t ldS, int64_t* row_nnz,
                                                    const int32_t* cx_ptr, const int32_t* cx_cols, const int32_t* cx_vcol,
                                                    int32_t n_complex, int32_t rule, void* stream) {
    LARIS_CHECK_ARG(n_complex >= 0 && (n_complex == 0 || (cx_ptr && cx_cols && cx_vcol)) && (rule == 0 || rule == 1),
                    "laris_diffuse_lr_score_dense_complex: bad complex table (n_complex=%d, rule=%d)", n_complex, rule);
    ComplexTable cx;
    cx.ptr = cx_ptr; cx.cols = cx_cols; cx.vcol = cx_vcol; cx.nc = n_complex; cx.rule = rule;
    return diffuse_lr_score_dense_impl(Xd, ldX, neg_flag, n, G_u, idx, w, k, order, lig, rec, P, S, ldS, row_nnz, cx,
                                       (cudaStream_t)stream);
}
