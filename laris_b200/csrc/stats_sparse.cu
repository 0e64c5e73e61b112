// K4 over SPARSE score rows: the observed statistic and the random-graph null from a packed copy of S.
//
// S = D_l * D_r gated by "ligand or receptor expressed in the cell" is mostly zeros (9.7 % non-zero at config 3; the
// reference stores it as a scipy CSR matrix for the same reason).  The dense passes (stats.cu) move 4*P bytes per
// gathered neighbour row whatever it holds; here a gathered row costs 8 bytes per NON-ZERO:
//
//   scores_pack_kernel        S (dense, read once) -> packed rows {column, value}, entries dealt round-robin over the 32
//                             shared-memory banks (the order of a row's entries is free: its columns are distinct)
//   spatial_stats_sparse_kernel   one CTA round = NW cells, one per warp.  Phase A: the warp adds its cell's k neighbour rows
//                             into a private slab T[0..P) of shared memory, row after row in neighbour order
//                             (T[col] = fma(w, x, T[col]) - bit for bit the sum the dense kernels form, the skipped terms
//                             are exact zeros), next row's entries in flight while the current one is added.  Phase B: the
//                             CTA's threads own columns; each reads its columns of the NW slabs (and zeroes them), scales by
//                             1/sum(w) for the L1-normalised null graphs, reads S[i, columns] (coalesced, L2 after the
//                             first graph) and accumulates sum S*T and sum T^2 (and sum S^2, sum S, nnz once) - fp32 over
//                             64 cells, fp64 beyond, per-CTA partials reduced in a fixed order (deterministic).
//                             Up to four graphs (observed + nulls) per launch share the read of S.
//
// HBM bytes: 8*nnz_row*k per (cell, graph) for random graphs (the observed graph's rows are spatial neighbours: L2 hits)
// + 4nP (S once per launch) instead of (1+k)*4nP per graph.
#include <cuda.h>
#include <stdlib.h>

#include "common.cuh"

namespace laris {

constexpr unsigned kFull = 0xffffffffu;

__device__ __forceinline__ int2 ld_entry8(const int2* p) {
    int2 r;
    asm volatile("ld.global.nc.L1::no_allocate.v2.s32 {%0,%1}, [%2];" : "=r"(r.x), "=r"(r.y) : "l"(p));
    return r;
}
// ------------------------------------------------------------------------------------------------ pack
// Warp per row.  Lane l reads the columns c = l (mod 32) - all in shared-memory bank l of a slab - and queues the block
// numbers of its non-zeros; round t of the emission writes the t-th non-zero of every lane that has one, so 32
// consecutive entries of a packed row hold at most two entries of a bank.  Rows are PADDED to a multiple of 32 entries
// (sp_ptr is the scan of the padded counts) with {column ldS + j, value 0}: the statistics kernel then treats every chunk
// of 32 entries alike, the padding adds zeros to 32 spare words behind its slab.
constexpr int kPackWarps = 8;
constexpr int kPackBatch = 16;

__global__ void __launch_bounds__(kPackWarps * 32)
scores_pack_kernel(const float* __restrict__ S, int64_t ldS, int64_t n, int P, const int64_t* __restrict__ sp_ptr,
                   int2* __restrict__ out, int64_t capacity) {
    extern __shared__ uint16_t pack_q[];                         // [warp][T][32]
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int T = (P + 31) / 32;
    uint16_t* q = pack_q + (size_t)warp * T * 32;
    for (int64_t row = (int64_t)blockIdx.x * kPackWarps + warp; row < n; row += (int64_t)gridDim.x * kPackWarps) {
        const float* srow = S + row * ldS;
        int cnt = 0;
        for (int t0 = 0; t0 < T; t0 += kPackBatch) {
            float v[kPackBatch];
#pragma unroll
            for (int u = 0; u < kPackBatch; ++u) {
                const int col = 32 * (t0 + u) + lane;
                v[u] = col < P ? __ldg(srow + col) : 0.f;
            }
#pragma unroll
            for (int u = 0; u < kPackBatch; ++u)
                if (v[u] != 0.f) {
                    q[cnt * 32 + lane] = (uint16_t)(t0 + u);
                    ++cnt;
                }
        }
        int maxc = cnt;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) maxc = max(maxc, __shfl_xor_sync(kFull, maxc, o));
        int64_t base = sp_ptr[row];
        // counts that do not describe S, or an entry array sized from an estimate, must not corrupt memory
        const int64_t end = min(sp_ptr[row + 1], capacity);
        const unsigned lt = (1u << lane) - 1u;
        for (int t0 = 0; t0 < maxc; t0 += 4) {                    // four rounds at a time: their value reads overlap
            int col[4];
            int64_t at[4];
            float v[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const bool has = cnt > t0 + u;
                const unsigned m = __ballot_sync(kFull, has);
                col[u] = has ? 32 * (int)q[(t0 + u) * 32 + lane] + lane : -1;
                at[u] = base + __popc(m & lt);
                base += __popc(m);
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) v[u] = col[u] >= 0 ? __ldg(srow + col[u]) : 0.f;
#pragma unroll
            for (int u = 0; u < 4; ++u)
                if (col[u] >= 0 && at[u] < end) out[at[u]] = make_int2(col[u], __float_as_int(v[u]));
        }
        if (base + lane < end) out[base + lane] = make_int2((int)ldS + (int)((base + lane - sp_ptr[row]) & 31), 0);   // padding
        __syncwarp();
    }
}

// ------------------------------------------------------------------------------------------------ statistics
constexpr int kSpMaxGraphs = 4;

// the graphs of one launch are slots slot0 .. slot0+NG-1 of the call; slot 0 is the observed graph when has_obs, the others
// are the random graphs (null_idx / null_w + r*nk), whose rows are L1-normalised
struct SparseGraphs {
    const int32_t* obs_idx;
    const float* obs_w;
    const int32_t* null_idx;
    const float* null_w;
    int64_t nk;
    int slot0, has_obs;
};

// prefetch loads of the builders' pipeline: volatile, so that they are ISSUED where they are written (the compiler would
// otherwise sink a plain load to its first use, a whole item later, and put its latency back on the critical path)
__device__ __forceinline__ int ldv_s32(const int32_t* p) {
    int v;
    asm volatile("ld.global.nc.s32 %0, [%1];" : "=r"(v) : "l"(p));
    return v;
}
__device__ __forceinline__ float ldv_f32(const float* p) {
    float v;
    asm volatile("ld.global.nc.f32 %0, [%1];" : "=f"(v) : "l"(p));
    return v;
}
__device__ __forceinline__ long long ldv_s64(const int64_t* p) {
    long long v;
    asm volatile("ld.global.nc.s64 %0, [%1];" : "=l"(v) : "l"(p));
    return v;
}

// A (cell, graph) item as seen by the warp that builds its slab: lane j holds neighbour j's packed row (pointer, length - a
// multiple of 32 entries, rows are padded) and weight.  The rows are consumed as a stream of SEGMENTS of <= CH chunks of 32
// entries: `start` = index of the row's first segment, `total` = segments of the item.  RING segments are in flight per warp
// (registers) - the pass gathers random HBM rows and lives on memory parallelism.  (CH, RING) = (8, 3) for long rows (a
// 195-entry row is one segment, ~4.7 KB in flight per warp), (2, 8) for short ones (sharded tables: 8 rows in flight).
struct SpItem {
    const int2* rowp;
    int len;
    float wv;
    int start, total;
};

template <int CH>
struct SpSeg {
    int2 e[CH];
    float w;
    int nch;                 // valid chunks of the segment (warp-uniform)
};

template <int CH>
__device__ __forceinline__ void sp_activate(SpItem& it, int lane) {
    const int nseg = (it.len + 32 * CH - 1) / (32 * CH);
    int incl = nseg;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int v = __shfl_up_sync(kFull, incl, o);
        if (lane >= o) incl += v;
    }
    it.start = incl - nseg;
    it.total = __shfl_sync(kFull, incl, 31);
}

// Groups of G chunks as ONE asm statement each: chunk c of the group is live iff c < n (warp-uniform), everything is
// predicated - no branches, and the read-modify-writes of a group issue as G shared loads, G fma, G shared stores
// (the columns of a row are distinct, so they are independent).
template <int G>
struct ChunkGroup;

template <>
struct ChunkGroup<4> {
    static __device__ __forceinline__ void load(int2& e0, int2& e1, int2& e2, int2& e3, const int2* p, int n) {
        asm volatile(
            "{\n\t.reg .pred q0, q1, q2, q3;\n\t"
            "setp.gt.s32 q0, %9, 0;\n\tsetp.gt.s32 q1, %9, 1;\n\tsetp.gt.s32 q2, %9, 2;\n\tsetp.gt.s32 q3, %9, 3;\n\t"
            "@q0 ld.global.nc.L1::no_allocate.v2.s32 {%0,%1}, [%8];\n\t"
            "@q1 ld.global.nc.L1::no_allocate.v2.s32 {%2,%3}, [%8+256];\n\t"
            "@q2 ld.global.nc.L1::no_allocate.v2.s32 {%4,%5}, [%8+512];\n\t"
            "@q3 ld.global.nc.L1::no_allocate.v2.s32 {%6,%7}, [%8+768];\n\t}"
            : "+r"(e0.x), "+r"(e0.y), "+r"(e1.x), "+r"(e1.y), "+r"(e2.x), "+r"(e2.y), "+r"(e3.x), "+r"(e3.y)
            : "l"(p), "r"(n));
    }
    static __device__ __forceinline__ void add(const int2& e0, const int2& e1, const int2& e2, const int2& e3, float w,
                                               uint32_t slab, int n) {
        asm volatile(
            "{\n\t.reg .pred q0, q1, q2, q3;\n\t.reg .f32 t0, t1, t2, t3;\n\t"
            "setp.gt.s32 q0, %9, 0;\n\tsetp.gt.s32 q1, %9, 1;\n\tsetp.gt.s32 q2, %9, 2;\n\tsetp.gt.s32 q3, %9, 3;\n\t"
            "@q0 ld.shared.f32 t0, [%0];\n\t@q1 ld.shared.f32 t1, [%1];\n\t@q2 ld.shared.f32 t2, [%2];\n\t@q3 ld.shared.f32 t3, [%3];\n\t"
            "@q0 fma.rn.f32 t0, %8, %4, t0;\n\t@q1 fma.rn.f32 t1, %8, %5, t1;\n\t@q2 fma.rn.f32 t2, %8, %6, t2;\n\t@q3 fma.rn.f32 t3, %8, %7, t3;\n\t"
            "@q0 st.shared.f32 [%0], t0;\n\t@q1 st.shared.f32 [%1], t1;\n\t@q2 st.shared.f32 [%2], t2;\n\t@q3 st.shared.f32 [%3], t3;\n\t}"
            :: "r"(slab + 4u * (uint32_t)e0.x), "r"(slab + 4u * (uint32_t)e1.x), "r"(slab + 4u * (uint32_t)e2.x),
               "r"(slab + 4u * (uint32_t)e3.x), "f"(__int_as_float(e0.y)), "f"(__int_as_float(e1.y)),
               "f"(__int_as_float(e2.y)), "f"(__int_as_float(e3.y)), "f"(w), "r"(n)
            : "memory");
    }
};

template <>
struct ChunkGroup<2> {
    static __device__ __forceinline__ void load(int2& e0, int2& e1, const int2* p, int n) {
        asm volatile(
            "{\n\t.reg .pred q0, q1;\n\tsetp.gt.s32 q0, %5, 0;\n\tsetp.gt.s32 q1, %5, 1;\n\t"
            "@q0 ld.global.nc.L1::no_allocate.v2.s32 {%0,%1}, [%4];\n\t"
            "@q1 ld.global.nc.L1::no_allocate.v2.s32 {%2,%3}, [%4+256];\n\t}"
            : "+r"(e0.x), "+r"(e0.y), "+r"(e1.x), "+r"(e1.y) : "l"(p), "r"(n));
    }
    static __device__ __forceinline__ void add(const int2& e0, const int2& e1, float w, uint32_t slab, int n) {
        asm volatile(
            "{\n\t.reg .pred q0, q1;\n\t.reg .f32 t0, t1;\n\tsetp.gt.s32 q0, %5, 0;\n\tsetp.gt.s32 q1, %5, 1;\n\t"
            "@q0 ld.shared.f32 t0, [%0];\n\t@q1 ld.shared.f32 t1, [%1];\n\t"
            "@q0 fma.rn.f32 t0, %4, %2, t0;\n\t@q1 fma.rn.f32 t1, %4, %3, t1;\n\t"
            "@q0 st.shared.f32 [%0], t0;\n\t@q1 st.shared.f32 [%1], t1;\n\t}"
            :: "r"(slab + 4u * (uint32_t)e0.x), "r"(slab + 4u * (uint32_t)e1.x), "f"(__int_as_float(e0.y)),
               "f"(__int_as_float(e1.y)), "f"(w), "r"(n)
            : "memory");
    }
};

template <int CH>
__device__ __forceinline__ void sp_seg_load(SpSeg<CH>& b, const SpItem& it, int s, int lane) {
    b.nch = 0;
    if (s < it.total) {                                              // warp-uniform
        const unsigned m = __ballot_sync(kFull, it.start <= s);      // the LAST lane whose first segment is <= s owns s
        const int row = 31 - __clz(m);
        const int off = (s - __shfl_sync(kFull, it.start, row)) * (32 * CH);
        const int2* p = reinterpret_cast<const int2*>(__shfl_sync(kFull, (unsigned long long)it.rowp, row)) + off + lane;
        b.nch = min(CH, (__shfl_sync(kFull, it.len, row) - off) >> 5);
        b.w = __shfl_sync(kFull, it.wv, row);
        if (CH == 2) {
            ChunkGroup<2>::load(b.e[0], b.e[1], p, b.nch);
        } else {
#pragma unroll
            for (int g0 = 0; g0 < CH; g0 += 4)
                ChunkGroup<4>::load(b.e[g0], b.e[g0 + 1], b.e[g0 + 2], b.e[g0 + 3], p + 32 * g0, b.nch - g0);
        }
    }
}

// T[col] = fma(w, x, T[col]) for the segment's entries (`slab`: shared-memory byte address of the warp's slab); the
// trailing __syncwarp orders the segment against the next row's
template <int CH>
__device__ __forceinline__ void sp_seg_add(const SpSeg<CH>& b, uint32_t slab) {
    if (CH == 2) {
        ChunkGroup<2>::add(b.e[0], b.e[1], b.w, slab, b.nch);
    } else {
#pragma unroll
        for (int g0 = 0; g0 < CH; g0 += 4)
            ChunkGroup<4>::add(b.e[g0], b.e[g0 + 1], b.e[g0 + 2], b.e[g0 + 3], b.w, slab, b.nch - g0);
    }
    __syncwarp();
}

// partial: [grid][2*n_slots + 3][ldw] doubles, zero-initialised.  Rows 2s, 2s+1 = sum S*T, sum T^2 of graph slot s; the
// last three = sum S^2, sum S, nnz (written when `first`).
//
// Warp-specialised, one 512-thread CTA per SM.  A STEP is (round of kSpCells cells, graph).  Warps 0-7 BUILD: warp w adds
// the k packed neighbour rows of cell (round, w) into its slab of the step's slab set.  Warps 8-15 OWN columns: thread t
// owns the column quads t, t+256, ... of all slabs; it folds a finished slab set and the cells' own dense rows into its
// running sums and clears the slabs.  Two slab sets alternate, so the builders gather the rows of step s+1 while the
// owners fold step s; full / empty hand-over through named barriers (bar.arrive / bar.sync, 512 participants).
constexpr int kSpCells = 8;              // builder warps = cells per step
constexpr int kSpOwners = 256;           // column-owner threads

__device__ __forceinline__ void named_sync(int id) { asm volatile("bar.sync %0, 512;" ::"r"(id) : "memory"); }
__device__ __forceinline__ void named_arrive(int id) { asm volatile("bar.arrive %0, 512;" ::"r"(id) : "memory"); }

template <int NQ, int CH, int RING>
__global__ void __launch_bounds__(512, 1)
spatial_stats_sparse_kernel(const int64_t* __restrict__ sp_ptr, const int2* __restrict__ entries,
                            const float4* __restrict__ S4, int64_t ld4, int64_t n, SparseGraphs gs, int NG, int k,
                            const int32_t* __restrict__ order, int n_slots, double* __restrict__ partial, int64_t capacity) {
    const bool first = gs.slot0 == 0;
    extern __shared__ __align__(16) unsigned char sp_smem[];
    __shared__ float inv_s[2][kSpCells];
    __shared__ int32_t cell_s[2][kSpCells];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int ldw = (int)ld4 * 4;
    // [2][kSpCells] slabs of ld4 + 8 quads (the last 8 take the padding entries), then [NG][2][ld4] owner sums
    float4* Tall4 = reinterpret_cast<float4*>(sp_smem);
    const int slab4 = (int)ld4 + 8;
    for (int t = tid; t < 2 * kSpCells * slab4 + 2 * NG * (int)ld4; t += 512) Tall4[t] = make_float4(0.f, 0.f, 0.f, 0.f);
    __syncthreads();
    const int64_t n_rounds = (n + kSpCells - 1) / kSpCells;

    if (warp < kSpCells) {
        // ============================== builders ==============================
        // Software pipeline over this warp's items t = (round, graph): when item t starts, its row pointers have arrived
        // and its first RING segments are in flight; the row pointers of t+1 are requested now (its neighbour ids
        // arrived during t-1) and the neighbour ids / weights of t+2 are requested now - the dependent chain
        // order -> idx -> sp_ptr -> entries never sits on the critical path.
        auto cell_of = [&](int64_t round) -> int64_t {
            const int64_t q = round * kSpCells + warp;
            if (round >= n_rounds || q >= n) return -1;
            return order ? (int64_t)order[q] : q;
        };
        auto fetch_edges = [&](int64_t i, int g, int j0, int& nb, float& wq) {
            nb = -1; wq = 0.f;
            if (i >= 0 && j0 + lane < k) {
                const int slot = gs.slot0 + g;
                const bool obs = gs.has_obs && slot == 0;
                const int64_t at = (obs ? 0 : (int64_t)(slot - gs.has_obs) * gs.nk) + i * k + j0 + lane;
                nb = ldv_s32((obs ? gs.obs_idx : gs.null_idx) + at);
                wq = ldv_f32((obs ? gs.obs_w : gs.null_w) + at);
            }
        };
        // rs / re: the two row pointers, still in flight after this call; finish() turns them into (rowp, len)
        auto resolve = [&](int nb, long long& rs, long long& re) {
            rs = re = 0;
            if (nb >= 0) {
                rs = ldv_s64(sp_ptr + nb);
                re = ldv_s64(sp_ptr + nb + 1);
            }
        };
        auto finish = [&](SpItem& it, long long rs, long long re) {
            it.len = re <= capacity ? (int)(re - rs) : 0;            // rows beyond an under-sized entry array are not read
            it.rowp = entries + rs;
        };
        SpSeg<CH> ring[RING];
        auto preload = [&](SpItem& it) {
            sp_activate<CH>(it, lane);
#pragma unroll
            for (int u = 0; u < RING; ++u) sp_seg_load<CH>(ring[u], it, u, lane);
        };
        auto consume = [&](const SpItem& it, uint32_t slab) {        // ring holds segments 0..RING-1 of `it`
            for (int s0 = 0; s0 < it.total; s0 += RING) {
#pragma unroll
                for (int u = 0; u < RING; ++u)
                    if (s0 + u < it.total) {
                        sp_seg_add<CH>(ring[u], slab);
                        sp_seg_load<CH>(ring[u], it, s0 + u + RING, lane);
                    }
            }
        };
        auto group_sum = [&](float wq, int kc, float& wsum) {        // same association as the dense kernels
            int j = 0;
            for (; j + 4 <= kc; j += 4)
                wsum += (__shfl_sync(kFull, wq, j) + __shfl_sync(kFull, wq, j + 1)) +
                        (__shfl_sync(kFull, wq, j + 2) + __shfl_sync(kFull, wq, j + 3));
            for (; j < kc; ++j) wsum += __shfl_sync(kFull, wq, j);
        };
        // the cell's own dense row is read by the owners one step later: bring it into L2 now
        auto prefetch_own = [&](int64_t i) {
            if (i < 0) return;
            const char* row = reinterpret_cast<const char*>(S4 + i * ld4);
            for (int b = lane * 128; b < ldw * 4; b += 32 * 128) asm volatile("prefetch.global.L2 [%0];" ::"l"(row + b));
        };

        int64_t i0 = cell_of(blockIdx.x), i1 = cell_of((int64_t)blockIdx.x + gridDim.x),
                i2 = cell_of((int64_t)blockIdx.x + 2 * (int64_t)gridDim.x);
        SpItem cur, nxt;                                             // items t and t+1 (first 32 neighbours)
        long long rs1, re1;                                          // row pointers of item t+1, in flight
        int nb2;                                                     // item t+2: neighbour ids and weights, in flight
        float wv2;
        {
            int nb;
            long long rs, re;
            fetch_edges(i0, 0, 0, nb, cur.wv);
            resolve(nb, rs, re);
            finish(cur, rs, re);
            fetch_edges(NG > 1 ? i0 : i1, NG > 1 ? 1 : 0, 0, nb2, wv2);   // item 1; resolved when item 0 starts
            prefetch_own(i0);
            preload(cur);
        }
        int step = 0;
        for (int64_t round = blockIdx.x; round < n_rounds; round += gridDim.x) {
            prefetch_own(i1);
            for (int g = 0; g < NG; ++g, ++step) {
                const int set = step & 1;
                const uint32_t Tw = (uint32_t)__cvta_generic_to_shared(sp_smem) + (uint32_t)(set * kSpCells + warp) * (uint32_t)slab4 * 16u;
                nxt.wv = wv2;
                resolve(nb2, rs1, re1);                              // item t+1
                {                                                    // item t+2
                    const int g2 = (g + 2) % NG, dr = (g + 2) / NG;
                    const int64_t ic = dr == 0 ? i0 : dr == 1 ? i1 : i2;
                    fetch_edges(ic, g2, 0, nb2, wv2);
                }
                float wsum = 0.f;
                group_sum(cur.wv, min(32, k), wsum);
                if (step >= 2) named_sync(3 + set);                  // the owners have folded and cleared this slab set
                consume(cur, Tw);
                for (int j0 = 32; j0 < k; j0 += 32) {                // k > 32: the later neighbours are resolved in line
                    SpItem more;
                    int nb;
                    long long rs, re;
                    fetch_edges(i0, g, j0, nb, more.wv);
                    resolve(nb, rs, re);
                    finish(more, rs, re);
                    group_sum(more.wv, min(32, k - j0), wsum);
                    preload(more);
                    consume(more, Tw);
                }
                if (lane == 0) {
                    const bool normalise = !(gs.has_obs && gs.slot0 + g == 0);
                    inv_s[set][warp] = (normalise && wsum != 0.f) ? 1.0f / wsum : 1.0f;
                    cell_s[set][warp] = (int32_t)i0;
                }
                finish(nxt, rs1, re1);
                cur = nxt;
                preload(cur);                                        // next item's rows: in flight from here on
                __threadfence_block();
                named_arrive(1 + set);
            }
            i0 = i1;
            i1 = i2;
            i2 = cell_of(round + 3 * (int64_t)gridDim.x);
        }
        // one owner arrival per used slab set is still pending
        if (step >= 1) named_sync(3 + ((step - 1) & 1));
        if (step >= 2) named_sync(3 + ((step - 2) & 1));
        return;
    }

    // ============================== column owners ==============================
    // sum S*T and sum T^2 of every graph live in shared memory (a thread's own quads: no conflicts, no atomics), fp32 over
    // <= 64 cells, then added to this CTA's fp64 partial; sum S^2 / sum S / nnz of the first launch stay in registers
    const int ot = tid - kSpOwners;
    float4* ACC4 = Tall4 + (size_t)2 * kSpCells * slab4;             // [NG][2][ld4]
    float p_ss[NQ][4], p_s[NQ][4];
    int p_nz[NQ][4];
#pragma unroll
    for (int m = 0; m < NQ; ++m)
#pragma unroll
        for (int c = 0; c < 4; ++c) { p_ss[m][c] = 0.f; p_s[m][c] = 0.f; p_nz[m][c] = 0; }
    const int n_rows = 2 * n_slots + 3;
    double* mine = partial + (size_t)blockIdx.x * n_rows * ldw;
    auto flush = [&]() {
#pragma unroll
        for (int m = 0; m < NQ; ++m) {
            const int qd = ot + kSpOwners * m;
            if (qd >= (int)ld4) continue;
            for (int g = 0; g < NG; ++g) {
                const float4 d = ACC4[(2 * g) * (int)ld4 + qd], t = ACC4[(2 * g + 1) * (int)ld4 + qd];
                ACC4[(2 * g) * (int)ld4 + qd] = make_float4(0.f, 0.f, 0.f, 0.f);
                ACC4[(2 * g + 1) * (int)ld4 + qd] = make_float4(0.f, 0.f, 0.f, 0.f);
                double* o = mine + (size_t)(2 * (gs.slot0 + g)) * ldw + 4 * qd;
                o[0] += (double)d.x; o[1] += (double)d.y; o[2] += (double)d.z; o[3] += (double)d.w;
                o += ldw;
                o[0] += (double)t.x; o[1] += (double)t.y; o[2] += (double)t.z; o[3] += (double)t.w;
            }
            if (first) {
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    const int col = 4 * qd + c;
                    mine[(size_t)(2 * n_slots) * ldw + col] += (double)p_ss[m][c];
                    mine[(size_t)(2 * n_slots + 1) * ldw + col] += (double)p_s[m][c];
                    mine[(size_t)(2 * n_slots + 2) * ldw + col] += (double)p_nz[m][c];
                    p_ss[m][c] = 0.f; p_s[m][c] = 0.f; p_nz[m][c] = 0;
                }
            }
        }
    };
    // The owners wait for the builders most of the time, and they have no registers tied up in rows: owner warp w runs
    // AHEAD of builder warp w and pulls the packed neighbour rows of its items into L2 (one bulk prefetch per row), so that
    // the builders' gathers - two rows in flight per warp is all their registers allow - see L2 latency instead of HBM's.
    const int ow = warp - kSpCells, olane = tid & 31;
    auto o_cell_of = [&](int64_t round) -> int64_t {
        const int64_t q = round * kSpCells + ow;
        if (round >= n_rounds || q >= n) return -1;
        return order ? (int64_t)order[q] : q;
    };
    auto o_neighbour = [&](int64_t i, int g) -> int {               // neighbour `olane` of cell i in graph g (first 32)
        if (i < 0 || olane >= k) return -1;
        const int slot = gs.slot0 + g;
        const bool obs = gs.has_obs && slot == 0;
        const int64_t at = (obs ? 0 : (int64_t)(slot - gs.has_obs) * gs.nk) + i * k + olane;
        return ldv_s32((obs ? gs.obs_idx : gs.null_idx) + at);
    };
    auto o_prefetch_row = [&](int nb) {
        if (nb < 0) return;
        const long long rs = ldv_s64(sp_ptr + nb), re = ldv_s64(sp_ptr + nb + 1);
        if (re > rs && re <= capacity)
            asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(entries + rs), "r"((uint32_t)((re - rs) * 8)) : "memory");
    };
    constexpr int kAhead = 3;                                        // items between the prefetch and the builder's use (5: slower)
    int64_t oi0 = o_cell_of(blockIdx.x), oi1 = o_cell_of((int64_t)blockIdx.x + gridDim.x),
            oi2 = o_cell_of((int64_t)blockIdx.x + 2 * (int64_t)gridDim.x), oi3 = o_cell_of((int64_t)blockIdx.x + 3 * (int64_t)gridDim.x);
    auto o_item_cell = [&](int g, int ahead, int& ga) -> int64_t {   // cell and graph of the item `ahead` steps after (round, g)
        ga = (g + ahead) % NG;
        const int dr = (g + ahead) / NG;
        return dr == 0 ? oi0 : dr == 1 ? oi1 : dr == 2 ? oi2 : dr == 3 ? oi3 : -1;
    };
    for (int a = 0; a < kAhead; ++a) {                              // the first items' rows
        int ga;
        const int64_t ic = o_item_cell(0, a, ga);
        o_prefetch_row(o_neighbour(ic, ga));
    }
    int step = 0, rounds_done = 0;
    for (int64_t round = blockIdx.x; round < n_rounds; round += gridDim.x) {
        for (int g = 0; g < NG; ++g, ++step) {
            const int set = step & 1;
            const bool stats1 = g == 0 && first;
            {
                int ga;
                const int64_t ic = o_item_cell(g, kAhead, ga);
                o_prefetch_row(o_neighbour(ic, ga));
            }
            named_sync(1 + set);                                     // the builders have finished this slab set
            float4* T4 = Tall4 + (size_t)set * kSpCells * slab4;
            float4 ad[NQ], at[NQ];
#pragma unroll
            for (int m = 0; m < NQ; ++m) {
                const int qd = ot + kSpOwners * m;
                ad[m] = at[m] = make_float4(0.f, 0.f, 0.f, 0.f);
                if (qd < (int)ld4) { ad[m] = ACC4[(2 * g) * (int)ld4 + qd]; at[m] = ACC4[(2 * g + 1) * (int)ld4 + qd]; }
            }
#pragma unroll
            for (int c0 = 0; c0 < kSpCells; c0 += 4) {
                float4 s[4][NQ];
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int ci = cell_s[set][c0 + u];
                    const float4* srow = S4 + (int64_t)(ci >= 0 ? ci : 0) * ld4 + ot;
#pragma unroll
                    for (int m = 0; m < NQ; ++m)
                        s[u][m] = (ci >= 0 && ot + kSpOwners * m < (int)ld4) ? __ldg(srow + kSpOwners * m) : make_float4(0.f, 0.f, 0.f, 0.f);
                }
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const float inv = inv_s[set][c0 + u];
                    float4* trow = T4 + (c0 + u) * slab4 + ot;
#pragma unroll
                    for (int m = 0; m < NQ; ++m) {
                        if (ot + kSpOwners * m >= (int)ld4) continue;
                        float4 t = trow[kSpOwners * m];
                        trow[kSpOwners * m] = make_float4(0.f, 0.f, 0.f, 0.f);
                        const float4 sv4 = s[u][m];
                        t.x *= inv; t.y *= inv; t.z *= inv; t.w *= inv;
                        ad[m].x = fmaf(sv4.x, t.x, ad[m].x); ad[m].y = fmaf(sv4.y, t.y, ad[m].y);
                        ad[m].z = fmaf(sv4.z, t.z, ad[m].z); ad[m].w = fmaf(sv4.w, t.w, ad[m].w);
                        at[m].x = fmaf(t.x, t.x, at[m].x); at[m].y = fmaf(t.y, t.y, at[m].y);
                        at[m].z = fmaf(t.z, t.z, at[m].z); at[m].w = fmaf(t.w, t.w, at[m].w);
                        if (stats1) {
                            const float sv[4] = {sv4.x, sv4.y, sv4.z, sv4.w};
#pragma unroll
                            for (int c = 0; c < 4; ++c) {
                                p_ss[m][c] = fmaf(sv[c], sv[c], p_ss[m][c]);
                                p_s[m][c] += sv[c];
                                p_nz[m][c] += sv[c] != 0.f ? 1 : 0;
                            }
                        }
                    }
                }
            }
#pragma unroll
            for (int m = 0; m < NQ; ++m) {
                const int qd = ot + kSpOwners * m;
                if (qd < (int)ld4) { ACC4[(2 * g) * (int)ld4 + qd] = ad[m]; ACC4[(2 * g + 1) * (int)ld4 + qd] = at[m]; }
            }
            __threadfence_block();
            named_arrive(3 + set);
        }
        if (++rounds_done * kSpCells >= 64) {                        // fp32 partials cover <= 64 cells
            flush();
            rounds_done = 0;
        }
        oi0 = oi1; oi1 = oi2; oi2 = oi3;
        oi3 = o_cell_of(round + 4 * (int64_t)gridDim.x);
    }
    flush();
}

// partial [grid][2*n_slots+3][ldw] -> out rows [n_slots][5][P]
__global__ void __launch_bounds__(256)
sparse_stats_reduce_kernel(const double* __restrict__ partial, int grid, int n_slots, int ldw, int P, double* __restrict__ out) {
    const int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;           // (slot, stat, col)
    if (e >= (int64_t)n_slots * 5 * P) return;
    const int col = (int)(e % P), stat = (int)((e / P) % 5), slot = (int)(e / ((int64_t)5 * P));
    const int n_rows = 2 * n_slots + 3;
    const int row = stat == 0 ? 2 * slot : stat == 2 ? 2 * slot + 1 : stat == 1 ? 2 * n_slots : stat == 3 ? 2 * n_slots + 1
                                                                                                         : 2 * n_slots + 2;
    double t = 0.0;
    for (int c = 0; c < grid; ++c) t += partial[((size_t)c * n_rows + row) * ldw + col];
    out[e] = t;
}

// shared memory of a launch: two sets of kSpCells slabs + the owners' sums of `ng` graphs, ldS floats each
static size_t sparse_smem(int64_t ldS, int ng) { return (size_t)(2 * kSpCells) * (ldS + 32) * 4 + (size_t)2 * ng * ldS * 4; }
constexpr size_t kSparseSmemMax = 200 * 1024;

template <int NQ, int CH, int RING>
static int launch_sparse_cfg(int64_t gs_capacity, int ng, const int64_t* sp_ptr, const int2* entries, const float* S, int64_t ld4, int64_t n,
                             const SparseGraphs& gs, int k, const int32_t* order, int n_slots, double* partial,
                             int grid, cudaStream_t st) {
    LARIS_CUDA(cudaFuncSetAttribute(spatial_stats_sparse_kernel<NQ, CH, RING>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    (int)kSparseSmemMax));
    spatial_stats_sparse_kernel<NQ, CH, RING><<<grid, 512, sparse_smem(ld4 * 4, ng), st>>>(
        sp_ptr, entries, (const float4*)S, ld4, n, gs, ng, k, order, n_slots, partial, gs_capacity);
    LARIS_LAUNCH_CHECK();
    return LARIS_OK;
}

// long_rows: segments of 8 chunks, 2 in flight (3 spill registers); else segments of 2 chunks, 8 in flight
template <int NQ>
static int launch_sparse(int64_t cap, bool long_rows, int ng, const int64_t* sp_ptr, const int2* entries, const float* S, int64_t ld4,
                         int64_t n, const SparseGraphs& gs, int k, const int32_t* order, int n_slots,
                         double* partial, int grid, cudaStream_t st) {
    static int ring = -1;                  // LARIS_B200_K4S_RING=3: three 8-chunk segments in flight instead of two (A/B runs; spills)
    if (ring < 0) {
        const char* e = getenv("LARIS_B200_K4S_RING");
        ring = e ? atoi(e) : 2;
    }
    if (long_rows && ring == 3)
        return launch_sparse_cfg<NQ, 8, 3>(cap, ng, sp_ptr, entries, S, ld4, n, gs, k, order, n_slots, partial, grid, st);
    return long_rows ? launch_sparse_cfg<NQ, 8, 2>(cap, ng, sp_ptr, entries, S, ld4, n, gs, k, order, n_slots, partial, grid, st)
                     : launch_sparse_cfg<NQ, 2, 8>(cap, ng, sp_ptr, entries, S, ld4, n, gs, k, order, n_slots, partial, grid, st);
}

}  // namespace laris

using namespace laris;

extern "C" int32_t laris_spatial_stats_sparse_max_ld(void) { return (int32_t)((kSparseSmemMax - 2 * kSpCells * 128) / (4 * (2 * kSpCells + 2)) / 4 * 4); }

extern "C" int laris_scores_pack(const float* S, int64_t ldS, int64_t n, int32_t P, const int64_t* sp_ptr, int64_t* entries,
                                 int64_t capacity, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    LARIS_CHECK_ARG(S && sp_ptr && entries, "laris_scores_pack: null pointer");
    LARIS_CHECK_ARG(n > 0 && P >= 1 && ldS >= P && capacity >= 0, "laris_scores_pack: need n > 0, ldS >= P >= 1, capacity >= 0");
    LARIS_CHECK_ARG(P <= 65536 * 32, "laris_scores_pack: P too large");
    const int T = (P + 31) / 32;
    const size_t smem = (size_t)kPackWarps * T * 32 * sizeof(uint16_t);
    LARIS_CHECK_ARG(smem <= 200 * 1024, "laris_scores_pack: P = %d exceeds the shared-memory queue (P <= %d)", P, 200 * 1024 / (kPackWarps * 2));
    LARIS_CUDA(cudaFuncSetAttribute(scores_pack_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int dev = 0, sms = kNumSMs;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    int per_sm = (int)((200 * 1024) / (smem + 1024));
    if (per_sm > 8) per_sm = 8;
    if (per_sm < 1) per_sm = 1;
    int64_t grid = (int64_t)sms * per_sm;
    if (grid > ceil_div(n, kPackWarps)) grid = ceil_div(n, kPackWarps);
    scores_pack_kernel<<<(unsigned)grid, kPackWarps * 32, smem, st>>>(S, ldS, n, P, sp_ptr, reinterpret_cast<int2*>(entries), capacity);
    LARIS_LAUNCH_CHECK();
    return LARIS_OK;
}

extern "C" int laris_spatial_stats_sparse(const float* S, int64_t ldS, int64_t n, int32_t P, const int64_t* sp_ptr,
                                          const int64_t* entries, int64_t n_entries, const int32_t* idx, const float* w, int k,
                                          const int32_t* order, const int32_t* null_idx, const float* null_w,
                                          int32_t n_repeats, double* out, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    LARIS_CHECK_ARG(S && sp_ptr && entries && out && (idx || n_repeats > 0), "laris_spatial_stats_sparse: null pointer");
    LARIS_CHECK_ARG(n > 0 && P >= 1 && ldS >= P && ldS % 4 == 0 && ((uintptr_t)S & 15) == 0,
                    "laris_spatial_stats_sparse: S needs ldS >= P, ldS %% 4 == 0 and 16-byte alignment");
    LARIS_CHECK_ARG(ldS <= laris_spatial_stats_sparse_max_ld(), "laris_spatial_stats_sparse: ldS = %lld exceeds %d (use laris_spatial_stats_batched)",
                    (long long)ldS, laris_spatial_stats_sparse_max_ld());
    LARIS_CHECK_ARG(k >= 1 && k <= 64, "laris_spatial_stats_sparse: k must be in [1,64] (got %d)", k);
    LARIS_CHECK_ARG(!idx || w, "laris_spatial_stats_sparse: null weights");
    LARIS_CHECK_ARG(n_repeats >= 0 && (n_repeats == 0 || (null_idx && null_w)), "laris_spatial_stats_sparse: bad null graphs");
    const int64_t ld4 = ldS / 4, nk = n * (int64_t)k;
    const int n_slots = (idx ? 1 : 0) + n_repeats;
    // column quads per owner thread (<= 4: ldS <= 4096 is implied by the shared-memory bound); graphs per launch: as many
    // as shared memory takes (they share the read of S), at most 4
    const int nq = (int)ceil_div(ld4, kSpOwners);
    const bool long_rows = n_entries >= 96 * n;                   // average padded row of >= 96 entries
    int max_g = 1;
    while (max_g < kSpMaxGraphs && sparse_smem(ldS, max_g + 1) <= kSparseSmemMax) ++max_g;
    if (const char* e = getenv("LARIS_B200_K4S_MAXG")) {          // A/B runs: fewer graphs per launch
        const int v = atoi(e);
        if (v >= 1 && v < max_g) max_g = v;
    }
    int dev = 0, sms = kNumSMs;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    int64_t grid = sms;
    if (grid > ceil_div(n, kSpCells)) grid = ceil_div(n, kSpCells);
    const int n_rows = 2 * n_slots + 3;
    Scratch<double> partial(st);
    LARIS_CUDA(partial.alloc((size_t)grid * n_rows * ldS));
    LARIS_CUDA(cudaMemsetAsync(partial.p, 0, (size_t)grid * n_rows * ldS * sizeof(double), st));
    const int2* ent = reinterpret_cast<const int2*>(entries);
    for (int s0 = 0; s0 < n_slots; s0 += max_g) {
        const int ng = n_slots - s0 < max_g ? n_slots - s0 : max_g;
        SparseGraphs gs;
        gs.obs_idx = idx; gs.obs_w = w; gs.null_idx = null_idx; gs.null_w = null_w;
        gs.nk = nk; gs.slot0 = s0; gs.has_obs = idx ? 1 : 0;
        int rc;
        if (nq == 1) rc = launch_sparse<1>(n_entries, long_rows, ng, sp_ptr, ent, S, ld4, n, gs, k, order, n_slots, partial.p, (int)grid, st);
        else if (nq == 2) rc = launch_sparse<2>(n_entries, long_rows, ng, sp_ptr, ent, S, ld4, n, gs, k, order, n_slots, partial.p, (int)grid, st);
        else if (nq == 3) rc = launch_sparse<3>(n_entries, long_rows, ng, sp_ptr, ent, S, ld4, n, gs, k, order, n_slots, partial.p, (int)grid, st);
        else rc = launch_sparse<4>(n_entries, long_rows, ng, sp_ptr, ent, S, ld4, n, gs, k, order, n_slots, partial.p, (int)grid, st);
        if (rc) return rc;
    }
    sparse_stats_reduce_kernel<<<(unsigned)ceil_div((int64_t)n_slots * 5 * P, 256), 256, 0, st>>>(partial.p, (int)grid, n_slots,
                                                                                                 (int)ldS, P, out);
    LARIS_LAUNCH_CHECK();
    return LARIS_OK;
}
