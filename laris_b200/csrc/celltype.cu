// K5 family: cell-type contractions.
//   neighborhood_composition  N[i,a] = sum of w over neighbours of type a
//   celltype_nonzero_count    cnt[c,p] = #{i in c : S[i,p] != 0}            (-> frac)
//   onehot_contract_csr       sum/cnt of Xu per (type, gene) + column sum of squares (-> COSG cosine)
//   celltype_pair_contract    num[p,(a,b)] = sum_i S[i,p] N[i,a] N[i,b]   (fp32 CUDA-core tiles; the
//                             tcgen05 variant lives in celltype_mma.cu)
#include "common.cuh"
#include "rng.cuh"
#include "scan.cuh"

namespace laris {

__global__ void neighborhood_composition_kernel(const int32_t* __restrict__ labels, const int32_t* __restrict__ idx,
                                                const float* __restrict__ w, int64_t n, int k, int C,
                                                float* __restrict__ N) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    float* row = N + i * C;                               // zeroed by the caller; this thread owns the row
    for (int j = 0; j < k; ++j) row[labels[idx[i * k + j]]] += w[i * k + j];
}

// ---- counts of non-zero scores per (cell type, pair): thread owns 4 columns, CTA-private
//      shared table [C][4][128] so increments never collide and never conflict on banks
constexpr int kCcThreads = 128;
constexpr int kCcUnroll = 8;

// cells grouped by type (device counting sort; the order inside a type does not matter for integer counts)
__global__ void label_histogram_kernel(const int32_t* __restrict__ labels, int64_t n, int C, int32_t* __restrict__ hist) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i < n) {
        const int c = labels[i];
        if (c >= 0 && c < C) atomicAdd(&hist[c], 1);
    }
}
__global__ void label_scatter_kernel(const int32_t* __restrict__ labels, int64_t n, int C, const int32_t* __restrict__ start,
                                     int32_t* __restrict__ cursor, int32_t* __restrict__ rows, int32_t* __restrict__ row_label) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i < n) {
        const int c = labels[i];
        if (c >= 0 && c < C) {
            const int pos = start[c] + atomicAdd(&cursor[c], 1);
            rows[pos] = (int32_t)i;
            row_label[pos] = c;
        }
    }
}

// A thread owns 4 adjacent pair columns and a CTA one contiguous range of the type-grouped cell list: the counts of
// the current type live in registers and are flushed (integer atomics, exact) only where the type changes, so the
// loop is nothing but kCcUnroll independent 128-bit streaming loads per step - an HBM-bound pass over S.
__global__ void __launch_bounds__(kCcThreads)
celltype_nonzero_count_kernel(const float4* __restrict__ S4, int64_t ld4, int64_t m, int P,
                              const int32_t* __restrict__ rows, const int32_t* __restrict__ row_label, int nchunks,
                              unsigned long long* __restrict__ counts /*[C][P]*/) {
    const int64_t col4 = (int64_t)blockIdx.x * kCcThreads + threadIdx.x;
    if (col4 >= ld4) return;
    const int64_t per = (m + nchunks - 1) / nchunks;
    const int64_t r0 = (int64_t)blockIdx.y * per, r1 = min(m, r0 + per);
    if (r0 >= r1) return;
    uint32_t c0 = 0, c1 = 0, c2 = 0, c3 = 0;
    int cur = row_label[r0];
    if (cur < 0) return;                                             // the chunk lies in the unused tail
    auto flush = [&](int lab) {
        const int64_t p = col4 * 4;
        unsigned long long* o = counts + (int64_t)lab * P + p;
        if (c0 && p < P) atomicAdd(o, (unsigned long long)c0);
        if (c1 && p + 1 < P) atomicAdd(o + 1, (unsigned long long)c1);
        if (c2 && p + 2 < P) atomicAdd(o + 2, (unsigned long long)c2);
        if (c3 && p + 3 < P) atomicAdd(o + 3, (unsigned long long)c3);
        c0 = c1 = c2 = c3 = 0;
    };
    for (int64_t r = r0; r < r1; r += kCcUnroll) {
        float4 v[kCcUnroll];
        int lab[kCcUnroll];
#pragma unroll
        for (int u = 0; u < kCcUnroll; ++u) {
            lab[u] = r + u < r1 ? row_label[r + u] : -1;
            const bool ok = lab[u] >= 0;
            v[u] = ok ? ld_stream_f4(S4 + (int64_t)rows[r + u] * ld4 + col4) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
#pragma unroll
        for (int u = 0; u < kCcUnroll; ++u) {
            if (lab[u] < 0) break;
            if (lab[u] != cur) { flush(cur); cur = lab[u]; }
            c0 += v[u].x != 0.f; c1 += v[u].y != 0.f; c2 += v[u].z != 0.f; c3 += v[u].w != 0.f;
        }
    }
    flush(cur);
}

// The same counts from the PACKED rows of S (csrc/stats_sparse.cu: entries {column, value bits}, rows padded to whole
// 32-entry chunks with zero-valued entries): 8 bytes per non-zero instead of 4P bytes per cell.  A CTA takes a contiguous
// range of the type-grouped cell list and walks it one type segment at a time; a warp owns a row, its lanes count the
// row's entries into a per-CTA shared-memory table [P] (a row's columns are distinct, so the adds of a row never collide
// inside the warp; rows of different warps do, hence atomics), flushed with global integer atomics at the end of a segment.
constexpr int kPcThreads = 256;

__global__ void __launch_bounds__(kPcThreads)
nonzero_count_packed_kernel(const int64_t* __restrict__ sp_ptr, const int2* __restrict__ entries, int64_t capacity, int P,
                            const int32_t* __restrict__ rows, const int32_t* __restrict__ start /*[C+1]*/, int C,
                            unsigned long long* __restrict__ counts /*[C][P]*/) {
    extern __shared__ uint32_t pc_cnt[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t m = start[C];                                       // validly labelled cells
    const int64_t per = (m + gridDim.x - 1) / gridDim.x;
    int64_t r0 = (int64_t)blockIdx.x * per;
    const int64_t r1 = min(m, r0 + per);
    if (r0 >= r1) return;
    for (int i = threadIdx.x; i < P; i += kPcThreads) pc_cnt[i] = 0u;
    int lo = 0, hi = C;                                               // type of cell r0: last c with start[c] <= r0
    while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if ((int64_t)start[mid] <= r0) lo = mid; else hi = mid;
    }
    int c = lo;
    __syncthreads();
    while (r0 < r1) {
        while (c + 1 < C && (int64_t)start[c + 1] <= r0) ++c;          // skip types without cells
        const int64_t seg_end = min(r1, (int64_t)start[c + 1]);
        for (int64_t r = r0 + warp; r < seg_end; r += kPcThreads / 32) {
            const int64_t i = rows[r];
            const int64_t rs = sp_ptr[i], re = min(sp_ptr[i + 1], capacity);
            for (int64_t e = rs + lane; e < re; e += 128) {            // four chunks in flight per lane
                int2 en[4];
#pragma unroll
                for (int u = 0; u < 4; ++u)
                    en[u] = e + 32 * u < re ? __ldg(entries + e + 32 * u) : make_int2(0, 0);
#pragma unroll
                for (int u = 0; u < 4; ++u)
                    if (en[u].y != 0 && en[u].x < P) atomicAdd(&pc_cnt[en[u].x], 1u);
            }
        }
        __syncthreads();
        for (int i = threadIdx.x; i < P; i += kPcThreads) {
            const uint32_t v = pc_cnt[i];
            if (v) {
                atomicAdd(counts + (int64_t)c * P + i, (unsigned long long)v);
                pc_cnt[i] = 0u;
            }
        }
        __syncthreads();
        r0 = seg_end;
    }
}

// ---- per (type, gene) sums straight from the packed CSR: warp per row
__global__ void onehot_contract_csr_kernel(const int64_t* __restrict__ row_start, const int64_t* __restrict__ row_end,
                                           const int64_t* __restrict__ packed,
                                           int64_t n, int G_u, const int32_t* __restrict__ labels,
                                           double* __restrict__ sum, double* __restrict__ cnt, double* __restrict__ colsq) {
    const int lane = threadIdx.x & 31;
    const int64_t row = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    if (row >= n) return;
    const int64_t base = (int64_t)labels[row] * G_u;
    for (int64_t t = row_start[row] + lane; t < row_end[row]; t += 32) {
        const int64_t pk = packed[t];
        const uint32_t g = (uint32_t)(pk & 0xffffffffll);
        const double v = (double)__uint_as_float((uint32_t)((uint64_t)pk >> 32));
        atomicAdd(&sum[base + g], v);
        atomicAdd(&cnt[base + g], 1.0);
        atomicAdd(&colsq[g], v * v);
    }
}

// ---- ||Q_(ab)||^2 = sum_i (N[i,a] N[i,b])^2
__global__ void pair_profile_norm_kernel(const float* __restrict__ N, int64_t n, int C, double* __restrict__ qnorm2) {
    extern __shared__ float nt[];                          // [rows][C]
    const int rows = 64;
    const int64_t r0 = (int64_t)blockIdx.x * rows;
    const int nr = (int)min((int64_t)rows, n - r0);
    for (int q = threadIdx.x; q < nr * C; q += blockDim.x) nt[q] = N[r0 * C + q];
    __syncthreads();
    for (int ab = threadIdx.x; ab < C * C; ab += blockDim.x) {
        const int a = ab / C, b = ab - a * C;
        double s = 0.0;
        for (int r = 0; r < nr; ++r) {
            const float q = nt[r * C + a] * nt[r * C + b];
            s += (double)(q * q);
        }
        if (s != 0.0) atomicAdd(&qnorm2[ab], s);
    }
}

// ---- fp32 CUDA-core contraction: 64 (pairs) x 64 (type pairs) tile per CTA, 4x4 per thread,
//      split over cells; partial sums are flushed to fp64 every kFlush cells
constexpr int kCtTile = 64, kCtKc = 16, kCtThreads = 256, kCtFlush = 512;

__global__ void __launch_bounds__(kCtThreads)
celltype_pair_contract_kernel(const float* __restrict__ S, int64_t ldS, int64_t n, int P, const float* __restrict__ N,
                              int C, int64_t cells_per_split, double* __restrict__ num /*[P][C*C]*/) {
    __shared__ float As[kCtKc][kCtTile + 4];               // [cell][pair]
    __shared__ float Bs[kCtKc][kCtTile + 4];               // [cell][type pair]
    extern __shared__ float Ns[];                          // [kCtKc][C]
    const int CC = C * C;
    const int p0 = blockIdx.x * kCtTile, q0 = blockIdx.y * kCtTile;
    const int64_t i_beg = (int64_t)blockIdx.z * cells_per_split;
    const int64_t i_end = min(n, i_beg + cells_per_split);
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;       // tx -> type-pair quad, ty -> pair quad
    float acc[4][4] = {};
    double dacc[4][4] = {};
    int since_flush = 0;
    for (int64_t i0 = i_beg; i0 < i_end; i0 += kCtKc) {
        __syncthreads();
        for (int q = tid; q < kCtKc * kCtTile; q += kCtThreads) {
            const int r = q / kCtTile, c = q - r * kCtTile;
            const int64_t i = i0 + r;
            As[r][c] = (i < i_end && p0 + c < P) ? S[i * ldS + p0 + c] : 0.f;
        }
        for (int q = tid; q < kCtKc * C; q += kCtThreads) {
            const int r = q / C;
            const int64_t i = i0 + r;
            Ns[q] = i < i_end ? N[i * C + (q - r * C)] : 0.f;
        }
        __syncthreads();
        for (int q = tid; q < kCtKc * kCtTile; q += kCtThreads) {
            const int r = q / kCtTile, c = q - r * kCtTile;
            const int ab = q0 + c;
            float v = 0.f;
            if (ab < CC) { const int a = ab / C; v = Ns[r * C + a] * Ns[r * C + (ab - a * C)]; }
            Bs[r][c] = v;
        }
        __syncthreads();
#pragma unroll
        for (int r = 0; r < kCtKc; ++r) {
            const float4 a = *reinterpret_cast<const float4*>(&As[r][ty * 4]);
            const float4 b = *reinterpret_cast<const float4*>(&Bs[r][tx * 4]);
            const float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
            for (int u = 0; u < 4; ++u)
#pragma unroll
                for (int v = 0; v < 4; ++v) acc[u][v] = fmaf(av[u], bv[v], acc[u][v]);
        }
        since_flush += kCtKc;
        if (since_flush >= kCtFlush) {
#pragma unroll
            for (int u = 0; u < 4; ++u)
#pragma unroll
                for (int v = 0; v < 4; ++v) { dacc[u][v] += (double)acc[u][v]; acc[u][v] = 0.f; }
            since_flush = 0;
        }
    }
#pragma unroll
    for (int u = 0; u < 4; ++u)
#pragma unroll
        for (int v = 0; v < 4; ++v) {
            const int p = p0 + ty * 4 + u, ab = q0 + tx * 4 + v;
            const double val = dacc[u][v] + (double)acc[u][v];
            if (p < P && ab < CC && val != 0.0) atomicAdd(&num[(int64_t)p * CC + ab], val);
        }
}

}  // namespace laris

using namespace laris;

extern "C" int laris_neighborhood_composition(const int32_t* labels, const int32_t* idx, const float* w, int64_t n, int k,
                                              int32_t C, float* Ncomp, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    LARIS_CHECK_ARG(labels && idx && w && Ncomp && n > 0 && k >= 1 && C >= 1, "laris_neighborhood_composition: bad arguments");
    LARIS_CUDA(cudaMemsetAsync(Ncomp, 0, (size_t)n * C * sizeof(float), st));
    neighborhood_composition_kernel<<<(unsigned)ceil_div(n, 256), 256, 0, st>>>(labels, idx, w, n, k, C, Ncomp);
    LARIS_LAUNCH_CHECK();
    return LARIS_OK;
}

// cells grouped by type: histogram -> scan -> scatter (cells with a label outside [0,C) are skipped); start [C+1],
// rows / row_label [n] (row_label = -1 in the unused tail)
static int group_cells_by_type(const int32_t* labels, int64_t n, int C, int32_t* hist, int32_t* start, int32_t* cursor,
                               int32_t* rows, int32_t* row_label, cudaStream_t st) {
    LARIS_CUDA(cudaMemsetAsync(hist, 0, (size_t)C * sizeof(int32_t), st));
    LARIS_CUDA(cudaMemsetAsync(cursor, 0, (size_t)C * sizeof(int32_t), st));
    LARIS_CUDA(cudaMemsetAsync(row_label, 0xff, (size_t)n * sizeof(int32_t), st));     // -1: unused tail
    const unsigned gn = (unsigned)ceil_div(n, 256);
    label_histogram_kernel<<<gn, 256, 0, st>>>(labels, n, C, hist);
    LARIS_LAUNCH_CHECK();
    int rc = exclusive_scan<int32_t>(hist, C, start, st);
    if (rc) return rc;
    label_scatter_kernel<<<gn, 256, 0, st>>>(labels, n, C, start, cursor, rows, row_label);
    LARIS_LAUNCH_CHECK();
    return LARIS_OK;
}

extern "C" int laris_celltype_nonzero_count(const float* S, int64_t ldS, int64_t n, int32_t P, const int32_t* labels,
                                            int32_t C, uint64_t* counts, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    LARIS_CHECK_ARG(S && labels && counts && n > 0 && P >= 1 && C >= 1, "laris_celltype_nonzero_count: bad arguments");
    LARIS_CHECK_ARG(ldS >= P && ldS % 4 == 0 && ((uintptr_t)S & 15) == 0, "laris_celltype_nonzero_count: S layout");
    LARIS_CHECK_ARG(n < (int64_t)INT32_MAX, "laris_celltype_nonzero_count: n must fit int32");
    LARIS_CUDA(cudaMemsetAsync(counts, 0, (size_t)C * P * sizeof(uint64_t), st));
    Scratch<int32_t> hist(st), start(st), cursor(st), rows(st), row_label(st);
    LARIS_CUDA(hist.alloc(C));
    LARIS_CUDA(start.alloc(C + 1));
    LARIS_CUDA(cursor.alloc(C));
    LARIS_CUDA(rows.alloc(n));
    LARIS_CUDA(row_label.alloc(n));
    int rc = group_cells_by_type(labels, n, C, hist.p, start.p, cursor.p, rows.p, row_label.p, st);
    if (rc) return rc;
    const int64_t ld4 = ldS / 4;
    const int slabs = (int)ceil_div(ld4, kCcThreads);
    int64_t nchunks = ceil_div((int64_t)kNumSMs * 16, slabs);          // ~16 resident 128-thread CTAs per SM
    if (nchunks > ceil_div(n, kCcUnroll)) nchunks = ceil_div(n, kCcUnroll);
    if (nchunks > 65535) nchunks = 65535;
    dim3 grid((unsigned)slabs, (unsigned)nchunks);
    // rows beyond the number of validly labelled cells carry label -1 and end the chunk loops
    celltype_nonzero_count_kernel<<<grid, kCcThreads, 0, st>>>((const float4*)S, ld4, n, P, rows.p, row_label.p, (int)nchunks,
                                                               (unsigned long long*)counts);
    LARIS_LAUNCH_CHECK();
    return LARIS_OK;
}

extern "C" int laris_celltype_nonzero_count_packed(const int64_t* sp_ptr, const void* entries, int64_t capacity, int64_t n,
                                                   int32_t P, const int32_t* labels, int32_t C, uint64_t* counts,
                                                   void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    LARIS_CHECK_ARG(sp_ptr && entries && labels && counts && n > 0 && P >= 1 && C >= 1 && capacity >= 0,
                    "laris_celltype_nonzero_count_packed: bad arguments");
    LARIS_CHECK_ARG(n < (int64_t)INT32_MAX, "laris_celltype_nonzero_count_packed: n must fit int32");
    LARIS_CHECK_ARG((size_t)P * sizeof(uint32_t) <= (size_t)200 * 1024,
                    "laris_celltype_nonzero_count_packed: P=%d does not fit shared memory", P);
    LARIS_CUDA(cudaMemsetAsync(counts, 0, (size_t)C * P * sizeof(uint64_t), st));
    Scratch<int32_t> hist(st), start(st), cursor(st), rows(st), row_label(st);
    LARIS_CUDA(hist.alloc(C));
    LARIS_CUDA(start.alloc(C + 1));
    LARIS_CUDA(cursor.alloc(C));
    LARIS_CUDA(rows.alloc(n));
    LARIS_CUDA(row_label.alloc(n));
    int rc = group_cells_by_type(labels, n, C, hist.p, start.p, cursor.p, rows.p, row_label.p, st);
    if (rc) return rc;
    const size_t smem = (size_t)P * sizeof(uint32_t);
    LARIS_CUDA(cudaFuncSetAttribute(nonzero_count_packed_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int64_t grid = (int64_t)kNumSMs * 8;                               // 8 x 256 threads per SM when the table is small
    if (grid > ceil_div(n, 64)) grid = ceil_div(n, 64);
    nonzero_count_packed_kernel<<<(unsigned)grid, kPcThreads, smem, st>>>(sp_ptr, (const int2*)entries, capacity, P, rows.p,
                                                                         start.p, C, (unsigned long long*)counts);
    LARIS_LAUNCH_CHECK();
    return LARIS_OK;
}

// ---- the same sums with the cells grouped by type: a CTA takes <= kOhRows rows of ONE type and accumulates its
// per-gene sums, squares and counts in shared memory, then adds each touched gene to the fp64 tables once - ~13x fewer
// global fp64 atomics than the row-at-a-time kernel above (5.4 -> ~1 ms at 1e6 cells x 1350 genes).  The rows of a CTA
// are taken one after the other by all threads (a row's genes are distinct, so the adds of one row never collide and
// need no atomics; a barrier separates rows; the next row's entries are already in registers), in fp64: the result
// differs from the row-at-a-time kernel only in fp64 summation order.
constexpr int kOhRows = 128, kOhThreads = 128, kOhMaxColumns = 8000, kOhPerThread = 4;

__global__ void onehot_chunk_base_kernel(const int32_t* __restrict__ start, int C, int32_t* __restrict__ chunk_base) {
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        int32_t acc = 0;
        for (int t = 0; t < C; ++t) {
            chunk_base[t] = acc;
            acc += (start[t + 1] - start[t] + kOhRows - 1) / kOhRows;
        }
        chunk_base[C] = acc;
    }
}

__global__ void __launch_bounds__(kOhThreads)
onehot_contract_grouped_kernel(const int64_t* __restrict__ row_start, const int64_t* __restrict__ row_end,
                               const int64_t* __restrict__ packed, int G_u, const int32_t* __restrict__ rows,
                               const int32_t* __restrict__ start, const int32_t* __restrict__ chunk_base, int C,
                               double* __restrict__ sum, double* __restrict__ cnt, double* __restrict__ colsq) {
    extern __shared__ double oh_smem[];
    double* s_sum = oh_smem;
    double* s_sq = oh_smem + G_u;
    uint32_t* s_cnt = reinterpret_cast<uint32_t*>(oh_smem + 2 * (size_t)G_u);
    const int b = blockIdx.x;
    if (b >= chunk_base[C]) return;
    int lo = 0, hi = C;                                     // type t with chunk_base[t] <= b < chunk_base[t+1]
    while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (chunk_base[mid] <= b) lo = mid; else hi = mid;
    }
    const int t = lo;
    const int p0 = start[t] + (b - chunk_base[t]) * kOhRows, p1 = min(start[t + 1], p0 + kOhRows);
    for (int g = threadIdx.x; g < G_u; g += kOhThreads) { s_sum[g] = 0.0; s_sq[g] = 0.0; s_cnt[g] = 0u; }
    // entries of a row: thread e, e + 128, ...; the first kOhPerThread of the NEXT row are fetched before this row is added
    int64_t nx[kOhPerThread];
    int64_t nx_s = 0, nx_e = 0;
    auto fetch = [&](int pos) {
        const int64_t row = rows[pos];
        nx_s = row_start[row];
        nx_e = row_end[row];
#pragma unroll
        for (int j = 0; j < kOhPerThread; ++j) {
            const int64_t e = nx_s + threadIdx.x + (int64_t)j * kOhThreads;
            nx[j] = e < nx_e ? packed[e] : -1;
        }
    };
    if (p0 < p1) fetch(p0);
    __syncthreads();
    for (int pos = p0; pos < p1; ++pos) {
        int64_t cur[kOhPerThread];
        const int64_t cs = nx_s, ce = nx_e;
#pragma unroll
        for (int j = 0; j < kOhPerThread; ++j) cur[j] = nx[j];
        if (pos + 1 < p1) fetch(pos + 1);
        auto add = [&](int64_t pk) {
            const uint32_t g = (uint32_t)(pk & 0xffffffffll);
            const double v = (double)__uint_as_float((uint32_t)((uint64_t)pk >> 32));
            s_sum[g] += v;
            s_sq[g] += v * v;
            s_cnt[g] += 1u;
        };
#pragma unroll
        for (int j = 0; j < kOhPerThread; ++j)
            if (cs + threadIdx.x + (int64_t)j * kOhThreads < ce) add(cur[j]);
        for (int64_t e = cs + threadIdx.x + (int64_t)kOhPerThread * kOhThreads; e < ce; e += kOhThreads) add(packed[e]);
        __syncthreads();                                    // the next row may touch the same genes
    }
    const int64_t base = (int64_t)t * G_u;
    for (int g = threadIdx.x; g < G_u; g += kOhThreads) {
        const uint32_t c = s_cnt[g];
        if (c) {
            atomicAdd(&sum[base + g], s_sum[g]);
            atomicAdd(&cnt[base + g], (double)c);
            atomicAdd(&colsq[g], s_sq[g]);
        }
    }
}

static int onehot_contract_impl(const int64_t* row_start, const int64_t* row_end, const int64_t* packed, int64_t n,
                                int32_t G_u, const int32_t* labels, int32_t C, double* sum, double* cnt, double* colsq,
                                cudaStream_t st, const char* who) {
    LARIS_CHECK_ARG(row_start && row_end && labels && sum && cnt && colsq && n > 0 && G_u >= 1 && C >= 1, "%s: bad arguments", who);
    LARIS_CUDA(cudaMemsetAsync(sum, 0, (size_t)C * G_u * sizeof(double), st));
    LARIS_CUDA(cudaMemsetAsync(cnt, 0, (size_t)C * G_u * sizeof(double), st));
    LARIS_CUDA(cudaMemsetAsync(colsq, 0, (size_t)G_u * sizeof(double), st));
    if (G_u > kOhMaxColumns || n >= (int64_t)INT32_MAX) {     // very wide panels: the row-at-a-time kernel
        onehot_contract_csr_kernel<<<(unsigned)ceil_div(n * 32, 256), 256, 0, st>>>(row_start, row_end, packed, n, G_u, labels,
                                                                                    sum, cnt, colsq);
        LARIS_LAUNCH_CHECK();
        return LARIS_OK;
    }
    // group the cells by type: histogram -> scan -> scatter (cells with a label outside [0,C) are skipped)
    Scratch<int32_t> hist(st), start(st), cursor(st), rows(st), row_label(st), chunk_base(st);
    LARIS_CUDA(hist.alloc(C));
    LARIS_CUDA(start.alloc(C + 1));
    LARIS_CUDA(cursor.alloc(C));
    LARIS_CUDA(rows.alloc(n));
    LARIS_CUDA(row_label.alloc(n));
    LARIS_CUDA(chunk_base.alloc(C + 1));
    LARIS_CUDA(cudaMemsetAsync(hist.p, 0, (size_t)C * sizeof(int32_t), st));
    LARIS_CUDA(cudaMemsetAsync(cursor.p, 0, (size_t)C * sizeof(int32_t), st));
    const unsigned gn = (unsigned)ceil_div(n, 256);
    label_histogram_kernel<<<gn, 256, 0, st>>>(labels, n, C, hist.p);
    LARIS_LAUNCH_CHECK();
    int rc = exclusive_scan<int32_t>(hist.p, C, start.p, st);
    if (rc) return rc;
    label_scatter_kernel<<<gn, 256, 0, st>>>(labels, n, C, start.p, cursor.p, rows.p, row_label.p);
    LARIS_LAUNCH_CHECK();
    onehot_chunk_base_kernel<<<1, 32, 0, st>>>(start.p, C, chunk_base.p);
    LARIS_LAUNCH_CHECK();
    const size_t smem = (size_t)G_u * (2 * sizeof(double) + sizeof(uint32_t)) + 16;
    LARIS_CUDA(cudaFuncSetAttribute(onehot_contract_grouped_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int64_t bound = ceil_div(n, kOhRows) + C;
    onehot_contract_grouped_kernel<<<(unsigned)bound, kOhThreads, smem, st>>>(row_start, row_end, packed, G_u, rows.p, start.p,
                                                                            chunk_base.p, C, sum, cnt, colsq);
    LARIS_LAUNCH_CHECK();
    return LARIS_OK;
}

extern "C" int laris_onehot_contract_csr(const int64_t* indptr, const int64_t* packed, int64_t n, int32_t G_u,
                                         const int32_t* labels, int32_t C, double* sum, double* cnt, double* colsq,
                                         void* stream) {
    LARIS_CHECK_ARG(indptr, "laris_onehot_contract_csr: bad arguments");
    return onehot_contract_impl(indptr, indptr + 1, packed, n, G_u, labels, C, sum, cnt, colsq, (cudaStream_t)stream,
                                "laris_onehot_contract_csr");
}

extern "C" int laris_onehot_contract_rows(const int64_t* row_start, const int64_t* row_end, const int64_t* packed, int64_t n,
                                          int32_t G_u, const int32_t* labels, int32_t C, double* sum, double* cnt,
                                          double* colsq, void* stream) {
    return onehot_contract_impl(row_start, row_end, packed, n, G_u, labels, C, sum, cnt, colsq, (cudaStream_t)stream,
                                "laris_onehot_contract_rows");
}

namespace laris {
int celltype_pair_contract_mma(const float* S, int64_t ldS, int64_t n, int32_t P, const float* Ncomp, int32_t C,
                               const int32_t* order, double* num, cudaStream_t st);   // celltype_mma.cu
}

// ---- label-permutation null (extension): labels_out[i] = labels[pi_t(i)], pi_t a keyed Feistel bijection of [0, n)
namespace laris {
__global__ void permute_labels_kernel(const int32_t* __restrict__ labels, int64_t n, FeistelKeys fk, int32_t* __restrict__ out) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i < n) out[i] = labels[feistel_perm((uint64_t)i, (uint64_t)n, fk)];
}
}  // namespace laris

extern "C" int laris_permute_labels(const int32_t* labels, int64_t n, int32_t permutation, uint64_t seed, int32_t* out,
                                    void* stream) {
    LARIS_CHECK_ARG(labels && out && n > 0 && permutation >= 0, "laris_permute_labels: bad arguments");
    const FeistelKeys fk = make_feistel_keys((uint64_t)n, kLabelPermBase + (uint32_t)permutation, seed);
    permute_labels_kernel<<<(unsigned)ceil_div(n, 256), 256, 0, (cudaStream_t)stream>>>(labels, n, fk, out);
    LARIS_LAUNCH_CHECK();
    return LARIS_OK;
}

extern "C" int laris_pair_profile_norm(const float* Ncomp, int64_t n, int32_t C, double* qnorm2, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    LARIS_CHECK_ARG(Ncomp && qnorm2 && n > 0 && C >= 1, "laris_pair_profile_norm: bad arguments");
    LARIS_CUDA(cudaMemsetAsync(qnorm2, 0, (size_t)C * C * sizeof(double), st));
    pair_profile_norm_kernel<<<(unsigned)ceil_div(n, 64), 256, 64 * C * sizeof(float), st>>>(Ncomp, n, C, qnorm2);
    LARIS_LAUNCH_CHECK();
    return LARIS_OK;
}

extern "C" int laris_celltype_pair_contract(const float* S, int64_t ldS, int64_t n, int32_t P, const float* Ncomp,
                                            int32_t C, const int32_t* order, double* num, double* qnorm2, int impl,
                                            void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    LARIS_CHECK_ARG(S && Ncomp && num && qnorm2 && n > 0 && P >= 1 && C >= 1 && ldS >= P, "laris_celltype_pair_contract: bad arguments");
    LARIS_CHECK_ARG(impl == 0 || impl == 1, "laris_celltype_pair_contract: impl must be 0 (tcgen05) or 1 (fp32 CUDA cores)");
    const int CC = C * C;
    LARIS_CUDA(cudaMemsetAsync(num, 0, (size_t)P * CC * sizeof(double), st));
    LARIS_CUDA(cudaMemsetAsync(qnorm2, 0, (size_t)CC * sizeof(double), st));
    pair_profile_norm_kernel<<<(unsigned)ceil_div(n, 64), 256, 64 * C * sizeof(float), st>>>(Ncomp, n, C, qnorm2);
    LARIS_LAUNCH_CHECK();
    if (impl == 0) return celltype_pair_contract_mma(S, ldS, n, P, Ncomp, C, order, num, st);
    const int tm = (int)ceil_div(P, kCtTile), tn = (int)ceil_div(CC, kCtTile);
    int64_t splits = ceil_div((int64_t)kNumSMs * 4, (int64_t)tm * tn);
    int64_t per = ceil_div(ceil_div(n, splits), kCtKc) * kCtKc;
    if (per < kCtKc) per = kCtKc;
    splits = ceil_div(n, per);
    if (splits > 65535) { splits = 65535; per = ceil_div(ceil_div(n, splits), kCtKc) * kCtKc; splits = ceil_div(n, per); }
    dim3 grid((unsigned)tm, (unsigned)tn, (unsigned)splits);
    celltype_pair_contract_kernel<<<grid, kCtThreads, (size_t)kCtKc * C * sizeof(float), st>>>(S, ldS, n, P, Ncomp, C, per, num);
    LARIS_LAUNCH_CHECK();
    return LARIS_OK;
}
