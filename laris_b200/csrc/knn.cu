// K1: exact Euclidean kNN on a uniform grid, emitted directly as fixed-degree CSR
// (+ the distance-decay weight transform).  See include/laris_b200.h for the
// reference call sites this replaces.
//
// Layout: points are binned into a uniform grid (~2.5 points per cell), a
// counting sort gives the cell-ordered point list (which is also the spatial
// processing order handed to the downstream kernels), and one thread per query
// walks Chebyshev rings of cells until the k-th best distance is provably
// smaller than anything outside the visited block.  Distances follow sklearn's
// KD-tree arithmetic bit for bit: d2 = ((0 + dx*dx) + dy*dy [+ dz*dz]) with
// separate fp64 multiply/add roundings (no FMA), then IEEE sqrt.
#include <float.h>
#include <math.h>

#include "common.cuh"
#include "scan.cuh"

namespace laris {

struct GridSpec {
    double lo[3];
    double inv_h;
    double h;
    int g[3];
    int dim;
    int64_t ncell;      // g[0]*g[1]*g[2]
    int64_t nslot;      // size of the strip-order slot table
    int strip_h;        // strip height in cells (min(kStripH, g[1]))
    int valid;          // 0: non-finite coordinates - every kernel leaves its outputs untouched except the
                        //    searches, which write index -1 / distance NaN
};

// ------------------------------------------------------------------ bbox --
__device__ void derive_grid_spec(const double* box, int64_t n, int dim, double per_cell, double min_h, int64_t cell_bound,
                                 GridSpec* out);

// bbox partials + (last CTA to finish) the final box and the grid shape, in ONE launch: the CTA that takes the last
// ticket sees every partial (threadfence before the ticket), reduces them in a fixed order and derives the GridSpec.
__global__ void __launch_bounds__(256)
bbox_grid_kernel(const double* __restrict__ xy, int64_t n, int dim, double per_cell, double min_h, int64_t cell_bound,
                 double* __restrict__ part, unsigned int* __restrict__ ticket, GridSpec* __restrict__ out) {
    double mn[3] = {DBL_MAX, DBL_MAX, DBL_MAX}, mx[3] = {-DBL_MAX, -DBL_MAX, -DBL_MAX};
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        for (int a = 0; a < dim; ++a) {
            double v = xy[i * dim + a];
            if (!isfinite(v)) v = -INFINITY;             // fmin/fmax drop NaN: make any non-finite value poison the box
            mn[a] = fmin(mn[a], v);
            mx[a] = fmax(mx[a], v);
        }
    }
    __shared__ double sm[6][32];
    __shared__ double box[6];
    __shared__ bool last;
    for (int a = 0; a < 3; ++a) {
        double lo = mn[a], hi = mx[a];
        for (int o = 16; o > 0; o >>= 1) {
            lo = fmin(lo, __shfl_xor_sync(0xffffffffu, lo, o));
            hi = fmax(hi, __shfl_xor_sync(0xffffffffu, hi, o));
        }
        if ((threadIdx.x & 31) == 0) {
            sm[a][threadIdx.x >> 5] = lo;
            sm[3 + a][threadIdx.x >> 5] = hi;
        }
    }
    __syncthreads();
    if (threadIdx.x < 6) {
        const int nw = blockDim.x >> 5;
        double v = sm[threadIdx.x][0];
        for (int w = 1; w < nw; ++w) v = threadIdx.x < 3 ? fmin(v, sm[threadIdx.x][w]) : fmax(v, sm[threadIdx.x][w]);
        part[blockIdx.x * 6 + threadIdx.x] = v;
        __threadfence();
    }
    __syncthreads();
    if (threadIdx.x == 0) last = atomicAdd(ticket, 1u) == gridDim.x - 1;
    __syncthreads();
    if (!last) return;
    __threadfence();
    const int t = threadIdx.x >> 5, lane = threadIdx.x & 31;      // one warp per box component
    if (t < 6) {
        const volatile double* vp = part;
        double v = t < 3 ? DBL_MAX : -DBL_MAX;
        for (int b = lane; b < (int)gridDim.x; b += 32) v = t < 3 ? fmin(v, vp[b * 6 + t]) : fmax(v, vp[b * 6 + t]);
        for (int o = 16; o > 0; o >>= 1) {
            const double u = __shfl_xor_sync(0xffffffffu, v, o);
            v = t < 3 ? fmin(v, u) : fmax(v, u);
        }
        if (lane == 0) box[t] = v;
    }
    __syncthreads();
    if (threadIdx.x == 0) derive_grid_spec(box, n, dim, per_cell, min_h, cell_bound, out);
}

// The grid shape is derived ON THE DEVICE from the bounding box (one thread), so the whole graph build is
// asynchronous: no host round trip between the bounding box and the binning.  Host-side buffers are sized by the
// bound ncell <= cell_bound that the loop below enforces.
constexpr int kStripH = 8;

__device__ void derive_grid_spec(const double* box, int64_t n, int dim, double per_cell, double min_h, int64_t cell_bound,
                                 GridSpec* out) {
    GridSpec gs;
    gs.dim = dim;
    gs.valid = 1;
    double ext[3] = {0, 0, 0}, vol = 1.0;
    int nz_dims = 0;
    for (int a = 0; a < 3; ++a) { gs.lo[a] = 0; gs.g[a] = 1; }
    for (int a = 0; a < dim; ++a) {
        if (!isfinite(box[a]) || !isfinite(box[3 + a])) gs.valid = 0;
        gs.lo[a] = box[a];
        ext[a] = box[3 + a] - box[a];
        if (ext[a] > 0) { vol *= ext[a]; ++nz_dims; }
    }
    double h = nz_dims ? pow(vol * per_cell / (double)n, 1.0 / nz_dims) : 1.0;
    if (!(h > 0) || !isfinite(h)) h = 1.0;
    if (h < min_h) h = min_h;
    int64_t ncell = 1;
    for (int it = 0; it < 4096 && gs.valid; ++it) {    // keep the grid within the buffers the host sized
        ncell = 1;
        for (int a = 0; a < dim; ++a) {
            double g = floor(ext[a] / h) + 1.0;
            if (g > 1e9) g = 1e9;
            gs.g[a] = (int)g;
            ncell *= gs.g[a];
        }
        if (ncell <= cell_bound) break;
        h *= 1.5;
    }
    if (!gs.valid) { gs.g[0] = gs.g[1] = gs.g[2] = 1; ncell = 1; }
    gs.h = h;
    gs.inv_h = 1.0 / h;
    gs.ncell = ncell;
    gs.strip_h = gs.g[1] < kStripH ? gs.g[1] : kStripH;
    const int64_t nstrip = (gs.g[1] + gs.strip_h - 1) / gs.strip_h;
    gs.nslot = (int64_t)gs.g[2] * nstrip * gs.strip_h * gs.g[0];       // < 2 * ncell
    *out = gs;
}

// --------------------------------------------------------------- binning --
__device__ __forceinline__ int cell_coord(double v, double lo, double inv_h, int g) {
    int c = (int)floor((v - lo) * inv_h);
    return min(max(c, 0), g - 1);
}

__global__ void cell_count_kernel(const double* __restrict__ xy, int64_t n, const GridSpec* __restrict__ gsp,
                                  int32_t* __restrict__ cell_of, int32_t* __restrict__ counts) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    const GridSpec gs = *gsp;
    if (!gs.valid) { cell_of[i] = 0; atomicAdd(&counts[0], 1); return; }
    int c = 0, stride = 1;
    for (int a = 0; a < gs.dim; ++a) {
        c += cell_coord(xy[i * gs.dim + a], gs.lo[a], gs.inv_h, gs.g[a]) * stride;
        stride *= gs.g[a];
    }
    cell_of[i] = c;
    atomicAdd(&counts[c], 1);
}

__global__ void cell_scatter_kernel(const int32_t* __restrict__ cell_of, int64_t n, const int32_t* __restrict__ cell_start,
                                    int32_t* __restrict__ fill, int32_t* __restrict__ sorted_id) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    int c = cell_of[i];
    sorted_id[cell_start[c] + atomicAdd(&fill[c], 1)] = (int32_t)i;
}

// order inside a cell came from atomics: sort each (tiny) cell segment by point id
// so that the emitted processing order is deterministic
__global__ void cell_sort_kernel(const int32_t* __restrict__ cell_start, const GridSpec* __restrict__ gsp,
                                 int32_t* __restrict__ sorted_id) {
    int64_t c = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (c >= gsp->ncell) return;
    const int s = cell_start[c], e = cell_start[c + 1], len = e - s;
    if (len < 2) return;
    constexpr int CAP = 16;                 // ~2.5 points per cell: the segment is sorted in a thread-private array
    if (len <= CAP) {                       // (one load and one store per point instead of a read-modify-write chain)
        int32_t v[CAP];
#pragma unroll
        for (int a = 0; a < CAP; ++a) v[a] = a < len ? sorted_id[s + a] : INT32_MAX;
#pragma unroll
        for (int a = 1; a < CAP; ++a) {
#pragma unroll
            for (int b = a; b > 0; --b) {
                const int32_t lo = min(v[b - 1], v[b]), hi = max(v[b - 1], v[b]);
                v[b - 1] = lo;
                v[b] = hi;
            }
        }
#pragma unroll
        for (int a = 0; a < CAP; ++a)
            if (a < len) sorted_id[s + a] = v[a];
        return;
    }
    // longer segments (clustered inputs, e.g. the (mean, variance) points of the background search): cell_sort_big_kernel
}

// cells with more than 16 points, sorted by a whole CTA: bitonic sort in shared memory (<= kBigCellCap points; beyond
// that one thread falls back to insertion sort in place).  A CTA scans 64 cells; almost all of them are small and skipped.
constexpr int kBigCellCap = 4096;
__global__ void __launch_bounds__(256)
cell_sort_big_kernel(const int32_t* __restrict__ cell_start, const GridSpec* __restrict__ gsp, int32_t* __restrict__ sorted_id) {
    __shared__ int32_t buf[kBigCellCap];
    const int64_t ncell = gsp->ncell;
    for (int q = 0; q < 64; ++q) {
        const int64_t c = (int64_t)blockIdx.x * 64 + q;
        if (c >= ncell) return;
        const int s = cell_start[c], e = cell_start[c + 1], len = e - s;
        if (len <= 16) continue;                                      // uniform over the CTA
        if (len > kBigCellCap) {
            if (threadIdx.x == 0)
                for (int a = s + 1; a < e; ++a) {
                    int v = sorted_id[a], b = a - 1;
                    while (b >= s && sorted_id[b] > v) { sorted_id[b + 1] = sorted_id[b]; --b; }
                    sorted_id[b + 1] = v;
                }
            __syncthreads();
            continue;
        }
        int cap = 32;
        while (cap < len) cap <<= 1;
        for (int i = threadIdx.x; i < cap; i += 256) buf[i] = i < len ? sorted_id[s + i] : INT32_MAX;
        __syncthreads();
        for (int kk = 2; kk <= cap; kk <<= 1)
            for (int j = kk >> 1; j > 0; j >>= 1) {
                for (int i = threadIdx.x; i < cap; i += 256) {
                    const int l = i ^ j;
                    if (l > i) {
                        const bool up = (i & kk) == 0;
                        const int32_t a = buf[i], b = buf[l];
                        if ((a > b) == up) { buf[i] = b; buf[l] = a; }
                    }
                }
                __syncthreads();
            }
        for (int i = threadIdx.x; i < len; i += 256) sorted_id[s + i] = buf[i];
        __syncthreads();
    }
}

__global__ void gather_sorted_kernel(const double* __restrict__ xy, int64_t n, int dim, const int32_t* __restrict__ sorted_id,
                                     const int32_t* __restrict__ cell_of, double* __restrict__ sx, double* __restrict__ sy,
                                     double* __restrict__ sz, int32_t* __restrict__ sorted_cell) {
    int64_t q = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (q >= n) return;
    int64_t i = sorted_id[q];
    sx[q] = xy[i * dim];
    sy[q] = xy[i * dim + 1];
    if (dim == 3) sz[q] = xy[i * dim + 2];
    sorted_cell[q] = cell_of[i];
}

// ------------------------------------------------------ processing order --
// The order handed to the downstream gather kernels walks the grid in horizontal strips of
// kStripH cells, column by column, boustrophedon in both directions: consecutive cells are
// spatial neighbours and the set of "recently visited" cells is a compact patch, so a kernel
// that lets one SM walk a contiguous range of the order finds most neighbour rows in its L1.
__device__ __forceinline__ int64_t strip_slot(int c, const GridSpec& gs) {
    const int gx = gs.g[0], gy = gs.g[1], H = gs.strip_h;
    const int cx = c % gx, cy = (c / gx) % gy, cz = c / (gx * gy);
    const int s = cy / H, yl = cy - s * H;
    const int xi = (s & 1) ? gx - 1 - cx : cx;
    const int yi = (xi & 1) ? H - 1 - yl : yl;
    const int64_t nstrip = (gy + H - 1) / H;
    return ((int64_t)cz * nstrip + s) * ((int64_t)H * gx) + (int64_t)xi * H + yi;
}

__global__ void strip_count_kernel(const int32_t* __restrict__ cell_start, const GridSpec* __restrict__ gsp,
                                   int32_t* __restrict__ slot_count) {
    int64_t c = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    const GridSpec gs = *gsp;
    if (c >= gs.ncell) return;
    slot_count[strip_slot((int)c, gs)] = cell_start[c + 1] - cell_start[c];
}

__global__ void strip_order_kernel(const int32_t* __restrict__ cell_start, const int32_t* __restrict__ sorted_id,
                                   const int32_t* __restrict__ sorted_cell, int64_t n, const GridSpec* __restrict__ gsp,
                                   const int32_t* __restrict__ slot_start, int32_t* __restrict__ order) {
    int64_t q = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (q >= n) return;
    const GridSpec gs = *gsp;
    const int c = sorted_cell[q];
    order[slot_start[strip_slot(c, gs)] + (int)(q - cell_start[c])] = sorted_id[q];
}

// ---------------------------------------------------------------- search --
// Sorted candidate list of one query.  SHARED: the lists of a CTA live in shared memory, entry-major
// ([entry][thread]): a thread's entries all sit in "its" bank(s), so the data-dependent positions of the insertion
// loop cost one wavefront per access whatever the lanes' positions are - in local memory the same loop touches up to 32
// different lines per instruction, which is what bounded the search.  Otherwise (k > 32) the list is thread-local.
constexpr int kKnnThreads = 128;

template <int KMAX, bool SHARED>
struct TopK {
    double d2_local[SHARED ? 1 : KMAX];
    int32_t id_local[SHARED ? 1 : KMAX];
    double* d2s;
    int32_t* ids;
    int count;
    int K;
    __device__ __forceinline__ void init(int k, double* sm_d2, int32_t* sm_id) {
        count = 0;
        K = k;
        d2s = sm_d2 + threadIdx.x;
        ids = sm_id + threadIdx.x;
    }
    __device__ __forceinline__ double& d2(int j) { return SHARED ? d2s[j * kKnnThreads] : d2_local[j]; }
    __device__ __forceinline__ int32_t& id(int j) { return SHARED ? ids[j * kKnnThreads] : id_local[j]; }
    __device__ __forceinline__ bool full() const { return count == K; }
    __device__ __forceinline__ double worst() { return d2(count - 1); }
    __device__ __forceinline__ void offer(double v, int32_t pid) {
        if (count == K) {
            double wv = d2(K - 1);
            if (v > wv || (v == wv && pid > id(K - 1))) return;
        }
        int pos = count < K ? count : K - 1;
        while (pos > 0) {
            const double dp = d2(pos - 1);
            const int32_t ip = id(pos - 1);
            if (!(dp > v || (dp == v && ip > pid))) break;
            d2(pos) = dp;
            id(pos) = ip;
            --pos;
        }
        d2(pos) = v;
        id(pos) = pid;
        if (count < K) ++count;
    }
};

template <int KMAX, int DIM>
__global__ void __launch_bounds__(kKnnThreads)
knn_search_kernel(int64_t n, const GridSpec* __restrict__ gsp, int kq, int k, int include_self, const int32_t* __restrict__ cell_start,
                  const int32_t* __restrict__ sorted_id, const int32_t* __restrict__ sorted_cell,
                  const double* __restrict__ sx, const double* __restrict__ sy, const double* __restrict__ sz,
                  int32_t* __restrict__ out_idx, double* __restrict__ out_dist) {
    constexpr bool SHARED = KMAX <= 32;                // 12 * KMAX * 128 bytes of shared memory: 24 KB / 48 KB
    __shared__ double sm_d2[SHARED ? KMAX * kKnnThreads : 1];
    __shared__ int32_t sm_id[SHARED ? KMAX * kKnnThreads : 1];
    int64_t q = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (q >= n) return;
    const GridSpec gs = *gsp;
    const int32_t self = sorted_id[q];
    if (!gs.valid) {                                   // non-finite coordinates: the caller sees index -1 / NaN
        for (int j = 0; j < k; ++j) { out_idx[(int64_t)self * k + j] = -1; out_dist[(int64_t)self * k + j] = nan(""); }
        return;
    }
    const double qx = sx[q], qy = sy[q], qz = DIM == 3 ? sz[q] : 0.0;
    int c = sorted_cell[q];
    const int cx = c % gs.g[0];
    const int cy = (c / gs.g[0]) % gs.g[1];
    const int cz = DIM == 3 ? c / (gs.g[0] * gs.g[1]) : 0;
    const int gx = gs.g[0], gy = gs.g[1], gz = DIM == 3 ? gs.g[2] : 1;

    TopK<KMAX, SHARED> top;
    top.init(kq, sm_d2, sm_id);
    const int rmax = max(max(max(cx, gx - 1 - cx), max(cy, gy - 1 - cy)), max(cz, gz - 1 - cz));
    for (int r = 0; r <= rmax; ++r) {
        const int z0 = DIM == 3 ? max(cz - r, 0) : 0, z1 = DIM == 3 ? min(cz + r, gz - 1) : 0;
        for (int z = z0; z <= z1; ++z) {
            const bool zface = DIM == 3 && (z == cz - r || z == cz + r);
            const int y0 = max(cy - r, 0), y1 = min(cy + r, gy - 1);
            for (int y = y0; y <= y1; ++y) {
                const bool full_row = zface || y == cy - r || y == cy + r;
                const int nseg = (full_row || r == 0) ? 1 : 2;
                for (int seg = 0; seg < nseg; ++seg) {
                    int xa, xb;
                    if (full_row || r == 0) { xa = max(cx - r, 0); xb = min(cx + r, gx - 1); }
                    else if (seg == 0) { xa = xb = cx - r; if (xa < 0) continue; }
                    else { xa = xb = cx + r; if (xa > gx - 1) continue; }
                    const int64_t base = ((int64_t)z * gy + y) * gx;
                    const int ps = cell_start[base + xa], pe = cell_start[base + xb + 1];
                    for (int p = ps; p < pe; ++p) {
                        const double dx = qx - sx[p], dy = qy - sy[p];
                        double d2 = __dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy));
                        if (DIM == 3) { const double dz = qz - sz[p]; d2 = __dadd_rn(d2, __dmul_rn(dz, dz)); }
                        top.offer(d2, sorted_id[p]);
                    }
                }
            }
        }
        if (top.full()) {
            // smallest possible distance to any point outside the visited block
            double bound = DBL_MAX;
            if (cx - r > 0) bound = fmin(bound, qx - (gs.lo[0] + (cx - r) * gs.h));
            if (cx + r < gx - 1) bound = fmin(bound, (gs.lo[0] + (cx + r + 1) * gs.h) - qx);
            if (cy - r > 0) bound = fmin(bound, qy - (gs.lo[1] + (cy - r) * gs.h));
            if (cy + r < gy - 1) bound = fmin(bound, (gs.lo[1] + (cy + r + 1) * gs.h) - qy);
            if (DIM == 3) {
                if (cz - r > 0) bound = fmin(bound, qz - (gs.lo[2] + (cz - r) * gs.h));
                if (cz + r < gz - 1) bound = fmin(bound, (gs.lo[2] + (cz + r + 1) * gs.h) - qz);
            }
            if (bound == DBL_MAX) break;
            bound -= gs.h * 1e-9;                      // slack for rounding in the cell assignment
            if (bound > 0.0 && top.worst() < bound * bound) break;
        }
    }

    int32_t* oi = out_idx + (int64_t)self * k;
    double* od = out_dist + (int64_t)self * k;
    if (include_self) {
        for (int j = 0; j < k; ++j) { oi[j] = top.id(j); od[j] = __dsqrt_rn(top.d2(j)); }
    } else {
        int drop = 0;                                  // sklearn: drop own index, else the first entry
        for (int j = 0; j < kq; ++j) if (top.id(j) == self) { drop = j; break; }
        int o = 0;
        for (int j = 0; j < kq; ++j) {
            if (j == drop) continue;
            oi[o] = top.id(j);
            od[o] = __dsqrt_rn(top.d2(j));
            ++o;
        }
    }
}


// ---------------------------------------------------------- radius graph --
// K1r (extension; oracle: sklearn.neighbors.radius_neighbors_graph(mode='distance')): every point within
// `radius` of the query, d^2 <= radius^2 with d^2 accumulated exactly as in the kNN search (sklearn compares
// reduced distances, sklearn/neighbors/_binary_tree.pxi.tp query_radius).  The grid's cell edge is >= radius,
// so the 3^dim block around the query's cell holds every candidate.  MODE 0 counts, MODE 1 writes the row
// (column-sorted: a canonical CSR) at its indptr offset.
template <int DIM, int MODE>
__global__ void __launch_bounds__(128)
radius_search_kernel(int64_t n, const GridSpec* __restrict__ gsp, double r2, int include_self, const int32_t* __restrict__ cell_start,
                     const int32_t* __restrict__ sorted_id, const int32_t* __restrict__ sorted_cell,
                     const double* __restrict__ sx, const double* __restrict__ sy, const double* __restrict__ sz,
                     int64_t* __restrict__ counts, const int64_t* __restrict__ indptr, int32_t* __restrict__ indices,
                     double* __restrict__ dist) {
    const int64_t q = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (q >= n) return;
    const GridSpec gs = *gsp;
    const int32_t self = sorted_id[q];
    if (!gs.valid) {                                   // non-finite coordinates: empty rows
        if (!MODE) counts[self] = 0;
        return;
    }
    const double qx = sx[q], qy = sy[q], qz = DIM == 3 ? sz[q] : 0.0;
    const int c = sorted_cell[q];
    const int gx = gs.g[0], gy = gs.g[1], gz = DIM == 3 ? gs.g[2] : 1;
    const int cx = c % gx, cy = (c / gx) % gy, cz = DIM == 3 ? c / (gx * gy) : 0;
    int64_t cnt = 0;
    const int64_t base_out = MODE ? indptr[self] : 0;
    for (int z = max(cz - 1, 0); z <= min(cz + 1, gz - 1); ++z)
        for (int y = max(cy - 1, 0); y <= min(cy + 1, gy - 1); ++y) {
            const int64_t base = ((int64_t)z * gy + y) * gx;
            const int ps = cell_start[base + max(cx - 1, 0)], pe = cell_start[base + min(cx + 1, gx - 1) + 1];
            for (int p = ps; p < pe; ++p) {
                const double dx = qx - sx[p], dy = qy - sy[p];
                double d2 = __dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy));
                if (DIM == 3) { const double dz = qz - sz[p]; d2 = __dadd_rn(d2, __dmul_rn(dz, dz)); }
                const int32_t id = sorted_id[p];
                if (d2 <= r2 && (include_self || id != self)) {
                    if (MODE) {                                  // insertion by column id into the row written so far
                        int64_t pos = base_out + cnt;
                        while (pos > base_out && indices[pos - 1] > id) {
                            indices[pos] = indices[pos - 1];
                            dist[pos] = dist[pos - 1];
                            --pos;
                        }
                        indices[pos] = id;
                        dist[pos] = __dsqrt_rn(d2);
                    }
                    ++cnt;
                }
            }
        }
    if (!MODE) counts[self] = cnt;
}

// variable-degree CSR rows -> fixed-degree (ELL) rows for the gather kernels: padding edges point at the row itself
// with weight 0, which adds nothing to any weighted sum
__global__ void csr_pad_rows_kernel(const int64_t* __restrict__ indptr, const int32_t* __restrict__ indices,
                                    const float* __restrict__ w, int64_t n, int kmax, int32_t* __restrict__ idx,
                                    float* __restrict__ wout) {
    const int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (e >= n * kmax) return;
    const int64_t i = e / kmax;
    const int j = (int)(e - i * kmax);
    const int64_t s = indptr[i], deg = indptr[i + 1] - s;
    idx[e] = j < deg ? indices[s + j] : (int32_t)i;
    wout[e] = j < deg ? w[s + j] : 0.f;
}

// --------------------------------------------------------------- weights --
constexpr int kSumBlocks = 592, kSumThreads = 256;

__global__ void sum_partial_kernel(const double* __restrict__ v, int64_t m, double* __restrict__ part) {
    double s = 0.0;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < m; i += (int64_t)gridDim.x * blockDim.x) s += v[i];
    __shared__ double sm[32];
    s = warp_sum(s);
    if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int w = 0; w < (blockDim.x >> 5); ++w) t += sm[w];
        part[blockIdx.x] = t;
    }
}

// Every CTA derives the scale itself from the block partials (warp 0: fixed lane-strided order + butterfly, so the
// sum is deterministic and identical in all CTAs) - no separate one-warp kernel between the sum and the weights.
__global__ void __launch_bounds__(256)
decay_weights_kernel(const double* __restrict__ dist, int64_t m, const double* __restrict__ part, int nb, double sigma,
                     double* __restrict__ w64, float* __restrict__ w32, double* __restrict__ scale_out) {
    __shared__ double s_scale;
    if (threadIdx.x < 32) {
        double sc = sigma;
        if (!(sigma > 0.0)) {
            double t = 0.0;
            for (int b = threadIdx.x; b < nb; b += 32) t += part[b];
            t = warp_sum(t);
            sc = (t / (double)m) / 2.0;
        }
        if (threadIdx.x == 0) {
            s_scale = sc;
            if (blockIdx.x == 0 && scale_out) *scale_out = sc;
        }
    }
    __syncthreads();
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= m) return;
    double w = 1.0 / exp(dist[i] / s_scale);
    if (w64) w64[i] = w;
    if (w32) w32[i] = (float)w;
}

}  // namespace laris

using namespace laris;

// ------------------------------------------------------------ point grid --
// Everything the searches need: the grid shape and the cell-sorted copies of the points.
struct PointGrid {
    int64_t cell_bound = 0;             // ncell <= cell_bound (enforced by derive_grid_spec), nslot < 2 * cell_bound
    Scratch<unsigned char> arena;       // ONE stream-ordered allocation for everything below (a dozen separate
                                        // allocations cost more host time than the small kernels they feed)
    GridSpec* gs = nullptr;
    int32_t *cell_of = nullptr, *counts = nullptr, *fill = nullptr, *slot_count = nullptr, *cell_start = nullptr,
            *slot_start = nullptr, *sorted_id = nullptr, *sorted_cell = nullptr;
    double *sx = nullptr, *sy = nullptr, *sz = nullptr, *part = nullptr;
    explicit PointGrid(cudaStream_t st) : arena(st) {}
};

// Bins the points on a uniform grid of about `per_cell` points per cell whose cell edge is at least `min_h`.
// Fully asynchronous: the grid shape stays on the device (GridSpec), buffers are sized by cell_bound = 2n + 1024.
static int build_point_grid(const double* xy, int64_t n, int dim, double per_cell, double min_h, bool want_order,
                            cudaStream_t st, PointGrid& G) {
    const int nb = kNumSMs * 2;
    const int64_t bound = 2 * n + 1024, slots = want_order ? 2 * bound : 0;
    G.cell_bound = bound;
    size_t off = 0;
    auto reserve = [&](size_t bytes) { const size_t o = off; off += (bytes + 255) & ~(size_t)255; return o; };
    const size_t o_gs = reserve(sizeof(GridSpec)), o_part = reserve((size_t)nb * 6 * 8);
    // ticket | counts | fill | slot_count are zero-initialised together
    const size_t o_ticket = reserve(4);
    const size_t o_counts = reserve((size_t)bound * 4), o_fill = reserve((size_t)bound * 4), o_slotc = reserve((size_t)slots * 4);
    const size_t zero_bytes = off - o_ticket;
    const size_t o_cstart = reserve((size_t)(bound + 1) * 4), o_sstart = reserve((size_t)(slots + 1) * 4);
    const size_t o_cellof = reserve((size_t)n * 4), o_sid = reserve((size_t)n * 4), o_scell = reserve((size_t)n * 4);
    const size_t o_sx = reserve((size_t)n * 8), o_sy = reserve((size_t)n * 8), o_sz = reserve((size_t)(dim == 3 ? n : 1) * 8);
    LARIS_CUDA(G.arena.alloc(off));
    unsigned char* base = G.arena.p;
    G.gs = (GridSpec*)(base + o_gs); G.part = (double*)(base + o_part);
    G.counts = (int32_t*)(base + o_counts); G.fill = (int32_t*)(base + o_fill); G.slot_count = (int32_t*)(base + o_slotc);
    G.cell_start = (int32_t*)(base + o_cstart); G.slot_start = (int32_t*)(base + o_sstart);
    G.cell_of = (int32_t*)(base + o_cellof); G.sorted_id = (int32_t*)(base + o_sid); G.sorted_cell = (int32_t*)(base + o_scell);
    G.sx = (double*)(base + o_sx); G.sy = (double*)(base + o_sy); G.sz = (double*)(base + o_sz);
    LARIS_CUDA(cudaMemsetAsync(base + o_ticket, 0, zero_bytes, st));
    bbox_grid_kernel<<<nb, 256, 0, st>>>(xy, n, dim, per_cell, min_h, bound, G.part, (unsigned int*)(base + o_ticket), G.gs);
    LARIS_LAUNCH_CHECK();
    const int T = 256;
    const unsigned gn = (unsigned)ceil_div(n, T), gc = (unsigned)ceil_div(bound, T);
    cell_count_kernel<<<gn, T, 0, st>>>(xy, n, G.gs, G.cell_of, G.counts);
    LARIS_LAUNCH_CHECK();
    int rc = exclusive_scan<int32_t>(G.counts, bound, G.cell_start, st);         // cells past ncell are empty
    if (rc) return rc;
    cell_scatter_kernel<<<gn, T, 0, st>>>(G.cell_of, n, G.cell_start, G.fill, G.sorted_id);
    LARIS_LAUNCH_CHECK();
    cell_sort_kernel<<<gc, T, 0, st>>>(G.cell_start, G.gs, G.sorted_id);
    LARIS_LAUNCH_CHECK();
    cell_sort_big_kernel<<<(unsigned)ceil_div(G.cell_bound, 64), 256, 0, st>>>(G.cell_start, G.gs, G.sorted_id);
    LARIS_LAUNCH_CHECK();
    gather_sorted_kernel<<<gn, T, 0, st>>>(xy, n, dim, G.sorted_id, G.cell_of, G.sx, G.sy, G.sz, G.sorted_cell);
    LARIS_LAUNCH_CHECK();
    return LARIS_OK;
}

// cell segments re-dealt in strip order -> the processing order handed to the gather kernels
static int emit_strip_order(const PointGrid& G, int64_t n, int32_t* out_order, cudaStream_t st) {
    const int T = 256;
    const unsigned gn = (unsigned)ceil_div(n, T), gc = (unsigned)ceil_div(G.cell_bound, T);
    const int64_t slot_bound = 2 * G.cell_bound;
    strip_count_kernel<<<gc, T, 0, st>>>(G.cell_start, G.gs, G.slot_count);
    LARIS_LAUNCH_CHECK();
    int rc = exclusive_scan<int32_t>(G.slot_count, slot_bound, G.slot_start, st);
    if (rc) return rc;
    strip_order_kernel<<<gn, T, 0, st>>>(G.cell_start, G.sorted_id, G.sorted_cell, n, G.gs, G.slot_start, out_order);
    LARIS_LAUNCH_CHECK();
    return LARIS_OK;
}

extern "C" int laris_knn_graph(const double* xy, int64_t n, int dim, int k, int include_self, int32_t* out_idx,
                               double* out_dist, int32_t* out_order, void* stream_) {
    cudaStream_t st = (cudaStream_t)stream_;
    LARIS_CHECK_ARG(xy && out_idx && out_dist, "laris_knn_graph: null pointer");
    LARIS_CHECK_ARG(dim == 2 || dim == 3, "laris_knn_graph: dim must be 2 or 3 (got %d)", dim);
    const int kq = include_self ? k : k + 1;
    LARIS_CHECK_ARG(k >= 1 && kq <= 64, "laris_knn_graph: k out of range (k=%d, limit 64 incl. self)", k);
    LARIS_CHECK_ARG(n >= kq && n < (int64_t)INT32_MAX, "laris_knn_graph: need k+1 <= n < 2^31 (n=%lld)", (long long)n);

    PointGrid G(st);
    int rc = build_point_grid(xy, n, dim, 2.5, 0.0, out_order != nullptr, st, G);
    if (rc) return rc;

    const unsigned gq = (unsigned)ceil_div(n, kKnnThreads);
#define LARIS_KNN_LAUNCH(KMAX, DIM)                                                                          \
    knn_search_kernel<KMAX, DIM><<<gq, kKnnThreads, 0, st>>>(n, G.gs, kq, k, include_self, G.cell_start, G.sorted_id, \
                                                     G.sorted_cell, G.sx, G.sy, G.sz, out_idx, out_dist)
    if (dim == 2) {
        if (kq <= 16) LARIS_KNN_LAUNCH(16, 2); else if (kq <= 32) LARIS_KNN_LAUNCH(32, 2); else LARIS_KNN_LAUNCH(64, 2);
    } else {
        if (kq <= 16) LARIS_KNN_LAUNCH(16, 3); else if (kq <= 32) LARIS_KNN_LAUNCH(32, 3); else LARIS_KNN_LAUNCH(64, 3);
    }
#undef LARIS_KNN_LAUNCH
    LARIS_LAUNCH_CHECK();
    if (out_order) {
        rc = emit_strip_order(G, n, out_order, st);
        if (rc) return rc;
    }
    return LARIS_OK;
}


static int radius_graph_pass(const double* xy, int64_t n, int dim, double radius, int include_self, int mode,
                             int64_t* counts, const int64_t* indptr, int32_t* indices, double* dist, int32_t* out_order,
                             cudaStream_t st, const char* who) {
    LARIS_CHECK_ARG(xy && n >= 1 && n < (int64_t)INT32_MAX, "%s: bad arguments (n=%lld)", who, (long long)n);
    LARIS_CHECK_ARG(dim == 2 || dim == 3, "%s: dim must be 2 or 3 (got %d)", who, dim);
    LARIS_CHECK_ARG(radius > 0.0 && isfinite(radius), "%s: radius must be positive and finite", who);
    PointGrid G(st);
    // cell edge >= radius (plus rounding slack in the cell assignment) so that one ring of cells is enough
    int rc = build_point_grid(xy, n, dim, 2.5, radius * (1.0 + 1e-9), out_order != nullptr, st, G);
    if (rc) return rc;
    const unsigned gq = (unsigned)ceil_div(n, 128);
    const double r2 = radius * radius;
#define LARIS_RADIUS_LAUNCH(DIM, MODE)                                                                              \
    radius_search_kernel<DIM, MODE><<<gq, 128, 0, st>>>(n, G.gs, r2, include_self, G.cell_start, G.sorted_id,         \
                                                        G.sorted_cell, G.sx, G.sy, G.sz, counts, indptr,            \
                                                        indices, dist)
    if (dim == 2) { if (mode) LARIS_RADIUS_LAUNCH(2, 1); else LARIS_RADIUS_LAUNCH(2, 0); }
    else { if (mode) LARIS_RADIUS_LAUNCH(3, 1); else LARIS_RADIUS_LAUNCH(3, 0); }
#undef LARIS_RADIUS_LAUNCH
    LARIS_LAUNCH_CHECK();
    if (out_order) {
        rc = emit_strip_order(G, n, out_order, st);
        if (rc) return rc;
    }
    return LARIS_OK;
}

extern "C" int laris_radius_graph_count(const double* xy, int64_t n, int dim, double radius, int include_self,
                                        int64_t* counts, int32_t* out_order, void* stream) {
    LARIS_CHECK_ARG(counts, "laris_radius_graph_count: null pointer");
    return radius_graph_pass(xy, n, dim, radius, include_self, 0, counts, nullptr, nullptr, nullptr, out_order,
                             (cudaStream_t)stream, "laris_radius_graph_count");
}

extern "C" int laris_radius_graph_fill(const double* xy, int64_t n, int dim, double radius, int include_self,
                                       const int64_t* indptr, int32_t* indices, double* dist, void* stream) {
    LARIS_CHECK_ARG(indptr && indices && dist, "laris_radius_graph_fill: null pointer");
    return radius_graph_pass(xy, n, dim, radius, include_self, 1, nullptr, indptr, indices, dist, nullptr,
                             (cudaStream_t)stream, "laris_radius_graph_fill");
}

extern "C" int laris_csr_pad_rows(const int64_t* indptr, const int32_t* indices, const float* w, int64_t n, int kmax,
                                  int32_t* idx, float* w_out, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    LARIS_CHECK_ARG(indptr && indices && w && idx && w_out && n >= 1 && kmax >= 1, "laris_csr_pad_rows: bad arguments");
    csr_pad_rows_kernel<<<(unsigned)ceil_div(n * kmax, 256), 256, 0, st>>>(indptr, indices, w, n, kmax, idx, w_out);
    LARIS_LAUNCH_CHECK();
    return LARIS_OK;
}

extern "C" int laris_decay_weights(const double* dist, int64_t m, double sigma, double* w64, float* w32,
                                   double* scale_out, void* stream_) {
    cudaStream_t st = (cudaStream_t)stream_;
    LARIS_CHECK_ARG(dist && m > 0, "laris_decay_weights: empty input");
    Scratch<double> part(st);
    LARIS_CUDA(part.alloc(kSumBlocks));
    if (!(sigma > 0.0)) {
        sum_partial_kernel<<<kSumBlocks, kSumThreads, 0, st>>>(dist, m, part.p);
        LARIS_LAUNCH_CHECK();
    }
    decay_weights_kernel<<<(unsigned)ceil_div(m, 256), 256, 0, st>>>(dist, m, part.p, kSumBlocks, sigma, w64, w32, scale_out);
    LARIS_LAUNCH_CHECK();
    return LARIS_OK;
}
