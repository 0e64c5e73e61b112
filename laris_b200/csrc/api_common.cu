// Error plumbing + small utility entry points of the C ABI.
#include <stdarg.h>

#include "common.cuh"
#include "scan.cuh"

#include <atomic>

namespace laris {
static std::atomic<long long> g_launches{0};
void note_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }
static std::atomic<unsigned long long> g_pool_ready{0};          // bit d = device d configured
void retain_pool_memory() {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return;
    const unsigned long long bit = 1ull << dev;
    if (g_pool_ready.load(std::memory_order_relaxed) & bit) return;
    cudaMemPool_t pool;
    if (cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess) {
        uint64_t keep = UINT64_MAX;
        cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
    }
    g_pool_ready.fetch_or(bit, std::memory_order_relaxed);
}
static thread_local char g_err[512] = "";
void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

__global__ void gather_f32_kernel(const float* __restrict__ w, const int64_t* __restrict__ perm, int64_t m, float* __restrict__ out) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i < m) out[i] = w[perm[i]];
}
}  // namespace laris

using namespace laris;

extern "C" int laris_abi_version(void) { return LARIS_ABI_VERSION; }
extern "C" int64_t laris_launch_count(void) { return (int64_t)g_launches.load(); }
extern "C" const char* laris_last_error(void) { return g_err; }

extern "C" int laris_exclusive_scan_i64(const int64_t* in, int64_t n, int64_t* out, void* stream) {
    LARIS_CHECK_ARG(in && out && n >= 0, "laris_exclusive_scan_i64: bad arguments");
    return exclusive_scan<int64_t>(in, n, out, (cudaStream_t)stream);
}

extern "C" int laris_gather_f32(const float* w, const int64_t* perm, int64_t m, float* wout, void* stream) {
    LARIS_CHECK_ARG(w && perm && wout && m >= 0, "laris_gather_f32: bad arguments");
    if (m == 0) return LARIS_OK;
    gather_f32_kernel<<<(unsigned)ceil_div(m, 256), 256, 0, (cudaStream_t)stream>>>(w, perm, m, wout);
    LARIS_LAUNCH_CHECK();
    return LARIS_OK;
}

// ---- generic helpers behind laris.pp (reference laris/preprocessing/__init__.py:28-38) ----
namespace laris {
__global__ void rowwise_cosine_kernel(const float* __restrict__ A, const float* __restrict__ B, int64_t rows, int64_t cols,
                                      int64_t lda, int64_t ldb, double* __restrict__ out) {
    const int lane = threadIdx.x & 31;
    const int64_t r = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    if (r >= rows) return;
    double dot = 0.0, aa = 0.0, bb = 0.0;
    for (int64_t c = lane; c < cols; c += 32) {
        const double a = A[r * lda + c], b = B[r * ldb + c];
        dot += a * b; aa += a * a; bb += b * b;
    }
    dot = warp_sum(dot); aa = warp_sum(aa); bb = warp_sum(bb);
    if (lane == 0) {
        const double den = sqrt(aa) * sqrt(bb);
        out[r] = den != 0.0 ? dot / den : 0.0;
    }
}

__global__ void pairwise_row_multiply_kernel(const float* __restrict__ M, int C, int64_t n, float* __restrict__ out) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    const int ab = blockIdx.y;
    if (i >= n) return;
    const int a = ab / C, b = ab - a * C;
    out[(int64_t)ab * n + i] = M[(int64_t)a * n + i] * M[(int64_t)b * n + i];
}
}  // namespace laris

// ---- per-cell QC metrics of the score matrix (scanpy.pp.calculate_qc_metrics at laris/tools/__init__.py:493-496) ----
namespace laris {
struct TopList { int n[4]; int count; };

// order-preserving map float -> uint32 (larger value <=> larger key)
__device__ __forceinline__ uint32_t order_key(float v) {
    const uint32_t b = __float_as_uint(v);
    return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}

// warp per row: the row is staged in shared memory; total / non-zero count in one pass, then for every N of the list
// the N-th largest value by a 32-step radix select on the ordered keys, and the sum of the N largest from it
__global__ void __launch_bounds__(256)
row_qc_kernel(const float* __restrict__ S, int64_t ldS, int64_t n, int P, TopList tops, double* __restrict__ total,
              int64_t* __restrict__ nnz, double* __restrict__ top_sum) {
    extern __shared__ float qc_rows[];
    constexpr unsigned FULL = 0xffffffffu;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float* row = qc_rows + (size_t)warp * P;
    const int64_t nwarps = (int64_t)gridDim.x * (blockDim.x >> 5);
    for (int64_t i = (int64_t)blockIdx.x * (blockDim.x >> 5) + warp; i < n; i += nwarps) {
        double s = 0.0;
        int c = 0;
        for (int p = lane; p < P; p += 32) {
            const float v = S[i * ldS + p];
            row[p] = v;
            s += (double)v;
            c += v != 0.f;
        }
        s = warp_sum(s);
        c = __reduce_add_sync(FULL, c);
        __syncwarp();
        if (lane == 0) { total[i] = s; nnz[i] = c; }
        for (int t = 0; t < tops.count; ++t) {
            const int N = tops.n[t];
            double ts = s;
            if (N < P) {
                uint32_t key = 0u;                                  // largest key with #{key(v) >= key} >= N
                for (int bit = 31; bit >= 0; --bit) {
                    const uint32_t cand = key | (1u << bit);
                    int cnt = 0;
                    for (int p = lane; p < P; p += 32) cnt += order_key(row[p]) >= cand;
                    if (__reduce_add_sync(FULL, cnt) >= N) key = cand;
                }
                double sg = 0.0;
                int cg = 0;
                float vN = 0.f;
                for (int p = lane; p < P; p += 32) {
                    const float v = row[p];
                    const uint32_t kv = order_key(v);
                    if (kv > key) { sg += (double)v; ++cg; }
                    if (kv == key) vN = v;
                }
                sg = warp_sum(sg);
                cg = __reduce_add_sync(FULL, cg);
                const unsigned has = __ballot_sync(FULL, order_key(vN) == key);
                vN = __shfl_sync(FULL, vN, has ? __ffs(has) - 1 : 0);
                ts = sg + (double)(N - cg) * (double)vN;
            }
            if (lane == 0) top_sum[i * tops.count + t] = ts;
        }
        __syncwarp();
    }
}
}  // namespace laris

extern "C" int laris_row_qc(const float* S, int64_t ldS, int64_t n, int32_t P, const int32_t* top_n, int32_t n_top,
                            double* total, int64_t* nnz, double* top_sum, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    LARIS_CHECK_ARG(S && total && nnz && n >= 0 && P >= 1 && ldS >= P, "laris_row_qc: bad arguments");
    LARIS_CHECK_ARG(n_top >= 0 && n_top <= 4 && (n_top == 0 || (top_n && top_sum)), "laris_row_qc: at most 4 top-N sizes");
    LARIS_CHECK_ARG((size_t)P * 8 * sizeof(float) <= (size_t)200 * 1024, "laris_row_qc: P=%d does not fit shared memory", P);
    if (n == 0) return LARIS_OK;
    TopList tl;
    tl.count = n_top;
    for (int t = 0; t < 4; ++t) tl.n[t] = t < n_top ? top_n[t] : 0;          // top_n is HOST memory
    for (int t = 0; t < n_top; ++t) LARIS_CHECK_ARG(tl.n[t] >= 1, "laris_row_qc: top sizes must be >= 1");
    const size_t smem = (size_t)P * 8 * sizeof(float);
    LARIS_CUDA(cudaFuncSetAttribute(row_qc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int64_t want = ceil_div(n, 8);
    const int64_t grid = want < (int64_t)kNumSMs * 4 ? want : (int64_t)kNumSMs * 4;
    row_qc_kernel<<<(unsigned)grid, 256, smem, st>>>(S, ldS, n, P, tl, total, nnz, top_sum);
    LARIS_LAUNCH_CHECK();
    return LARIS_OK;
}

extern "C" int laris_rowwise_cosine(const float* A, const float* B, int64_t rows, int64_t cols, int64_t lda, int64_t ldb,
                                    double* out, void* stream) {
    LARIS_CHECK_ARG(A && B && out && rows >= 0 && cols >= 0 && lda >= cols && ldb >= cols, "laris_rowwise_cosine: bad arguments");
    if (rows == 0) return LARIS_OK;
    rowwise_cosine_kernel<<<(unsigned)ceil_div(rows * 32, 256), 256, 0, (cudaStream_t)stream>>>(A, B, rows, cols, lda, ldb, out);
    LARIS_LAUNCH_CHECK();
    return LARIS_OK;
}

extern "C" int laris_pairwise_row_multiply(const float* M, int32_t C, int64_t n, float* out, void* stream) {
    LARIS_CHECK_ARG(M && out && C >= 1 && C * C <= 65535 && n >= 0, "laris_pairwise_row_multiply: bad arguments");
    if (n == 0) return LARIS_OK;
    dim3 grid((unsigned)ceil_div(n, 256), (unsigned)(C * C));
    pairwise_row_multiply_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(M, C, n, out);
    LARIS_LAUNCH_CHECK();
    return LARIS_OK;
}
