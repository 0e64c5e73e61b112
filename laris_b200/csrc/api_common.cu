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

// warp per row.  The row is read once (128-bit loads, two in flight per lane); its NON-ZERO values are compacted into
// shared memory on the way (ballot + prefix popcount - the order inside the list is irrelevant), together with the total,
// the non-zero count and the sum / count of the positive values.  The sum of the N largest values then needs no search at
// all when the positives number at most N and the zeros fill the rest (the common case for mostly-zero score rows);
// otherwise the N-th largest value comes from a 32-step radix select on the ordered keys of the LIST - the P - m zeros
// join every count at the key of 0.0f - i.e. ~m/32 instead of P/32 shared loads per step.
template <bool VEC>
__global__ void __launch_bounds__(256)
row_qc_kernel(const float* __restrict__ S, int64_t ldS, int64_t n, int P, TopList tops, double* __restrict__ total,
              int64_t* __restrict__ nnz, double* __restrict__ top_sum) {
    extern __shared__ float qc_rows[];
    constexpr unsigned FULL = 0xffffffffu;
    constexpr uint32_t KEY0 = 0x80000000u;                          // order_key(0.0f)
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const unsigned lt = (1u << lane) - 1u;
    float* list = qc_rows + (size_t)warp * P;
    const int64_t nwarps = (int64_t)gridDim.x * (blockDim.x >> 5);
    for (int64_t i = (int64_t)blockIdx.x * (blockDim.x >> 5) + warp; i < n; i += nwarps) {
        const float* row = S + i * ldS;
        double s = 0.0, sp = 0.0;
        int m = 0, cp = 0;
        auto take = [&](float v) {                                  // warp-uniform call
            const bool nz = v != 0.f;
            const unsigned mask = __ballot_sync(FULL, nz);
            if (nz) {
                list[m + __popc(mask & lt)] = v;
                s += (double)v;
                if (v > 0.f) { sp += (double)v; ++cp; }
            }
            m += __popc(mask);
        };
        if (VEC) {
            const int nq = (P + 3) >> 2;                            // ldS % 4 == 0 and ldS >= P: quads stay inside the row
            const float4* rq = reinterpret_cast<const float4*>(row);
            for (int q0 = 0; q0 < nq; q0 += 64) {
                const int qa = q0 + lane, qb = q0 + 32 + lane;
                float4 a = make_float4(0.f, 0.f, 0.f, 0.f), b = a;
                if (qa < nq) a = ld_stream_f4(rq + qa);
                if (qb < nq) b = ld_stream_f4(rq + qb);
                const int pa = 4 * qa, pb = 4 * qb;                 // columns >= P do not belong to the table
                take(pa + 0 < P ? a.x : 0.f); take(pa + 1 < P ? a.y : 0.f);
                take(pa + 2 < P ? a.z : 0.f); take(pa + 3 < P ? a.w : 0.f);
                take(pb + 0 < P ? b.x : 0.f); take(pb + 1 < P ? b.y : 0.f);
                take(pb + 2 < P ? b.z : 0.f); take(pb + 3 < P ? b.w : 0.f);
            }
        } else {
            for (int p0 = 0; p0 < P; p0 += 32) take(p0 + lane < P ? row[p0 + lane] : 0.f);
        }
        s = warp_sum(s);
        sp = warp_sum(sp);
        cp = __reduce_add_sync(FULL, cp);
        __syncwarp();
        if (lane == 0) { total[i] = s; nnz[i] = m; }
        const int z = P - m;                                        // zeros of the row
        for (int t = 0; t < tops.count; ++t) {
            const int N = tops.n[t];
            double ts;
            if (N >= P) {
                ts = s;
            } else if (cp <= N && N <= cp + z) {
                ts = sp;                                            // every positive value, then zeros
            } else {
                uint32_t key = 0u;                                  // largest key with #{key(v) >= key} >= N
                for (int bit = 31; bit >= 0; --bit) {
                    const uint32_t cand = key | (1u << bit);
                    int cnt = 0;
                    for (int q = lane; q < m; q += 32) cnt += order_key(list[q]) >= cand;
                    if (__reduce_add_sync(FULL, cnt) + (KEY0 >= cand ? z : 0) >= N) key = cand;
                }
                double sg = 0.0;
                int cg = 0;
                float vN = 0.f;
                bool found = false;
                for (int q = lane; q < m; q += 32) {
                    const float v = list[q];
                    const uint32_t kv = order_key(v);
                    if (kv > key) { sg += (double)v; ++cg; }
                    if (kv == key) { vN = v; found = true; }
                }
                sg = warp_sum(sg);
                cg = __reduce_add_sync(FULL, cg) + (KEY0 > key ? z : 0);
                const unsigned has = __ballot_sync(FULL, found);
                vN = has ? __shfl_sync(FULL, vN, __ffs(has) - 1) : 0.f;     // no list entry at the key: the N-th is a zero
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
    const int64_t want = ceil_div(n, 8);
    const int64_t grid = want < (int64_t)kNumSMs * 4 ? want : (int64_t)kNumSMs * 4;
    if (ldS % 4 == 0 && (reinterpret_cast<uintptr_t>(S) & 15) == 0) {
        LARIS_CUDA(cudaFuncSetAttribute(row_qc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        row_qc_kernel<true><<<(unsigned)grid, 256, smem, st>>>(S, ldS, n, P, tl, total, nnz, top_sum);
    } else {
        LARIS_CUDA(cudaFuncSetAttribute(row_qc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        row_qc_kernel<false><<<(unsigned)grid, 256, smem, st>>>(S, ldS, n, P, tl, total, nnz, top_sum);
    }
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
