// Table arithmetic of runLARIS step 2 on the device: everything between the S-sized passes (K4/K5/K5b) and the
// p-value kernels (K6/K7) works on P x C^2-sized fp64 tables that never leave HBM:
//   T1  cosine + COSG regularisation of the co-localisation numerators      (reference laris/tools/_utils.py:862-880)
//   T2  per-column IQR / log1p / max normalisation (stand-in of cosg.iqrLogNormalize, :596, :975; parity unpinned)
//   T3  interaction scores, mean of the 100 largest, rescale, co-localisation weight, threshold   (:1351-1489)
//   T5  p-values from the exceedance counts, the rounded copy and the selection mask BH-FDR takes, and the
//       clip / NaN / -log10 epilogue                                                               (:1573-1658)
// All kernels are elementwise or small per-row / per-column reductions with a fixed summation order
// (bitwise reproducible).  Order statistics use an exact 8-pass radix select on order-preserving 64-bit keys,
// so there is no limit on P.
#include <math.h>

#include "common.cuh"

namespace laris {

constexpr int kTbThreads = 256;

// deterministic block-wide sum (fixed lane / warp order); every thread gets the total
__device__ __forceinline__ double block_sum(double v, double* red) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    v = warp_sum(v);
    __syncthreads();
    if (lane == 0) red[warp] = v;
    __syncthreads();
    double t = 0.0;
    for (int i = 0; i < nw; ++i) t += red[i];
    return t;
}

// ------------------------------------------------------------------ T1 --
// one CTA per pair row: cos = num / (sqrt(ss_p) * sqrt(qn2_l)) (0 where the denominator is 0), then
// reg = cos^2 / ((1-mu) cos^2 + mu * sum_l cos^2) * cos   (mu == 1: / sum_l cos^2); exact zeros stay zero
__global__ void __launch_bounds__(kTbThreads)
pair_cosine_regularise_kernel(const double* __restrict__ num, const double* __restrict__ ss,
                              const double* __restrict__ qn2, int L, double mu, double* __restrict__ cos_out,
                              double* __restrict__ reg_out) {
    __shared__ double red[kTbThreads / 32];
    const int64_t p = blockIdx.x;
    const double sp = sqrt(ss[p]);
    const double* nrow = num + p * L;
    double acc = 0.0;
    for (int l = threadIdx.x; l < L; l += kTbThreads) {
        const double den = sp * sqrt(qn2[l]);
        const double c = den != 0.0 ? nrow[l] / den : 0.0;
        if (cos_out) cos_out[p * L + l] = c;
        acc += c * c;
    }
    const double tot = block_sum(acc, red);
    for (int l = threadIdx.x; l < L; l += kTbThreads) {
        const double den = sp * sqrt(qn2[l]);
        const double c = den != 0.0 ? nrow[l] / den : 0.0;
        const double sq = c * c;
        const double d2 = mu == 1.0 ? tot : (1.0 - mu) * sq + mu * tot;
        reg_out[p * L + l] = c != 0.0 ? (sq / d2) * c : 0.0;
    }
}

// ------------------------------------------------------- radix select --
// order-preserving key of a double (larger value <=> larger key); NaN sorts below everything
__device__ __forceinline__ uint64_t f64_key(double v) {
    if (v != v) return 0ull;
    const uint64_t b = (uint64_t)__double_as_longlong(v);
    return (b >> 63) ? ~b : (b | 0x8000000000000000ull);
}
__device__ __forceinline__ double key_f64(uint64_t k) {
    const uint64_t b = (k >> 63) ? (k & 0x7fffffffffffffffull) : ~k;
    return __longlong_as_double((long long)b);
}

// Key of the element of ascending rank `rank` (0-based) among the elements of v[0..m) with include(v) true,
// by 8 passes of 8-bit digits (most significant first).  CTA-wide; `hist` is 256 ints of shared memory and
// `bc` two 64-bit broadcast slots.  POSITIVE_ONLY: only v > 0 take part.  DESC: rank counts from the largest.
template <bool POSITIVE_ONLY, bool DESC>
__device__ uint64_t cta_select_key(const double* __restrict__ v, int64_t m, int64_t stride, int64_t rank, int* hist,
                                   unsigned long long* bc) {
    uint64_t prefix = 0ull, mask = 0ull;
    for (int byte = 7; byte >= 0; --byte) {
        for (int i = threadIdx.x; i < 256; i += blockDim.x) hist[i] = 0;
        __syncthreads();
        for (int64_t i = threadIdx.x; i < m; i += blockDim.x) {
            const double x = v[i * stride];
            if (POSITIVE_ONLY && !(x > 0.0)) continue;
            uint64_t k = f64_key(x);
            if (DESC) k = ~k;
            if ((k & mask) == prefix) atomicAdd(&hist[(int)((k >> (8 * byte)) & 255ull)], 1);
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            int64_t r = rank;
            int b = 0;
            for (; b < 255; ++b) {
                if (r < hist[b]) break;
                r -= hist[b];
            }
            bc[0] = (unsigned long long)b;
            bc[1] = (unsigned long long)r;
        }
        __syncthreads();
        prefix |= (uint64_t)bc[0] << (8 * byte);
        mask |= 255ull << (8 * byte);
        rank = (int64_t)bc[1];
        __syncthreads();
    }
    return prefix;
}

// (count of participating keys <= key, smallest participating key > key) - gives the next order statistic
template <bool POSITIVE_ONLY>
__device__ void cta_next_key(const double* __restrict__ v, int64_t m, int64_t stride, uint64_t key, int* hist,
                             unsigned long long* bc, int64_t* n_le, uint64_t* next) {
    if (threadIdx.x == 0) { hist[0] = 0; bc[0] = ~0ull; }
    __syncthreads();
    int c = 0;
    unsigned long long mn = ~0ull;
    for (int64_t i = threadIdx.x; i < m; i += blockDim.x) {
        const double x = v[i * stride];
        if (POSITIVE_ONLY && !(x > 0.0)) continue;
        const uint64_t k = f64_key(x);
        if (k <= key) ++c; else if (k < mn) mn = k;
    }
    atomicAdd(&hist[0], c);
    atomicMin(&bc[0], mn);
    __syncthreads();
    *n_le = hist[0];
    *next = (uint64_t)bc[0];
    __syncthreads();
}

// numpy.quantile(..., method='linear') of the positive entries: h = (m-1) q, lerp of the two order statistics
template <bool POSITIVE_ONLY>
__device__ double cta_quantile(const double* __restrict__ v, int64_t len, int64_t stride, int64_t m, double q, int* hist,
                               unsigned long long* bc) {
    const double h = (double)(m - 1) * q;
    const int64_t i = (int64_t)floor(h);
    const double t = h - (double)i;
    const uint64_t ka = cta_select_key<POSITIVE_ONLY, false>(v, len, stride, i, hist, bc);
    const double a = key_f64(ka);
    if (!(t > 0.0) || i + 1 >= m) return a;
    int64_t n_le;
    uint64_t kb;
    cta_next_key<POSITIVE_ONLY>(v, len, stride, ka, hist, bc, &n_le, &kb);
    const double b = i + 1 < n_le ? a : key_f64(kb);
    const double d = b - a;
    return t >= 0.5 ? b - d * (1.0 - t) : a + d * t;          // numpy's _lerp
}

// ------------------------------------------------------------------ T2 --
// one CTA per column j of a [rows][cols] table: scale = Q_upper - Q_lower of the positive entries (their maximum
// when that is not positive), y = log1p(max(x, 0) / scale), out = y / max(y); a column without positive entries is 0.
__global__ void __launch_bounds__(kTbThreads)
column_iqr_log_kernel(const double* __restrict__ in, int64_t rows, int64_t cols, double q_upper, double q_lower,
                      double* __restrict__ out) {
    __shared__ int hist[256];
    __shared__ unsigned long long bc[2];
    __shared__ double red[kTbThreads / 32];
    __shared__ double s_mx;
    const int64_t j = blockIdx.x;
    const double* col = in + j;
    double cnt = 0.0, mx = 0.0;
    for (int64_t i = threadIdx.x; i < rows; i += kTbThreads) {
        const double x = col[i * cols];
        if (x > 0.0) { cnt += 1.0; mx = fmax(mx, x); }
    }
    const int64_t m = (int64_t)block_sum(cnt, red);
    // block max through the same scratch
    {
        const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
        for (int o = 16; o > 0; o >>= 1) mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
        __syncthreads();
        if (lane == 0) red[warp] = mx;
        __syncthreads();
        if (threadIdx.x == 0) {
            double t = 0.0;
            for (int i = 0; i < kTbThreads / 32; ++i) t = fmax(t, red[i]);
            s_mx = t;
        }
        __syncthreads();
        mx = s_mx;
    }
    if (m == 0) {
        for (int64_t i = threadIdx.x; i < rows; i += kTbThreads) out[i * cols + j] = 0.0;
        return;
    }
    const double qu = cta_quantile<true>(col, rows, cols, m, q_upper, hist, bc);
    const double ql = cta_quantile<true>(col, rows, cols, m, q_lower, hist, bc);
    double scale = qu - ql;
    if (!(scale > 0.0)) scale = mx;
    const double ymax = log1p(mx / scale);
    for (int64_t i = threadIdx.x; i < rows; i += kTbThreads) {
        const double x = col[i * cols];
        out[i * cols + j] = log1p(fmax(x, 0.0) / scale) / ymax;
    }
}

// ------------------------------------------------------------------ T3 --
// inter[t][s][r] = (frac[p,s] * frac[p,r]) * ((sl[t,s] * sr[t,r]) * factor[t]),  p = pair_id[t],
// frac[p,c] = cnt[c][p] / n_c[c]   (same operation order as the host formula it replaces)
__global__ void interaction_scores_kernel(const double* __restrict__ sl, const double* __restrict__ sr,
                                          const double* __restrict__ factor, const int64_t* __restrict__ cnt,
                                          const double* __restrict__ n_c, const int32_t* __restrict__ pair_id, int Pt,
                                          int P, int C, double* __restrict__ inter) {
    const int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    const int64_t L = (int64_t)C * C;
    if (e >= (int64_t)Pt * L) return;
    const int t = (int)(e / L), g = (int)(e - (int64_t)t * L), s = g / C, r = g - s * C;
    const int p = pair_id[t];
    const double fs = (double)cnt[(int64_t)s * P + p] / n_c[s], fr = (double)cnt[(int64_t)r * P + p] / n_c[r];
    inter[e] = (fs * fr) * ((sl[(int64_t)t * C + s] * sr[(int64_t)t * C + r]) * factor[t]);
}

// stage 1 of the mean of the k largest: every CTA reduces a chunk to its own k largest (ties filled with the threshold)
__global__ void __launch_bounds__(kTbThreads)
topk_candidates_kernel(const double* __restrict__ v, int64_t m, int64_t chunk, int k, double* __restrict__ cand,
                       int* __restrict__ cand_count) {
    __shared__ int hist[256];
    __shared__ unsigned long long bc[2];
    __shared__ int s_pos;
    const int64_t beg = (int64_t)blockIdx.x * chunk;
    const int64_t len = min(chunk, m - beg);
    const double* x = v + beg;
    double* out = cand + (int64_t)blockIdx.x * k;
    if (len <= k) {
        for (int64_t i = threadIdx.x; i < len; i += kTbThreads) out[i] = x[i];
        if (threadIdx.x == 0) cand_count[blockIdx.x] = (int)len;
        return;
    }
    const uint64_t kth = ~cta_select_key<false, true>(x, len, 1, k - 1, hist, bc);      // key of the k-th largest
    if (threadIdx.x == 0) s_pos = 0;
    __syncthreads();
    for (int64_t i = threadIdx.x; i < len; i += kTbThreads)
        if (f64_key(x[i]) > kth) out[atomicAdd(&s_pos, 1)] = x[i];
    __syncthreads();
    const double thr = key_f64(kth);
    for (int i = s_pos + threadIdx.x; i < k; i += kTbThreads) out[i] = thr;
    if (threadIdx.x == 0) cand_count[blockIdx.x] = k;
}

// stage 2 (one CTA): mean of the k largest candidates -> out[0]; out[1] = the rescale 0.1 / mean (1 if mean <= 0)
__global__ void __launch_bounds__(kTbThreads)
topk_mean_kernel(const double* __restrict__ cand, const int* __restrict__ cand_count, int n_chunks, int k, int64_t m,
                 double* __restrict__ packed, double* __restrict__ out) {
    __shared__ int hist[256];
    __shared__ unsigned long long bc[2];
    __shared__ double red[kTbThreads / 32];
    __shared__ int s_total;
    // compact the per-chunk candidate lists (each CTA wrote cand_count[c] <= k values)
    if (threadIdx.x == 0) {
        int t = 0;
        for (int c = 0; c < n_chunks; ++c) t += cand_count[c];
        s_total = t;
    }
    __syncthreads();
    const int total = s_total;
    for (int c = 0, o = 0; c < n_chunks; o += cand_count[c], ++c)
        for (int i = threadIdx.x; i < cand_count[c]; i += kTbThreads) packed[o + i] = cand[(int64_t)c * k + i];
    __syncthreads();
    const int kk = (int)min((int64_t)k, m);
    if (kk <= 0) {
        if (threadIdx.x == 0) { out[0] = 0.0; out[1] = 1.0; }
        return;
    }
    double sum = 0.0, gt = 0.0, thr = 0.0;
    if (total <= kk) {
        for (int i = threadIdx.x; i < total; i += kTbThreads) sum += packed[i];
        sum = block_sum(sum, red);
    } else {
        const uint64_t kth = ~cta_select_key<false, true>(packed, total, 1, kk - 1, hist, bc);
        thr = key_f64(kth);
        for (int i = threadIdx.x; i < total; i += kTbThreads)
            if (f64_key(packed[i]) > kth) { sum += packed[i]; gt += 1.0; }
        sum = block_sum(sum, red);
        gt = block_sum(gt, red);
        sum += ((double)kk - gt) * thr;
    }
    if (threadIdx.x == 0) {
        const double mean = sum / (double)kk;
        out[0] = mean;
        out[1] = mean > 0.0 ? 0.1 / mean : 1.0;
    }
}

// flat = (inter * rescale) * ctc[p][g]; flat < threshold -> 0; rows layout [t][g] and the group-major
// score table table[g][p] (+ active[p]) the p-value kernel reads
__global__ void finalize_scores_kernel(const double* __restrict__ inter, const double* __restrict__ rescale,
                                       const double* __restrict__ ctc, const int32_t* __restrict__ pair_id, int Pt, int P,
                                       int L, double score_threshold, double* __restrict__ flat,
                                       double* __restrict__ table, uint8_t* __restrict__ active) {
    const int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (e >= (int64_t)Pt * L) return;
    const int t = (int)(e / L), g = (int)(e - (int64_t)t * L);
    const int p = pair_id[t];
    double f = (inter[e] * rescale[1]) * ctc[(int64_t)p * L + g];
    if (f < score_threshold) f = 0.0;
    flat[e] = f;
    table[(int64_t)g * P + p] = f;
    if (g == 0) active[p] = 1;
}

// ------------------------------------------------------------------ T5 --
// p-values from the exceedance counts (group-major [L][P]), the 6-decimal copy BH-FDR ranks, and its selection mask
__global__ void pvalues_kernel(const double* __restrict__ table, const uint8_t* __restrict__ active,
                               const int32_t* __restrict__ ge, const int32_t* __restrict__ nz,
                               const int32_t* __restrict__ ge_nz, int64_t total, int P, int n_perm, int conditional,
                               int prefilter, double prefilter_threshold, double* __restrict__ p_out,
                               double* __restrict__ p_round, uint8_t* __restrict__ sel) {
    const int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (e >= total) return;
    const double sc = table[e];
    double pv;
    if (conditional) pv = sc > 0.0 ? ((double)ge_nz[e] + 1.0) / ((double)nz[e] + 1.0) : 1.0;
    else pv = ((double)ge[e] + 1.0) / ((double)n_perm + 1.0);
    p_out[e] = pv;
    p_round[e] = rint(pv * 1e6) / 1e6;                              // numpy.round(p, 6)
    const int pcol = (int)(e % P);
    sel[e] = (uint8_t)((active[pcol] != 0) && (!prefilter || sc > prefilter_threshold));
}

// rows layout [t][g] of the group-major p / fdr tables, with the reference's epilogue on the fdr:
// min(fdr, 1), NaN -> 1, nlog10 = max(-log10(fdr + 1e-10), 0)
__global__ void gather_rows_kernel(const double* __restrict__ p_tab, const double* __restrict__ fdr_tab,
                                   const int32_t* __restrict__ pair_id, int Pt, int P, int L, double* __restrict__ p_rows,
                                   double* __restrict__ fdr_rows, double* __restrict__ nlog_rows) {
    const int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (e >= (int64_t)Pt * L) return;
    const int t = (int)(e / L), g = (int)(e - (int64_t)t * L);
    const int64_t src = (int64_t)g * P + pair_id[t];
    p_rows[e] = p_tab[src];
    double f = fmin(fdr_tab[src], 1.0);
    if (f != f) f = 1.0;
    fdr_rows[e] = f;
    nlog_rows[e] = fmax(-log10(f + 1e-10), 0.0);
}

// exceedance counts of a permutation null over a table: ge[e] += (val[e] >= obs[e])
__global__ void count_exceed_kernel(const double* __restrict__ obs, const double* __restrict__ val, int64_t total,
                                    int32_t* __restrict__ ge) {
    const int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (e < total && val[e] >= obs[e]) ge[e] += 1;
}

}  // namespace laris

using namespace laris;

extern "C" int laris_count_exceed(const double* obs, const double* val, int64_t total, int32_t* ge, void* stream) {
    LARIS_CHECK_ARG(obs && val && ge && total >= 0, "laris_count_exceed: bad arguments");
    if (total == 0) return LARIS_OK;
    count_exceed_kernel<<<(unsigned)ceil_div(total, 256), 256, 0, (cudaStream_t)stream>>>(obs, val, total, ge);
    LARIS_LAUNCH_CHECK();
    return LARIS_OK;
}

extern "C" int laris_pair_cosine_regularise(const double* num, const double* ss, const double* qn2, int32_t P, int32_t L,
                                            double mu, double* cos_out, double* reg_out, void* stream) {
    LARIS_CHECK_ARG(num && ss && qn2 && reg_out && P >= 1 && L >= 1, "laris_pair_cosine_regularise: bad arguments");
    pair_cosine_regularise_kernel<<<(unsigned)P, kTbThreads, 0, (cudaStream_t)stream>>>(num, ss, qn2, L, mu, cos_out,
                                                                                       reg_out);
    LARIS_LAUNCH_CHECK();
    return LARIS_OK;
}

extern "C" int laris_column_iqr_log(const double* table, int64_t rows, int64_t cols, double q_upper, double q_lower,
                                    double* out, void* stream) {
    LARIS_CHECK_ARG(table && out && rows >= 1 && cols >= 1 && cols <= 0x7fffffff, "laris_column_iqr_log: bad arguments");
    LARIS_CHECK_ARG(q_upper >= 0.0 && q_upper <= 1.0 && q_lower >= 0.0 && q_lower <= 1.0,
                    "laris_column_iqr_log: quantiles must lie in [0,1]");
    column_iqr_log_kernel<<<(unsigned)cols, kTbThreads, 0, (cudaStream_t)stream>>>(table, rows, cols, q_upper, q_lower, out);
    LARIS_LAUNCH_CHECK();
    return LARIS_OK;
}

extern "C" int laris_interaction_scores(const double* spec_l, const double* spec_r, const double* factor,
                                        const int64_t* cnt, const double* n_c, const int32_t* pair_id, int32_t Pt,
                                        int32_t P, int32_t C, double* inter, void* stream) {
    LARIS_CHECK_ARG(spec_l && spec_r && factor && cnt && n_c && pair_id && inter && Pt >= 1 && P >= 1 && C >= 1,
                    "laris_interaction_scores: bad arguments");
    const int64_t total = (int64_t)Pt * C * C;
    interaction_scores_kernel<<<(unsigned)ceil_div(total, 256), 256, 0, (cudaStream_t)stream>>>(
        spec_l, spec_r, factor, cnt, n_c, pair_id, Pt, P, C, inter);
    LARIS_LAUNCH_CHECK();
    return LARIS_OK;
}

extern "C" int laris_topk_mean(const double* values, int64_t m, int32_t k, double* out2, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    LARIS_CHECK_ARG(values && out2 && m >= 0 && k >= 1 && k <= 4096, "laris_topk_mean: bad arguments (k in [1,4096])");
    const int64_t chunk = 16384;
    const int n_chunks = (int)(m > 0 ? ceil_div(m, chunk) : 1);
    Scratch<double> cand(st), packed(st);
    Scratch<int> count(st);
    LARIS_CUDA(cand.alloc((size_t)n_chunks * k));
    LARIS_CUDA(packed.alloc((size_t)n_chunks * k));
    LARIS_CUDA(count.alloc((size_t)n_chunks));
    LARIS_CUDA(cudaMemsetAsync(count.p, 0, sizeof(int) * (size_t)n_chunks, st));
    if (m > 0) {
        topk_candidates_kernel<<<(unsigned)n_chunks, kTbThreads, 0, st>>>(values, m, chunk, k, cand.p, count.p);
        LARIS_LAUNCH_CHECK();
    }
    topk_mean_kernel<<<1, kTbThreads, 0, st>>>(cand.p, count.p, n_chunks, k, m, packed.p, out2);
    LARIS_LAUNCH_CHECK();
    return LARIS_OK;
}

extern "C" int laris_finalize_scores(const double* inter, const double* rescale2, const double* ctc,
                                     const int32_t* pair_id, int32_t Pt, int32_t P, int32_t L, double score_threshold,
                                     double* flat, double* table, uint8_t* active, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    LARIS_CHECK_ARG(inter && rescale2 && ctc && pair_id && flat && table && active && Pt >= 1 && P >= 1 && L >= 1,
                    "laris_finalize_scores: bad arguments");
    LARIS_CUDA(cudaMemsetAsync(table, 0, sizeof(double) * (size_t)P * L, st));
    LARIS_CUDA(cudaMemsetAsync(active, 0, (size_t)P, st));
    const int64_t total = (int64_t)Pt * L;
    finalize_scores_kernel<<<(unsigned)ceil_div(total, 256), 256, 0, st>>>(inter, rescale2, ctc, pair_id, Pt, P, L,
                                                                         score_threshold, flat, table, active);
    LARIS_LAUNCH_CHECK();
    return LARIS_OK;
}

extern "C" int laris_pvalues_from_counts(const double* table, const uint8_t* active, const int32_t* ge, const int32_t* nz,
                                         const int32_t* ge_nz, int32_t L, int32_t P, int32_t n_perm, int conditional,
                                         int prefilter, double prefilter_threshold, double* p_out, double* p_round,
                                         uint8_t* sel, void* stream) {
    LARIS_CHECK_ARG(table && active && ge && p_out && p_round && sel && L >= 1 && P >= 1 && n_perm >= 0,
                    "laris_pvalues_from_counts: bad arguments");
    LARIS_CHECK_ARG(!conditional || (nz && ge_nz), "laris_pvalues_from_counts: the conditional p-value needs nz and ge_nz");
    const int64_t total = (int64_t)L * P;
    pvalues_kernel<<<(unsigned)ceil_div(total, 256), 256, 0, (cudaStream_t)stream>>>(
        table, active, ge, nz, ge_nz, total, P, n_perm, conditional, prefilter, prefilter_threshold, p_out, p_round, sel);
    LARIS_LAUNCH_CHECK();
    return LARIS_OK;
}

extern "C" int laris_gather_result_rows(const double* p_tab, const double* fdr_tab, const int32_t* pair_id, int32_t Pt,
                                        int32_t P, int32_t L, double* p_rows, double* fdr_rows, double* nlog_rows,
                                        void* stream) {
    LARIS_CHECK_ARG(p_tab && fdr_tab && pair_id && p_rows && fdr_rows && nlog_rows && Pt >= 1 && P >= 1 && L >= 1,
                    "laris_gather_result_rows: bad arguments");
    const int64_t total = (int64_t)Pt * L;
    gather_rows_kernel<<<(unsigned)ceil_div(total, 256), 256, 0, (cudaStream_t)stream>>>(p_tab, fdr_tab, pair_id, Pt, P, L,
                                                                                       p_rows, fdr_rows, nlog_rows);
    LARIS_LAUNCH_CHECK();
    return LARIS_OK;
}
