// K5 on the 5th-generation tensor cores: num[p,(a,b)] = sum_i S[i,p] * N[i,a] * N[i,b]
// (reference: _pairwise_row_multiply + cosine_similarity numerators, laris/tools/_utils.py:859-866).
//
// GEMM view: D[M = pairs][N = type pairs] = A * B^T with the contraction over cells,
//   A[p][i] = S[i][p]                 (S is cell-major, so A is "MN-major": p is the contiguous index)
//   B[(a,b)][i] = N[i][a] * N[i][b]   (built on the fly from the C values of each cell)
// fp32 inputs are split into bf16 hi + bf16 lo; three tcgen05.mma (lo*hi, hi*lo, hi*hi) accumulate in
// fp32 in TMEM, which restores ~16 mantissa bits per operand.  Tensor-core accumulation truncates, so a
// TMEM accumulator only ever sums 256 cells (48 MMAs); the epilogue warps add each such chunk into fp32
// running sums in registers (round-to-nearest), and the K-split partials are combined in fp64.
//
// CTA = one 128 x 128 output tile x one range of cells; 16 warps, one CTA per SM:
//   warps 0-7   epilogue: tcgen05.ld the finished TMEM chunk (two accumulator buffers, so the next chunk
//               is being multiplied meanwhile) and add it to the running sums; at the end write the partial
//   warp  8     one elected thread issues the MMAs and commits them to mbarriers
//   warps 9-15  producers: read S (fp32, 32-byte pieces of 8 cell rows per quarter-warp) and N, split,
//               and store the bf16 operand tiles of a 64-cell stage in the canonical no-swizzle
//               MN-major layout (16-byte slots, conflict-free), then fence.proxy.async + mbarrier arrive
// Shared memory: 2 operand stages (A_hi, A_lo, B_hi, B_lo, 16 KB each), 2 raw fp32 S tiles and 3 raw N blocks
// in flight by cp.async.
#include "common.cuh"

#include <cuda_bf16.h>

namespace laris {
namespace mma {

constexpr int TM = 128, TN = 128, SK = 64;
constexpr int FLUSH_STAGES = 4;                        // 256 cells per TMEM accumulation chunk
constexpr int N_EPI = 8, N_PROD = 7;
constexpr int THREADS = (N_EPI + 1 + N_PROD) * 32;     // 512
constexpr int PROD_THREADS = N_PROD * 32;
constexpr int OPER_BYTES = SK * TM * 2;                // 16 KB
constexpr int STAGE_BYTES = 4 * OPER_BYTES;
constexpr int UNITS = SK * TM / 8;                     // 16-byte slots per operand tile (1024)
constexpr uint32_t IDESC = (1u << 4)      // D format fp32
                         | (1u << 7)      // A format bf16
                         | (1u << 10)     // B format bf16
                         | (1u << 15)     // A is MN-major
                         | (1u << 16)     // B is MN-major
                         | ((uint32_t)(TN >> 3) << 17) | ((uint32_t)(TM >> 4) << 24);
constexpr uint32_t LBO = (TM / 8) * 128;              // bytes between 8-cell groups (K direction)
constexpr uint32_t SBO = 128;                          // bytes between 8-column groups (MN direction)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.shared::cta.b64 st, [%0];\n\t}" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    do {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    } while (!ok);
}
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ uint64_t make_desc(uint32_t saddr) {
    const uint32_t lo = ((saddr >> 4) & 0x3fffu) | ((LBO >> 4) << 16);
    const uint32_t hi = (SBO >> 4) | (1u << 14);      // version 1 (Blackwell), no swizzle, base offset 0
    return ((uint64_t)hi << 32) | lo;
}
__device__ __forceinline__ void umma(uint32_t tacc, uint64_t adesc, uint64_t bdesc, uint32_t accumulate) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                 ::"r"(tacc), "l"(adesc), "l"(bdesc), "r"(IDESC), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}

// 8 fp32 -> 8 bf16 "hi" and 8 bf16 "lo" (x ~ hi + lo to ~16 mantissa bits)
__device__ __forceinline__ void split8(const float (&x)[8], uint4& hi, uint4& lo) {
    uint32_t h[4], l[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const __nv_bfloat162 hb = __floats2bfloat162_rn(x[2 * q], x[2 * q + 1]);
        h[q] = *reinterpret_cast<const uint32_t*>(&hb);
        const float h0 = __uint_as_float(h[q] << 16), h1 = __uint_as_float(h[q] & 0xffff0000u);
        const __nv_bfloat162 lb = __floats2bfloat162_rn(x[2 * q] - h0, x[2 * q + 1] - h1);
        l[q] = *reinterpret_cast<const uint32_t*>(&lb);
    }
    hi = make_uint4(h[0], h[1], h[2], h[3]);
    lo = make_uint4(l[0], l[1], l[2], l[3]);
}

constexpr int RAW_A_BYTES = SK * TM * 4;               // fp32 S tile of a stage as it arrives by cp.async (32 KB)
constexpr int RAW_A_SLOTS = 2, RAW_N_SLOTS = 3;

struct Smem {
    // dynamic layout (1024-byte aligned): [nstages][STAGE_BYTES] bf16 operand tiles, [2][RAW_A_BYTES] fp32 S tiles in
    // flight, [3][SK*C] fp32 N rows in flight, the (a,b) table of the tile's 128 columns, barriers, TMEM base address
    static __host__ __device__ size_t n_slot_bytes(int C) { return ((size_t)SK * C * 4 + 15) & ~(size_t)15; }
    static __host__ __device__ size_t bytes(int nstages, int C) {
        return (size_t)nstages * STAGE_BYTES + (size_t)RAW_A_SLOTS * RAW_A_BYTES + RAW_N_SLOTS * n_slot_bytes(C) + TN * 2 +
               16 * 8 + 16 + 1024;
    }
};

__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, uint32_t src_bytes) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

__global__ void __launch_bounds__(THREADS, 1)
pair_contract_mma_kernel(const float* __restrict__ S, int64_t ldS, int64_t n, int P, const float* __restrict__ Ncomp, int C,
                         int tiles_n, int tiles_mn, int64_t cells_per_split, int nstages,
                         float* __restrict__ partial /*[splits][P][C*C]*/) {
    extern __shared__ unsigned char smem_dyn[];
    unsigned char* base = reinterpret_cast<unsigned char*>(((uintptr_t)smem_dyn + 1023) & ~(uintptr_t)1023);
    const int CC = C * C;
    unsigned char* rawA = base + (size_t)nstages * STAGE_BYTES;
    unsigned char* rawN = rawA + (size_t)RAW_A_SLOTS * RAW_A_BYTES;
    const size_t n_slot = Smem::n_slot_bytes(C);
    uint16_t* ab_s = reinterpret_cast<uint16_t*>(rawN + RAW_N_SLOTS * n_slot);
    uint64_t* bars = reinterpret_cast<uint64_t*>((reinterpret_cast<uintptr_t>(ab_s + TN) + 7) & ~(uintptr_t)7);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 16);
    // barriers: full[0..3], empty[4..7], acc_full[8..9], acc_empty[10..11]
    const uint32_t bar0 = smem_u32(bars);
    auto BAR = [&](int i) { return bar0 + 8u * i; };

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int tile = blockIdx.x % tiles_mn, split = blockIdx.x / tiles_mn;
    const int tn = tile % tiles_n, tm = tile / tiles_n;
    const int p0 = tm * TM, n0 = tn * TN;
    const int64_t i_beg = (int64_t)split * cells_per_split;
    const int64_t i_end = min(n, i_beg + cells_per_split);
    const int nst = (int)((i_end - i_beg + SK - 1) / SK);
    const int nchunk = (nst + FLUSH_STAGES - 1) / FLUSH_STAGES;

    for (int c = tid; c < TN; c += THREADS) {
        const int ab = n0 + c;
        ab_s[c] = ab < CC ? (uint16_t)((ab / C) | ((ab % C) << 8)) : (uint16_t)0xffffu;
    }
    if (tid == 0) {
        for (int s = 0; s < nstages; ++s) { mbar_init(BAR(s), N_PROD); mbar_init(BAR(4 + s), 1); }
        mbar_init(BAR(8), 1); mbar_init(BAR(9), 1);
        mbar_init(BAR(10), N_EPI); mbar_init(BAR(11), N_EPI);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(2 * TN) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp < N_EPI) {
        // ================= epilogue: TMEM chunk -> fp32 running sums in registers =================
        const int quarter = warp & 3, half = warp >> 2;              // TMEM lanes 32*quarter.., columns 64*half..
        float sums[64];
#pragma unroll
        for (int j = 0; j < 64; ++j) sums[j] = 0.f;
        for (int c = 0; c < nchunk; ++c) {
            const int acc = c & 1;
            mbar_wait(BAR(8 + acc), (uint32_t)((c >> 1) & 1));
            tc_fence_after();
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                uint32_t v[32];
                const uint32_t taddr = tmem_base + ((uint32_t)(32 * quarter) << 16) + (uint32_t)(acc * TN + 64 * half + 32 * h);
                asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                             "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
                             "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
                             : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                               "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
                               "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
                               "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
                             : "r"(taddr) : "memory");
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
                for (int j = 0; j < 32; ++j) sums[32 * h + j] += __uint_as_float(v[j]);
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(BAR(10 + acc));
        }
        const int p = p0 + 32 * quarter + lane;
        if (p < P) {
            float* out = partial + ((int64_t)split * P + p) * CC + n0 + 64 * half;
#pragma unroll
            for (int j = 0; j < 64; ++j)
                if (n0 + 64 * half + j < CC) out[j] = sums[j];
        }
    } else if (warp == N_EPI) {
        // ================= MMA issuer: one elected thread =================
        if (lane == 0) {
            for (int s = 0; s < nst; ++s) {
                const int stage = s % nstages;
                const uint32_t ph = (uint32_t)((s / nstages) & 1);
                const int c = s / FLUSH_STAGES, acc = c & 1;
                const bool first = (s % FLUSH_STAGES) == 0;
                if (first) { mbar_wait(BAR(10 + acc), (uint32_t)(((c >> 1) & 1) ^ 1)); tc_fence_after(); }
                mbar_wait(BAR(stage), ph);
                tc_fence_after();
                const uint32_t sa = smem_u32(base + (size_t)stage * STAGE_BYTES);
                const uint32_t tacc = tmem_base + (uint32_t)(acc * TN);
#pragma unroll
                for (int kk = 0; kk < SK / 16; ++kk) {
                    const uint32_t off = (uint32_t)kk * 2u * LBO;
                    const uint64_t a_hi = make_desc(sa + off), a_lo = make_desc(sa + OPER_BYTES + off);
                    const uint64_t b_hi = make_desc(sa + 2 * OPER_BYTES + off), b_lo = make_desc(sa + 3 * OPER_BYTES + off);
                    umma(tacc, a_lo, b_hi, (first && kk == 0) ? 0u : 1u);
                    umma(tacc, a_hi, b_lo, 1u);
                    umma(tacc, a_hi, b_hi, 1u);
                }
                umma_commit(BAR(4 + stage));                         // smem stage is free once these MMAs retire
                if ((s % FLUSH_STAGES) == FLUSH_STAGES - 1 || s == nst - 1) umma_commit(BAR(8 + acc));
            }
        }
    } else {
        // ================= producers: fp32 -> split bf16 operand tiles =================
        // The raw fp32 inputs of stage s+1 are in flight (cp.async, no registers) while stage s is converted.
        // An S unit is fetched and consumed by the same thread, so only the N rows need a producer barrier.
        const int ptid = tid - (N_EPI + 1) * 32;
        const int n_chunks16 = (int)(n_slot / 16);
        auto prefetch = [&](int s) {
            const int64_t i0 = i_beg + (int64_t)s * SK;
            const uint32_t ra = smem_u32(rawA + (size_t)(s % RAW_A_SLOTS) * RAW_A_BYTES);
            for (int u = ptid; u < UNITS; u += PROD_THREADS) {
                const int r = u & 7, g = (u >> 3) & 15, kg = u >> 7;
                const int64_t i = i0 + kg * 8 + r;
                const int pc = p0 + 8 * g;
                const bool row = i < i_end;
                const float* src = S + (row ? i : 0) * ldS;
                const bool in0 = row && pc + 4 <= ldS, in1 = row && pc + 8 <= ldS;   // columns [P,ldS) hold zeros
                cp_async16(ra + 16 * u, in0 ? src + pc : S, in0 ? 16u : 0u);
                cp_async16(ra + UNITS * 16 + 16 * u, in1 ? src + pc + 4 : S, in1 ? 16u : 0u);
            }
            const uint32_t rn = smem_u32(rawN + (size_t)(s % RAW_N_SLOTS) * n_slot);
            const int64_t f0 = i0 * C, fend = i_end * C;             // the stage's N rows are one contiguous block
            for (int q = ptid; q < n_chunks16; q += PROD_THREADS) {
                const int64_t f = f0 + 4 * q;
                const int64_t left = fend - f;
                const uint32_t nb = left >= 4 ? 16u : left > 0 ? (uint32_t)(4 * left) : 0u;
                cp_async16(rn + 16 * q, nb ? Ncomp + f : Ncomp, nb);
            }
            cp_async_commit();
        };
        if (nst > 0) prefetch(0);
        for (int s = 0; s < nst; ++s) {
            const int stage = s % nstages;
            const uint32_t ph = (uint32_t)((s / nstages) & 1);
            if (s + 1 < nst) { prefetch(s + 1); cp_async_wait<1>(); } else { cp_async_wait<0>(); }
            if (lane == 0) mbar_wait(BAR(4 + stage), ph ^ 1u);
            __syncwarp();
            unsigned char* st = base + (size_t)stage * STAGE_BYTES;
            const unsigned char* ra = rawA + (size_t)(s % RAW_A_SLOTS) * RAW_A_BYTES;
            // A: unit u = 16-byte slot ((cell/8)*16 + colgroup)*8 + cell%8  <-  S[i0+cell][p0 + 8*colgroup ..+7]
            for (int u = ptid; u < UNITS; u += PROD_THREADS) {
                const float4 v0 = *reinterpret_cast<const float4*>(ra + 16 * u);
                const float4 v1 = *reinterpret_cast<const float4*>(ra + UNITS * 16 + 16 * u);
                const float x[8] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w};
                uint4 hi, lo;
                split8(x, hi, lo);
                *reinterpret_cast<uint4*>(st + 16 * u) = hi;
                *reinterpret_cast<uint4*>(st + OPER_BYTES + 16 * u) = lo;
            }
            asm volatile("bar.sync 1, %0;" ::"n"(PROD_THREADS) : "memory");     // every thread's N chunks of this stage landed
            const float* nsr = reinterpret_cast<const float*>(rawN + (size_t)(s % RAW_N_SLOTS) * n_slot);
            // B: same slot numbering with type-pair columns  <-  N[i][a] * N[i][b]
            for (int u = ptid; u < UNITS; u += PROD_THREADS) {
                const int r = u & 7, g = (u >> 3) & 15, kg = u >> 7;
                const float* nrow = nsr + (kg * 8 + r) * C;
                const uint4 abv = *reinterpret_cast<const uint4*>(ab_s + 8 * g);
                const uint32_t abw[4] = {abv.x, abv.y, abv.z, abv.w};
                float x[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const uint32_t ab = (abw[j >> 1] >> (16 * (j & 1))) & 0xffffu;
                    x[j] = ab == 0xffffu ? 0.f : nrow[ab & 0xffu] * nrow[ab >> 8];
                }
                uint4 hi, lo;
                split8(x, hi, lo);
                *reinterpret_cast<uint4*>(st + 2 * OPER_BYTES + 16 * u) = hi;
                *reinterpret_cast<uint4*>(st + 3 * OPER_BYTES + 16 * u) = lo;
            }
            fence_async_smem();                                      // generic-proxy writes -> visible to the tensor core
            __syncwarp();
            if (lane == 0) mbar_arrive(BAR(stage));
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(2 * TN) : "memory");
    }
}

__global__ void partial_reduce_kernel(const float* __restrict__ partial, int splits, int64_t total, double* __restrict__ num) {
    const int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (e >= total) return;
    double s = 0.0;
    for (int k = 0; k < splits; ++k) s += (double)partial[(int64_t)k * total + e];
    num[e] = s;
}

}  // namespace mma

int celltype_pair_contract_mma(const float* S, int64_t ldS, int64_t n, int32_t P, const float* Ncomp, int32_t C,
                               double* num, cudaStream_t st) {
    using namespace mma;
    if (C > 255 || ldS % 4 != 0 || ((uintptr_t)S & 15) != 0) {
        set_error("laris_celltype_pair_contract(impl=0): needs C <= 255, ldS %% 4 == 0 and a 16-byte aligned S");
        return LARIS_E_LIMIT;
    }
    int nstages = 2;
    if (Smem::bytes(nstages, C) > (size_t)227 * 1024) {
        set_error("laris_celltype_pair_contract(impl=0): C=%d does not fit shared memory", C);
        return LARIS_E_LIMIT;
    }
    const int CC = C * C;
    const int tiles_m = (int)ceil_div(P, TM), tiles_n = (int)ceil_div(CC, TN), tiles = tiles_m * tiles_n;
    int dev = 0, sms = kNumSMs;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    // K split: enough CTAs for whole waves of the SMs, at least 16 stages of work each
    const int64_t max_splits = ceil_div(n, (int64_t)16 * SK) < 1 ? 1 : ceil_div(n, (int64_t)16 * SK);
    int best = 1;
    double best_eff = 0.0;
    for (int s = 1; s <= 64 && s <= max_splits; ++s) {
        const double ctas = (double)tiles * s;
        const double eff = ctas / (ceil(ctas / sms) * sms);
        if (eff > best_eff + 0.02) { best_eff = eff; best = s; }
    }
    int64_t per = ceil_div(ceil_div(n, (int64_t)best), (int64_t)SK) * SK;
    const int splits = (int)ceil_div(n, per);
    Scratch<float> partial(st);
    LARIS_CUDA(partial.alloc((size_t)splits * P * CC));
    const size_t smem = Smem::bytes(nstages, C);
    LARIS_CUDA(cudaFuncSetAttribute(pair_contract_mma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    pair_contract_mma_kernel<<<(unsigned)(tiles * splits), THREADS, smem, st>>>(S, ldS, n, P, Ncomp, C, tiles_n, tiles, per, nstages,
                                                                               partial.p);
    LARIS_LAUNCH_CHECK();
    const int64_t total = (int64_t)P * CC;
    partial_reduce_kernel<<<(unsigned)ceil_div(total, 256), 256, 0, st>>>(partial.p, splits, total, num);
    LARIS_LAUNCH_CHECK();
    return LARIS_OK;
}

}  // namespace laris
