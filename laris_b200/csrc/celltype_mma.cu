// K5 on the 5th-generation tensor cores: num[p,(a,b)] = sum_i S[i,p] * N[i,a] * N[i,b]
// (reference: _pairwise_row_multiply + cosine_similarity numerators, laris/tools/_utils.py:859-866).
//
// GEMM view: D[M = pairs][N = type pairs] = A * B^T with the contraction over cells,
//   A[p][i] = S[i][p]                 (S is cell-major, so A is "MN-major": p is the contiguous index)
//   B[(a,b)][i] = N[i][a] * N[i][b]   (built on the fly from the C values of each cell)
// fp32 inputs are split three ways into bf16 hi + mid + lo (x = hi + mid + lo to 2^-24 relative: a two-way
// split only reaches 2^-17 per operand in the worst case, which breaks the 1e-5 bar on sums of a few terms);
// six tcgen05.mma (mid*mid, lo*hi, hi*lo, mid*hi, hi*mid, hi*hi; the dropped products are < 2^-23) accumulate
// in fp32 in TMEM.
// Tensor-core accumulation truncates (measured: bias grows linearly with the number of accumulated MMAs), so a
// TMEM accumulator only ever sums 256 cells (64 MMAs); the epilogue warps add each such chunk into fp32
// running sums in registers (round-to-nearest), and the K-split partials are combined in fp64.
//
// Two kernels:
//  split_scores_kernel   one pass over S: fp32 -> bf16 hi / mid / lo, written tile by tile in exactly the shared-memory
//                        operand layout (canonical no-swizzle MN-major: 16-byte slots [cell/8][pair/8][cell%8][8 pairs]),
//                        cells taken in the spatial processing order when one is given; also gathers the N rows
//                        in that order.  Cost: S read once, the same number of bytes written.
//  pair_contract_mma_kernel   CTA = one 128 x 128 output tile x one range of cells; 16 warps, one CTA per SM:
//   warps 0-7   epilogue: tcgen05.ld the finished TMEM chunk (two accumulator buffers, so the next chunk
//               is being multiplied meanwhile) and add it to the running sums; at the end write the partial
//   warp  8     one elected thread issues the MMAs and commits them to mbarriers
//   warps 9-15  producers: one thread issues the bulk copies (cp.async.bulk -> mbarrier complete_tx)
//               that bring the stage's A_hi / A_mid / A_lo tiles straight into their operand slots; all of them build
//               B = N[i,a]*N[i,b] (split hi/mid/lo) for the stage's 64 cells from the N rows (cp.async, one stage
//               ahead).  N is mostly zeros (a neighbourhood holds few cell types), so a unit whose N[i,a] is
//               zero is stored as zeros without touching the b side; with spatially ordered cells whole warps
//               take that path.  Then fence.proxy.async + mbarrier arrive.
// Shared memory: NSTAGES operand stages of six tiles (A_hi, A_mid, A_lo, B_hi, B_mid, B_lo; SK*256 bytes each).
#include "common.cuh"

#include <cuda_bf16.h>

namespace laris {
namespace mma {

#ifndef LARIS_MMA_SK
#define LARIS_MMA_SK 64
#endif
constexpr int TM = 128, TN = 128, SK = LARIS_MMA_SK;  // SK = cells per pipeline stage (32 or 64)
constexpr int NSTAGES = SK == 64 ? 2 : 4;              // 192 KB of operand tiles either way
constexpr int FLUSH_STAGES = 256 / SK;                 // 256 cells per TMEM accumulation chunk
constexpr int N_EPI = 8, N_PROD = 7;
constexpr int THREADS = (N_EPI + 1 + N_PROD) * 32;     // 512
constexpr int PROD_THREADS = N_PROD * 32;
constexpr int OPER_BYTES = SK * TM * 2;                // 16 KB
constexpr int NSPLIT = 3;                              // bf16 terms per fp32 operand
constexpr int STAGE_BYTES = 2 * NSPLIT * OPER_BYTES;
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

// 8 fp32 -> 8 bf16 "hi", "mid" and "lo": x = hi + mid + lo to 24 mantissa bits (each remainder is exact in fp32)
__device__ __forceinline__ void split8(const float (&x)[8], uint4& hi, uint4& mid, uint4& lo) {
    uint32_t h[4], m[4], l[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const __nv_bfloat162 hb = __floats2bfloat162_rn(x[2 * q], x[2 * q + 1]);
        h[q] = *reinterpret_cast<const uint32_t*>(&hb);
        const float r0 = x[2 * q] - __uint_as_float(h[q] << 16), r1 = x[2 * q + 1] - __uint_as_float(h[q] & 0xffff0000u);
        const __nv_bfloat162 mb = __floats2bfloat162_rn(r0, r1);
        m[q] = *reinterpret_cast<const uint32_t*>(&mb);
        const __nv_bfloat162 lb = __floats2bfloat162_rn(r0 - __uint_as_float(m[q] << 16), r1 - __uint_as_float(m[q] & 0xffff0000u));
        l[q] = *reinterpret_cast<const uint32_t*>(&lb);
    }
    hi = make_uint4(h[0], h[1], h[2], h[3]);
    mid = make_uint4(m[0], m[1], m[2], m[3]);
    lo = make_uint4(l[0], l[1], l[2], l[3]);
}

constexpr uint32_t A_TX_BYTES = NSPLIT * OPER_BYTES;

constexpr int LIST_CAP = 384;                          // non-zero B elements a warp remembers per stage buffer
constexpr int CELLS_PER_WARP = (SK + N_PROD - 1) / N_PROD;      // 10
constexpr int MAX_C = 64;                              // N row of a cell = two values per lane

struct Smem {
    // dynamic layout (1024-byte aligned): [nstages][STAGE_BYTES] bf16 operand tiles, [nstages][N_PROD][LIST_CAP] u16
    // offsets of the non-zero B elements each producer warp wrote into each stage buffer, control words, barriers
    static __host__ __device__ size_t bytes(int nstages) {
        return (size_t)nstages * STAGE_BYTES + (size_t)nstages * N_PROD * LIST_CAP * 2 + 256 + 16 * 8 + 1024;
    }
};

__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1;\n\t}" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait_sleep(uint32_t bar, uint32_t parity) {      // long waits: do not burn issue slots
    uint32_t ok;
    for (;;) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
        if (ok) break;
        __nanosleep(128);
    }
}

__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, uint32_t src_bytes) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// S (fp32, cell-major) -> A_hi / A_mid / A_lo tiles in operand layout: block (sg, tm) = cells [SK sg, SK sg + SK) of the
// processing order x pairs [128 tm, 128 tm + 128), hi tile, mid tile, lo tile.  One thread per 16-byte slot.
__global__ void __launch_bounds__(256)
split_scores_kernel(const float* __restrict__ S, int64_t ldS, int64_t n, int P, const int32_t* __restrict__ order,
                    int tiles_m, unsigned char* __restrict__ Asplit) {
    constexpr int CTAS_PER_TILE = UNITS / 256;
    const int64_t blk = blockIdx.x / CTAS_PER_TILE;
    const int u = (int)(blockIdx.x % CTAS_PER_TILE) * 256 + threadIdx.x;
    const int64_t sg = blk / tiles_m;
    const int tm = (int)(blk - sg * tiles_m);
    const int r = u & 7, g = (u >> 3) & 15, kg = u >> 7;
    const int64_t q = sg * SK + kg * 8 + r;
    const int pc = tm * TM + 8 * g;
    float x[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) x[j] = 0.f;
    if (q < n) {
        const int64_t i = order ? (int64_t)order[q] : q;
        const float* row = S + i * ldS + pc;
        if (pc + 8 <= ldS) {                                          // columns [P, ldS) hold zeros
            const float4 v0 = ld_stream_f4(reinterpret_cast<const float4*>(row));
            const float4 v1 = ld_stream_f4(reinterpret_cast<const float4*>(row + 4));
            x[0] = v0.x; x[1] = v0.y; x[2] = v0.z; x[3] = v0.w; x[4] = v1.x; x[5] = v1.y; x[6] = v1.z; x[7] = v1.w;
        } else {
#pragma unroll
            for (int j = 0; j < 8; ++j) if (pc + j < P) x[j] = row[j];
        }
    }
    uint4 hi, mid, lo;
    split8(x, hi, mid, lo);
    unsigned char* dst = Asplit + blk * (int64_t)A_TX_BYTES + 16 * u;
    *reinterpret_cast<uint4*>(dst) = hi;
    *reinterpret_cast<uint4*>(dst + OPER_BYTES) = mid;
    *reinterpret_cast<uint4*>(dst + 2 * OPER_BYTES) = lo;
}

__global__ void gather_rows_kernel(const float* __restrict__ N, int64_t n, int C, const int32_t* __restrict__ order,
                                   float* __restrict__ out) {
    const int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (e >= n * C) return;
    const int64_t q = e / C;
    out[e] = N[(int64_t)order[q] * C + (e - q * C)];
}

// which cell types occur (N != 0) in each 64-cell stage: one warp per stage, 64-bit mask
__global__ void stage_type_mask_kernel(const float* __restrict__ N, int64_t n, int C, int64_t nsg,
                                       unsigned long long* __restrict__ mask) {
    const int lane = threadIdx.x & 31;
    const int64_t sg = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    if (sg >= nsg) return;
    unsigned long long m = 0ull;
    for (int r = lane; r < SK; r += 32) {
        const int64_t i = sg * SK + r;
        if (i >= n) break;
        const float* row = N + i * C;
        for (int a = 0; a < C; ++a)
            if (row[a] != 0.f) m |= 1ull << a;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m |= __shfl_xor_sync(0xffffffffu, m, o);
    if (lane == 0) mask[sg] = m;
}

// per column tile: the ascending list of stages in which one of the tile's cell types a occurs (a superset of the
// stages with a non-zero B tile; the others contribute nothing and are never visited).  One CTA per tile.
__global__ void __launch_bounds__(1024)
active_stage_list_kernel(const unsigned long long* __restrict__ mask, int64_t nsg, int C, int32_t* __restrict__ list /*[tiles_n][nsg]*/,
                         int32_t* __restrict__ count /*[tiles_n]*/) {
    __shared__ int32_t warp_tot[32];
    __shared__ int32_t running;
    const int tn = blockIdx.x, tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    const int a_lo = tn * TN / C, a_hi = min(C - 1, (tn * TN + TN - 1) / C);
    unsigned long long am = 0ull;
    for (int a = a_lo; a <= a_hi; ++a) am |= 1ull << a;
    if (tid == 0) running = 0;
    __syncthreads();
    int32_t* out = list + (int64_t)tn * nsg;
    for (int64_t s0 = 0; s0 < nsg; s0 += 1024) {
        const int64_t sg = s0 + tid;
        const bool act = sg < nsg && (mask[sg] & am) != 0ull;
        const unsigned bm = __ballot_sync(0xffffffffu, act);
        if (lane == 0) warp_tot[w] = __popc(bm);
        __syncthreads();
        int32_t before = running;
        for (int q = 0; q < w; ++q) before += warp_tot[q];
        if (act) out[before + __popc(bm & ((1u << lane) - 1u))] = (int32_t)sg;
        __syncthreads();
        if (tid == 0) { int32_t t = 0; for (int q = 0; q < 32; ++q) t += warp_tot[q]; running += t; }
        __syncthreads();
    }
    if (tid == 0) count[tn] = running;
}

__global__ void __launch_bounds__(THREADS, 1)
pair_contract_mma_kernel(const unsigned char* __restrict__ Asplit, int tiles_m, int64_t n, int P,
                         const float* __restrict__ Ncomp, int C, int tiles_n, int tiles_mn, int splits,
                         const int32_t* __restrict__ alist, const int32_t* __restrict__ acount, int64_t nsg,
                         int nstages, float* __restrict__ partial /*[splits][P][C*C]*/) {
    extern __shared__ unsigned char smem_dyn[];
    unsigned char* base = reinterpret_cast<unsigned char*>(((uintptr_t)smem_dyn + 1023) & ~(uintptr_t)1023);
    const int CC = C * C;
    uint16_t* blist = reinterpret_cast<uint16_t*>(base + (size_t)nstages * STAGE_BYTES);   // [nstages][N_PROD][LIST_CAP]
    // control words: [4..7] non-zero count of the stage being built, [8..11] stage-skip flag, [12..13]
    // accumulator-holds-data flag, [14] TMEM base address, [16 + 8*stage + warp] per-warp list lengths
    uint32_t* ctrl = reinterpret_cast<uint32_t*>(blist + (size_t)nstages * N_PROD * LIST_CAP);
    uint64_t* bars = reinterpret_cast<uint64_t*>(ctrl + 64);
    uint32_t* tmem_slot = ctrl + 14;
    // barriers: full[0..3], empty[4..7], acc_full[8..9], acc_empty[10..11]
    const uint32_t bar0 = smem_u32(bars);
    auto BAR = [&](int i) { return bar0 + 8u * i; };

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int tile = blockIdx.x % tiles_mn, split = blockIdx.x / tiles_mn;
    const int tn = tile % tiles_n, tm = tile / tiles_n;
    const int p0 = tm * TM, n0 = tn * TN;
    // this CTA's share of the tile's active stages
    const int n_act = acount[tn];
    const int s_lo = (int)((int64_t)n_act * split / splits), s_hi = (int)((int64_t)n_act * (split + 1) / splits);
    const int32_t* my_stages = alist + (int64_t)tn * nsg + s_lo;
    const int nst = s_hi - s_lo;
    const int nchunk = (nst + FLUSH_STAGES - 1) / FLUSH_STAGES;

    if (tid < 64) ctrl[tid] = 0u;
    for (int q = tid; q < nstages * (NSPLIT * OPER_BYTES / 16); q += THREADS)   // B tiles start as zeros
        *reinterpret_cast<uint4*>(base + (size_t)(q / (NSPLIT * OPER_BYTES / 16)) * STAGE_BYTES + NSPLIT * OPER_BYTES +
                                  16 * (q % (NSPLIT * OPER_BYTES / 16))) = make_uint4(0u, 0u, 0u, 0u);
    fence_async_smem();
    if (tid == 0) {
        for (int s = 0; s < nstages; ++s) { mbar_init(BAR(s), N_PROD + 1); mbar_init(BAR(4 + s), 1); }
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
            mbar_wait_sleep(BAR(8 + acc), (uint32_t)((c >> 1) & 1));
            tc_fence_after();
            const bool has_data = *reinterpret_cast<volatile uint32_t*>(ctrl + 12 + acc) != 0u;
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                if (!has_data) break;                                // every stage of this chunk was skipped
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
            bool chunk_data = false;
            for (int s = 0; s < nst; ++s) {
                const int stage = s % nstages;
                const uint32_t ph = (uint32_t)((s / nstages) & 1);
                const int c = s / FLUSH_STAGES, acc = c & 1;
                if ((s % FLUSH_STAGES) == 0) {
                    mbar_wait(BAR(10 + acc), (uint32_t)(((c >> 1) & 1) ^ 1));
                    tc_fence_after();
                    chunk_data = false;
                }
                mbar_wait(BAR(stage), ph);
                tc_fence_after();
                if (*reinterpret_cast<volatile uint32_t*>(ctrl + 8 + stage) == 0u) {      // B of this stage is not all zeros
                    const uint32_t sa = smem_u32(base + (size_t)stage * STAGE_BYTES);
                    const uint32_t tacc = tmem_base + (uint32_t)(acc * TN);
#pragma unroll
                    for (int kk = 0; kk < SK / 16; ++kk) {
                        const uint32_t off = (uint32_t)kk * 2u * LBO;
                        const uint64_t a_hi = make_desc(sa + off), a_mi = make_desc(sa + OPER_BYTES + off),
                                       a_lo = make_desc(sa + 2 * OPER_BYTES + off);
                        const uint64_t b_hi = make_desc(sa + 3 * OPER_BYTES + off), b_mi = make_desc(sa + 4 * OPER_BYTES + off),
                                       b_lo = make_desc(sa + 5 * OPER_BYTES + off);
                        umma(tacc, a_mi, b_mi, (!chunk_data && kk == 0) ? 0u : 1u);     // smallest terms first
                        umma(tacc, a_lo, b_hi, 1u);
                        umma(tacc, a_hi, b_lo, 1u);
                        umma(tacc, a_mi, b_hi, 1u);
                        umma(tacc, a_hi, b_mi, 1u);
                        umma(tacc, a_hi, b_hi, 1u);
                    }
                    chunk_data = true;
                }
                umma_commit(BAR(4 + stage));                         // smem stage is free once these MMAs retire
                if ((s % FLUSH_STAGES) == FLUSH_STAGES - 1 || s == nst - 1) {
                    ctrl[12 + acc] = chunk_data ? 1u : 0u;
                    __threadfence_block();
                    umma_commit(BAR(8 + acc));
                }
            }
        }
    } else {
        // ================= producers: A tiles by bulk copy, B tiles scattered from the non-zeros of N =================
        // B[(a,b)][i] = N[i,a] N[i,b] is non-zero only where both types occur in cell i's neighbourhood: a handful of
        // elements per cell.  A stage buffer's B tiles are kept all-zero except for the elements its lists name;
        // re-using the buffer clears exactly those, then scatters the new ones.  Each producer warp owns every 7th
        // cell of a stage: it prefetches those cells' N rows into registers a stage ahead (lane = cell type), finds
        // the non-zero types by ballot and lets the lanes write the products.  A stage whose B stays empty is flagged:
        // no A copy is issued and the MMA thread skips it.
        const int pw = warp - (N_EPI + 1);
        const int a_lo = n0 / C, a_hi = min(C - 1, (n0 + TN - 1) / C);
        const unsigned a_mask_lo = a_lo < 32 ? (a_hi >= 31 ? 0xffffffffu : ((1u << (a_hi + 1)) - 1u)) & ~((1u << a_lo) - 1u) : 0u;
        const unsigned a_mask_hi = a_hi >= 32 ? ((a_hi - 32 >= 31 ? 0xffffffffu : ((1u << (a_hi - 31)) - 1u)) &
                                                ~(a_lo > 32 ? ((1u << (a_lo - 32)) - 1u) : 0u)) : 0u;
        // N rows of this warp's cells: stage s in cv*, s+1 in nv*, s+2 in fv* (two stages of loads in flight)
        float nv0[CELLS_PER_WARP], nv1[CELLS_PER_WARP], fv0[CELLS_PER_WARP], fv1[CELLS_PER_WARP];
        auto load_n = [&](int s, float (&v0)[CELLS_PER_WARP], float (&v1)[CELLS_PER_WARP]) {
            const int64_t i0 = s < nst ? (int64_t)my_stages[s] * SK : n;
#pragma unroll
            for (int j = 0; j < CELLS_PER_WARP; ++j) {
                const int cell = pw + N_PROD * j;
                const int64_t i = i0 + cell;
                const bool ok = cell < SK && i < n;
                v0[j] = ok && lane < C ? Ncomp[i * C + lane] : 0.f;
                v1[j] = ok && lane + 32 < C ? Ncomp[i * C + lane + 32] : 0.f;
            }
        };
        load_n(0, nv0, nv1);
        load_n(1, fv0, fv1);
        for (int s = 0; s < nst; ++s) {
            const int stage = s % nstages;
            const uint32_t ph = (uint32_t)((s / nstages) & 1);
            float cv0[CELLS_PER_WARP], cv1[CELLS_PER_WARP];
#pragma unroll
            for (int j = 0; j < CELLS_PER_WARP; ++j) { cv0[j] = nv0[j]; cv1[j] = nv1[j]; nv0[j] = fv0[j]; nv1[j] = fv1[j]; }
            load_n(s + 2, fv0, fv1);
            const int64_t sg = my_stages[s];
            if (lane == 0) mbar_wait(BAR(4 + stage), ph ^ 1u);       // the MMAs that read this buffer have retired
            __syncwarp();
            unsigned char* st = base + (size_t)stage * STAGE_BYTES;
            unsigned char* bh = st + NSPLIT * OPER_BYTES;
            uint16_t* lst = blist + ((size_t)stage * N_PROD + pw) * LIST_CAP;
            uint32_t* my_len = ctrl + 16 + 8 * stage + pw;
            // -- clear what this warp wrote into the buffer last time
            const uint32_t old = *my_len;
            if (old > LIST_CAP) {                                    // list overflowed: zero all columns of this warp's cells
                for (int q = lane; q < CELLS_PER_WARP * (TN / 8); q += 32) {
                    const int cell = pw + N_PROD * (q / (TN / 8));
                    if (cell < SK) {
                        const uint32_t off = (uint32_t)(cell >> 3) * LBO + (uint32_t)(cell & 7) * 16u + (uint32_t)(q % (TN / 8)) * SBO;
                        *reinterpret_cast<uint4*>(bh + off) = make_uint4(0u, 0u, 0u, 0u);
                        *reinterpret_cast<uint4*>(bh + OPER_BYTES + off) = make_uint4(0u, 0u, 0u, 0u);
                        *reinterpret_cast<uint4*>(bh + 2 * OPER_BYTES + off) = make_uint4(0u, 0u, 0u, 0u);
                    }
                }
            } else {
                for (uint32_t q = lane; q < old; q += 32) {
                    const uint32_t off = 2u * lst[q];
                    *reinterpret_cast<uint16_t*>(bh + off) = 0;
                    *reinterpret_cast<uint16_t*>(bh + OPER_BYTES + off) = 0;
                    *reinterpret_cast<uint16_t*>(bh + 2 * OPER_BYTES + off) = 0;
                }
            }
            __syncwarp();
            // -- scatter the non-zeros of this warp's cells
            uint32_t len = 0;
#pragma unroll
            for (int j = 0; j < CELLS_PER_WARP; ++j) {
                const int cell = pw + N_PROD * j;
                const unsigned nz0 = __ballot_sync(0xffffffffu, cv0[j] != 0.f), nz1 = __ballot_sync(0xffffffffu, cv1[j] != 0.f);
                unsigned am0 = nz0 & a_mask_lo, am1 = nz1 & a_mask_hi;
                const uint32_t koff = (uint32_t)(cell >> 3) * LBO + (uint32_t)(cell & 7) * 16u;
                while (am0 | am1) {                                  // each non-zero type a of this tile's column range
                    int a;
                    float va;
                    if (am0) { a = __ffs(am0) - 1; am0 &= am0 - 1; va = __shfl_sync(0xffffffffu, cv0[j], a); }
                    else { a = 32 + __ffs(am1) - 1; am1 &= am1 - 1; va = __shfl_sync(0xffffffffu, cv1[j], a - 32); }
#pragma unroll
                    for (int hb = 0; hb < 2; ++hb) {                 // lane = cell type b (two halves)
                        const int b = lane + 32 * hb;
                        const float q = va * (hb ? cv1[j] : cv0[j]);
                        const int nl = a * C + b - n0;
                        const bool hit = q != 0.f && b < C && nl >= 0 && nl < TN;
                        const unsigned hm = __ballot_sync(0xffffffffu, hit);
                        if (hit) {
                            const uint32_t off = koff + (uint32_t)(nl >> 3) * SBO + (uint32_t)(nl & 7) * 2u;
                            const __nv_bfloat16 h = __float2bfloat16_rn(q);
                            const float r1 = q - __bfloat162float(h);
                            const __nv_bfloat16 m = __float2bfloat16_rn(r1);
                            const __nv_bfloat16 l = __float2bfloat16_rn(r1 - __bfloat162float(m));
                            *reinterpret_cast<__nv_bfloat16*>(bh + off) = h;
                            *reinterpret_cast<__nv_bfloat16*>(bh + OPER_BYTES + off) = m;
                            *reinterpret_cast<__nv_bfloat16*>(bh + 2 * OPER_BYTES + off) = l;
                            const uint32_t pos = len + __popc(hm & ((1u << lane) - 1u));
                            if (pos < LIST_CAP) lst[pos] = (uint16_t)(off >> 1);
                        }
                        len += __popc(hm);
                    }
                }
            }
            if (lane == 0) {
                *my_len = len;                                       // > LIST_CAP means "overflowed" to the next user of the buffer
                if (len) atomicAdd(ctrl + 4 + stage, len);
            }
            fence_async_smem();                                      // generic-proxy writes -> visible to the tensor core
            asm volatile("bar.sync 1, %0;" ::"n"(PROD_THREADS) : "memory");     // the stage's count is final
            if (pw == 0 && lane == 0) {
                const uint32_t cnt = ctrl[4 + stage];
                ctrl[4 + stage] = 0u;
                ctrl[8 + stage] = cnt == 0u ? 1u : 0u;
                if (cnt) {                                           // A_hi | A_mid | A_lo of this stage: one contiguous block
                    mbar_expect_tx(BAR(stage), A_TX_BYTES);
                    const unsigned char* src = Asplit + (sg * tiles_m + tm) * (int64_t)A_TX_BYTES;
#pragma unroll
                    for (int t = 0; t < NSPLIT; ++t)
                        bulk_g2s(smem_u32(st + t * OPER_BYTES), src + t * OPER_BYTES, OPER_BYTES, BAR(stage));
                } else {
                    mbar_arrive(BAR(stage));
                }
            }
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
                               const int32_t* order, double* num, cudaStream_t st) {
    using namespace mma;
    if (C < 1 || C > MAX_C || ldS % 4 != 0 || ((uintptr_t)S & 15) != 0) {
        set_error("laris_celltype_pair_contract(impl=0): needs C <= %d, ldS %% 4 == 0 and a 16-byte aligned S", MAX_C);
        return LARIS_E_LIMIT;
    }
    const int nstages = NSTAGES;
    const int CC = C * C;
    const int tiles_m = (int)ceil_div(P, TM), tiles_n = (int)ceil_div(CC, TN), tiles = tiles_m * tiles_n;
    int dev = 0, sms = kNumSMs;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    // K split: the active stages of a column tile are dealt evenly to `splits` CTAs; aim at >= 4 waves of CTAs so
    // the hardware scheduler evens out tiles with few / many active stages
    const int64_t nsg = ceil_div(n, SK);
    int splits = (int)ceil_div((int64_t)sms * 4, tiles);
    if (splits > nsg / 8) splits = (int)(nsg / 8);
    if (splits < 1) splits = 1;
    // pass 1: split S once (and put the N rows in the same cell order); active stages of every column tile
    Scratch<unsigned char> asplit(st);
    Scratch<float> nsorted(st), partial(st);
    Scratch<unsigned long long> smask(st);
    Scratch<int32_t> alist(st), acount(st);
    LARIS_CUDA(asplit.alloc((size_t)nsg * tiles_m * A_TX_BYTES));
    split_scores_kernel<<<(unsigned)(nsg * tiles_m * (UNITS / 256)), 256, 0, st>>>(S, ldS, n, P, order, tiles_m, asplit.p);
    LARIS_LAUNCH_CHECK();
    const float* nrows = Ncomp;
    if (order) {
        LARIS_CUDA(nsorted.alloc((size_t)n * C));
        gather_rows_kernel<<<(unsigned)ceil_div(n * C, 256), 256, 0, st>>>(Ncomp, n, C, order, nsorted.p);
        LARIS_LAUNCH_CHECK();
        nrows = nsorted.p;
    }
    LARIS_CUDA(smask.alloc((size_t)nsg));
    LARIS_CUDA(alist.alloc((size_t)tiles_n * nsg));
    LARIS_CUDA(acount.alloc((size_t)tiles_n));
    stage_type_mask_kernel<<<(unsigned)ceil_div(nsg * 32, 256), 256, 0, st>>>(nrows, n, C, nsg, smask.p);
    LARIS_LAUNCH_CHECK();
    active_stage_list_kernel<<<(unsigned)tiles_n, 1024, 0, st>>>(smask.p, nsg, C, alist.p, acount.p);
    LARIS_LAUNCH_CHECK();
    // pass 2: the contraction
    LARIS_CUDA(partial.alloc((size_t)splits * P * CC));
    const size_t smem = Smem::bytes(nstages);
    LARIS_CUDA(cudaFuncSetAttribute(pair_contract_mma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    pair_contract_mma_kernel<<<(unsigned)(tiles * splits), THREADS, smem, st>>>(asplit.p, tiles_m, n, P, nrows, C, tiles_n, tiles,
                                                                               splits, alist.p, acount.p, nsg, nstages, partial.p);
    LARIS_LAUNCH_CHECK();
    const int64_t total = (int64_t)P * CC;
    partial_reduce_kernel<<<(unsigned)ceil_div(total, 256), 256, 0, st>>>(partial.p, splits, total, num);
    LARIS_LAUNCH_CHECK();
    return LARIS_OK;
}

}  // namespace laris
