// Device exclusive scan (out[0]=0, out[i+1]=sum in[0..i]) used for CSR row
// pointers and grid-cell offsets.  Three phases: per-block sums, a recursive
// scan of those sums, then a per-block rescan with the block offset added.
#pragma once
#include "common.cuh"

namespace laris {

constexpr int kScanThreads = 256, kScanItems = 8, kScanTile = kScanThreads * kScanItems;

template <typename T>
__device__ __forceinline__ T block_exclusive_scan(T v, T* total, T* smem /*[32]*/) {
    // returns the exclusive prefix of v over the block; *total = block sum
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    T inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        T t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += t;
    }
    if (lane == 31) smem[wid] = inc;
    __syncthreads();
    if (wid == 0) {
        T w = lane < (kScanThreads >> 5) ? smem[lane] : T(0);
        T winc = w;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            T t = __shfl_up_sync(0xffffffffu, winc, o);
            if (lane >= o) winc += t;
        }
        smem[lane] = winc - w;                                   // exclusive warp offsets
        if (lane == 31) smem[32] = winc;                         // block total
    }
    __syncthreads();
    T res = inc - v + smem[wid];
    *total = smem[32];
    return res;
}

template <typename T>
__global__ void __launch_bounds__(kScanThreads) scan_block_sums_kernel(const T* __restrict__ in, int64_t n, T* __restrict__ sums) {
    __shared__ T sm[33];
    const int64_t base = (int64_t)blockIdx.x * kScanTile + (int64_t)threadIdx.x * kScanItems;
    T s = 0;
#pragma unroll
    for (int j = 0; j < kScanItems; ++j) if (base + j < n) s += in[base + j];
    T total;
    block_exclusive_scan<T>(s, &total, sm);
    if (threadIdx.x == 0) sums[blockIdx.x] = total;
}

template <typename T>
__global__ void __launch_bounds__(kScanThreads) scan_apply_kernel(const T* __restrict__ in, int64_t n, const T* __restrict__ block_off,
                                                                  T* __restrict__ out) {
    __shared__ T sm[33];
    const int64_t base = (int64_t)blockIdx.x * kScanTile + (int64_t)threadIdx.x * kScanItems;
    T v[kScanItems];
    T s = 0;
#pragma unroll
    for (int j = 0; j < kScanItems; ++j) { v[j] = base + j < n ? in[base + j] : T(0); s += v[j]; }
    T total;
    T pre = block_exclusive_scan<T>(s, &total, sm) + (block_off ? block_off[blockIdx.x] : T(0));
    if (blockIdx.x == 0 && threadIdx.x == 0) out[0] = 0;
#pragma unroll
    for (int j = 0; j < kScanItems; ++j) {
        pre += v[j];
        if (base + j < n) out[base + j + 1] = pre;
    }
}

template <typename T>
int exclusive_scan(const T* in, int64_t n, T* out, cudaStream_t st) {
    if (n <= 0) { LARIS_CUDA(cudaMemsetAsync(out, 0, sizeof(T), st)); return LARIS_OK; }
    const int64_t nb = ceil_div(n, kScanTile);
    if (nb == 1) {
        scan_apply_kernel<T><<<1, kScanThreads, 0, st>>>(in, n, nullptr, out);
        LARIS_LAUNCH_CHECK();
        return LARIS_OK;
    }
    Scratch<T> sums(st), offs(st);
    LARIS_CUDA(sums.alloc(nb));
    LARIS_CUDA(offs.alloc(nb + 1));
    scan_block_sums_kernel<T><<<(unsigned)nb, kScanThreads, 0, st>>>(in, n, sums.p);
    LARIS_LAUNCH_CHECK();
    int rc = exclusive_scan<T>(sums.p, nb, offs.p, st);
    if (rc) return rc;
    scan_apply_kernel<T><<<(unsigned)nb, kScanThreads, 0, st>>>(in, n, offs.p, out);
    LARIS_LAUNCH_CHECK();
    return LARIS_OK;
}

}  // namespace laris
