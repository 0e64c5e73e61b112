// Shared host/device helpers for the laris_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "../../include/laris_b200.h"

namespace laris {

void set_error(const char* fmt, ...);

#define LARIS_CHECK_ARG(cond, ...)                         \
    do {                                                   \
        if (!(cond)) {                                     \
            ::laris::set_error(__VA_ARGS__);               \
            return LARIS_E_ARG;                            \
        }                                                  \
    } while (0)

#define LARIS_CUDA(call)                                                             \
    do {                                                                             \
        cudaError_t e__ = (call);                                                    \
        if (e__ != cudaSuccess) {                                                    \
            ::laris::set_error("%s:%d %s: %s", __FILE__, __LINE__, #call,            \
                               cudaGetErrorString(e__));                             \
            return LARIS_E_CUDA;                                                     \
        }                                                                            \
    } while (0)

// exactly one LARIS_LAUNCH_CHECK() follows every kernel launch: it also feeds the launch counter
// that bench.py reports as gpu_launches
void note_launch();
#define LARIS_LAUNCH_CHECK()          \
    do {                              \
        ::laris::note_launch();       \
        LARIS_CUDA(cudaGetLastError()); \
    } while (0)

constexpr int kNumSMs = 148;   // B200: 2 dies x 74 SMs

inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }

// keep freed scratch cached in the device's default memory pool instead of handing it back to the
// driver at every synchronisation (the pool's default release threshold is 0)
void retain_pool_memory();

// stream-ordered scratch; freed on the same stream (no host sync)
template <typename T>
struct Scratch {
    T* p = nullptr;
    cudaStream_t s;
    explicit Scratch(cudaStream_t st) : s(st) {}
    cudaError_t alloc(size_t count) {
        retain_pool_memory();
        return cudaMallocAsync(&p, (count ? count : 1) * sizeof(T), s);
    }
    ~Scratch() { if (p) cudaFreeAsync(p, s); }
    Scratch(const Scratch&) = delete;
    Scratch& operator=(const Scratch&) = delete;
};

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ int warp_sum(int v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// streaming (read-once / write-once) 128-bit accesses that do not pollute L1
__device__ __forceinline__ float4 ld_stream_f4(const float4* p) {
    float4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
    return r;
}
__device__ __forceinline__ void st_stream_f4(float4* p, float4 v) {
    asm volatile("st.global.L1::no_allocate.v4.f32 [%0], {%1,%2,%3,%4};"
                 :: "l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

}  // namespace laris
