// Microbenchmark behind DESIGN.md §9.2: bandwidth of RANDOM row gathers (row = W bytes) from a working set of a
// given size - does a matrix block that fits the 126 MB L2 serve random gathers faster than HBM does?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o l2_gather_probe l2_gather_probe.cu && ./l2_gather_probe
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

__device__ __forceinline__ uint32_t mix(uint32_t x) {
    x ^= x >> 16; x *= 0x85EBCA6Bu; x ^= x >> 13; x *= 0xC2B2AE35u; x ^= x >> 16;
    return x;
}

// QUADS float4 per row; a thread owns one quad lane and gathers `per_thread` random rows, 8 loads in flight
template <int QUADS>
__global__ void probe(const float4* __restrict__ M, uint32_t rows, int per_thread, uint32_t salt, float4* __restrict__ out) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t q = t % QUADS, g = t / QUADS;
    float4 acc = make_float4(0, 0, 0, 0);
    for (int i = 0; i < per_thread; i += 8) {
        float4 v[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const uint32_t r = mix(g * 977u + (uint32_t)(i + j) * 0x9E3779B1u + salt) % rows;
            v[j] = __ldg(M + (size_t)r * QUADS + q);
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) { acc.x += v[j].x; acc.y += v[j].y; acc.z += v[j].z; acc.w += v[j].w; }
    }
    if (acc.x == 1234.5f) out[t] = acc;
}

template <int QUADS>
void run(size_t bytes) {
    const uint32_t rows = (uint32_t)(bytes / (QUADS * 16));
    float4 *M, *out;
    cudaMalloc(&M, (size_t)rows * QUADS * 16);
    cudaMalloc(&out, 1 << 20);
    cudaMemset(M, 0, (size_t)rows * QUADS * 16);
    const int threads = 256, blocks = 148 * 16, per_thread = 2048;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e30f;
    for (int rep = 0; rep < 4; ++rep) {
        cudaEventRecord(e0);
        probe<QUADS><<<blocks, threads>>>(M, rows, per_thread, 17u * rep, out);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        if (rep > 0 && ms < best) best = ms;
    }
    const double gathered = (double)blocks * threads * per_thread * 16.0;
    printf("row %4d B  working set %7.1f MB : %8.1f GB/s of gathered bytes (%.3f ms)\n", QUADS * 16, bytes / 1048576.0,
           gathered / best / 1e6, best);
    cudaFree(M); cudaFree(out);
}

int main() {
    const size_t mb = 1 << 20;
    for (size_t ws : {16 * mb, 32 * mb, 48 * mb, 64 * mb, 80 * mb, 96 * mb, 128 * mb, 256 * mb, 1024 * mb, 8192 * mb}) {
        run<4>(ws);      // 64-byte rows  (16 columns)
    }
    for (size_t ws : {32 * mb, 64 * mb, 96 * mb, 128 * mb, 1024 * mb, 8192 * mb}) run<8>(ws);     // 128-byte rows
    for (size_t ws : {64 * mb, 8192 * mb}) run<128>(ws);                                          // 2 KB rows (today's K4b pieces)
    return 0;
}
