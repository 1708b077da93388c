// TMEM as per-lane scratch (no MMA): allocate all 512 columns, every warp of a 128-thread CTA writes / reads its own
// 32 lanes x 512 columns (2 KB per thread), verifies the round trip and times tcgen05.st / tcgen05.ld.
#include <cstdio>
#include <cuda_runtime.h>
#include "../automatic-speech-recognition_b200/csrc/tmem_ops_gen.h"
using namespace fe;

__global__ void __launch_bounds__(128, 1) k(float* out, long long* clk, int iters) {
    __shared__ uint32_t s_base;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"((uint32_t)__cvta_generic_to_shared(&s_base)), "r"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");
    const uint32_t base = s_base + ((uint32_t)(warp * 32) << 16);
    float v[64], w[64];
    int bad = 0;
    // round trip: column c of lane (warp, lane) <- tag
    for (int c0 = 0; c0 < 512; c0 += 64) {
#pragma unroll
        for (int i = 0; i < 64; ++i) v[i] = (float)(threadIdx.x * 1000 + c0 + i);
        tmem_st64(base + c0, v);
    }
    tmem_wait_st();
    for (int c0 = 0; c0 < 512; c0 += 64) {
        tmem_ld64(base + c0, w);
        tmem_wait_ld();
#pragma unroll
        for (int i = 0; i < 64; ++i) bad += (w[i] != (float)(threadIdx.x * 1000 + c0 + i));
    }
    // x16 at odd column offsets
    tmem_ld16(base + 37, w); tmem_wait_ld();
#pragma unroll
    for (int i = 0; i < 16; ++i) bad += (w[i] != (float)(threadIdx.x * 1000 + 37 + i));
    // timing: st 512 columns + wait
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
        for (int c0 = 0; c0 < 512; c0 += 64) tmem_st64(base + c0, v);
        tmem_wait_st();
    }
    long long t1 = clock64();
    float acc = 0.f;
    for (int it = 0; it < iters; ++it) {
        for (int c0 = 0; c0 < 512; c0 += 64) { tmem_ld64(base + c0, w); tmem_wait_ld(); acc += w[it & 63]; }
    }
    long long t2 = clock64();
    for (int it = 0; it < iters; ++it) {
        for (int c0 = 0; c0 < 512; c0 += 16) { tmem_ld16(base + c0, w); tmem_wait_ld(); acc += w[it & 15]; }
    }
    long long t3 = clock64();
    // latency of one dependent ld16 + wait
    for (int it = 0; it < iters; ++it) { tmem_ld16(base + (it & 255), w); tmem_wait_ld(); acc += w[0]; }
    long long t4 = clock64();
    out[blockIdx.x * 128 + threadIdx.x] = acc + bad;
    if (threadIdx.x == 0 && blockIdx.x == 0) { clk[0] = t1 - t0; clk[1] = t2 - t1; clk[2] = t3 - t2; clk[3] = t4 - t3; clk[4] = bad; }
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(s_base), "r"(512));
}

int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    float* d; cudaMalloc(&d, p.multiProcessorCount * 128 * 4);
    long long* dc; cudaMalloc(&dc, 64);
    const int iters = 200;
    k<<<p.multiProcessorCount, 128>>>(d, dc, iters);
    cudaError_t e = cudaDeviceSynchronize();
    long long c[5]; cudaMemcpy(c, dc, 40, cudaMemcpyDeviceToHost);
    printf("{\"status\": \"%s\", \"mismatches\": %lld, \"st512_cycles\": %.1f, \"ld512_x64_cycles\": %.1f, \"ld512_x16_cycles\": %.1f, \"ld16_wait_latency\": %.1f, "
           "\"note\": \"4 warps per SM, each 32 lanes x 512 columns (64 KB per warp, 2 KB per thread)\"}\n",
           cudaGetErrorString(e), c[4], (double)c[0] / iters, (double)c[1] / iters, (double)c[2] / iters, (double)c[3] / iters);
    return 0;
}
