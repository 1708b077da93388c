// Shared-memory wavefront micro-benchmark: how many cycles per warp-instruction for address patterns
// that the classic 32-bank model calls conflict-free.
#include <cstdio>
#include <cuda_runtime.h>
#include <stdint.h>
__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
template <int MODE> __global__ void k(float* out, int iters, long long* cyc) {
    extern __shared__ __align__(2048) unsigned char sm[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int fs = lane >> 3, t = lane & 7;
    uint32_t base = s32(sm) + warp * 8192;
    uint32_t a;
    if (MODE == 0) a = base + lane * 4;                                        // STS.32 one line
    if (MODE == 1) a = base + fs * 2048 + (fs * 8 + t) * 4;                    // STS.32 4 lines, distinct banks
    if (MODE == 2) a = base + fs * 2048 + t * 256 + ((t ^ ((fs & 1) << 2)) * 16);   // LDS.128 v2 stage B
    if (MODE == 3) a = base + fs * 2048 + t * 128 + ((t ^ ((fs & 1) << 2)) * 16);   // LDS.128 v1 stage B
    if (MODE == 4) a = base + lane * 16;                                       // LDS.128 contiguous
    if (MODE == 5) a = base + (fs * 80 + t + 8 * (fs >> 1)) * 4;               // LDS.32 raw pattern
    if (MODE == 6) a = base + lane * 4;                                        // LDS.32 one line
    if (MODE == 7) a = base + fs * 2048 + t * 16 ;                             // LDS.128: 4 lines (one per quarter), contiguous within quarter
    if (MODE == 8) a = base + fs * 512 + t * 256 + (t * 16);                   // LDS.128: 8 lines per quarter, quads distinct, no frame xor
    if (MODE == 9) a = base + (lane >> 4) * 2048 + (lane & 15) * 4 + (lane >> 4) * 64;   // STS.32 2 lines distinct banks
    float4 acc = make_float4(0, 0, 0, 0);
    float v = lane;
    __syncthreads();
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            if (MODE == 0 || MODE == 1 || MODE == 9) { asm volatile("st.volatile.shared.f32 [%0], %1;" :: "r"(a), "f"(v) : "memory"); }
            else if (MODE == 5 || MODE == 6) { float x; asm volatile("ld.volatile.shared.f32 %0, [%1];" : "=f"(x) : "r"(a) : "memory"); acc.x += x; }
            else { float4 x; asm volatile("ld.volatile.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(x.x), "=f"(x.y), "=f"(x.z), "=f"(x.w) : "r"(a) : "memory"); acc.x += x.x; acc.y += x.w; }
        }
    }
    long long t1 = clock64();
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc.x + acc.y + v;
}
template <int MODE> void run(const char* name, float* d, long long* dc) {
    const int iters = 2000, block = 512, grid = 148;
    cudaFuncSetAttribute(k<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 16 * 8192 + 2048);
    k<MODE><<<grid, block, 16 * 8192 + 2048>>>(d, 10, dc);
    k<MODE><<<grid, block, 16 * 8192 + 2048>>>(d, iters, dc);
    long long c; cudaMemcpy(&c, dc, 8, cudaMemcpyDeviceToHost);
    // 16 warps per SM each issuing iters*8 instructions
    printf("%-44s %6.2f cycles per warp-instruction (SM-wide)\n", name, (double)c / (iters * 8.0 * 16));
}
int main() {
    float* d; long long* dc; cudaMalloc(&d, 148 * 512 * 4); cudaMalloc(&dc, 8);
    run<0>("STS.32 one line", d, dc);
    run<9>("STS.32 2 lines, distinct banks", d, dc);
    run<1>("STS.32 4 lines, distinct banks", d, dc);
    run<6>("LDS.32 one line", d, dc);
    run<5>("LDS.32 raw-PCM pattern (4 frames)", d, dc);
    run<4>("LDS.128 contiguous", d, dc);
    run<7>("LDS.128 4 lines (1/quarter) contiguous", d, dc);
    run<3>("LDS.128 v1 stageB (rows 128B apart)", d, dc);
    run<2>("LDS.128 v2 stageB (rows 256B apart)", d, dc);
    run<8>("LDS.128 8 lines/quarter 256B apart no xor", d, dc);
    return 0;
}
