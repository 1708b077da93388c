// How fast does the packed FFT16 codelet run at a given number of resident warps?  (registers only, no memory)
#include <cstdio>
#include <cuda_runtime.h>
#include "../automatic-speech-recognition_b200/csrc/fe_core.cuh"
using namespace fe;
template <int REGCAP>
__global__ void __launch_bounds__(128, 1) k(float* out, int iters, long long* cyc) {
    float2 r[16], i[16];
#pragma unroll
    for (int a = 0; a < 16; ++a) { r[a] = make_float2(threadIdx.x * 1e-3f + a, 0.5f * a); i[a] = make_float2(0.25f * a, threadIdx.x * 2e-3f - a); }
    __syncthreads();
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
        fft16(r, i);
#pragma unroll
        for (int a = 0; a < 16; ++a) { r[a] = pmul(r[a], pbc(0.25f)); }   // keep magnitudes bounded (16 FMUL2)
    }
    long long t1 = clock64();
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
    float s = 0; for (int a = 0; a < 16; ++a) s += r[a].x + r[a].y + i[a].x + i[a].y;
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
int main() {
    float* d; long long* dc; cudaMalloc(&d, 148 * 32 * 128 * 4); cudaMalloc(&dc, 8);
    const int iters = 2000;
    for (int ctas = 1; ctas <= 5; ++ctas) {       // 4 warps per CTA
        // occupancy controlled by dynamic smem so that exactly `ctas` CTAs fit per SM
        int smem = (220 * 1024) / ctas - 2048;
        cudaFuncSetAttribute(k<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        k<0><<<148 * ctas, 128, smem>>>(d, 10, dc);
        k<0><<<148 * ctas, 128, smem>>>(d, iters, dc);
        long long c; cudaMemcpy(&c, dc, 8, cudaMemcpyDeviceToHost);
        cudaError_t e = cudaGetLastError();
        // per fft16 call: 164 packed + 16 FMUL2 = 180 packed instrs per warp
        double packed_per_clk_sm = (double)iters * 180.0 * 4 * ctas / (double)c;
        printf("warps/SM=%2d  cycles/iter/warp=%.1f  packed instr/clk/SM=%.2f (pipe peak 2.0)  %s\n", 4 * ctas, (double)c / iters, packed_per_clk_sm, cudaGetErrorString(e));
    }
    return 0;
}
