// Micro-benchmark: scalar FFMA/FADD vs packed fma.rn.f32x2 / add.rn.f32x2 throughput on sm_100a.
#include <cstdio>
#include <cuda_runtime.h>
template <int MODE> __global__ void k(float* out, int iters, float a, float b) {
    float2 x[8];
    for (int i = 0; i < 8; ++i) x[i] = make_float2(threadIdx.x * 0.001f + i, threadIdx.x * 0.002f - i);
    float2 A = make_float2(a, a * 0.5f), B = make_float2(b, b * 0.25f);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (MODE == 0) { x[i].x = fmaf(x[i].x, A.x, B.x); x[i].y = fmaf(x[i].y, A.y, B.y); }      // 2 FFMA
            if (MODE == 1) { x[i] = __ffma2_rn(x[i], A, B); }                                           // 1 FFMA2
            if (MODE == 2) { x[i].x = x[i].x + B.x; x[i].y = x[i].y + B.y; }                           // 2 FADD
            if (MODE == 3) { x[i] = __fadd2_rn(x[i], B); }                                              // 1 FADD2
            if (MODE == 4) { x[i] = __fmul2_rn(x[i], A); }                                              // 1 FMUL2
        }
    }
    float s = 0; for (int i = 0; i < 8; ++i) s += x[i].x + x[i].y;
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int MODE> void run(const char* name, float* d, int sms) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    int iters = 20000, grid = sms * 8, block = 256;
    k<MODE><<<grid, block>>>(d, 100, 1.0001f, 0.5f);
    cudaEventRecord(e0); k<MODE><<<grid, block>>>(d, iters, 1.0001f, 0.5f); cudaEventRecord(e1);
    cudaEventSynchronize(e1); float ms; cudaEventElapsedTime(&ms, e0, e1);
    double lane_ops = (double)grid * block * iters * 16.0;      // scalar-equivalent ops (2 per element pair x 8)
    printf("%-8s %.3f ms  %.2f T lane-ops/s  (%.1f lane-ops/clk/SM at 1.9 GHz)\n", name, ms, lane_ops / ms / 1e9,
           lane_ops / (ms * 1e-3) / sms / 1.9e9);
}
int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    float* d; cudaMalloc(&d, p.multiProcessorCount * 8 * 256 * 4);
    printf("%s SMs=%d clock=%d kHz\n", p.name, p.multiProcessorCount, p.clockRate);
    run<0>("FFMA", d, p.multiProcessorCount); run<1>("FFMA2", d, p.multiProcessorCount);
    run<2>("FADD", d, p.multiProcessorCount); run<3>("FADD2", d, p.multiProcessorCount);
    run<4>("FMUL2", d, p.multiProcessorCount);
    run<0>("FFMA", d, p.multiProcessorCount); run<1>("FFMA2", d, p.multiProcessorCount);
    return 0;
}
