// Micro-benchmark: FP32 issue rate on sm_100a, the denominator of K1's roofline.
//   scalar FFMA / FADD (register operands), packed fma.rn.f32x2 / add.rn.f32x2 / mul.rn.f32x2,
//   each with 8 or 16 independent chains per thread and the inner body unrolled 8x (loop overhead < 2 %).
// Prints T lane-ops/s (1 FFMA = 1 lane-op = 2 FLOP) and lane-ops/clk/SM at the clock the run measured.
#include <cstdio>
#include <cuda_runtime.h>

template <int MODE, int CH>
__global__ void __launch_bounds__(256) k(float* out, int iters, float a, float b, long long* clk) {
    float2 x[CH];
#pragma unroll
    for (int i = 0; i < CH; ++i) x[i] = make_float2(threadIdx.x * 0.001f + i, threadIdx.x * 0.002f - i);
    // all operands in registers: the multiplier / addend differ per chain so that nothing is a uniform or constant operand
    float2 A[CH], B[CH];
#pragma unroll
    for (int i = 0; i < CH; ++i) { A[i] = make_float2(a + i * 1e-7f, a * 0.5f + threadIdx.x * 1e-9f); B[i] = make_float2(b + threadIdx.x * 1e-6f, b * 0.25f + i); }
    const long long c0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
#pragma unroll
            for (int i = 0; i < CH; ++i) {
                if (MODE == 0) { x[i].x = fmaf(x[i].x, A[i].x, B[i].x); x[i].y = fmaf(x[i].y, A[i].y, B[i].y); }   // 2 FFMA
                if (MODE == 1) { x[i] = __ffma2_rn(x[i], A[i], B[i]); }                                             // 1 FFMA2
                if (MODE == 2) { x[i].x = x[i].x + B[i].x; x[i].y = x[i].y + B[i].y; }                              // 2 FADD
                if (MODE == 3) { x[i] = __fadd2_rn(x[i], B[i]); }                                                   // 1 FADD2
                if (MODE == 4) { x[i] = __fmul2_rn(x[i], A[i]); }                                                   // 1 FMUL2
            }
        }
    }
    const long long c1 = clock64();
    float s = 0;
#pragma unroll
    for (int i = 0; i < CH; ++i) s += x[i].x + x[i].y;
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *clk = c1 - c0;
}

template <int MODE, int CH> void run(const char* name, float* d, long long* dclk, int sms, int ctas_per_sm) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int iters = 4000, grid = sms * ctas_per_sm, block = 256;
    k<MODE, CH><<<grid, block>>>(d, 50, 1.0001f, 0.5f, dclk);
    float best = 1e30f; long long clk = 0;
    for (int rep = 0; rep < 3; ++rep) {
        cudaEventRecord(e0); k<MODE, CH><<<grid, block>>>(d, iters, 1.0001f, 0.5f, dclk); cudaEventRecord(e1);
        cudaEventSynchronize(e1); float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) { best = ms; cudaMemcpy(&clk, dclk, 8, cudaMemcpyDeviceToHost); }
    }
    const double lane_ops = (double)grid * block * iters * 8.0 * CH * 2.0;      // scalar-equivalent ops
    const double ghz = clk / (best * 1e-3) / 1e9 * (1.0 / ctas_per_sm) ;        // cycles of one CTA / wall: only meaningful at 1 wave
    printf("{\"op\": \"%s\", \"chains\": %d, \"ctas_per_sm\": %d, \"ms\": %.3f, \"tera_lane_ops\": %.2f, \"tflops_if_fma\": %.2f, "
           "\"cta_cycles\": %lld}\n", name, CH, ctas_per_sm, best, lane_ops / best / 1e9, 2 * lane_ops / best / 1e9, clk);
    (void)ghz;
}

int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    float* d; cudaMalloc(&d, (size_t)p.multiProcessorCount * 8 * 256 * 4);
    long long* dclk; cudaMalloc(&dclk, 8);
    int khz = 0; cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
    printf("{\"device\": \"%s\", \"sms\": %d, \"clock_khz\": %d, \"nominal_tflops\": %.2f}\n", p.name, p.multiProcessorCount, khz,
           p.multiProcessorCount * 128.0 * 2.0 * khz * 1e3 / 1e12);
    const int S = p.multiProcessorCount;
    for (int c : {4, 8}) {
        if (c == 4) { run<0, 8>("FFMA", d, dclk, S, c); run<0, 16>("FFMA", d, dclk, S, c); run<1, 8>("FFMA2", d, dclk, S, c); run<1, 16>("FFMA2", d, dclk, S, c);
                      run<2, 8>("FADD", d, dclk, S, c); run<3, 8>("FADD2", d, dclk, S, c); run<4, 8>("FMUL2", d, dclk, S, c); }
        else        { run<0, 8>("FFMA", d, dclk, S, c); run<1, 8>("FFMA2", d, dclk, S, c); run<1, 16>("FFMA2", d, dclk, S, c); }
    }
    return 0;
}
