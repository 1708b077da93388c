// Issue-cost model of the FP32 instructions K1 is made of (sm_100a).  Every kernel runs an unrolled
// body of inline-PTX instructions over distinct registers, 4 warps per SMSP (CTA = 512 threads, 1 CTA / SM),
// and reports cycles per warp-instruction per SMSP (clock64 of one warp / instructions / warps per SMSP).
#include <cstdio>
#include <cuda_runtime.h>

#define F2(d, a, b, c) asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c))
#define A2(d, a, b)    asm volatile("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b))
#define M2(d, a, b)    asm volatile("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b))
#define F1(d, a, b, c) asm volatile("fma.rn.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c))
#define A1(d, a, b)    asm volatile("add.rn.f32 %0, %1, %2;" : "=f"(d) : "f"(a), "f"(b))

typedef unsigned long long u64;
__device__ __forceinline__ u64 pk(float a, float b) { u64 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ float lo(u64 v) { float a, b; asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); return a + b; }

template <int MODE>
__global__ void __launch_bounds__(512, 1) k(float* out, int iters, float fa, float fb, long long* clk) {
    const float t = threadIdx.x * 1e-6f;
    u64 x[8], a[8], b[8];
    float s[16], sa[8], sb[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) { x[i] = pk(t + i, t - i); a[i] = pk(fa + i * 1e-7f + t, fa - t); b[i] = pk(fb + t, fb + i); sa[i] = fa + i * 1e-7f + t; sb[i] = fb + i + t; }
#pragma unroll
    for (int i = 0; i < 16; ++i) s[i] = t + i;
    const u64 ua = pk(fa, fa), ub = pk(fb, fb);         // warp-uniform operands (kernel arguments)
    const long long c0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 4; ++u) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                if (MODE == 0) F2(x[i], x[i], a[i], b[i]);                        // FFMA2 three distinct register pairs
                if (MODE == 1) F2(x[i], a[i], b[i], x[i]);                        // FFMA2 accumulate
                if (MODE == 2) F2(x[i], x[i], ua, ub);                            // FFMA2 two uniform operands
                if (MODE == 3) F2(x[i], x[i], ua, b[i]);                          // FFMA2 one uniform multiplier
                if (MODE == 4) F2(x[i], x[i], a[0], b[0]);                        // FFMA2 same register operands every time (reuse)
                if (MODE == 5) A2(x[i], x[i], b[i]);                              // FADD2
                if (MODE == 6) M2(x[i], x[i], a[i]);                              // FMUL2
                if (MODE == 7) { F1(s[2 * i], s[2 * i], sa[i], sb[i]); F1(s[2 * i + 1], s[2 * i + 1], sb[i], sa[i]); }  // 2 scalar FFMA
                if (MODE == 8) { A2(x[i], x[i], b[i]); F2(a[i], a[i], ua, b[i]); }   // FADD2 + FFMA2(uniform) interleaved
                if (MODE == 9) { A2(x[i], x[i], b[i]); F1(s[2 * i], s[2 * i], sa[i], sb[i]); F1(s[2 * i + 1], s[2 * i + 1], sb[i], sa[i]); }  // FADD2 + 2 FFMA
                if (MODE == 10) { A2(x[i], x[i], b[i]); F2(a[i], a[i], x[i], b[i]); }  // FADD2 + dependent FFMA2 (3 reg)
                if (MODE == 11) { F1(s[2 * i], s[2 * i], fa, sb[i]); F1(s[2 * i + 1], s[2 * i + 1], fa, sa[i]); }       // scalar FFMA, uniform multiplier
                if (MODE == 12) { A1(s[2 * i], s[2 * i], sb[i]); A1(s[2 * i + 1], s[2 * i + 1], sa[i]); }               // 2 scalar FADD
            }
        }
    }
    const long long c1 = clock64();
    float r = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) r += lo(x[i]) + lo(a[i]) + s[2 * i] + s[2 * i + 1];
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
    if (threadIdx.x == 0 && blockIdx.x == 0) *clk = c1 - c0;
}

template <int MODE> void run(const char* name, int per_iter, float* d, long long* dclk, int sms) {
    const int iters = 2000;
    k<MODE><<<sms, 512>>>(d, 20, 1.0001f, 0.5f, dclk);
    long long best = 1LL << 62;
    for (int rep = 0; rep < 3; ++rep) {
        k<MODE><<<sms, 512>>>(d, iters, 1.0001f, 0.5f, dclk);
        long long c; cudaMemcpy(&c, dclk, 8, cudaMemcpyDeviceToHost);
        if (c < best) best = c;
    }
    // 4 warps per SMSP share the issue port: cycles per warp-instruction = cycles / (instructions per warp * 4)
    const double per = (double)best / ((double)iters * 4 * 8 * per_iter * 4);
    printf("{\"pattern\": \"%s\", \"instr_per_step\": %d, \"cycles_per_step_per_smsp\": %.3f, \"cycles_per_instr\": %.3f}\n",
           name, per_iter, per * per_iter, per);
}

int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    float* d; cudaMalloc(&d, (size_t)p.multiProcessorCount * 512 * 4);
    long long* dclk; cudaMalloc(&dclk, 8);
    const int S = p.multiProcessorCount;
    run<0>("FFMA2 d=d*a+b (3 register pairs)", 1, d, dclk, S);
    run<1>("FFMA2 d=a*b+d (3 register pairs)", 1, d, dclk, S);
    run<2>("FFMA2 d=d*U+U (2 uniform)", 1, d, dclk, S);
    run<3>("FFMA2 d=d*U+b (1 uniform)", 1, d, dclk, S);
    run<4>("FFMA2 d=d*a0+b0 (repeated operands)", 1, d, dclk, S);
    run<5>("FADD2", 1, d, dclk, S);
    run<6>("FMUL2", 1, d, dclk, S);
    run<7>("2 x FFMA (register operands)", 2, d, dclk, S);
    run<8>("FADD2 + FFMA2(1 uniform)", 2, d, dclk, S);
    run<9>("FADD2 + 2 x FFMA", 3, d, dclk, S);
    run<10>("FADD2 + FFMA2 (3 register pairs)", 2, d, dclk, S);
    run<11>("2 x FFMA (uniform multiplier)", 2, d, dclk, S);
    run<12>("2 x FADD", 2, d, dclk, S);
    return 0;
}
