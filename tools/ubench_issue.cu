// Micro-benchmark: does a packed f32x2 instruction (2 FMA-pipe cycles) block the SMSP's issue port for its
// second cycle, or can an ALU / LSU instruction issue in its shadow?  Cycles per loop body for mixes of
// K packed FADD2 (independent chains) and M LOP3 / M LDS per FADD2, at 1..4 warps per SMSP.
#include <cstdio>
#include <cuda_runtime.h>
template <int M_ALU, int M_LDS>
__global__ void __launch_bounds__(512, 1) k(float* out, int iters, long long* cyc) {
    __shared__ float sm[4096];
    for (int i = threadIdx.x; i < 4096; i += blockDim.x) sm[i] = i;
    __syncthreads();
    float2 x[8];
    unsigned a[8];
    float l[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) { x[i] = make_float2(threadIdx.x + i, i); a[i] = threadIdx.x * 7 + i; l[i] = 0.f; }
    const float2 B = make_float2(0.5f, 0.25f);
    const float* p = sm + threadIdx.x;
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            x[i] = __fadd2_rn(x[i], B);
#pragma unroll
            for (int m = 0; m < M_ALU; ++m) a[i] = (a[i] ^ 0x5bd1e995u) & (a[(i + 1 + m) & 7] | 0x10u);
#pragma unroll
            for (int m = 0; m < M_LDS; ++m) l[i] += p[((it + i + m) & 63) * 32];
        }
    }
    long long t1 = clock64();
    float s = 0; unsigned u = 0;
    for (int i = 0; i < 8; ++i) { s += x[i].x + x[i].y + l[i]; u ^= a[i]; }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s + u;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
template <int A, int L> void run(float* d, long long* dc, int warps) {
    const int iters = 4000;
    k<A, L><<<148, warps * 32>>>(d, 10, dc);
    k<A, L><<<148, warps * 32>>>(d, iters, dc);
    long long c; cudaMemcpy(&c, dc, 8, cudaMemcpyDeviceToHost);
    double per = (double)c / iters / 8.0;                 // cycles per (FADD2 + extras) seen by one warp
    printf("FADD2 + %d LOP3-pair + %d LDS, %2d warps/SM (%d/SMSP): %.2f cycles per group per warp -> %.2f cycles per group per SMSP\n",
           A, L, warps, warps / 4, per, per / (warps / 4));
}
int main() {
    float* d; long long* dc; cudaMalloc(&d, 148 * 512 * 4); cudaMalloc(&dc, 8);
    for (int w : {4, 8, 16}) {
        run<0, 0>(d, dc, w); run<1, 0>(d, dc, w); run<2, 0>(d, dc, w); run<0, 1>(d, dc, w); run<1, 1>(d, dc, w);
    }
    return 0;
}
