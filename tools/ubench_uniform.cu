// When does ptxas (CUDA 12.9, sm_100a) keep warp-uniform values in the uniform datapath?  Compile-only probe: every
// variant runs the same inner loop -- packed FMAs whose multiplier comes from constant memory at an index that depends
// only on the loop counter p -- under a different control structure; tools/uniform_probe.sh counts, per variant,
// the FFMA2 that take a uniform-register operand (`FFMA2 R, R, UR.F32x2, R`, fed by LDCU) against those that take a
// vector register (fed by LDC with a register index).  The findings shaped k_frames_to_statics_u (fe_kernels.cuh).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -cubin -o ubench_uniform.cubin tools/ubench_uniform.cu
#include <cuda_runtime.h>
#include <cstdint>
__constant__ float4 c_tab[128];

__device__ __forceinline__ bool mbar_try(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ void body(int p, float2& acc, float2& b) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const float4 w = c_tab[p * 16 + j];
        asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(*(unsigned long long*)&acc) : "l"(*(unsigned long long*)&b), "l"(*(const unsigned long long*)&w.x));
        asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(*(unsigned long long*)&b) : "l"(*(unsigned long long*)&acc), "l"(*(const unsigned long long*)&w.z));
    }
}
// V: 0 mask from threadIdx >> 5, constant-bound loop          1 mask from __shfl_sync           2 mask from %warpid
//    3 tile loop that starts at a warp-derived index           4 uniform tile loop, plain guard  5 uniform tile loop, vote guard
//    6 + mbarrier spin with a per-thread exit                  7 + mbarrier spin that exits on vote.all
//    8 tcgen05.st under the warp-derived branch                9 tcgen05.st outside any non-uniform branch
template <int V>
__global__ void __launch_bounds__(256, 1) k(float* out, int n_tiles) {
    __shared__ unsigned long long bars[8];
    int warp = threadIdx.x >> 5;
    if (V == 1) warp = __shfl_sync(0xffffffffu, warp, 0);
    if (V == 2) { unsigned w; asm("mov.u32 %0, %%warpid;" : "=r"(w)); warp = (int)w; }
    const int group = warp & 3, half = warp >> 2;
    const unsigned mine = half ? 0x2au : 0xd5u;
    const uint32_t bar = (uint32_t)__cvta_generic_to_shared(&bars[group]);
    float2 acc = make_float2(threadIdx.x, 1.f), b = make_float2(2.f, threadIdx.x);
    uint32_t phase = 0;
    auto pairs = [&]() {
#pragma unroll 1
        for (int p = 0; p < 8; ++p) {
            if (V == 9) asm volatile("tcgen05.st.sync.aligned.32x32b.x2.b32 [%0], {%1, %2};" :: "r"((uint32_t)(64 * p)), "f"(acc.x), "f"(b.x) : "memory");
            if (!((mine >> p) & 1u)) continue;
            body(p, acc, b);
            if (V == 8) asm volatile("tcgen05.st.sync.aligned.32x32b.x2.b32 [%0], {%1, %2};" :: "r"((uint32_t)(64 * p)), "f"(acc.x), "f"(b.x) : "memory");
        }
    };
    if (V <= 2 || V >= 8) pairs();
    if (V == 3) for (int t = blockIdx.x * 4 + group; t < n_tiles; t += gridDim.x * 4) pairs();
    if (V >= 4 && V <= 7) {
        const int n_iter = (n_tiles + gridDim.x * 4 - 1) / (gridDim.x * 4);
#pragma unroll 1
        for (int it = 0; it < n_iter; ++it) {
            const int t = (blockIdx.x + it * gridDim.x) * 4 + group;
            const bool live = V == 4 ? (t < n_tiles) : __all_sync(0xffffffffu, t < n_tiles);
            if (!live) continue;
            if (V == 6) { while (!mbar_try(bar, phase & 1)) __nanosleep(32); phase ^= 1; }
            if (V == 7) { while (!__all_sync(0xffffffffu, mbar_try(bar, phase & 1))) __nanosleep(32); phase ^= 1; }
            pairs();
        }
    }
    out[threadIdx.x + blockIdx.x * 256] = acc.x + acc.y + b.x + b.y;
}
template __global__ void k<0>(float*, int);
template __global__ void k<1>(float*, int);
template __global__ void k<2>(float*, int);
template __global__ void k<3>(float*, int);
template __global__ void k<4>(float*, int);
template __global__ void k<5>(float*, int);
template __global__ void k<6>(float*, int);
template __global__ void k<7>(float*, int);
template __global__ void k<8>(float*, int);
template __global__ void k<9>(float*, int);
