// Per-warp phases of the frames->statics kernel (K1), written as host/device
// functions so that tests/host_sim can replay the exact lane-level dataflow on a
// CPU (lanes as a loop, shared memory as an array) before anything runs on a GPU.
//
// Replaces, per frame, what speechpy does for /root/reference/preprocess.py:72-82:
//   stack_frames (400/160, rectangular or table window) -> rfft(512) -> |X|^2/512
//   -> frame energy -> mel filterbank -> zero_handling -> log -> DCT-II(ortho)[:D]
//   -> c0 <- log(energy).
//
// Geometry: one frame = 8 lanes (t = 0..7), one warp = 4 consecutive frames (fs = 0..3).
// The 512-point real FFT is a 256-point complex FFT (z[m] = x[2m] + i x[2m+1]) split
// 16 x 16.  Every lane runs TWO 16-point FFTs per stage, carried in the two halves of
// packed f32x2 registers (Blackwell FADD2 / FMUL2 / FFMA2: one issue slot, two lanes
// of FP32 work; negate / half-swap / broadcast operand modifiers are free):
//   stage A  halves = complex points j = t and t+8 (m = j + 16 a, a = 0..15)
//   exchange through a swizzled shared-memory region (scalar stores, 16-byte loads)
//   stage B  halves = output rows k1 = rx and ry = 16 - rx, so Z[k] and Z[256-k] of
//            the real-FFT split sit in the SAME lane (half-swapped operand) and no
//            second exchange is needed.
#pragma once
#include <stdint.h>
#include <math.h>
#include <type_traits>
#include "fe_plans_gen.h"

#if defined(__CUDACC__)
#define FE_HD __host__ __device__ __forceinline__
#else
#define FE_HD inline
struct float2 { float x, y; };
struct float4 { float x, y, z, w; };
struct uint2 { unsigned x, y; };
struct uint4 { unsigned x, y, z, w; };
static inline uint4 make_uint4(unsigned a, unsigned b, unsigned c, unsigned d) { uint4 v = {a, b, c, d}; return v; }
static inline float4 make_float4(float a, float b, float c, float d) { float4 v = {a, b, c, d}; return v; }
static inline float2 make_float2(float a, float b) { float2 v = {a, b}; return v; }
#endif

namespace fe {

constexpr int kNfft = 512;
constexpr int kBins = 257;
constexpr int kWarpFrames = 4;          // frames per warp pass
constexpr int kTileFrames = 32;         // frames per tile: one lane per frame in the epilogue
constexpr int kTileGroups = kTileFrames / kWarpFrames;   // 4-frame groups (one FFT warp pass each) per tile
constexpr int kFftWarps = 12;           // producer warps of a K1 CTA (three warpgroups)
constexpr int kEpiWarps = 4;            // consumer warps (one warpgroup): mel / log / DCT for whole tiles
constexpr int kERegion = 512;           // floats of exchange buffer per frame (2 KB, 2 KB aligned)
constexpr int kPStride = 36;            // floats per bin row of the CTA's power buffer: [bin][frame], stride 36 makes
                                        // both the post-pass scatter (8 bins x 4 frames per store) and the
                                        // epilogue's column reads (32 frames of one bin) bank-conflict free
constexpr int kMaxFilters = 128;
constexpr float kEpsF64 = 2.220446049250313e-16f;   // np.finfo(float).eps, as float

// ---------------------------------------------------------------------------
// packed f32x2 arithmetic
// ---------------------------------------------------------------------------
#if defined(__CUDA_ARCH__)
FE_HD float2 padd(float2 a, float2 b) { return __fadd2_rn(a, b); }
FE_HD float2 psub(float2 a, float2 b) { return __fadd2_rn(a, make_float2(-b.x, -b.y)); }
FE_HD float2 pmul(float2 a, float2 b) { return __fmul2_rn(a, b); }
FE_HD float2 pfma(float2 a, float2 b, float2 c) { return __ffma2_rn(a, b, c); }
#else
FE_HD float2 padd(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
FE_HD float2 psub(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }
FE_HD float2 pmul(float2 a, float2 b) { return make_float2(a.x * b.x, a.y * b.y); }
FE_HD float2 pfma(float2 a, float2 b, float2 c) { return make_float2(fmaf(a.x, b.x, c.x), fmaf(a.y, b.y, c.y)); }
#endif
// FE_FMA_MODE (measured with tools/ubench_issue2.cu): an FFMA2 whose three operands are
// distinct register pairs occupies the issue port longer than two scalar FFMAs; 1 = such FMAs are issued
// as scalar pairs, 2 = every FMA is, 0 = everything packed.
#ifndef FE_FMA_MODE
#define FE_FMA_MODE 0
#endif
FE_HD float2 pfma_s(float2 a, float2 b, float2 c) { return make_float2(fmaf(a.x, b.x, c.x), fmaf(a.y, b.y, c.y)); }
// all three operands live in registers (table twiddles, squares)
FE_HD float2 pfma_rr(float2 a, float2 b, float2 c) { return FE_FMA_MODE >= 1 ? pfma_s(a, b, c) : pfma(a, b, c); }
// one operand is a compile-time constant broadcast to both halves
FE_HD float2 pfma_k(float2 a, float k, float2 c) { return FE_FMA_MODE >= 2 ? make_float2(fmaf(a.x, k, c.x), fmaf(a.y, k, c.y)) : pfma(a, make_float2(k, k), c); }
FE_HD float2 pbc(float s) { return make_float2(s, s); }
FE_HD float2 pswap(float2 a) { return make_float2(a.y, a.x); }
FE_HD float2 pneg(float2 a) { return make_float2(-a.x, -a.y); }
FE_HD float2 pnfma(float2 a, float2 b, float2 c) { return pfma(pneg(a), b, c); }   // c - a*b
FE_HD float2 pmul_k(float2 a, float k) { return pmul(a, make_float2(k, k)); }
// scalar twins: the lane = frame kernel (fe_k1t.cuh) runs the same codelets on plain floats -- its twiddles are
// warp-uniform (constant-bank / uniform-register operands), and a scalar FFMA with such an operand issues every cycle
FE_HD float padd(float a, float b) { return a + b; }
FE_HD float psub(float a, float b) { return a - b; }
FE_HD float pmul(float a, float b) { return a * b; }
FE_HD float pfma(float a, float b, float c) { return fmaf(a, b, c); }
FE_HD float pfma_rr(float a, float b, float c) { return fmaf(a, b, c); }
FE_HD float pfma_k(float a, float k, float c) { return fmaf(a, k, c); }
FE_HD float pmul_k(float a, float k) { return a * k; }
FE_HD float pneg(float a) { return -a; }

// ---------------------------------------------------------------------------
// Shared-memory tables of one CTA
// ---------------------------------------------------------------------------
struct SmemTables {
    // twiddles are generated from a few per-lane bases (shared-memory wavefronts are the scarce
    // resource of this kernel, packed FP32 issue slots are not):
    // twiddles come straight from tables: K1 is bound by instruction issue (a packed f32x2 instruction takes two
    // issue cycles), so one 16-byte load per twiddle pair beats regenerating it from a base with four packed FMAs
    const float4* tw256;      // [15][16] row j-1, cfg = flip*8 + t: (wr(j rx), wr(j ry), wi(j rx), wi(j ry)) of W_256^(j row), applied by stage B
    const float4* tw512;      // [8][16]  row k2, cfg = flip*8 + t: (cos, cos, sin, sin) of 2 pi (rx + 16 k2) / 512 and (ry + 16 k2)
    const float2* window;     // [ROWS*16] (w[2m], w[2m+1]) per complex point m, or nullptr
    // mel plan: every epilogue warp streams its own flat list of 4-bin weight groups (its filters back to
    // back, runs padded with zero weights to a multiple of 4): mel_desc[i] = first bin | last-group-of-
    // filter << 10 | filter << 16, weights mel_w4[i].  A warp's slice is [off, off + cnt) (+ 8 zero groups of
    // prefetch slack).  lane = frame, so descriptors and weights are warp-uniform loads.
    const int*    mel_desc;
    const float4* mel_w4;         // weight groups, pre-scaled by pscale / 2048
    const float*  dctf;           // [D][dct_stride] folded DCT rows (n < ceil(nf/2)), zero padded to dct_stride
    int nf, D, dct_stride, nh;    // nh = ceil(nf / 2), dct_stride = nh rounded up to 4
    int full_spectrum;        // filterbank touches bins > 128
    int is_mfcc, fbank_log, dc_elim;
    float pscale;             // 2^-30 when samples are raw int16 counts, 1 for float PCM
};

// ---------------------------------------------------------------------------
// radix-4 butterfly and 16-point forward FFT on packed registers
// ---------------------------------------------------------------------------
template <class V>
FE_HD void bfly4(V& r0, V& i0, V& r1, V& i1, V& r2, V& i2, V& r3, V& i3) {
    V t0r = padd(r0, r2), t0i = padd(i0, i2);
    V t1r = psub(r0, r2), t1i = psub(i0, i2);
    V t2r = padd(r1, r3), t2i = padd(i1, i3);
    V t3r = psub(r1, r3), t3i = psub(i1, i3);
    r0 = padd(t0r, t2r); i0 = padd(t0i, t2i);
    r2 = psub(t0r, t2r); i2 = psub(t0i, t2i);
    r1 = padd(t1r, t3i); i1 = psub(t1i, t3r);     // t1 - i t3
    r3 = psub(t1r, t3i); i3 = padd(t1i, t3r);     // t1 + i t3
}

// bfly4 with a zero fourth input (the zero padding of the frame: rows 13..15 of stage A)
template <class V>
FE_HD void bfly4_z3(V& r0, V& i0, V& r1, V& i1, V& r2, V& i2, V& r3, V& i3) {
    V t0r = padd(r0, r2), t0i = padd(i0, i2);
    V t1r = psub(r0, r2), t1i = psub(i0, i2);
    r0 = padd(t0r, r1); i0 = padd(t0i, i1);           // t2 = t3 = x1
    r2 = psub(t0r, r1); i2 = psub(t0i, i1);
    V t3r = r1, t3i = i1;
    r1 = padd(t1r, t3i); i1 = psub(t1i, t3r);         // t1 - i t3
    r3 = psub(t1r, t3i); i3 = padd(t1i, t3r);         // t1 + i t3
}

// In-place FFT16: input natural order x[n]; output X[k] lands at slot pos16(k).
FE_HD constexpr int pos16(int k) { return (k >> 2) + ((k & 3) << 2); }

// (ar, ai) = (r, i) * (wr + i wi)                       2 FMUL2 + 2 FFMA2
template <class V>
FE_HD void cmul_c(V r, V i, float wr, float wi, V& ar, V& ai) {
    ar = pfma_k(i, -wi, pmul_k(r, wr));
    ai = pfma_k(r, wi, pmul_k(i, wr));
}
// (tr, ti) = (br, bi) + (r, i) * (wr + i wi)            4 FFMA2: the product is never materialised
template <class V>
FE_HD void cmac_c(V r, V i, float wr, float wi, V br, V bi, V& tr, V& ti) {
    tr = pfma_k(r, wr, pfma_k(i, -wi, br));
    ti = pfma_k(r, wi, pfma_k(i, wr, bi));
}
// same with per-half twiddles (packed operands)
template <class V>
FE_HD void cmul_p(V r, V i, V wr, V wi, V& ar, V& ai) {
    ar = pfma_rr(pneg(i), wi, pmul(r, wr));
    ai = pfma_rr(r, wi, pmul(i, wr));
}
template <class V>
FE_HD void cmac_p(V r, V i, V wr, V wi, V br, V bi, V& tr, V& ti) {
    tr = pfma_rr(r, wr, pfma_rr(pneg(i), wi, br));
    ti = pfma_rr(r, wi, pfma_rr(i, wr, bi));
}
template <class V>
FE_HD V ptwice_minus(V a, V t) { return pfma_k(a, 2.f, pneg(t)); }     // 2 a - t

// second half of a radix-4 butterfly: (t0, t1, t2, t3) -> outputs
template <class V>
FE_HD void bfly4_out(V t0r, V t0i, V t1r, V t1i, V t2r, V t2i, V t3r, V t3i,
                     V& r0, V& i0, V& r1, V& i1, V& r2, V& i2, V& r3, V& i3) {
    r0 = padd(t0r, t2r); i0 = padd(t0i, t2i);
    r2 = psub(t0r, t2r); i2 = psub(t0i, t2i);
    r1 = padd(t1r, t3i); i1 = psub(t1i, t3r);     // t1 - i t3
    r3 = psub(t1r, t3i); i3 = padd(t1i, t3r);     // t1 + i t3
}

// Second level of the FFT16 with the W_16^(n1 k2) twiddles folded into the butterflies' additions (the
// kernel is bound by instruction issue, so a multiply that can ride on the following add as an FMA is free):
//   x0 + w x2 and x0 - w x2 with w = +-H(1 -+ i) cost 2 adds + 4 FMAs instead of 2 adds + 2 muls + 4 adds,
//   a = w1 x1 (4 ops), t2 = a + w3 x3 (4 FMAs), t3 = 2 a - t2 (2 FMAs) instead of 4 + 4 + 4.
// 84 packed operations for the level instead of 96.
template <class V>
FE_HD void fft16_level2(V (&xr)[16], V (&xi)[16]) {
    constexpr float C1 = 0.92387953251128674f, S1 = 0.38268343236508977f, H = 0.70710678118654752f;
    bfly4(xr[0], xi[0], xr[1], xi[1], xr[2], xi[2], xr[3], xi[3]);                    // k2 = 0: no twiddles
    {   // k2 = 1: x1 W^1, x2 W^2 = H (1 - i), x3 W^3
        const V s2 = padd(xr[6], xi[6]), d2 = psub(xi[6], xr[6]);
        const V t0r = pfma_k(s2, H, xr[4]), t0i = pfma_k(d2, H, xi[4]);
        const V t1r = pfma_k(s2, -H, xr[4]), t1i = pfma_k(d2, -H, xi[4]);
        V ar, ai, t2r, t2i;
        cmul_c(xr[5], xi[5], C1, -S1, ar, ai);
        cmac_c(xr[7], xi[7], S1, -C1, ar, ai, t2r, t2i);
        const V t3r = ptwice_minus(ar, t2r), t3i = ptwice_minus(ai, t2i);
        bfly4_out(t0r, t0i, t1r, t1i, t2r, t2i, t3r, t3i, xr[4], xi[4], xr[5], xi[5], xr[6], xi[6], xr[7], xi[7]);
    }
    {   // k2 = 2: x1 W^2 = H (s1, d1), x2 W^4 = -i x2, x3 W^6 = H (d3, -s3); the H rides on the output additions
        const V s1 = padd(xr[9], xi[9]), d1 = psub(xi[9], xr[9]);
        const V s3 = padd(xr[11], xi[11]), d3 = psub(xi[11], xr[11]);
        const V t0r = padd(xr[8], xi[10]), t0i = psub(xi[8], xr[10]);
        const V t1r = psub(xr[8], xi[10]), t1i = padd(xi[8], xr[10]);
        const V ur = padd(s1, d3), ui = psub(d1, s3), vr = psub(s1, d3), vi = padd(d1, s3);
        xr[8] = pfma_k(ur, H, t0r);   xi[8] = pfma_k(ui, H, t0i);
        xr[10] = pfma_k(ur, -H, t0r); xi[10] = pfma_k(ui, -H, t0i);
        xr[9] = pfma_k(vi, H, t1r);   xi[9] = pfma_k(vr, -H, t1i);
        xr[11] = pfma_k(vi, -H, t1r); xi[11] = pfma_k(vr, H, t1i);
    }
    {   // k2 = 3: x1 W^3, x2 W^6 = H (d2, -s2), x3 W^9
        const V s2 = padd(xr[14], xi[14]), d2 = psub(xi[14], xr[14]);
        const V t0r = pfma_k(d2, H, xr[12]), t0i = pfma_k(s2, -H, xi[12]);
        const V t1r = pfma_k(d2, -H, xr[12]), t1i = pfma_k(s2, H, xi[12]);
        V ar, ai, t2r, t2i;
        cmul_c(xr[13], xi[13], S1, -C1, ar, ai);
        cmac_c(xr[15], xi[15], -C1, S1, ar, ai, t2r, t2i);
        const V t3r = ptwice_minus(ar, t2r), t3i = ptwice_minus(ai, t2i);
        bfly4_out(t0r, t0i, t1r, t1i, t2r, t2i, t3r, t3i, xr[12], xi[12], xr[13], xi[13], xr[14], xi[14], xr[15], xi[15]);
    }
}

// NZ: inputs n >= NZ are known to be zero (12 <= NZ <= 16 supported: only the fourth butterfly input can vanish)
template <int NZ = 16, class V>
FE_HD void fft16(V (&xr)[16], V (&xi)[16]) {
    static_assert(NZ >= 12 && NZ <= 16, "fft16: only the last four inputs may be structurally zero");
#pragma unroll
    for (int n1 = 0; n1 < 4; ++n1) {
        if (n1 + 12 >= NZ) bfly4_z3(xr[n1], xi[n1], xr[n1 + 4], xi[n1 + 4], xr[n1 + 8], xi[n1 + 8], xr[n1 + 12], xi[n1 + 12]);
        else bfly4(xr[n1], xi[n1], xr[n1 + 4], xi[n1 + 4], xr[n1 + 8], xi[n1 + 8], xr[n1 + 12], xi[n1 + 12]);
    }
    fft16_level2(xr, xi);       // slot n1 + 4 k2 holds A[n1][k2]; twiddle W_16^(n1 k2)
}

// FFT16 of x[n] * w[n] (w[0] = 1): the input twiddles are folded into the first level's additions.
//   a0 = x0 w0 (4 ops, none for n1 = 0), t0 = a0 + x2 w2 (4 FMAs), t1 = 2 a0 - t0 (2), same for (x1, x3):
//   24 / 28 operations per butterfly instead of 28 / 32 with separate complex multiplies.
// tw(n) returns (wr, wi) of input n as packed pairs (one twiddle per half).
template <class V, class TW>
FE_HD void fft16_twiddled(V (&xr)[16], V (&xi)[16], TW&& tw) {
#pragma unroll
    for (int n1 = 0; n1 < 4; ++n1) {
        V a0r = xr[n1], a0i = xi[n1], wr, wi;
        if (n1 > 0) { tw(n1, wr, wi); cmul_p(xr[n1], xi[n1], wr, wi, a0r, a0i); }
        V t0r, t0i, a1r, a1i, t2r, t2i;
        tw(n1 + 8, wr, wi);  cmac_p(xr[n1 + 8], xi[n1 + 8], wr, wi, a0r, a0i, t0r, t0i);
        const V t1r = ptwice_minus(a0r, t0r), t1i = ptwice_minus(a0i, t0i);
        tw(n1 + 4, wr, wi);  cmul_p(xr[n1 + 4], xi[n1 + 4], wr, wi, a1r, a1i);
        tw(n1 + 12, wr, wi); cmac_p(xr[n1 + 12], xi[n1 + 12], wr, wi, a1r, a1i, t2r, t2i);
        const V t3r = ptwice_minus(a1r, t2r), t3i = ptwice_minus(a1i, t2i);
        bfly4_out(t0r, t0i, t1r, t1i, t2r, t2i, t3r, t3i,
                  xr[n1], xi[n1], xr[n1 + 4], xi[n1 + 4], xr[n1 + 8], xi[n1 + 8], xr[n1 + 12], xi[n1 + 12]);
    }
    fft16_level2(xr, xi);
}

// ---------------------------------------------------------------------------
// Exchange region of one frame (512 floats, 2 KB aligned).  Element (row k1, column j,
// plane re/im) lives at float index
//     p*64 + plane*32 + 4*((j>>1) ^ p ^ X) + 2*(j&1) + (slot ^ F)
// p = pair-row of k1 ({8,0} -> 0, {k,16-k} -> k), slot = which row of the pair,
// X = 4*(fs&1) and F = (fs>>1 when p != 0) spread the four frames of a warp over the
// banks.  All lane-dependent terms are XORs into the low 5 index bits, so an access is
// "lane base XOR compile-time constant" plus an immediate offset.
// ---------------------------------------------------------------------------
FE_HD constexpr int e_prow(int k1) { return (k1 == 0 || k1 == 8) ? 0 : (k1 < 8 ? k1 : 16 - k1); }
FE_HD constexpr int e_slot(int k1) { return (k1 == 0) ? 1 : (k1 == 8 ? 0 : (k1 < 8 ? 0 : 1)); }

// XOR-addressable pointer into a 2 KB aligned shared region
struct EPtr {
#if defined(__CUDA_ARCH__)
    uint32_t a;                              // shared-window byte address | lane bits
    __device__ __forceinline__ EPtr x(int c) const { EPtr r; r.a = a ^ (uint32_t)(c << 2); return r; }
#else
    float* base; int l;
    EPtr x(int c) const { EPtr r; r.base = base; r.l = l ^ c; return r; }
#endif
};

#if defined(__CUDA_ARCH__)
__device__ __forceinline__ void e_st(EPtr p, int off, float v) {
    asm volatile("st.shared.f32 [%0], %1;" :: "r"(p.a + (uint32_t)(off << 2)), "f"(v) : "memory");
}
__device__ __forceinline__ float4 e_ld4(EPtr p, int off) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
                 : "r"(p.a + (uint32_t)(off << 2)) : "memory");
    return v;
}
__device__ __forceinline__ EPtr e_make(const float* region, int lane_bits) {
    EPtr p; p.a = (uint32_t)__cvta_generic_to_shared(region) | (uint32_t)(lane_bits << 2); return p;
}
#else
inline void e_st(EPtr p, int off, float v) { p.base[p.l + off] = v; }
inline float4 e_ld4(EPtr p, int off) { const float* q = p.base + p.l + off; return make_float4(q[0], q[1], q[2], q[3]); }
inline EPtr e_make(float* region, int lane_bits) { EPtr p; p.base = region; p.l = lane_bits; return p; }
#endif

// ---------------------------------------------------------------------------
// Phase 1 (stage A)
//   raw_f : this frame's samples as staged by the bulk copy (int16 words or floats)
//   e_f   : this frame's exchange region
//   halves: .x = complex point jx = t + 8*swap, .y = jy = t + 8*(1-swap), swap = fs>>1
//           (the swap makes the two 4-byte loads of a row bank-conflict free)
//   returns the lane's partial sum of squares (Parseval frame energy, sample units)
// ---------------------------------------------------------------------------
template <int FRAME_LEN, int IN_F32, int HAS_WINDOW>
FE_HD float stage_a(const void* raw_f, float* e_f, const SmemTables& tb, int t, int fs) {
    float2 re[16], im[16];
    float2 ss = make_float2(0.f, 0.f);
    constexpr int ROWS = (FRAME_LEN + 31) / 32;
    const int swap = fs >> 1;
    const int jx = t + 8 * swap, jy = t + 8 * (1 - swap);
#pragma unroll
    for (int a = 0; a < 16; ++a) {
        if (a < ROWS) {
            float2 vr, vi;
            if (IN_F32) {
                const float2* pf = reinterpret_cast<const float2*>(raw_f);
                float2 u = pf[16 * a + jx], v = pf[16 * a + jy];
                vr = make_float2(u.x, v.x); vi = make_float2(u.y, v.y);
            } else {
                const uint32_t* pw = reinterpret_cast<const uint32_t*>(raw_f);
                uint32_t u = pw[16 * a + jx], v = pw[16 * a + jy];
#if defined(__CUDA_ARCH__) && defined(FE_I2F_HALFSEL)
                // both halves straight from the 32-bit word (I2F.S16 with a half selector: no shift instruction)
                asm("{\n\t.reg .b16 lo, hi;\n\tmov.b32 {lo, hi}, %2;\n\tcvt.rn.f32.s16 %0, lo;\n\tcvt.rn.f32.s16 %1, hi;\n\t}"
                    : "=f"(vr.x), "=f"(vi.x) : "r"(u));
                asm("{\n\t.reg .b16 lo, hi;\n\tmov.b32 {lo, hi}, %2;\n\tcvt.rn.f32.s16 %0, lo;\n\tcvt.rn.f32.s16 %1, hi;\n\t}"
                    : "=f"(vr.y), "=f"(vi.y) : "r"(v));
#else
                vr = make_float2((float)(short)(u & 0xffffu), (float)(short)(v & 0xffffu));
                vi = make_float2((float)((int)u >> 16), (float)((int)v >> 16));
#endif
            }
            if (HAS_WINDOW) {
                float2 w0 = tb.window[16 * a + jx], w1 = tb.window[16 * a + jy];
                vr = pmul(vr, make_float2(w0.x, w1.x)); vi = pmul(vi, make_float2(w0.y, w1.y));
            }
            if (32 * a + 32 > FRAME_LEN) {      // zero padding inside the last row (sample n = 2 (j + 16 a) + c)
                const int nx = 2 * (jx + 16 * a), ny = 2 * (jy + 16 * a);
                if (nx >= FRAME_LEN) vr.x = 0.f;
                if (nx + 1 >= FRAME_LEN) vi.x = 0.f;
                if (ny >= FRAME_LEN) vr.y = 0.f;
                if (ny + 1 >= FRAME_LEN) vi.y = 0.f;
            }
            re[a] = vr; im[a] = vi;
            ss = pfma_rr(vr, vr, ss); ss = pfma_rr(vi, vi, ss);
        } else {
            re[a] = make_float2(0.f, 0.f); im[a] = make_float2(0.f, 0.f);
        }
    }
    fft16<(ROWS > 12 ? ROWS : 16)>(re, im);
    // scatter (the W_256^(j k1) twiddles are applied by stage B, folded into its first butterflies).
    // Lane part of the index (low 5 bits):
    //   4*((t>>1) ^ X ^ 4*hb) + 2*(t&1) + F   with hb = j>>3 of the half being stored
    const int X = (fs & 1) << 2, F = fs >> 1;
    const int lb = (((t >> 1) ^ X) << 2) | ((t & 1) << 1);
    const int hbx = swap << 4, hby = (1 - swap) << 4;           // chunk bit 2 = index bit 4
    const EPtr x0 = e_make(e_f, lb ^ hbx), y0 = e_make(e_f, lb ^ hby);              // pair-row 0: no slot flip
    const EPtr xF = e_make(e_f, (lb ^ hbx) | F), yF = e_make(e_f, (lb ^ hby) | F);
#pragma unroll
    for (int k1 = 0; k1 < 16; ++k1) {
        const int s = pos16(k1);
        const float2 yr = re[s], yi = im[s];
        const int p = e_prow(k1), c = (p << 2) | e_slot(k1), off = p * 64;
        const EPtr qx = (p == 0 ? x0 : xF).x(c), qy = (p == 0 ? y0 : yF).x(c);
        e_st(qx, off, yr.x);
        e_st(qy, off, yr.y);
        e_st(qx, off + 32, yi.x);
        e_st(qy, off + 32, yi.y);
    }
    return ss.x + ss.y;
}

// ---------------------------------------------------------------------------
// Phase 2 (stage B): pair-row p = t of the exchange region -> packed FFT16 over j.
// Halves: .x = row rx, .y = row ry; slot pos16(k2) holds Z[r + 16 k2].
// ---------------------------------------------------------------------------
struct LaneZ { float2 r[16], i[16]; };

FE_HD int lane_flip(int t, int fs) { return t ? (fs >> 1) : 0; }
FE_HD int row_x(int t, int fs) { return t == 0 ? 8 : (lane_flip(t, fs) ? 16 - t : t); }
FE_HD int row_y(int t, int fs) { return t == 0 ? 0 : (lane_flip(t, fs) ? t : 16 - t); }

FE_HD void stage_b(float* e_f, LaneZ& z, const SmemTables& tb, int t, int fs) {
    const int X = (fs & 1) << 2;
    const EPtr b = e_make(e_f, t * 64 + ((t ^ X) << 2));
#pragma unroll
    for (int q = 0; q < 8; ++q) {
        float4 vr = e_ld4(b.x(q << 2), 0);
        float4 vi = e_ld4(b.x(q << 2), 32);
        z.r[2 * q] = make_float2(vr.x, vr.y); z.r[2 * q + 1] = make_float2(vr.z, vr.w);
        z.i[2 * q] = make_float2(vi.x, vi.y); z.i[2 * q + 1] = make_float2(vi.z, vi.w);
    }
    // column j carries W_256^(j row): rows (rx, ry) are a per-lane constant, the table row is the column
    const float4* tw = tb.tw256 + lane_flip(t, fs) * 8 + t;
    fft16_twiddled(z.r, z.i, [&](int j, float2& wr, float2& wi) {
        const float4 w = tw[(j - 1) * 16];
        wr = make_float2(w.x, w.y); wi = make_float2(w.z, w.w);
    });
}

// ---------------------------------------------------------------------------
// Phase 3 (post-pass): real-FFT split, power |2X|^2 (the 1/2048 = 1/(4*512) lives in
// the filterbank weights), scatter to this frame's column of the CTA's power buffer
// (p_f = pbuf + frame; bin k at p_f[k * kPStride]).
//   lanes t >= 1: partner of half .x is half .y of slot 15-k2 and vice versa (free swap)
//   lanes t == 0: rows 8 (.x) and 0 (.y) are self-paired
// ---------------------------------------------------------------------------
FE_HD void post_pass(const LaneZ& z, float* p_f, const SmemTables& tb, int t, int fs, float& x0, float& x256) {
    const bool t0 = (t == 0);
    const int rx = row_x(t, fs), ry = row_y(t, fs);
    const int cfg = lane_flip(t, fs) * 8 + t;
#pragma unroll
    for (int k2 = 0; k2 < 8; ++k2) {
        const int sa = pos16(k2), sp = pos16(15 - k2), sq = pos16((16 - k2) & 15);
        float2 ar = z.r[sa], ai = z.i[sa];
        // partner Z[256 - k]
        float2 pr, pi;
        pr.x = t0 ? z.r[sp].x : z.r[sp].y;   pr.y = t0 ? z.r[sq].y : z.r[sp].x;
        pi.x = t0 ? z.i[sp].x : z.i[sp].y;   pi.y = t0 ? z.i[sq].y : z.i[sp].x;
        const float4 wb = tb.tw512[k2 * 16 + cfg];             // angle of bins rx + 16 k2, ry + 16 k2
        const float2 c = make_float2(wb.x, wb.y), s = make_float2(wb.z, wb.w);
        float2 er = padd(ar, pr), ei = psub(ai, pi);          // 2E  (B = conj(partner))
        float2 orr = psub(ar, pr), oi = padd(ai, pi);         // 2O
        // T = i w O, w = (c, -s): Tr = s Or - c Oi, Ti = c Or + s Oi; 2 X[k] = 2E - T with the products folded
        // into the subtraction (2 FFMA2 per component instead of FMUL2 + FFMA2 + FADD2)
        float2 xr = pfma_rr(pneg(s), orr, pfma_rr(c, oi, er));
        float2 xi = pfma_rr(pneg(s), oi, pfma_rr(pneg(c), orr, ei));
        float2 plo = pfma_rr(xi, xi, pmul(xr, xr));
        p_f[(rx + 16 * k2) * kPStride] = plo.x;
        p_f[(ry + 16 * k2) * kPStride] = plo.y;
        if (tb.full_spectrum) {
            float2 yr = ptwice_minus(er, xr), yi = ptwice_minus(ei, xi);      // conj(2 X[256-k]) = 2E + T = 2 (2E) - 2X
            float2 phi = pfma_rr(yi, yi, pmul(yr, yr));
            p_f[(256 - (rx + 16 * k2)) * kPStride] = phi.x;
            p_f[(256 - (ry + 16 * k2)) * kPStride] = phi.y;   // lane 0, k2 = 0 writes bin 256
        }
    }
    // bin 128 = row 0, k2 = 8 (self-paired): 2 X[128] = 2 conj(Z[128]); lane 0 holds row 0 in .y
    float zr = z.r[pos16(8)].y, zi = z.i[pos16(8)].y;
    if (t0) p_f[128 * kPStride] = 4.f * (zr * zr + zi * zi);
    x0 = z.r[0].y + z.i[0].y;         // X[0]   = Re Z[0] + Im Z[0]   (lane 0 only)
    x256 = z.r[0].y - z.i[0].y;       // X[256] = Re Z[0] - Im Z[0]
}

// Parseval frame energy: sum_{k=0..256} |X_k|^2 / 512 = sum x^2 / 2 + (X0^2 + X256^2) / 1024
FE_HD float frame_energy(float sumsq, float x0, float x256, float pscale) {
    float e = (0.5f * sumsq + (x0 * x0 + x256 * x256) * (1.0f / 1024.0f)) * pscale;
    return e == 0.f ? kEpsF64 : e;
}

// natural log of a normal positive float (inputs are >= 2.2e-16 after zero handling, so the
// denormal pre-scaling of __logf is dead weight): lg2.approx * ln 2
FE_HD float fe_log(float x) {
#if defined(__CUDA_ARCH__)
    float r;
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r * 0.693147180559945309f;
#else
    return logf(x);
#endif
}

// ---------------------------------------------------------------------------
// Epilogue (the consumer warps of K1): lane = frame of the 32-frame tile, warp = filter / coefficient
// group.  Everything a warp reads besides its own frame column is warp-uniform (broadcast loads), every
// column access is 32 consecutive floats (one wavefront, no conflicts).
//   pbuf  [bins + 3][kPStride]   power columns written by the post-pass (pad rows stay zero)
//   lm    [nf + 4][32]           (log-)mel rows of the tile (mfcc only; fbank rows go straight to out_t)
//   out_t [D][32]                the tile's statics, coefficient-major (what K2 reads)
// ---------------------------------------------------------------------------
// Phase 4: mel filterbank (+ zero handling, + log).  The warp's weight groups are streamed through a
// 4-deep software pipeline (descriptors 8 groups ahead, weights and power values 4 ahead): a serial
// load -> FMA -> log chain per filter would leave the warp idle for hundreds of cycles per filter.
// The only loop-carried dependency is the running sum of the current filter.
FE_HD void epi_mel(const float* pbuf, float* dst, const SmemTables& tb, int off, int G, int lane) {
    const float* pcol = pbuf + lane;
    const bool want_log = tb.is_mfcc || tb.fbank_log;
    const int* ent = tb.mel_desc + off;
    const float4* w4 = tb.mel_w4 + off;
    int e[4], en[4];
    float4 w[4];
    float pv[4][4];
#pragma unroll
    for (int s = 0; s < 4; ++s) {
        e[s] = ent[s]; en[s] = ent[4 + s]; w[s] = w4[s];
        const float* p = pcol + (e[s] & 1023) * kPStride;
#pragma unroll
        for (int j = 0; j < 4; ++j) pv[s][j] = p[j * kPStride];
    }
    float acc = 0.f;
    for (int i = 0; i < G; i += 4) {
#pragma unroll
        for (int s = 0; s < 4; ++s) {
            const float a = fmaf(w[s].x, pv[s][0], w[s].z * pv[s][2]) + fmaf(w[s].y, pv[s][1], w[s].w * pv[s][3]);
            acc += a;
            const int ec = e[s];
            // refill this stage with group i + s + 4 (the list has 8 zero groups of slack)
            e[s] = en[s]; en[s] = ent[i + s + 8]; w[s] = w4[i + s + 4];
            const float* p = pcol + (e[s] & 1023) * kPStride;
#pragma unroll
            for (int j = 0; j < 4; ++j) pv[s][j] = p[j * kPStride];
            if (ec & 1024) {                                    // last group of a filter (warp-uniform)
                float v = acc == 0.f ? kEpsF64 : acc;           // speechpy.functions.zero_handling
                if (want_log) v = fe_log(v);
                dst[(ec >> 16) * 32 + lane] = v;
                acc = 0.f;
            }
        }
    }
}

// Phase 5 (mfcc): DCT-II (ortho), y_c = sum_{n < nh} C[c][n] (x[n] + (-1)^c x[nf-1-n]).  Warp computes
// coefficients warp, warp + 4, ... (all of one parity -> one folded input), four at a time so that every
// log-mel row is loaded once per four coefficients; c0 <- log(frame energy) when dc_elim.  Weights are
// warp-uniform.  N4 > 0: number of 4-row groups known at compile time (5 for the reference's 40 filters).
template <int N4>
FE_HD void epi_dct_n(const float* lm, const float* energies, float* out_t, const SmemTables& tb, int warp, int lane) {
    const int nh4 = (tb.nh + 3) & ~3, n4 = N4 > 0 ? N4 : (nh4 >> 2);
    const float* in = lm + lane;
    const float sgn = (warp & 1) ? -1.f : 1.f;
    const int nf1 = tb.nf - 1;
    for (int cb = warp; cb < tb.D; cb += 4 * kEpiWarps) {
        const float4* d[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int c = cb + j * kEpiWarps;
            d[j] = reinterpret_cast<const float4*>(tb.dctf + (c < tb.D ? c : cb) * tb.dct_stride);
        }
        float acc[4] = {0.f, 0.f, 0.f, 0.f}, bcc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int q = 0; q < n4; ++q) {
            float x[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int n = 4 * q + j, m = nf1 - n;           // rows past nh carry zero weights; keep the reads in range
                const float a = in[n * 32];
                const float b = in[(m > 0 ? m : 0) * 32];
                x[j] = (m == n) ? (sgn > 0.f ? a : 0.f) : fmaf(sgn, b, a);
            }
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float4 w = d[j][q];
                acc[j] = fmaf(w.x, x[0], acc[j]); bcc[j] = fmaf(w.y, x[1], bcc[j]);
                acc[j] = fmaf(w.z, x[2], acc[j]); bcc[j] = fmaf(w.w, x[3], bcc[j]);
            }
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int c = cb + j * kEpiWarps;
            float v = acc[j] + bcc[j];
            if (c == 0 && tb.dc_elim) v = fe_log(energies[lane]);
            if (c < tb.D) out_t[c * 32 + lane] = v;
        }
    }
}

FE_HD void epi_dct(const float* lm, const float* energies, float* out_t, const SmemTables& tb, int warp, int lane) {
    if (tb.nh == 20) epi_dct_n<5>(lm, energies, out_t, tb, warp, lane);
    else epi_dct_n<0>(lm, energies, out_t, tb, warp, lane);
}

// ---------------------------------------------------------------------------
// Specialised epilogue for the reference's two filterbanks (fe_plans_gen.h): ONE warp takes a whole tile,
// lane = frame.  The filterbank structure is a compile-time plan, so the whole mel -> log -> fold -> DCT
// chain is straight-line code over registers: one shared-memory load and one FMA per non-zero weight
// (weights and DCT rows are compile-time offsets into `w`, i.e. constant-bank operands on the device),
// no descriptors, no barrier, no log-mel round trip through shared memory.
//   w: [Plan::NNZ] mel weights in CSR order (pre-scaled), then [D][NF/2] folded DCT rows (mfcc)
// ---------------------------------------------------------------------------
template <int I, int N, class F>
FE_HD void static_for(F&& f) {
    if constexpr (I < N) {
        f(std::integral_constant<int, I>{});
        static_for<I + 1, N>(f);
    }
}

template <class Plan, int M, bool LOG, int PS = kPStride>
FE_HD float mel_spec(const float* pcol, const float* w) {
    constexpr int b0 = Plan::B0[M], n = Plan::N[M], off = Plan::OFF[M];
    float a0 = 0.f, a1 = 0.f;
    static_for<0, n>([&](auto ii) {
        constexpr int i = decltype(ii)::value;
        if (i & 1) a1 = fmaf(w[off + i], pcol[(b0 + i) * PS], a1);
        else a0 = fmaf(w[off + i], pcol[(b0 + i) * PS], a0);
    });
    float v = a0 + a1;
    v = v == 0.f ? kEpsF64 : v;                 // speechpy.functions.zero_handling
    if (LOG) v = fe_log(v);
    return v;
}

// Four filters at a time with two accumulators each: eight independent FMA chains (the epilogue runs in ONE warp per
// lane group; with the two chains of mel_spec it is eligible to issue less than half of the time).
template <class Plan, int M0, int M1, int M2, int M3, bool LOG, int PS = kPStride>
FE_HD void mel_spec4(const float* pcol, const float* w, float (&out)[4]) {
    constexpr int n0 = Plan::N[M0], n1 = Plan::N[M1], n2 = Plan::N[M2], n3 = Plan::N[M3];
    constexpr int n01 = n0 > n1 ? n0 : n1, n23 = n2 > n3 ? n2 : n3, nmax = n01 > n23 ? n01 : n23;
    float a0[4] = {0.f, 0.f, 0.f, 0.f}, a1[4] = {0.f, 0.f, 0.f, 0.f};
    static_for<0, nmax>([&](auto ii) {
        constexpr int i = decltype(ii)::value;
        static_for<0, 4>([&](auto jj) {
            constexpr int j = decltype(jj)::value;
            constexpr int m = j == 0 ? M0 : (j == 1 ? M1 : (j == 2 ? M2 : M3));
            constexpr int b0 = Plan::B0[m], nn = Plan::N[m], off = Plan::OFF[m];
            if constexpr (i < nn) {
                if (i & 1) a1[j] = fmaf(w[off + i], pcol[(b0 + i) * PS], a1[j]);
                else a0[j] = fmaf(w[off + i], pcol[(b0 + i) * PS], a0[j]);
            }
        });
    });
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        float v = a0[j] + a1[j];
        v = v == 0.f ? kEpsF64 : v;             // speechpy.functions.zero_handling
        if (LOG) v = fe_log(v);
        out[j] = v;
    }
}

// PS: row stride of the power buffer in floats; `energy`: this lane's (zero-handled) frame energy
// `lane` selects the column of the power buffer, `frame` the position inside the tile's [D][32] statics block
template <class Plan, int D, bool MFCC, bool LOG, int PS = kPStride>
FE_HD void epi_tile_spec_e(const float* pbuf, float energy, float* out_t, const float* w, bool dc_elim, int lane, int frame) {
    const float* pcol = pbuf + lane;
    if constexpr (!MFCC) {
        // measured on K1U (60 audio-h): the four-filter form helps the 80-filter bank (6.19 -> 6.06 ms), not the 40-filter
        // ones (fbank-40 5.88 -> 5.97, 13 cepstra 13.02 -> 13.24 on the bench shard): their epilogue warp is not the limit
        if constexpr (Plan::NF >= 80) {
            static_assert(Plan::NF % 4 == 0, "filters are processed four at a time");
            static_for<0, Plan::NF / 4>([&](auto mi) {
                constexpr int M = 4 * decltype(mi)::value;
                float v[4];
                mel_spec4<Plan, M, M + 1, M + 2, M + 3, LOG, PS>(pcol, w, v);
#pragma unroll
                for (int j = 0; j < 4; ++j) out_t[(M + j) * 32 + frame] = v[j];
            });
        } else {
            static_for<0, Plan::NF>([&](auto mi) {
                constexpr int M = decltype(mi)::value;
                out_t[M * 32 + frame] = mel_spec<Plan, M, LOG, PS>(pcol, w);
            });
        }
    } else {
        constexpr int NH = Plan::NF / 2;
        static_assert(Plan::NF % 2 == 0, "the specialised plans have an even number of filters");
        float s[NH], d[NH];                     // folded log-mel: x[n] +- x[NF-1-n]
        static_for<0, NH>([&](auto ni) {
            constexpr int n = decltype(ni)::value;
            const float lo = mel_spec<Plan, n, LOG, PS>(pcol, w), hi = mel_spec<Plan, Plan::NF - 1 - n, LOG, PS>(pcol, w);
            s[n] = lo + hi; d[n] = lo - hi;
        });
        const float* dw = w + Plan::NNZ;
        // Second fold for the even coefficients: cos(pi (2n + 1) c / (2 NF)) is symmetric (c = 0 mod 4) or antisymmetric
        // (c = 2 mod 4) under n -> NH - 1 - n, so they need NH / 2 products of s[n] +- s[NH - 1 - n] (the table row's first
        // half) instead of NH.  Odd coefficients keep the NH products of d[n].
        static_assert(NH % 2 == 0, "second fold");
        constexpr int NQ = NH / 2;
        float ss[NQ], sd[NQ];
#pragma unroll
        for (int n = 0; n < NQ; ++n) { ss[n] = s[n] + s[NH - 1 - n]; sd[n] = s[n] - s[NH - 1 - n]; }
        // four coefficients at a time, two accumulators each: eight independent FMA chains (a single warp runs this code;
        // with two chains it issues one FMA every other cycle at best)
        constexpr int CB = 4;
#pragma unroll
        for (int c0 = 0; c0 < D; c0 += CB) {
            float a0[CB], a1[CB];
#pragma unroll
            for (int j = 0; j < CB; ++j) { a0[j] = 0.f; a1[j] = 0.f; }
#pragma unroll
            for (int n = 0; n < NH; n += 2) {
#pragma unroll
                for (int j = 0; j < CB; ++j) {
                    const int c = c0 + j;
                    if (c < D) {
                        if (c & 1) {
                            a0[j] = fmaf(dw[c * NH + n], d[n], a0[j]);
                            a1[j] = fmaf(dw[c * NH + n + 1], d[n + 1], a1[j]);
                        } else if (n < NQ) {
                            a0[j] = fmaf(dw[c * NH + n], (c & 2) ? sd[n] : ss[n], a0[j]);
                            a1[j] = fmaf(dw[c * NH + n + 1], (c & 2) ? sd[n + 1] : ss[n + 1], a1[j]);
                        }
                    }
                }
            }
#pragma unroll
            for (int j = 0; j < CB; ++j) {
                const int c = c0 + j;
                if (c < D) {
                    float v = a0[j] + a1[j];
                    if (c == 0 && dc_elim) v = fe_log(energy);
                    out_t[c * 32 + frame] = v;
                }
            }
        }
    }
}
template <class Plan, int D, bool MFCC, bool LOG>
FE_HD void epi_tile_spec(const float* pbuf, const float* energies, float* out_t, const float* w, bool dc_elim, int lane) {
    epi_tile_spec_e<Plan, D, MFCC, LOG, kPStride>(pbuf, energies[lane], out_t, w, dc_elim, lane, lane);
}

// which specialised epilogue (if any) serves a configuration: 0 = generic, 1 = mfcc 40 filters -> 13, 2 = fbank 80
template <class Plan>
inline bool plan_matches(const int* row_start, const int* first_bin, int nf) {
    if (nf != Plan::NF || row_start[nf] != Plan::NNZ) return false;
    for (int m = 0; m < nf; ++m)
        if (first_bin[m] != Plan::B0[m] || row_start[m + 1] - row_start[m] != Plan::N[m] || row_start[m] != Plan::OFF[m]) return false;
    return true;
}

}  // namespace fe
