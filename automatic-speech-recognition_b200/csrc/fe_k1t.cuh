// K1T: frames -> statics with LANE = FRAME and the FFT's 16 x 16 exchange in TENSOR MEMORY.
//
// Replaces, per frame, the same reference calls as K1 (preprocess.py:72-82: stack_frames -> rfft(512) -> |X|^2/512
// -> mel filterbank -> zero_handling -> log -> DCT-II -> c0 <- log(energy)).
//
// Why a second kernel.  K1 (fe_kernels.cuh) gives a frame to 8 lanes; its twiddles differ from lane to lane, so
// they are register operands loaded from shared memory, the 16 x 16 transposition between the two FFT16 stages goes
// through 64 scalar shared-memory stores per lane, and the real-FFT pairing needs per-lane selects.  Measured on the
// B200 (tools/ubench_issue2.cu, profiles/r02_issue_cost_model.jsonl): the register file delivers about two operand
// words per cycle and lane, so an FMA whose three operands are distinct registers costs 1.5 issue cycles while one with
// a warp-uniform operand (uniform register / constant bank) costs 1; K1 runs at 96 % of that issue model, i.e. it
// cannot get faster without doing fewer register reads.
//
// Here one lane owns one frame from the raw samples to the 13 cepstra:
//   * every twiddle, filterbank weight and DCT coefficient is the same for all 32 lanes of an instruction ->
//     uniform operands, no twiddle loads, no selects, no lane-dependent addressing;
//   * the 256 complex values between the two FFT16 stages (2 KB per frame) live in the lane's 512 columns of
//     TENSOR MEMORY (tcgen05.st / tcgen05.ld, shape 32x32b: thread i of a warp owns TMEM lane 32 (warp % 4) + i) --
//     Blackwell's 256 KB of TMEM per SM is exactly 128 lanes x 512 x 32 bit, one frame per lane for 4 warps.  Shared
//     memory could not hold it (4 warps x 32 frames x 2 KB = 256 KB);
//   * nothing is exchanged between lanes: no shared-memory transposition, no selects, no producer / consumer slots.  A
//     tile = 32 consecutive frames of one utterance; the kernel (fe_kernels.cuh: k_frames_to_statics_t) lets the two
//     warps that share a scheduler and its 32 TMEM lanes split a tile's columns / row pairs between them.
// Shared memory only stages the raw samples (one bulk copy per tile, double-buffered) and the lane's 129 power bins
// ([bin][lane]).
//
// Two kernels are built from these phases (fe_kernels.cuh):
//   * k_frames_to_statics_u (K1U, the default): 16 FFT warps + 4 epilogue warps per SM, four warps share a tile (one
//     column quad / two row pairs each), every warp-uniform quantity in the uniform datapath, conflict-free chunked raw
//     layout: 13.0 ms on the bench shard = 0.66 of nominal FP32 (K1: 16.7 ms) -- profiles/r02b_k1u.md;
//   * k_frames_to_statics_t (K1T, FE_K1T=1): the first form, two warps per tile, 19.2 ms -- TMEM holds one frame per
//     lane for four warps per SM, and two warps per scheduler cannot keep the issue port busy (profiles/r02_k1t.md).
//
// The per-lane phases are host/device functions over an exchange accessor, so tests/host_sim replays them on the CPU
// (exchange = a float[512]) against the float64 oracle before anything runs on a GPU.
#pragma once
#include "fe_core.cuh"

namespace fe {

// Raw samples of a tile sit in shared memory as they do in HBM (one bulk copy per tile): frame f starts 160 samples
// = 320 bytes = 20 sixteen-byte vectors after frame f - 1.  The 16-byte loads of 8 neighbouring lanes then hit only
// two bank groups (4-way conflict): 16 instead of 4 wavefronts per load, 832 per tile -- the LSU is far from
// saturated (28 %), whereas 32 per-frame copies into padded rows cost the issuing warp > 1 000 cycles per tile.
#ifndef FE_TFRAMEVECS
#define FE_TFRAMEVECS 20           // 21: timing experiment only (conflict-free lane stride, wrong samples)
#endif
constexpr int kTFrameVecs = FE_TFRAMEVECS;
constexpr int kTRawBytes = ((kTileFrames - 1) * 160 + 400) * 2 + 32;      // 10 752 bytes per buffer

// Twiddles.  The two rows of a pair share every instruction of stage B (packed f32x2 halves), so their twiddles are
// stored as pairs: tw256p[p * 16 + j] = (wr_a, wr_b, wi_a, wi_b) of W_256^(j ra), W_256^(j rb) = (cos, -sin), and
// tw512p[p * 8 + k2] = (c_a, c_b, s_a, s_b) = (cos, sin)(2 pi k / 512) of the bins k = ra + 16 k2, rb + 16 k2, where pair
// p = 0 is rows (0, 8) and pair p >= 1 rows (p, 16 - p); tw512[k] (cos, sin), k = 0 .. 128, serves the self-paired
// rows 0 and 8.  On the device they sit in constant memory; every index is warp-uniform, so an FFMA2 takes its
// twiddle as a uniform-register PAIR operand (`FFMA2 R, R, UR.F32x2, R`).
struct TTwiddles {
    const float4* tw256p;
    const float4* tw512p;
    const float2* tw512;
};

FE_HD constexpr int k1t_row_a(int p) { return p == 0 ? 0 : p; }
FE_HD constexpr int k1t_row_b(int p) { return p == 0 ? 8 : 16 - p; }

// ---------------------------------------------------------------------------
// Exchange layout (512 words per lane): pair p, plane (re / im), column j, slot (row a / row b) at
//     64 p + 32 plane + 2 j + slot
// so that stage B reads a pair's plane with ONE 32-column load whose consecutive register pairs are the packed
// halves (row a, row b) of column j.
//
// Stage A: 16 column FFT16s over the rows of z[m] = x[2m] + i x[2m+1], m = j + 16 a (a < 13: 400 samples), from the
// lane's raw samples (int16 pairs = one 32-bit word per complex point).  Plain scalar code: no twiddles here, and
// scalar results can be stored as the (row a, row b) pairs the layout wants (8 pairs x 2 planes x st2 per column).
// Returns the lane's sum of squares over the columns it processed (Parseval frame energy, sample units).
// ---------------------------------------------------------------------------
// ENERGY = false skips the sum of squares (filterbank features do not use the frame energy) and returns 0.
template <bool ENERGY = true, class EX>
FE_HD float k1t_stage_a(const uint4* raw4, EX& ex, int q_begin = 0, int q_end = 4) {
    float ss0 = 0.f, ss1 = 0.f, ss2 = 0.f, ss3 = 0.f;
#if defined(__CUDA_ARCH__)
#pragma unroll 1
#endif
    for (int q = q_begin; q < q_end; ++q) {                     // columns 4q .. 4q + 3
        uint4 w[13];
#pragma unroll
        for (int a = 0; a < 12; ++a) w[a] = raw4[4 * a + q];   // words 16 a + 4 q .. + 3
        w[12] = make_uint4(0u, 0u, 0u, 0u);
        if (q < 2) w[12] = raw4[48 + q];                        // row 12: points 192 .. 199 only (samples 384 .. 399)
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            float re[16], im[16];
#pragma unroll
            for (int a = 0; a < 16; ++a) {
                if (a < 13) {
                    const uint32_t u = c == 0 ? w[a].x : (c == 1 ? w[a].y : (c == 2 ? w[a].z : w[a].w));
#if defined(__CUDA_ARCH__)
                    // high half: shift + an OPAQUE convert.  Left to itself the compiler moves the first butterfly level of
                    // the imaginary parts into integer arithmetic (26 IADD3 per column) and then converts both the inputs
                    // (for the sum of squares) and the sums: 13 more instructions per column (13.52 -> 13.16 ms).  The low
                    // half converts with one I2F.S16 (XU pipe); both halves through the XU are slower (14.09), both
                    // through PRMT / I2FP as well (13.68) -- profiles/r02b_k1u.md.
                    re[a] = (float)(short)(u & 0xffffu);
                    asm("cvt.rn.f32.s32 %0, %1;" : "=f"(im[a]) : "r"((int)u >> 16));
#else
                    re[a] = (float)(short)(u & 0xffffu);
                    im[a] = (float)((int)u >> 16);
#endif
                } else {
                    re[a] = 0.f; im[a] = 0.f;
                }
            }
            if (ENERGY) {
#pragma unroll
                for (int a = 0; a < 13; ++a) {
                    if (a & 1) { ss1 = fmaf(re[a], re[a], ss1); ss3 = fmaf(im[a], im[a], ss3); }
                    else { ss0 = fmaf(re[a], re[a], ss0); ss2 = fmaf(im[a], im[a], ss2); }
                }
            }
            fft16<13>(re, im);
            const int j = 4 * q + c;
#pragma unroll
            for (int p = 0; p < 8; ++p) {
                const int sa = pos16(k1t_row_a(p)), sb = pos16(k1t_row_b(p));
                ex.st2(64 * p + 2 * j, re[sa], re[sb]);
                ex.st2(64 * p + 32 + 2 * j, im[sa], im[sb]);
            }
        }
    }
    return (ss0 + ss1) + (ss2 + ss3);
}

// 2 X[k] = 2E - i w O for the pair (A = Z[k], P = Z[256 - k]); returns |2 X[k]|^2 (the 1/2048 lives in the weights)
FE_HD float k1t_bin_power(float ar, float ai, float pr, float pi, float c, float s) {
    const float er = ar + pr, ei = ai - pi;         // 2E  (B = conj(partner))
    const float orr = ar - pr, oi = ai + pi;        // 2O
    const float xr = fmaf(-s, orr, fmaf(c, oi, er));
    const float xi = fmaf(-s, oi, fmaf(-c, orr, ei));
    return fmaf(xi, xi, xr * xr);
}

// ---------------------------------------------------------------------------
// Stage B + real-FFT split of pair p: ONE packed twiddled FFT16 over the columns (halves = the pair's two rows), then
//   p >= 1: the 16 bins p + 16 k2 and 16 - p + 16 k2 (k2 < 8): Z[k] and Z[256 - k] are the two halves of slots k2 and
//           15 - k2, so the partner is a half-swapped operand (as in K1's post-pass);
//   p == 0: rows 0 and 8 are their own partners (index (16 - k2) & 15 of row 0, 15 - k2 of row 8): 17 bins in scalar
//           code, plus X[0] and X[256] for the Parseval frame energy.
// pcol: this lane's column of the power buffer (bin k at pcol[k * PS]).
// ---------------------------------------------------------------------------
// after_loads(): called once the pair's exchange values are in registers (the kernel signals "exchange drained" there).
struct K1TNoop { FE_HD void operator()() const {} };
// BIN_LO .. BIN_HI: the bins the caller's power buffer has rows for (others are computed but not stored)
template <int PS = kPStride, int BIN_LO = 0, int BIN_HI = 128, class EX, class AL = K1TNoop>
FE_HD void k1t_pair(EX& ex, int p, const TTwiddles& tw, float* pcol, float& x0, float& x256, AL&& after_loads = AL()) {
    float va[32], vb[32];
    ex.ld32(64 * p, va);
    ex.ld32(64 * p + 32, vb);
    ex.wait_ld();
    after_loads();
    float2 zr[16], zi[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) { zr[j] = make_float2(va[2 * j], va[2 * j + 1]); zi[j] = make_float2(vb[2 * j], vb[2 * j + 1]); }
    const float4* t256 = tw.tw256p + p * 16;
    fft16_twiddled(zr, zi, [&](int j, float2& wr, float2& wi) {
        const float4 w = t256[j];
        wr = make_float2(w.x, w.y); wi = make_float2(w.z, w.w);
    });
    if (p == 0) {
#pragma unroll
        for (int k2 = 0; k2 <= 8; ++k2) {
            const int sa = pos16(k2), sp = pos16((16 - k2) & 15);
            const float2 w = tw.tw512[16 * k2];
            if (16 * k2 >= BIN_LO && 16 * k2 <= BIN_HI)
                pcol[(16 * k2) * PS] = k1t_bin_power(zr[sa].x, zi[sa].x, zr[sp].x, zi[sp].x, w.x, w.y);
        }
#pragma unroll
        for (int k2 = 0; k2 < 8; ++k2) {
            const int sa = pos16(k2), sp = pos16(15 - k2);
            const float2 w = tw.tw512[8 + 16 * k2];
            pcol[(8 + 16 * k2) * PS] = k1t_bin_power(zr[sa].y, zi[sa].y, zr[sp].y, zi[sp].y, w.x, w.y);
        }
        x0 = zr[0].x + zi[0].x;           // X[0]   = Re Z[0] + Im Z[0]
        x256 = zr[0].x - zi[0].x;         // X[256] = Re Z[0] - Im Z[0]
        return;
    }
    const float4* t512 = tw.tw512p + p * 8;
    float* pa = pcol + p * PS;
    float* pb = pcol + (16 - p) * PS;
#pragma unroll
    for (int k2 = 0; k2 < 8; ++k2) {
        const int sa = pos16(k2), sp = pos16(15 - k2);
        const float2 ar = zr[sa], ai = zi[sa], pr = pswap(zr[sp]), pi = pswap(zi[sp]);
        const float4 w = t512[k2];
        const float2 c = make_float2(w.x, w.y), s = make_float2(w.z, w.w);
        const float2 er = padd(ar, pr), ei = psub(ai, pi);          // 2E  (B = conj(partner))
        const float2 orr = psub(ar, pr), oi = padd(ai, pi);         // 2O
        const float2 xr = pfma_rr(pneg(s), orr, pfma_rr(c, oi, er));
        const float2 xi = pfma_rr(pneg(s), oi, pfma_rr(pneg(c), orr, ei));
        const float2 plo = pfma_rr(xi, xi, pmul(xr, xr));
        if (BIN_LO <= 1 || k2 > 0 || p >= BIN_LO) pa[16 * k2 * PS] = plo.x;      // bin p + 16 k2
        pb[16 * k2 * PS] = plo.y;                                                 // bin 16 - p + 16 k2 (9 .. 127)
    }
}

// the whole frame (host replay; the kernel splits the same phases over two warps): raw samples -> exchange -> power
// column; returns the frame energy (zero-handled)
template <class EX>
FE_HD float k1t_frame(const uint4* raw4, EX& ex, const TTwiddles& tw, float* pcol, float pscale) {
    const float ss = k1t_stage_a(raw4, ex);
    ex.wait_st();
    float x0 = 0.f, x256 = 0.f, d0, d1;
#if defined(__CUDA_ARCH__)
#pragma unroll 1
#endif
    for (int p = 0; p < 8; ++p) {
        if (p == 0) k1t_pair(ex, p, tw, pcol, x0, x256);
        else k1t_pair(ex, p, tw, pcol, d0, d1);
    }
    return frame_energy(ss, x0, x256, pscale);
}

// host-side construction of the packed twiddle tables (float64, rounded once)
template <class F4, class F2>
inline void k1t_build_twiddles(F4* tw256p /* [128] */, F4* tw512p /* [64] */, F2* tw512 /* [129] */) {
    const double kPi = 3.14159265358979323846;
    for (int p = 0; p < 8; ++p) {
        const int ra = k1t_row_a(p), rb = k1t_row_b(p);
        for (int j = 0; j < 16; ++j) {
            const double aa = 2.0 * kPi * ((ra * j) % 256) / 256.0, ab = 2.0 * kPi * ((rb * j) % 256) / 256.0;
            tw256p[p * 16 + j].x = (float)cos(aa); tw256p[p * 16 + j].y = (float)cos(ab);
            tw256p[p * 16 + j].z = (float)-sin(aa); tw256p[p * 16 + j].w = (float)-sin(ab);
        }
        for (int k2 = 0; k2 < 8; ++k2) {
            const double aa = 2.0 * kPi * (ra + 16 * k2) / 512.0, ab = 2.0 * kPi * (rb + 16 * k2) / 512.0;
            tw512p[p * 8 + k2].x = (float)cos(aa); tw512p[p * 8 + k2].y = (float)cos(ab);
            tw512p[p * 8 + k2].z = (float)sin(aa); tw512p[p * 8 + k2].w = (float)sin(ab);
        }
    }
    for (int k = 0; k <= 128; ++k) { const double a = 2.0 * kPi * k / 512.0; tw512[k].x = (float)cos(a); tw512[k].y = (float)sin(a); }
}

}  // namespace fe
