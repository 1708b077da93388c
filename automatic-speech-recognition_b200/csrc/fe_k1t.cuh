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
// Status: parity-green, SLOWER than K1 (20.6 vs 16.7 ms on the bench shard) -- TMEM holds one frame per lane for four
// warps per SM, and scalar straight-line code of this size is bound by instruction fetch, not issue
// (profiles/r02_k1t.md).  Selected with FE_K1T=1 only.
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
constexpr int kTFrameVecs = 20;
constexpr int kTRawBytes = ((kTileFrames - 1) * 160 + 400) * 2 + 32;      // 10 752 bytes per buffer

// Twiddles as plain arrays of (re, im): tw256[r * 16 + j] = W_256^(j r) = (cos, -sin)(2 pi j r / 256), r, j in 0..15;
// tw512[k] = (cos, sin)(2 pi k / 512), k in 0..128.  On the device they sit in constant memory and every index below
// is warp-uniform.
struct TTwiddles {
    const float2* tw256;
    const float2* tw512;
};

// ---------------------------------------------------------------------------
// Stage A: 16 column FFT16s over the rows of z[m] = x[2m] + i x[2m+1], m = j + 16 a (a < 13: 400 samples), from the
// lane's raw row (int16 pairs = one 32-bit word per complex point).  Column j's output row k1 goes to exchange
// words 32 k1 + 2 j (+1 for the imaginary part): row-major, so stage B reads a row with one 32-column load.
// Returns the lane's sum of squares (Parseval frame energy, sample units).
// ---------------------------------------------------------------------------
template <class EX>
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
                    re[a] = (float)(short)(u & 0xffffu);
                    im[a] = (float)((int)u >> 16);
                } else {
                    re[a] = 0.f; im[a] = 0.f;
                }
            }
#pragma unroll
            for (int a = 0; a < 13; ++a) {
                if (a & 1) { ss1 = fmaf(re[a], re[a], ss1); ss3 = fmaf(im[a], im[a], ss3); }
                else { ss0 = fmaf(re[a], re[a], ss0); ss2 = fmaf(im[a], im[a], ss2); }
            }
            fft16<13>(re, im);
            const int j = 4 * q + c;
#pragma unroll
            for (int k1 = 0; k1 < 16; ++k1) ex.st2(32 * k1 + 2 * j, re[pos16(k1)], im[pos16(k1)]);
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
// Stage B + real-FFT split for the row pair (r, 16 - r), r = 1 .. 7: two twiddled FFT16s over the columns, then the
// 16 bins r + 16 k2 and 16 - r + 16 k2 (k2 < 8) -- Z[k] and Z[256 - k] sit in the two rows of the pair.
// pcol: this lane's column of the power buffer (bin k at pcol[k * kPStride]).
// ---------------------------------------------------------------------------
template <class EX>
FE_HD void k1t_row_pair(EX& ex, int r, const TTwiddles& tw, float* pcol) {
    const int s = 16 - r;
    float va[32], vb[32];
    ex.ld32(32 * r, va);
    ex.ld32(32 * s, vb);
    ex.wait_ld();
    float ar[16], ai[16], br[16], bi[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) { ar[j] = va[2 * j]; ai[j] = va[2 * j + 1]; br[j] = vb[2 * j]; bi[j] = vb[2 * j + 1]; }
    const float2* ta = tw.tw256 + r * 16;
    const float2* tb = tw.tw256 + s * 16;
    fft16_twiddled(ar, ai, [&](int j, float& wr, float& wi) { const float2 w = ta[j]; wr = w.x; wi = w.y; });
    fft16_twiddled(br, bi, [&](int j, float& wr, float& wi) { const float2 w = tb[j]; wr = w.x; wi = w.y; });
#pragma unroll
    for (int k2 = 0; k2 < 8; ++k2) {
        const int sa = pos16(k2), sp = pos16(15 - k2);
        const float2 wa = tw.tw512[r + 16 * k2], wb = tw.tw512[s + 16 * k2];
        pcol[(r + 16 * k2) * kPStride] = k1t_bin_power(ar[sa], ai[sa], br[sp], bi[sp], wa.x, wa.y);
        pcol[(s + 16 * k2) * kPStride] = k1t_bin_power(br[sa], bi[sa], ar[sp], ai[sp], wb.x, wb.y);
    }
}

// Rows 0 and 8 are their own partners: bins 16 k2 (k2 = 0 .. 8, partner index (16 - k2) & 15 of row 0) and
// 8 + 16 k2 (k2 < 8, partner index 15 - k2 of row 8).  Also X[0] and X[256] for the Parseval frame energy.
template <class EX>
FE_HD void k1t_rows_0_8(EX& ex, const TTwiddles& tw, float* pcol, float& x0, float& x256) {
    float va[32], vb[32];
    ex.ld32(0, va);
    ex.ld32(32 * 8, vb);
    ex.wait_ld();
    float ar[16], ai[16], br[16], bi[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) { ar[j] = va[2 * j]; ai[j] = va[2 * j + 1]; br[j] = vb[2 * j]; bi[j] = vb[2 * j + 1]; }
    fft16(ar, ai);                                              // row 0: W_256^0 = 1
    const float2* tb = tw.tw256 + 8 * 16;
    fft16_twiddled(br, bi, [&](int j, float& wr, float& wi) { const float2 w = tb[j]; wr = w.x; wi = w.y; });
#pragma unroll
    for (int k2 = 0; k2 <= 8; ++k2) {
        const int sa = pos16(k2), sp = pos16((16 - k2) & 15);
        const float2 w = tw.tw512[16 * k2];
        pcol[(16 * k2) * kPStride] = k1t_bin_power(ar[sa], ai[sa], ar[sp], ai[sp], w.x, w.y);
    }
#pragma unroll
    for (int k2 = 0; k2 < 8; ++k2) {
        const int sa = pos16(k2), sp = pos16(15 - k2);
        const float2 w = tw.tw512[8 + 16 * k2];
        pcol[(8 + 16 * k2) * kPStride] = k1t_bin_power(br[sa], bi[sa], br[sp], bi[sp], w.x, w.y);
    }
    x0 = ar[0] + ai[0];           // X[0]   = Re Z[0] + Im Z[0]
    x256 = ar[0] - ai[0];         // X[256] = Re Z[0] - Im Z[0]
}

// the whole frame: raw row -> exchange -> power column; returns the frame energy (zero-handled)
template <class EX>
FE_HD float k1t_frame(const uint4* raw4, EX& ex, const TTwiddles& tw, float* pcol, float pscale) {
    const float ss = k1t_stage_a(raw4, ex);
    ex.wait_st();
    float x0, x256;
    k1t_rows_0_8(ex, tw, pcol, x0, x256);
#if defined(__CUDA_ARCH__)
#pragma unroll 1
#endif
    for (int r = 1; r < 8; ++r) k1t_row_pair(ex, r, tw, pcol);
    return frame_energy(ss, x0, x256, pscale);
}

}  // namespace fe
