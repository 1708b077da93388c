// Per-warp phases of the frames->statics kernel (K1), written as host/device
// functions so that tests/host_sim can replay the exact lane-level dataflow on a
// CPU (lanes as a loop, shared memory as an array) before anything runs on a GPU.
//
// Replaces, per frame, what speechpy does for /root/reference/preprocess.py:72-82:
//   stack_frames (400/160, rectangular or table window) -> rfft(512) -> |X|^2/512
//   -> frame energy -> mel filterbank -> zero_handling -> log -> DCT-II(ortho)[:D]
//   -> c0 <- log(energy).
//
// Geometry: one frame = 8 lanes, one warp = 4 consecutive frames of one utterance.
// The 512-point real FFT is a 256-point complex FFT (z[m] = x[2m] + i x[2m+1])
// split 16 x 16: stage A = two in-register FFT16 per lane over a (m = j + 16a,
// j = t and t+8), twiddle W_256^(j k1), exchange through a swizzled shared
// buffer, stage B = two in-register FFT16 per lane over j for rows k1 = {t, 16-t}
// (lane 0: {8, 0}), so that Z[k] and Z[256-k] of the real-FFT post-pass sit in the
// SAME lane and no second exchange is needed.
#pragma once
#include <stdint.h>
#include <math.h>

#if defined(__CUDACC__)
#define FE_HD __host__ __device__ __forceinline__
#else
#define FE_HD inline
struct float2 { float x, y; };
struct float4 { float x, y, z, w; };
static inline float4 make_float4(float a, float b, float c, float d) { float4 v = {a, b, c, d}; return v; }
static inline float2 make_float2(float a, float b) { float2 v = {a, b}; return v; }
#endif

namespace fe {

constexpr int kNfft = 512;
constexpr int kBins = 257;
constexpr int kWarpFrames = 4;          // frames per warp pass
constexpr int kCtaWarps = 8;
constexpr int kCtaFrames = kWarpFrames * kCtaWarps;   // frames per tile-table entry
constexpr int kERegion = 512;           // floats of exchange buffer per frame
constexpr int kPStagger = 8;            // per-frame-slot float offset of the power row
constexpr int kLogmelOff = 320;         // log-mel row offset inside the frame's region
constexpr int kMaxFilters = 128;
constexpr int kTw256Stride = 17;        // float2 per row of the padded W_256 table
constexpr float kEpsF64 = 2.220446049250313e-16f;   // np.finfo(float).eps, as float

// ---------------------------------------------------------------------------
// Shared-memory tables of one CTA (pointers into dynamic smem)
// ---------------------------------------------------------------------------
struct SmemTables {
    const float2* tw256;      // [16][kTw256Stride]  W_256^(j*k) = (cos, -sin)
    const float2* tw512;      // [257]               (cos, sin)(2 pi k / 512)
    const float*  window;     // [pcm layout of one frame, 13*32 floats] or nullptr
    const int*    fb_start;   // [nf + 1]
    const int*    fb_bin0;    // [nf]
    const float*  fb_w;       // [nnz], pre-scaled by 1/2048
    const float*  dct;        // [D][dct_stride]
    int nf, D, dct_stride;
    int full_spectrum;        // filterbank touches bins > 128
    int is_mfcc, fbank_log, dc_elim;
};

// position of sample r (0..31) inside a 32-float block of the staged PCM tile:
// complex point j' = r/2 (re/im = r&1), u = j' & 7 (lane), h = j' >> 3 (which of
// the lane's two FFT16) -> 4u + 2c + h, so one 16-byte load at 32a + 4t yields
// (re_t, re_{t+8}, im_t, im_{t+8}) of row a.
FE_HD int pcm_pos(int r) { int jp = r >> 1; return ((jp & 7) << 2) + ((r & 1) << 1) + (jp >> 3); }

// ---------------------------------------------------------------------------
// radix-4 butterfly and 16-point forward FFT on registers
// ---------------------------------------------------------------------------
FE_HD void bfly4(float& r0, float& i0, float& r1, float& i1, float& r2, float& i2, float& r3, float& i3) {
    float t0r = r0 + r2, t0i = i0 + i2;
    float t1r = r0 - r2, t1i = i0 - i2;
    float t2r = r1 + r3, t2i = i1 + i3;
    float t3r = r1 - r3, t3i = i1 - i3;
    r0 = t0r + t2r; i0 = t0i + t2i;
    r2 = t0r - t2r; i2 = t0i - t2i;
    r1 = t1r + t3i; i1 = t1i - t3r;     // t1 - i t3
    r3 = t1r - t3i; i3 = t1i + t3r;     // t1 + i t3
}

// multiply (r, i) by exp(-2 pi i M / 16)
template <int M> FE_HD void mul_w16(float& r, float& i) {
    constexpr float C1 = 0.92387953251128674f, S1 = 0.38268343236508977f, H = 0.70710678118654752f;
    if (M == 1)      { float t = r * C1 + i * S1; i = i * C1 - r * S1; r = t; }
    else if (M == 2) { float t = (r + i) * H;     i = (i - r) * H;     r = t; }
    else if (M == 3) { float t = r * S1 + i * C1; i = i * S1 - r * C1; r = t; }
    else if (M == 4) { float t = r; r = i; i = -t; }
    else if (M == 6) { float t = (i - r) * H;     i = -(r + i) * H;    r = t; }
    else if (M == 9) { float t = -r * C1 - i * S1; i = r * S1 - i * C1;  r = t; }
}

// In-place FFT16: input natural order x[n]; output X[k] lands at slot pos16(k).
FE_HD constexpr int pos16(int k) { return (k >> 2) + ((k & 3) << 2); }

FE_HD void fft16(float (&xr)[16], float (&xi)[16]) {
#pragma unroll
    for (int n1 = 0; n1 < 4; ++n1)
        bfly4(xr[n1], xi[n1], xr[n1 + 4], xi[n1 + 4], xr[n1 + 8], xi[n1 + 8], xr[n1 + 12], xi[n1 + 12]);
    // slot n1 + 4 k2 holds A[n1][k2]; twiddle W_16^(n1 k2)
    mul_w16<1>(xr[5], xi[5]);   mul_w16<2>(xr[9], xi[9]);   mul_w16<3>(xr[13], xi[13]);
    mul_w16<2>(xr[6], xi[6]);   mul_w16<4>(xr[10], xi[10]); mul_w16<6>(xr[14], xi[14]);
    mul_w16<3>(xr[7], xi[7]);   mul_w16<6>(xr[11], xi[11]); mul_w16<9>(xr[15], xi[15]);
#pragma unroll
    for (int k2 = 0; k2 < 4; ++k2)
        bfly4(xr[4 * k2], xi[4 * k2], xr[4 * k2 + 1], xi[4 * k2 + 1],
              xr[4 * k2 + 2], xi[4 * k2 + 2], xr[4 * k2 + 3], xi[4 * k2 + 3]);
}

// swizzled float offset of the 16-byte chunk holding (row k1, columns 2c, 2c+1)
FE_HD int e_chunk(int k1, int c, int fs) { return (k1 << 5) + (((c ^ (k1 & 7) ^ ((fs & 1) << 2))) << 2); }

// ---------------------------------------------------------------------------
// Phase 1 (stage A): load 13 rows of the frame, window, sum of squares, two FFT16,
// twiddle, scatter into the frame's exchange region.
//   pcm_f : staged PCM of this frame (tile base + fs*HOP), permuted layout
//   e_f   : this frame's 512-float exchange region
//   returns the lane's partial sum of squares (Parseval frame energy)
// ---------------------------------------------------------------------------
template <int FRAME_LEN>
FE_HD float stage_a(const float* pcm_f, float* e_f, const SmemTables& tb, int t, int fs) {
    float r0[16], i0[16], r1[16], i1[16];
    float ss = 0.f;
    constexpr int ROWS = (FRAME_LEN + 31) / 32;
#pragma unroll
    for (int a = 0; a < 16; ++a) {
        if (a < ROWS) {
            float4 v = *reinterpret_cast<const float4*>(pcm_f + 32 * a + 4 * t);
            if (tb.window) {
                float4 w = *reinterpret_cast<const float4*>(tb.window + 32 * a + 4 * t);
                v.x *= w.x; v.y *= w.y; v.z *= w.z; v.w *= w.w;
            }
            // validity of sample n = 2 (j + 16 a) + c against FRAME_LEN (t-dependent only in the last row)
            int n0 = 2 * (t + 16 * a), n1 = 2 * (t + 8 + 16 * a);
            if (32 * a + 32 > FRAME_LEN) {
                if (n0 >= FRAME_LEN) v.x = 0.f;
                if (n0 + 1 >= FRAME_LEN) v.z = 0.f;
                if (n1 >= FRAME_LEN) v.y = 0.f;
                if (n1 + 1 >= FRAME_LEN) v.w = 0.f;
            }
            r0[a] = v.x; r1[a] = v.y; i0[a] = v.z; i1[a] = v.w;
            ss = fmaf(v.x, v.x, ss); ss = fmaf(v.y, v.y, ss);
            ss = fmaf(v.z, v.z, ss); ss = fmaf(v.w, v.w, ss);
        } else {
            r0[a] = 0.f; r1[a] = 0.f; i0[a] = 0.f; i1[a] = 0.f;
        }
    }
    fft16(r0, i0);
    fft16(r1, i1);
    const float2* tw0 = tb.tw256 + t * kTw256Stride;
    const float2* tw1 = tb.tw256 + (t + 8) * kTw256Stride;
    const int c0 = t >> 1, c1 = (t + 8) >> 1, half = (t & 1) << 1;
#pragma unroll
    for (int k1 = 0; k1 < 16; ++k1) {
        const int s = pos16(k1);
        float2 w0 = tw0[k1], w1 = tw1[k1];
        float yr0 = r0[s] * w0.x - i0[s] * w0.y, yi0 = r0[s] * w0.y + i0[s] * w0.x;
        float yr1 = r1[s] * w1.x - i1[s] * w1.y, yi1 = r1[s] * w1.y + i1[s] * w1.x;
        *reinterpret_cast<float2*>(e_f + e_chunk(k1, c0, fs) + half) = make_float2(yr0, yi0);
        *reinterpret_cast<float2*>(e_f + e_chunk(k1, c1, fs) + half) = make_float2(yr1, yi1);
    }
    return ss;
}

// ---------------------------------------------------------------------------
// Phase 2 (stage B): rows ra/rb of the exchange region -> two FFT16 over j.
// Slot pos16(k2) of (ar, ai) holds Z[ra + 16 k2]; same for b.
// ---------------------------------------------------------------------------
struct LaneZ { float ar[16], ai[16], br[16], bi[16]; };

FE_HD int row_a(int t) { return t ? t : 8; }
FE_HD int row_b(int t) { return t ? 16 - t : 0; }

FE_HD void stage_b(const float* e_f, LaneZ& z, int t, int fs) {
    const int ra = row_a(t), rb = row_b(t);
#pragma unroll
    for (int c = 0; c < 8; ++c) {
        float4 va = *reinterpret_cast<const float4*>(e_f + e_chunk(ra, c, fs));
        float4 vb = *reinterpret_cast<const float4*>(e_f + e_chunk(rb, c, fs));
        z.ar[2 * c] = va.x; z.ai[2 * c] = va.y; z.ar[2 * c + 1] = va.z; z.ai[2 * c + 1] = va.w;
        z.br[2 * c] = vb.x; z.bi[2 * c] = vb.y; z.br[2 * c + 1] = vb.z; z.bi[2 * c + 1] = vb.w;
    }
    fft16(z.ar, z.ai);
    fft16(z.br, z.bi);
}

// ---------------------------------------------------------------------------
// Phase 3 (post-pass): real-FFT split, power (scaled by 4, i.e. |2X|^2; the
// 1/2048 = 1/(4*512) lives in the filterbank weights), store the power row.
//   p_f : power row of this frame (e_f + fs*kPStagger), indexed by bin
//   returns |2 X[0]|^2 + |2 X[256]|^2 contribution pieces via x0/x256 (lane t==0)
// ---------------------------------------------------------------------------
FE_HD void pair_power(float Ar, float Ai, float Pr, float Pi, float c, float s,
                      float& plo, float& phi) {
    // A = Z[k], partner Zp = Z[256-k]; B = conj(Zp)
    float er = Ar + Pr, ei = Ai - Pi;         // 2E
    float orr = Ar - Pr, oi = Ai + Pi;        // 2O
    float tr = s * orr - c * oi;              // T = i w O, w = (c, -s)
    float ti = c * orr + s * oi;
    float xr = er - tr, xi = ei - ti;         // 2 X[k]
    float yr = er + tr, yi = ei + ti;         // conj(2 X[256-k])
    plo = xr * xr + xi * xi;
    phi = yr * yr + yi * yi;
}

FE_HD void post_pass(const LaneZ& z, float* p_f, const SmemTables& tb, int t,
                     float& x0, float& x256) {
    const bool t0 = (t == 0);
    const int ra = row_a(t), rb = row_b(t);
    // set 1: bins ra + 16 k2, partner Vb[15 - k2]; set 2: bins rb + 16 k2, partner Va[15 - k2]
#pragma unroll
    for (int k2 = 0; k2 < 8; ++k2) {
        const int sa = pos16(k2), sp = pos16(15 - k2);
        // partner of set 1 = Vb[15-k2] = t0 ? Za[15-k2] : Zb[15-k2]      (upper half of Vb)
        float p1r = t0 ? z.ar[sp] : z.br[sp];
        float p1i = t0 ? z.ai[sp] : z.bi[sp];
        // partner of set 2 = Va[15-k2] = t0 ? Zb[(16-k2)&15] : Za[15-k2] (upper half of Va)
        const int sq = pos16((16 - k2) & 15);
        float p2r = t0 ? z.br[sq] : z.ar[sp];
        float p2i = t0 ? z.bi[sq] : z.ai[sp];
        const int k_1 = ra + 16 * k2, k_2 = rb + 16 * k2;
        float2 w1 = tb.tw512[k_1], w2 = tb.tw512[k_2];
        float lo1, hi1, lo2, hi2;
        pair_power(z.ar[sa], z.ai[sa], p1r, p1i, w1.x, w1.y, lo1, hi1);
        pair_power(z.br[sa], z.bi[sa], p2r, p2i, w2.x, w2.y, lo2, hi2);
        p_f[k_1] = lo1;
        p_f[k_2] = lo2;
        if (tb.full_spectrum) {
            p_f[256 - k_1] = hi1;
            p_f[256 - k_2] = hi2;      // k_2 == 0 (lane 0) writes bin 256
        }
    }
    // bin 128 = row 0, k2 = 8 (self-paired): 2 X[128] = 2 conj(Z[128]); lane 0 holds row 0 in b
    float zr = z.br[pos16(8)], zi = z.bi[pos16(8)];
    if (t0) p_f[128] = 4.f * (zr * zr + zi * zi);
    // X[0] = Zr + Zi, X[256] = Zr - Zi of Z[0] (lane 0, row 0 slot 0)
    x0 = z.br[0] + z.bi[0];
    x256 = z.br[0] - z.bi[0];
}

// Parseval frame energy: sum_{k=0..256} |X_k|^2 / 512 = sum x^2 / 2 + (X0^2 + X256^2) / 1024
FE_HD float frame_energy(float sumsq, float x0, float x256) {
    float e = 0.5f * sumsq + (x0 * x0 + x256 * x256) * (1.0f / 1024.0f);
    return e == 0.f ? kEpsF64 : e;
}

FE_HD float fe_log(float x) {
#if defined(__CUDA_ARCH__)
    return __logf(x);
#else
    return logf(x);
#endif
}

// ---------------------------------------------------------------------------
// Phase 4: mel filterbank of one (frame, filter) task from the power row.
// ---------------------------------------------------------------------------
FE_HD float mel_task(const float* p_f, const SmemTables& tb, int m) {
    const int s = tb.fb_start[m], e = tb.fb_start[m + 1];
    const float* p = p_f + tb.fb_bin0[m];
    float acc = 0.f;
    for (int i = s; i < e; ++i) acc = fmaf(tb.fb_w[i], p[i - s], acc);
    return acc == 0.f ? kEpsF64 : acc;
}

// ---------------------------------------------------------------------------
// Phase 5: one cepstral coefficient from the log-mel row (rows padded with zeros
// to a multiple of 4 floats).
// ---------------------------------------------------------------------------
FE_HD float dct_task(const float* logmel, const SmemTables& tb, int c) {
    const float* d = tb.dct + c * tb.dct_stride;
    float acc = 0.f;
    const int n4 = (tb.nf + 3) >> 2;
    for (int q = 0; q < n4; ++q) {
        float4 a = *reinterpret_cast<const float4*>(d + 4 * q);
        float4 b = *reinterpret_cast<const float4*>(logmel + 4 * q);
        acc = fmaf(a.x, b.x, acc); acc = fmaf(a.y, b.y, acc);
        acc = fmaf(a.z, b.z, acc); acc = fmaf(a.w, b.w, acc);
    }
    return acc;
}

// ---------------------------------------------------------------------------
// Phase 0: stage one item (two sample pairs 16 apart) of the warp's PCM span.
//   p      first sample of the warp's first frame;  n_samp  samples to stage (even)
//   id     item index: block q = id >> 3, lane slot u = id & 7
// ---------------------------------------------------------------------------
FE_HD void stage_store(float* pcm_w, int id, float a0, float a1, float b0, float b1) {
    *reinterpret_cast<float4*>(pcm_w + ((id >> 3) << 5) + ((id & 7) << 2)) = make_float4(a0, b0, a1, b1);
}

FE_HD void stage_item_i16(const short* p, int n_samp, int id, float* pcm_w) {
    const int sA = ((id >> 3) << 5) + ((id & 7) << 1), sB = sA + 16;
    float a0 = 0.f, a1 = 0.f, b0 = 0.f, b1 = 0.f;
    const float k = 1.0f / 32768.0f;
#if defined(__CUDA_ARCH__)
    if (sA < n_samp) { short2 v = __ldg(reinterpret_cast<const short2*>(p + sA)); a0 = (float)v.x * k; a1 = (float)v.y * k; }
    if (sB < n_samp) { short2 v = __ldg(reinterpret_cast<const short2*>(p + sB)); b0 = (float)v.x * k; b1 = (float)v.y * k; }
#else
    if (sA < n_samp) { a0 = (float)p[sA] * k; a1 = (float)p[sA + 1] * k; }
    if (sB < n_samp) { b0 = (float)p[sB] * k; b1 = (float)p[sB + 1] * k; }
#endif
    stage_store(pcm_w, id, a0, a1, b0, b1);
}

FE_HD void stage_item_f32(const float* p, int n_samp, int id, float* pcm_w) {
    const int sA = ((id >> 3) << 5) + ((id & 7) << 1), sB = sA + 16;
    float a0 = 0.f, a1 = 0.f, b0 = 0.f, b1 = 0.f;
    if (sA < n_samp) { a0 = p[sA]; a1 = p[sA + 1]; }
    if (sB < n_samp) { b0 = p[sB]; b1 = p[sB + 1]; }
    stage_store(pcm_w, id, a0, a1, b0, b1);
}

// rows inside the warp's exchange buffer e_w (kWarpFrames regions of kERegion floats)
FE_HD float* power_row(float* e_w, int f) { return e_w + f * kERegion + f * kPStagger; }
FE_HD float* logmel_row(float* e_w, int f) { return e_w + f * kERegion + kLogmelOff + f * kPStagger; }

// Phase 4, one task: id -> (filter m = id >> 2, frame slot f = id & 3)
FE_HD void mel_phase(float* e_w, const SmemTables& tb, int id, int nfw) {
    const int m = id >> 2, f = id & 3;
    if (f >= nfw) return;
    float v = 0.f;
    if (m < tb.nf) {
        v = mel_task(power_row(e_w, f), tb, m);
        if (tb.is_mfcc || tb.fbank_log) v = fe_log(v);
    }
    logmel_row(e_w, f)[m] = v;
}

// Phase 5, one task: id = f * D + c -> statics value of (frame slot f, coefficient c)
FE_HD float emit_phase(float* e_w, const float* energies, const SmemTables& tb, int id) {
    const int f = id / tb.D, c = id - f * tb.D;
    const float* row = logmel_row(e_w, f);
    if (!tb.is_mfcc) return row[c];
    if (c == 0 && tb.dc_elim) return fe_log(energies[f]);
    return dct_task(row, tb, c);
}

}  // namespace fe
