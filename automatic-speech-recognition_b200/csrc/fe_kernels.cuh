// sm_100a kernels of the acoustic front-end.
//   K0  k_resample           speed perturbation (polyphase Kaiser-sinc) + requantise   utils/augmentation.py:6-31
//   K1  k_frames_to_statics  framing -> rFFT512 -> power -> mel -> log -> DCT          preprocess.py:72-82
//   K2  k_cmvn_delta_pack    per-utterance CMVN, delta, delta-delta, cube (L, D, 3)    preprocess.py:85-88
//   k_build_tiles            (utterance, first frame) table for K1's persistent tile loop
#pragma once
#include <cuda_runtime.h>
#include "fe_core.cuh"

namespace fe {

struct UttDesc {
    long long pcm_off;      // element offset of the samples K1 frames (source or K0 scratch)
    long long src_off;      // element offset in the caller's PCM buffer
    long long stat_off;     // float offset of frame 0's statics
    long long out_off;      // float offset of the utterance's output
    int n_samples;          // samples K1 frames (after speed perturbation)
    int n_src;              // samples in the caller's buffer
    int n_frames;
    int src_sel;            // 0: caller's PCM, 1: K0 scratch (int16)
    int speed_idx;          // -1 = none
    float gain;             // 1 = none
};

struct DevTables {
    const float2* tw256;    // [256]
    const float2* tw512;    // [257]
    const float* window;    // [13*32] permuted frame layout, or nullptr
    const int* fb_start; const int* fb_bin0; const float* fb_w;
    const float* dct;       // [D][dct_stride]
    int nf, nnz, D, dct_stride, full_spectrum, is_mfcc, fbank_log, dc_elim;
};

// ---------------------------------------------------------------------------
__global__ void k_build_tiles(const UttDesc* __restrict__ utts, const long long* __restrict__ tile_prefix,
                              int n_utts, int frames_per_tile, int2* __restrict__ tiles) {
    int u = blockIdx.x * blockDim.x + threadIdx.x;
    if (u >= n_utts) return;
    long long b = tile_prefix[u];
    int nt = (utts[u].n_frames + frames_per_tile - 1) / frames_per_tile;
    for (int i = 0; i < nt; ++i) tiles[b + i] = make_int2(u, i * frames_per_tile);
}

// ---------------------------------------------------------------------------
// K1 shared-memory carve-up (bytes), shared by host (size) and device (pointers)
// ---------------------------------------------------------------------------
struct K1Smem {
    int off_tw256, off_tw512, off_window, off_fb_start, off_fb_bin0, off_fb_w, off_dct, off_warp;
    int warp_pcm_floats, warp_bytes, total;
};

__host__ __device__ inline int align16(int x) { return (x + 15) & ~15; }

__host__ __device__ inline K1Smem k1_smem_layout(int nf, int nnz, int D, int dct_stride, int has_window,
                                                 int frame_len, int hop, int is_mfcc) {
    K1Smem s;
    int o = 0;
    s.off_tw256 = o;    o = align16(o + 16 * kTw256Stride * 8);
    s.off_tw512 = o;    o = align16(o + kBins * 8);
    s.off_window = o;   o = align16(o + (has_window ? ((frame_len + 31) / 32) * 32 * 4 : 0));
    s.off_fb_start = o; o = align16(o + (nf + 1) * 4);
    s.off_fb_bin0 = o;  o = align16(o + nf * 4);
    s.off_fb_w = o;     o = align16(o + nnz * 4);
    s.off_dct = o;      o = align16(o + (is_mfcc ? D * dct_stride * 4 : 0));
    s.off_warp = o;
    s.warp_pcm_floats = (kWarpFrames - 1) * hop + ((frame_len + 31) / 32) * 32;
    s.warp_bytes = align16(s.warp_pcm_floats * 4) + kWarpFrames * kERegion * 4 + 64 * 4;
    s.total = o + kCtaWarps * s.warp_bytes;
    return s;
}

// one staged sample of utterance `u`: gain + requantise (int16 path) and scale to [-1, 1)
template <int PCM_F32>
__device__ __forceinline__ float fetch_sample(const void* __restrict__ src, long long idx, float gain) {
    if (PCM_F32) {
        float v = __ldg(reinterpret_cast<const float*>(src) + idx);
        if (gain != 1.f) v = fminf(fmaxf(v * gain, -1.f), 32767.f / 32768.f);
        return v;
    } else {
        float v = (float)__ldg(reinterpret_cast<const short*>(src) + idx);
        if (gain != 1.f) v = fminf(fmaxf(rintf(v * gain), -32768.f), 32767.f);
        return v * (1.0f / 32768.0f);
    }
}

template <int FRAME_LEN, int HOP, int PCM_F32>
__global__ void __launch_bounds__(kCtaWarps * 32, 2)
k_frames_to_statics(const void* __restrict__ pcm, const short* __restrict__ scratch,
                    const UttDesc* __restrict__ utts, const int2* __restrict__ tiles, int n_tiles,
                    DevTables dt, float* __restrict__ statics, float preemph) {
    static_assert(HOP % 32 == 0, "staged layout needs frame starts on 32-float blocks");
    static_assert(FRAME_LEN % 2 == 0 && FRAME_LEN <= kNfft, "frame must fit the 512-point FFT");
    extern __shared__ __align__(16) unsigned char smem[];
    const K1Smem L = k1_smem_layout(dt.nf, dt.nnz, dt.D, dt.dct_stride, dt.window != nullptr,
                                    FRAME_LEN, HOP, dt.is_mfcc);
    float2* s_tw256 = reinterpret_cast<float2*>(smem + L.off_tw256);
    float2* s_tw512 = reinterpret_cast<float2*>(smem + L.off_tw512);
    float* s_window = reinterpret_cast<float*>(smem + L.off_window);
    int* s_fb_start = reinterpret_cast<int*>(smem + L.off_fb_start);
    int* s_fb_bin0 = reinterpret_cast<int*>(smem + L.off_fb_bin0);
    float* s_fb_w = reinterpret_cast<float*>(smem + L.off_fb_w);
    float* s_dct = reinterpret_cast<float*>(smem + L.off_dct);

    const int tid = threadIdx.x;
    for (int i = tid; i < 256; i += blockDim.x) s_tw256[(i >> 4) * kTw256Stride + (i & 15)] = dt.tw256[i];
    for (int i = tid; i < kBins; i += blockDim.x) s_tw512[i] = dt.tw512[i];
    if (dt.window) for (int i = tid; i < ((FRAME_LEN + 31) / 32) * 32; i += blockDim.x) s_window[i] = dt.window[i];
    for (int i = tid; i <= dt.nf; i += blockDim.x) s_fb_start[i] = dt.fb_start[i];
    for (int i = tid; i < dt.nf; i += blockDim.x) s_fb_bin0[i] = dt.fb_bin0[i];
    for (int i = tid; i < dt.nnz; i += blockDim.x) s_fb_w[i] = dt.fb_w[i];
    if (dt.is_mfcc) for (int i = tid; i < dt.D * dt.dct_stride; i += blockDim.x) s_dct[i] = dt.dct[i];
    __syncthreads();

    SmemTables tb;
    tb.tw256 = s_tw256; tb.tw512 = s_tw512; tb.window = dt.window ? s_window : nullptr;
    tb.fb_start = s_fb_start; tb.fb_bin0 = s_fb_bin0; tb.fb_w = s_fb_w; tb.dct = s_dct;
    tb.nf = dt.nf; tb.D = dt.D; tb.dct_stride = dt.dct_stride; tb.full_spectrum = dt.full_spectrum;
    tb.is_mfcc = dt.is_mfcc; tb.fbank_log = dt.fbank_log; tb.dc_elim = dt.dc_elim;

    const int warp = tid >> 5, lane = tid & 31;
    const int fs = lane >> 3, t = lane & 7;
    unsigned char* wbase = smem + L.off_warp + warp * L.warp_bytes;
    float* pcm_w = reinterpret_cast<float*>(wbase);
    float* e_w = reinterpret_cast<float*>(wbase + align16(L.warp_pcm_floats * 4));
    float* scr_w = e_w + kWarpFrames * kERegion;      // [0..31] sum-of-squares partials, [32..35] energies
    const int D = dt.D;
    const int nf4 = (dt.nf + 3) & ~3;

    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int2 te = tiles[tile];
        const UttDesc u = utts[te.x];
        const int f0 = te.y + warp * kWarpFrames;
        const int nfw = min(kWarpFrames, u.n_frames - f0);
        if (nfw <= 0) continue;                                   // warp-uniform

        // ---- phase 0: stage PCM of frames [f0, f0+nfw) as float, permuted 32-blocks ----
        {
            const long long s0 = (long long)f0 * HOP;
            const int n_samp = (nfw - 1) * HOP + FRAME_LEN;
            const int items = ((n_samp + 31) >> 5) << 3;
            const void* src = u.src_sel ? (const void*)(scratch + u.pcm_off)
                                        : (PCM_F32 ? (const void*)(reinterpret_cast<const float*>(pcm) + u.pcm_off)
                                                   : (const void*)(reinterpret_cast<const short*>(pcm) + u.pcm_off));
            const bool f32 = PCM_F32 && !u.src_sel;
            const bool plain = (u.gain == 1.f) && (preemph == 0.f);
            for (int id = lane; id < items; id += 32) {
                const int sA = ((id >> 3) << 5) + ((id & 7) << 1), sB = sA + 16;
                if (plain) {
                    if (f32) stage_item_f32(reinterpret_cast<const float*>(src) + s0, n_samp, id, pcm_w);
                    else     stage_item_i16(reinterpret_cast<const short*>(src) + s0, n_samp, id, pcm_w);
                } else {
                    // gain / pre-emphasis path: sample by sample (pre-emphasis is circular over the utterance)
                    float v[4]; const int sidx[4] = {sA, sA + 1, sB, sB + 1};
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        v[e] = 0.f;
                        if (sidx[e] < n_samp) {
                            long long n = s0 + sidx[e];
                            float x = f32 ? fetch_sample<1>(src, n, u.gain) : fetch_sample<0>(src, n, u.gain);
                            if (preemph != 0.f) {
                                long long pn = n > 0 ? n - 1 : (long long)u.n_samples - 1;
                                float xp = f32 ? fetch_sample<1>(src, pn, u.gain) : fetch_sample<0>(src, pn, u.gain);
                                x = x - preemph * xp;
                            }
                            v[e] = x;
                        }
                    }
                    stage_store(pcm_w, id, v[0], v[1], v[2], v[3]);
                }
            }
        }
        __syncwarp();

        const bool active = fs < nfw;
        float* e_f = e_w + fs * kERegion;
        float* p_f = e_f + fs * kPStagger;

        // ---- phase 1: stage A ----
        if (active) {
            float ss = stage_a<FRAME_LEN>(pcm_w + fs * HOP, e_f, tb, t, fs);
            scr_w[lane] = ss;
        }
        __syncwarp();

        // ---- phase 2: stage B ----
        LaneZ z;
        if (active) stage_b(e_f, z, t, fs);
        __syncwarp();

        // ---- phase 3: post-pass, power row, frame energy ----
        if (active) {
            float x0, x256;
            post_pass(z, p_f, tb, t, x0, x256);
            if (t == 0) {
                float s = 0.f;
#pragma unroll
                for (int i = 0; i < 8; ++i) s += scr_w[fs * 8 + i];
                scr_w[32 + fs] = frame_energy(s, x0, x256);
            }
        }
        __syncwarp();

        // ---- phase 4: mel filterbank (+ log) into the per-frame row ----
        for (int id = lane; id < kWarpFrames * nf4; id += 32) mel_phase(e_w, tb, id, nfw);
        __syncwarp();

        // ---- phase 5: DCT (mfcc) or copy (fbank) -> statics[(f0 + f) * D + c] ----
        {
            float* dst = statics + u.stat_off + (long long)f0 * D;
            const int ntask = nfw * D;
            for (int id = lane; id < ntask; id += 32) dst[id] = emit_phase(e_w, scr_w + 32, tb, id);
        }
        __syncwarp();
    }
}

// ---------------------------------------------------------------------------
// K2: per-utterance CMVN (two-pass mean / population std, eps = 2^-30), deltas,
// cube pack.  One CTA per utterance (grid-stride).
// ---------------------------------------------------------------------------
constexpr int kK2Threads = 256;

__host__ __device__ inline int k2_rows_per_chunk(int D) {
    int rb = 8192 / (3 * D);
    if (rb > 64) rb = 64;
    rb &= ~3;
    return rb < 4 ? 4 : rb;
}
__host__ __device__ inline int k2_smem_floats(int D) {
    int rb = k2_rows_per_chunk(D);
    return kK2Threads + 2 * D + (rb + 8) * D + (rb + 4) * D + rb * 3 * D + 8;
}

__global__ void __launch_bounds__(kK2Threads)
k_cmvn_delta_pack(const UttDesc* __restrict__ utts, int n_utts, const float* __restrict__ statics,
                  float* __restrict__ out, int D, int delta_mode) {
    extern __shared__ __align__(16) float sm2[];
    const int RB = k2_rows_per_chunk(D);
    float* red = sm2;
    float* mean = red + kK2Threads;
    float* inv = mean + D;
    float* cube = inv + D + ((4 - ((kK2Threads + 2 * D) & 3)) & 3);   // 16-byte aligned
    float* vt = cube + RB * 3 * D;
    float* d1 = vt + (RB + 8) * D;
    const int tid = threadIdx.x;
    const int R = kK2Threads / D;
    const int c = tid % D, r = tid / D;
    const bool act = r < R;
    const int cp1 = min(c + 1, D - 1), cp2 = min(c + 2, D - 1);

    for (int ui = blockIdx.x; ui < n_utts; ui += gridDim.x) {
        const int L = utts[ui].n_frames;
        if (L <= 0) continue;
        const float* x = statics + utts[ui].stat_off;
        float* o = out + utts[ui].out_off;

        float s = 0.f;
        if (act) for (int tt = r; tt < L; tt += R) s += x[(long long)tt * D + c];
        red[tid] = s;
        __syncthreads();
        if (tid < D) {
            float m = 0.f;
            for (int rr = 0; rr < R; ++rr) m += red[rr * D + tid];
            mean[tid] = m / (float)L;
        }
        __syncthreads();
        const float mu = act ? mean[c] : 0.f;
        float q = 0.f;
        if (act) for (int tt = r; tt < L; tt += R) { float d = x[(long long)tt * D + c] - mu; q = fmaf(d, d, q); }
        red[tid] = q;
        __syncthreads();
        if (tid < D) {
            float v = 0.f;
            for (int rr = 0; rr < R; ++rr) v += red[rr * D + tid];
            inv[tid] = 1.0f / (sqrtf(v / (float)L) + 9.313225746154785e-10f);   // 2^-30
        }
        __syncthreads();
        const float iv = act ? inv[c] : 0.f;

        for (int t0 = 0; t0 < L; t0 += RB) {
            const int nrow = min(RB, L - t0);
            if (delta_mode == 0) {
                // speechpy as shipped: deltas slide along the coefficient axis of the same frame
                if (act) for (int row = r; row < nrow; row += R)
                    vt[row * D + c] = (x[(long long)(t0 + row) * D + c] - mu) * iv;
                __syncthreads();
                if (act) for (int row = r; row < nrow; row += R)
                    d1[row * D + c] = (vt[row * D + cp1] + 2.f * vt[row * D + cp2]) / 10.f;
                __syncthreads();
                if (act) for (int row = r; row < nrow; row += R) {
                    float dd = (d1[row * D + cp1] + 2.f * d1[row * D + cp2]) / 10.f;
                    float* q3 = cube + (row * D + c) * 3;
                    q3[0] = vt[row * D + c]; q3[1] = d1[row * D + c]; q3[2] = dd;
                }
            } else {
                // textbook regression along time, edge replication; vt row i <-> frame clamp(t0-4+i),
                // d1 row i <-> frame clamp(t0-2+i)
                if (act) for (int i = r; i < nrow + 8; i += R) {
                    int a = min(max(t0 - 4 + i, 0), L - 1);
                    vt[i * D + c] = (x[(long long)a * D + c] - mu) * iv;
                }
                __syncthreads();
                if (act) for (int i = r; i < nrow + 4; i += R) {
                    int sfr = min(max(t0 - 2 + i, 0), L - 1);
                    float acc = 0.f;
#pragma unroll
                    for (int k = 1; k <= 2; ++k) {
                        int ap = min(sfr + k, L - 1) - (t0 - 4), am = max(sfr - k, 0) - (t0 - 4);
                        acc += (float)k * (vt[ap * D + c] - vt[am * D + c]);
                    }
                    d1[i * D + c] = acc / 10.f;
                }
                __syncthreads();
                if (act) for (int row = r; row < nrow; row += R) {
                    int tt = t0 + row;
                    float acc = 0.f;
#pragma unroll
                    for (int k = 1; k <= 2; ++k) {
                        int ap = min(tt + k, L - 1) - (t0 - 2), am = max(tt - k, 0) - (t0 - 2);
                        acc += (float)k * (d1[ap * D + c] - d1[am * D + c]);
                    }
                    float* q3 = cube + (row * D + c) * 3;
                    q3[0] = vt[(row + 4) * D + c]; q3[1] = d1[(row + 2) * D + c]; q3[2] = acc / 10.f;
                }
            }
            __syncthreads();
            // coalesced copy of nrow * 3D floats; chunk start is 16-byte aligned (RB % 4 == 0)
            const int total = nrow * 3 * D;
            float* dst = o + (long long)t0 * 3 * D;
            const int n4 = total >> 2;
            for (int i = tid; i < n4; i += kK2Threads)
                reinterpret_cast<float4*>(dst)[i] = reinterpret_cast<const float4*>(cube)[i];
            for (int i = (n4 << 2) + tid; i < total; i += kK2Threads) dst[i] = cube[i];
            __syncthreads();
        }
    }
}

// ---------------------------------------------------------------------------
// K0: speed perturbation.  y[j] = sum_t taps[(j*down) % up][t] * x[floor(j*down/up) - 15 + t],
// x = 0 off the ends, int16 in -> round-half-even, saturate -> int16 out.
// ---------------------------------------------------------------------------
constexpr int kK0Outputs = 1024;

__global__ void __launch_bounds__(256)
k_resample(const short* __restrict__ pcm, const UttDesc* __restrict__ utts,
           const int2* __restrict__ atiles, int n_atiles,
           const int* __restrict__ sp_up, const int* __restrict__ sp_down, const int* __restrict__ sp_tap_off,
           const float* __restrict__ taps_all, short* __restrict__ dst, int use_dst_off) {
    for (int tile = blockIdx.x; tile < n_atiles; tile += gridDim.x) {
        const int2 te = atiles[tile];
        const UttDesc u = utts[te.x];
        const short* x = pcm + u.src_off;
        short* y = dst + (use_dst_off ? u.out_off : u.pcm_off);
        const int j1 = min(te.y + kK0Outputs, u.n_samples);
        if (u.speed_idx < 0) {
            for (int j = te.y + threadIdx.x; j < j1; j += blockDim.x) {
                float v = (float)x[j];
                if (u.gain != 1.f) v = fminf(fmaxf(rintf(v * u.gain), -32768.f), 32767.f);
                y[j] = (short)v;
            }
            continue;
        }
        const int up = sp_up[u.speed_idx], down = sp_down[u.speed_idx];
        const float* taps = taps_all + sp_tap_off[u.speed_idx];
        for (int j = te.y + threadIdx.x; j < j1; j += blockDim.x) {
            const long long pos = (long long)j * down;
            const int base = (int)(pos / up) - 15;
            const int ph = (int)(pos % up);
            const float* tp = taps + ph * 32;
            float acc = 0.f;
#pragma unroll 8
            for (int k = 0; k < 32; ++k) {
                int i = base + k;
                float xv = (i >= 0 && i < u.n_src) ? (float)__ldg(x + i) : 0.f;
                acc = fmaf(__ldg(tp + k), xv, acc);
            }
            if (u.gain != 1.f) acc *= u.gain;
            y[j] = (short)fminf(fmaxf(rintf(acc), -32768.f), 32767.f);
        }
    }
}

}  // namespace fe
