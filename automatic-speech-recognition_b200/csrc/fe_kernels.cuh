// sm_100a kernels of the acoustic front-end.
//   K0  k_resample           speed perturbation (polyphase Kaiser-sinc) + gain + requantise   utils/augmentation.py:6-56
//   K0' k_preemph            optional pre-emphasis (speechpy.processing.preemphasis) to float scratch
//   K1  k_frames_to_statics  framing -> rFFT512 -> power -> mel -> log -> DCT          preprocess.py:72-82
//   K2a k_utt_stats          per-utterance mean / std of the statics (CMVN)            preprocess.py:85
//   K2b k_norm_delta_pack    normalise, delta, delta-delta, cube (L, D, 3)             preprocess.py:86-88
//   k_build_tiles            tile descriptors for K1's persistent tile loop
#pragma once
#include <cuda_runtime.h>
#include "fe_core.cuh"
#include "fe_k1t.cuh"
#include "tmem_ops_gen.h"

namespace fe {

struct UttDesc {
    long long pcm_off;      // element offset of the samples K1 frames (source or scratch)
    long long src_off;      // element offset in the caller's PCM buffer
    long long stat_off;     // float offset of frame 0's statics
    long long out_off;      // float offset of the utterance's output
    int n_samples;          // samples K1 frames (after speed perturbation)
    int n_src;              // samples in the caller's buffer
    int n_frames;
    int src_sel;            // 0: caller's PCM, 1: scratch
    int speed_idx;          // -1 = none
    float gain;             // 1 = none
};

// one entry per tile (32 consecutive frames of one utterance; 48 bytes, three 16-byte loads)
struct __align__(16) TileDesc {
    long long pcm_off;      // element offset of the tile's first sample
    long long stat_off;     // float offset of the tile's statics block
    long long out_off;      // float offset of the utterance's output (frame 0)
    int n_frames;           // 1..32
    int src_sel;
    int utt;
    int first_frame;        // of this tile inside the utterance
    int utt_frames;
    int pad;
};

struct DevTables {
    const float4* tw256;    // [15][16]
    const float4* tw512;    // [8][16]
    const float2* window;   // [ROWS*16] or nullptr
    const int* mel_desc; const float* mel_w;
    const float* dctf;      // [D][dct_stride]
    int mel_groups, p_rows;
    int epi_off[kEpiWarps], epi_cnt[kEpiWarps];
    int nf, D, dct_stride, nh, full_spectrum, is_mfcc, fbank_log, dc_elim;
    float pscale;
};

// ---------------------------------------------------------------------------
// Statics between K1 and K2 are stored tile by tile, coefficient-major: tile j of an utterance
// is a [D][32] block (frame 32 j + l of coefficient c at block[c * 32 + l]); the last block of
// an utterance is padded to 32 frames.  K1's epilogue writes it with one 128-byte store per
// (warp, coefficient); K2 reads it with the same coalescing.
// ---------------------------------------------------------------------------
__host__ __device__ inline long long statics_tiles(long long n_frames) { return (n_frames + kTileFrames - 1) / kTileFrames; }

__global__ void k_build_tiles(const UttDesc* __restrict__ utts, const long long* __restrict__ tile_prefix,
                              int n_utts, int hop, int D, TileDesc* __restrict__ tiles) {
    int u = blockIdx.x * blockDim.x + threadIdx.x;
    if (u >= n_utts) return;
    const UttDesc d = utts[u];
    long long b = tile_prefix[u];
    for (int f = 0; f < d.n_frames; f += kTileFrames) {
        TileDesc t;
        t.pcm_off = d.pcm_off + (long long)f * hop;
        t.stat_off = d.stat_off + (long long)f * D;        // f is a multiple of 32: whole [D][32] blocks
        t.n_frames = min(kTileFrames, d.n_frames - f);
        t.out_off = d.out_off; t.first_frame = f; t.utt_frames = d.n_frames;
        t.src_sel = d.src_sel; t.utt = u; t.pad = 0;
        tiles[b++] = t;
    }
}

// K0's tile table (one entry per kK0Outputs output samples of a perturbed utterance), grouped by resampler
// class on the host: a_start[u] = index of the utterance's first entry (< 0: none)
__global__ void k_build_atiles(const UttDesc* __restrict__ utts, const long long* __restrict__ a_start, int n_utts,
                               int tile_outputs, int2* __restrict__ atiles) {
    int u = blockIdx.x * blockDim.x + threadIdx.x;
    if (u >= n_utts) return;
    long long b = a_start[u];
    if (b < 0) return;
    const int n = utts[u].n_samples;
    for (int j = 0; j < n; j += tile_outputs) atiles[b++] = make_int2(u, j);
}

// ---------------------------------------------------------------------------
// K1 shared-memory carve-up (bytes), shared by host (size) and device (pointers).
// The exchange regions come first so that they are 2 KB aligned (the kernel rounds
// the dynamic shared base up to 2 KB; the host adds 2 KB of slack).
// ---------------------------------------------------------------------------
constexpr int kK1Threads = (kFftWarps + kEpiWarps) * 32;
constexpr int kMaxSlots = 4;            // power-buffer slots between the FFT warps and the epilogue warps
constexpr int kK1SmemMax = 227 * 1024;

struct K1Smem {
    int off_e, off_raw, off_scr, off_tw256, off_tw512, off_window, off_desc, off_melw, off_dct, off_pbuf, off_sd,
        off_energy, off_bar;
    int raw_bytes;          // one raw buffer of one warp
    int raw_bufs;           // 2 = double-buffered bulk copies, 1 when shared memory is short (float PCM)
    int slots;              // power-buffer slots (3, or 2 when the full 257-bin spectrum is kept)
    int pbuf_floats, sd_floats;     // per slot / per sd buffer
    int total;
};

__host__ __device__ inline int align16(int x) { return (x + 15) & ~15; }

__host__ __device__ inline K1Smem k1_smem_layout_n(int nf, int mel_groups, int p_rows, int D, int dct_stride, int has_window,
                                                   int frame_len, int hop, int is_mfcc, int in_f32, int raw_bufs, int slots,
                                                   int spec) {
    K1Smem s;
    s.raw_bufs = raw_bufs; s.slots = slots;
    const int rows = (frame_len + 31) / 32;
    int o = 0;
    s.off_e = o;      o += kFftWarps * kWarpFrames * kERegion * 4;
    s.raw_bytes = align16(((kWarpFrames - 1) * hop + rows * 32) * (in_f32 ? 4 : 2));
    s.off_raw = o;    o += kFftWarps * raw_bufs * s.raw_bytes;
    s.off_scr = o;    o += kFftWarps * 32 * 4;
    s.off_tw256 = o;  o += 15 * 16 * 16;
    s.off_tw512 = o;  o += 8 * 16 * 16;
    s.off_window = o; o = align16(o + (has_window ? rows * 16 * 8 : 0));
    // (the specialised epilogue reads its weights from the kernel parameters and keeps log-mel in registers)
    s.off_desc = o;   o = align16(o + (spec ? 0 : (mel_groups > 0 ? mel_groups : 1) * 4));
    s.off_melw = o;   o = align16(o + (spec ? 0 : (mel_groups > 0 ? mel_groups : 1) * 16));
    s.off_dct = o;    o = align16(o + (is_mfcc && !spec ? D * dct_stride * 4 : 0));
    s.pbuf_floats = p_rows * kPStride;
    s.off_pbuf = o;   o = align16(o + slots * s.pbuf_floats * 4);
    s.sd_floats = is_mfcc && !spec ? (nf + 4) * 32 : 0; // log-mel rows of one tile
    s.off_sd = o;     o = align16(o + 2 * s.sd_floats * 4);
    s.off_energy = o; o += kMaxSlots * kTileFrames * 4;
    s.off_bar = o;    o += (kFftWarps * 2 + 2 * kMaxSlots) * 8;
    s.total = o + 2048;     // slack for the 2 KB round-up
    return s;
}

// the richest buffering that fits one SM's shared memory (total > kK1SmemMax: the configuration does not fit).
// spec: layout for a specialised epilogue -- consumer warp w owns slot w, so exactly 4 slots are needed.
__host__ inline K1Smem k1_smem_layout(int nf, int mel_groups, int p_rows, int D, int dct_stride, int has_window,
                                      int frame_len, int hop, int is_mfcc, int in_f32, int spec) {
    const int gen[4][2] = {{2, 3}, {2, 2}, {1, 3}, {1, 2}}, sp[2][2] = {{2, 4}, {1, 4}};
    K1Smem s{};
    for (int i = 0; i < (spec ? 2 : 4); ++i) {
        const int* o = spec ? sp[i] : gen[i];
        s = k1_smem_layout_n(nf, mel_groups, p_rows, D, dct_stride, has_window, frame_len, hop, is_mfcc, in_f32, o[0], o[1], spec);
        if (s.total <= kK1SmemMax) break;
    }
    return s;
}

// ---------------------------------------------------------------------------
// mbarrier / bulk-copy (TMA 1-D) helpers
// ---------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(bar) : "memory");
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 :: "r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ bool mbar_try(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    return ok != 0;
}
// K1 is bound by instruction issue, so a waiting warp must not spin: after a failed try it sleeps
// (the 12 FFT warps and the other consumers keep the issue ports busy meanwhile)
template <int NS = 128>
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    if (mbar_try(bar, parity)) return;
    do { __nanosleep(NS); } while (!mbar_try(bar, parity));
}

// "slot is full" goes through hardware named barriers (ids 2 .. 2 + slots - 1): the 8 producer warps of a tile
// bar.arrive (non-blocking), the consumer warp(s) bar.sync -- a blocked warp costs no issue slots, whereas a
// consumer polling an mbarrier burnt ~8 % of the SM's issue cycles.  "slot is empty" stays an mbarrier: the
// producers must not wait for each other there.
template <int THREADS>
__device__ __forceinline__ void full_arrive(int slot) {
    asm volatile("bar.arrive %0, %1;" :: "r"(2 + slot), "n"(THREADS) : "memory");
}
template <int THREADS>
__device__ __forceinline__ void full_wait(int slot) {
    asm volatile("bar.sync %0, %1;" :: "r"(2 + slot), "n"(THREADS) : "memory");
}

// everything K1 needs besides the data pointers; lives in the constant bank
constexpr int kEpiWCap = 1024;           // floats of specialised-epilogue weights carried in the kernel parameters

struct K1Params {
    DevTables dt;
    K1Smem L;
    float epi_w[kEpiWCap];          // EPI != 0: mel CSR weights + folded DCT rows (constant-bank operands)
    int dbg;                        // phase ablation for profiling (FE_K1_DBG): 1 = no mel/DCT
};

#define FE_OPAQUE(v) asm volatile("" : "+r"(v))

// FE_K1_DBG & 8: per-phase clock64 totals of lane 0 of every warp (profiling aid, read with fe_debug_counters)
//   [0] epilogue: wait for a full slot  [1] mel  [2] named barrier  [3] DCT  [4] tiles
//   [8] FFT: wait for raw samples  [9] stage A  [10] stage B  [11] wait for an empty slot  [12] post-pass  [13] groups
// Compiled in only with -DFE_K1_PROF (the counters cost 14 registers in the FFT loop).
__device__ unsigned long long g_k1_prof[16];
#ifdef FE_K1_PROF
#define FE_PROF_DECL(n) long long prof[n] = {0}, tick = clock64()
#define FE_TICK(slot)                                                             \
    do { if (P.dbg & 8) { const long long now__ = clock64(); prof[slot] += now__ - tick; tick = now__; } } while (0)
#else
#define FE_PROF_DECL(n)
#define FE_TICK(slot) do { } while (0)
#endif

// ---------------------------------------------------------------------------
// K1.  Persistent, warp-specialised: one CTA of 16 warps per SM.
//   warps 0..11  (three warpgroups, 144 registers/thread after setmaxnreg): FFT producers.  Each pass
//     takes one 4-frame group (8 lanes per frame): TMA bulk copy of the raw samples (double-buffered per
//     warp) -> stage A -> exchange -> stage B -> real-FFT split -> the group's 4 columns of a [bin][frame]
//     power-buffer slot.  No CTA barrier: producers only meet the consumers through mbarriers.
//   warps 12..15 (one warpgroup, 80 registers/thread): epilogue consumers.  For every tile (8 groups =
//     32 frames) lane = frame, warp = filter / coefficient group: mel (+ log) -> log-mel rows ->
//     named barrier of the 4 warps -> DCT -> the tile's [D][32] statics block (128-byte stores).
//   tile k of the CTA lives in slot k % 3; full[slot] (8 arrivals, one per group) / empty[slot] (4).
// The latency-bound, LSU-heavy epilogue so runs in the issue slots the FP32-bound FFT warps leave idle.
// ---------------------------------------------------------------------------
// EPI: 0 = generic epilogue (run-time mel plan, the 4 consumer warps share every tile),
//      1 / 2 / 3 / 4 = specialised for the reference's filterbanks: 40 filters -> 13 cepstra, fbank-80, 40 filters -> 39
//      cepstra (run.sh's default feat_dim), fbank-40: consumer warp w takes tiles w, w + 4, ... whole.
template <int FRAME_LEN, int HOP, int IN_F32, int HAS_WINDOW, int EPI>
__global__ void __launch_bounds__(kK1Threads, 1)
k_frames_to_statics(const void* __restrict__ pcm, const void* __restrict__ scratch,
                    const TileDesc* __restrict__ tiles, int n_tiles,
                    const __grid_constant__ K1Params P, float* __restrict__ statics) {
    static_assert(HOP % 8 == 0 && FRAME_LEN % 8 == 0 && FRAME_LEN <= kNfft, "bulk copies need 16-byte granules");
    extern __shared__ unsigned char smem_dyn[];
    const DevTables& dt = P.dt;
    const K1Smem& L = P.L;
    unsigned char* smem = smem_dyn + ((2048u - (smem_u32(smem_dyn) & 2047u)) & 2047u);
    float4* s_tw256 = reinterpret_cast<float4*>(smem + L.off_tw256);
    float4* s_tw512 = reinterpret_cast<float4*>(smem + L.off_tw512);
    float2* s_window = reinterpret_cast<float2*>(smem + L.off_window);
    int* s_desc = reinterpret_cast<int*>(smem + L.off_desc);
    float* s_melw = reinterpret_cast<float*>(smem + L.off_melw);
    float* s_dct = reinterpret_cast<float*>(smem + L.off_dct);
    float* s_pbuf = reinterpret_cast<float*>(smem + L.off_pbuf);
    float* s_sd = reinterpret_cast<float*>(smem + L.off_sd);
    float* s_energy = reinterpret_cast<float*>(smem + L.off_energy);

    const int tid = threadIdx.x;
    constexpr int ROWS = (FRAME_LEN + 31) / 32;
    for (int i = tid; i < 15 * 16; i += blockDim.x) s_tw256[i] = dt.tw256[i];
    for (int i = tid; i < 8 * 16; i += blockDim.x) s_tw512[i] = dt.tw512[i];
    if (dt.window) for (int i = tid; i < ROWS * 16; i += blockDim.x) s_window[i] = dt.window[i];
    if (EPI == 0) {
        for (int i = tid; i < dt.mel_groups; i += blockDim.x) s_desc[i] = dt.mel_desc[i];
        for (int i = tid; i < dt.mel_groups * 4; i += blockDim.x) s_melw[i] = dt.mel_w[i];
        if (dt.is_mfcc) for (int i = tid; i < dt.D * dt.dct_stride; i += blockDim.x) s_dct[i] = dt.dctf[i];
    }
    // pad rows of the power slots and of the folded rows are read (times zero weights): keep them finite
    for (int i = tid; i < L.slots * L.pbuf_floats; i += blockDim.x) s_pbuf[i] = 0.f;
    for (int i = tid; i < 2 * L.sd_floats; i += blockDim.x) s_sd[i] = 0.f;
    for (int i = tid; i < kMaxSlots * kTileFrames; i += blockDim.x) s_energy[i] = 1.f;

    int lane = tid & 31, warp = tid >> 5;
    const uint32_t bar_base = smem_u32(smem + L.off_bar);
    const uint32_t bar_full = bar_base + kFftWarps * 16, bar_empty = bar_full + kMaxSlots * 8;
    if (tid == 0) {
        for (int i = 0; i < kFftWarps * 2; ++i) mbar_init(bar_base + 8 * i, 1);
        for (int i = 0; i < kMaxSlots; ++i) { mbar_init(bar_full + 8 * i, kTileGroups); mbar_init(bar_empty + 8 * i, EPI ? 1 : kEpiWarps); }
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    __syncthreads();

    SmemTables tb;
    tb.tw256 = s_tw256; tb.tw512 = s_tw512; tb.window = dt.window ? s_window : nullptr;
    tb.mel_desc = s_desc; tb.mel_w4 = reinterpret_cast<const float4*>(s_melw); tb.dctf = s_dct;
    tb.nf = dt.nf; tb.D = dt.D; tb.dct_stride = dt.dct_stride; tb.nh = dt.nh;
    tb.full_spectrum = EPI ? 0 : dt.full_spectrum;        // the baked plans only use bins <= 128
    tb.is_mfcc = dt.is_mfcc; tb.fbank_log = dt.fbank_log;
    tb.dc_elim = dt.dc_elim; tb.pscale = dt.pscale;

    int bx = blockIdx.x, gx = gridDim.x;
    FE_OPAQUE(bx); FE_OPAQUE(gx);
    const int nk = (n_tiles - bx + gx - 1) / gx;      // tiles of this CTA
    constexpr int kFullThreads = (kTileGroups + (EPI ? 1 : kEpiWarps)) * 32;          // producers + consumer(s) of a tile

    if (warp >= kFftWarps) {
        // =========================== epilogue warpgroup ===========================
        asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;\n" :: "n"(80));
        const int we = warp - kFftWarps;
        if (EPI != 0) {
            // one warp per tile: k = we, we + 4, ... in slot `we` (4 slots: a warp sees every phase of its barrier in order)
            const int slot = we;
            uint32_t par = 0;
            for (int k = we; k < nk; k += kEpiWarps, par ^= 1u) {
                float* out_t = statics + tiles[blockIdx.x + (long long)k * gridDim.x].stat_off;
                const float* pb = s_pbuf + slot * L.pbuf_floats;
                full_wait<kFullThreads>(slot);
                if (!(P.dbg & 1)) {
                    const float* en = s_energy + slot * kTileFrames;
                    if (EPI == 1) epi_tile_spec<PlanMfcc40, 13, true, true>(pb, en, out_t, P.epi_w, tb.dc_elim, lane);
                    else if (EPI == 3) epi_tile_spec<PlanMfcc40, 39, true, true>(pb, en, out_t, P.epi_w, tb.dc_elim, lane);
                    else if (EPI == 2 && tb.fbank_log) epi_tile_spec<PlanFbank80, 80, false, true>(pb, en, out_t, P.epi_w, false, lane);
                    else if (EPI == 2) epi_tile_spec<PlanFbank80, 80, false, false>(pb, en, out_t, P.epi_w, false, lane);
                    else if (tb.fbank_log) epi_tile_spec<PlanMfcc40, 40, false, true>(pb, en, out_t, P.epi_w, false, lane);
                    else epi_tile_spec<PlanMfcc40, 40, false, false>(pb, en, out_t, P.epi_w, false, lane);
                } else {
                    out_t[lane] = pb[(5 + (lane & 7)) * kPStride + lane] + s_energy[slot * kTileFrames + lane];
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(bar_empty + 8 * slot);
            }
            return;
        }
        const int my_off = dt.epi_off[we], my_cnt = dt.epi_cnt[we];        // this warp's slice of the mel plan
        int slot = 0;
        uint32_t par = 0;
        long long stat_next = tiles[blockIdx.x].stat_off;
        FE_PROF_DECL(5);
        for (int k = 0; k < nk; ++k) {
            float* out_t = statics + stat_next;
            if (k + 1 < nk) stat_next = tiles[blockIdx.x + (long long)(k + 1) * gridDim.x].stat_off;
            float* sd = s_sd + (k & 1) * L.sd_floats;
            const float* pb = s_pbuf + slot * L.pbuf_floats;
            full_wait<kFullThreads>(slot);
            FE_TICK(0);
            if (!(P.dbg & 1)) {
                // ---- phase 4: mel filterbank (+ log, + fold), lane = frame ----
                if (tb.is_mfcc) epi_mel(pb, sd, tb, my_off, my_cnt, lane);        // two calls: shared vs global stores
                else epi_mel(pb, out_t, tb, my_off, my_cnt, lane);
                FE_TICK(1);
                if (tb.is_mfcc) {
                    asm volatile("bar.sync 1, %0;" :: "n"(kEpiWarps * 32) : "memory");
                    FE_TICK(2);
                    // ---- phase 5: DCT, lane = frame ----
                    epi_dct(sd, s_energy + slot * kTileFrames, out_t, tb, we, lane);
                    FE_TICK(3);
                }
            } else if (we == 0) {
                out_t[lane] = pb[(5 + (lane & 7)) * kPStride + lane] + s_energy[slot * kTileFrames + lane];
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_empty + 8 * slot);
            if (++slot == L.slots) { slot = 0; par ^= 1u; }
        }
#ifdef FE_K1_PROF
        if ((P.dbg & 8) && lane == 0) {
            for (int i = 0; i < 4; ++i) atomicAdd(&g_k1_prof[i], (unsigned long long)prof[i]);
            atomicAdd(&g_k1_prof[4], (unsigned long long)nk);
        }
#endif
        return;
    }

    // =============================== FFT warps ===============================
    asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;\n" :: "n"(144));
    // per-thread constants, pinned in registers (FE_OPAQUE stops the compiler from
    // re-deriving them from threadIdx inside the loop)
    FE_OPAQUE(lane); FE_OPAQUE(warp);
    const int fs = lane >> 3, t = lane & 7;
    constexpr int ESZ = IN_F32 ? 4 : 2;
    uint32_t o_e = (uint32_t)(smem - smem_dyn) + L.off_e + warp * (kWarpFrames * kERegion * 4);   // warp's exchange buffer
    uint32_t o_raw = (uint32_t)(smem - smem_dyn) + L.off_raw + warp * L.raw_bufs * L.raw_bytes;
    const int dbl = L.raw_bufs - 1;                                    // 1: double-buffered raw samples
    uint32_t o_scr = (uint32_t)(smem - smem_dyn) + L.off_scr + warp * 128;
    FE_OPAQUE(o_e); FE_OPAQUE(o_raw); FE_OPAQUE(o_scr);
    const uint32_t bar0 = bar_base + warp * 16;                        // two mbarriers per warp (raw double buffer)

    // group g of this CTA: tile blockIdx.x + (g >> 3) * gridDim.x, frames 4 (g & 7) .. of it.
    // The FFT warps need three fields of a tile descriptor only (4 registers instead of 12 per descriptor in flight).
    struct GDesc { long long pcm_off; int n_frames, src_sel; };
    const int n_groups = nk * kTileGroups;
    auto load_desc = [&](int g, GDesc& td) {
        td.n_frames = 0; td.pcm_off = 0; td.src_sel = 0;
        if (g < n_groups) {
            const TileDesc* t = tiles + (bx + (long long)(g >> 3) * gx);
            td.pcm_off = t->pcm_off;
            const int2 ns = *reinterpret_cast<const int2*>(&t->n_frames);       // n_frames, src_sel (8-byte aligned pair)
            td.n_frames = ns.x; td.src_sel = ns.y;
        }
    };
    // issue the bulk copy of group g's samples into raw buffer `buf`
    auto prefetch = [&](const GDesc& td, int g, int buf) {
        const int q = g & (kTileGroups - 1);
        const int nfw = (P.dbg & 4) ? 0 : min(kWarpFrames, td.n_frames - q * kWarpFrames);
        if (nfw > 0 && lane == 0) {
            const unsigned char* base = reinterpret_cast<const unsigned char*>(td.src_sel ? scratch : pcm);
            const unsigned char* src = base + (td.pcm_off + (long long)q * kWarpFrames * HOP) * ESZ;
            const uint32_t bytes = (uint32_t)(((nfw - 1) * HOP + FRAME_LEN) * ESZ);
            mbar_expect_tx(bar0 + 8 * buf, bytes);
            bulk_g2s(smem_u32(smem_dyn + o_raw + buf * L.raw_bytes), src, bytes, bar0 + 8 * buf);
        }
    };

    // Raw samples run two passes ahead of the FFT when double-buffered: a buffer is refilled as soon as
    // stage A has consumed it (with the samples of the group two passes later), descriptors three ahead.
    int g = warp;
    GDesc cur, next, nn;
    load_desc(g, cur);
    prefetch(cur, g, 0);
    load_desc(g + kFftWarps, next);
    if (L.raw_bufs > 1) prefetch(next, g + kFftWarps, 1);
    load_desc(g + 2 * kFftWarps, nn);
    uint32_t phase = 0;      // bit b = parity to wait for on raw barrier b
    int buf = 0;
    // tile k = g >> 3 lives in slot k % slots; `use` = k / slots = how often the slot has been used before.
    // Tracked incrementally (g advances by 12 = one tile and a half): no integer division in the loop.
    int slot = (warp >> 3) % L.slots;
    uint32_t use = (uint32_t)((warp >> 3) / L.slots);
    FE_PROF_DECL(6);
    for (; g < n_groups; g += kFftWarps) {
        GDesc n3;                                               // descriptor three passes ahead: in flight during this pass
        load_desc(g + 3 * kFftWarps, n3);
        const int q = g & (kTileGroups - 1);
        const int nfw = (P.dbg & 4) ? 0 : min(kWarpFrames, cur.n_frames - q * kWarpFrames);
        if (nfw > 0) {
            float* e_w = reinterpret_cast<float*>(smem_dyn + o_e);
            float* scr_w = reinterpret_cast<float*>(smem_dyn + o_scr);       // the lanes' partial sums of squares
            float* e_f = e_w + fs * kERegion;
            FE_TICK(5);
            mbar_wait(bar0 + 8 * buf, (phase >> buf) & 1u);
            phase ^= 1u << buf;
            FE_TICK(0);
            // Lanes of frame slots beyond nfw (last, partial group of an utterance) run the same code on
            // stale shared memory: their columns of the power slot only feed their own (padding) lanes
            // of the epilogue.
            // ---- phase 1: stage A ----
            const unsigned char* raw_f = smem_dyn + o_raw + buf * L.raw_bytes + fs * HOP * ESZ;
            scr_w[lane] = stage_a<FRAME_LEN, IN_F32, HAS_WINDOW>(raw_f, e_f, tb, t, fs);
            __syncwarp();
            // the buffer is free again: refill it (double-buffered: for the group two passes ahead)
            if (dbl) prefetch(nn, g + 2 * kFftWarps, buf); else prefetch(next, g + kFftWarps, 0);
            FE_TICK(1);
            // ---- phase 2: stage B ----
            LaneZ z;
            stage_b(e_f, z, tb, t, fs);
            FE_TICK(2);
            // ---- phase 3: post-pass, power columns, frame energy (the slot must have been drained) ----
            if (use > 0) mbar_wait(bar_empty + 8 * slot, (use - 1u) & 1u);
            FE_TICK(3);
            float x0, x256;
            post_pass(z, s_pbuf + slot * L.pbuf_floats + q * kWarpFrames + fs, tb, t, fs, x0, x256);
            if (t == 0) {
                float s = 0.f;
#pragma unroll
                for (int i = 0; i < 8; ++i) s += scr_w[fs * 8 + i];
                s_energy[slot * kTileFrames + q * kWarpFrames + fs] = frame_energy(s, x0, x256, tb.pscale);
            }
            __syncwarp();
            FE_TICK(4);
        } else {
            // an empty group (partial last tile of an utterance) still has to arrive in the slot's CURRENT round
            if (dbl) prefetch(nn, g + 2 * kFftWarps, buf); else prefetch(next, g + kFftWarps, 0);
            if (use > 0) mbar_wait(bar_empty + 8 * slot, (use - 1u) & 1u);
        }
        full_arrive<kFullThreads>(slot);
        cur = next;
        next = nn;
        nn = n3;
        buf = (buf ^ 1) & dbl;
        slot += 1 + ((q + 4) >> 3);                             // g += 12: k advances by 1, or by 2 when q wraps
        if (slot >= L.slots) { slot -= L.slots; ++use; }
    }
#ifdef FE_K1_PROF
    if ((P.dbg & 8) && lane == 0) {
        for (int i = 0; i < 5; ++i) atomicAdd(&g_k1_prof[8 + i], (unsigned long long)prof[i]);
        atomicAdd(&g_k1_prof[13], (unsigned long long)((n_groups - warp + kFftWarps - 1) / kFftWarps));
        atomicAdd(&g_k1_prof[14], (unsigned long long)prof[5]);
    }
#endif
}

// ---------------------------------------------------------------------------
// K1T (fe_k1t.cuh): lane = frame, exchange in tensor memory.  One CTA of 8 warps per SM (persistent).  Warps q and
// q + 4 sit on the same SM sub-partition and own the same 32 tensor-memory lanes: together they are one "lane group"
// working on one tile (32 consecutive frames of one utterance) at a time -- two warps per scheduler, because a single
// warp cannot keep the issue port busy (profiles/r02_k1t_summary.md: 45 % issue-active with one).  The group splits
// the WORK, not the data:
//     stage A   warp 0: columns 0-3            warp 1: columns 4-15
//     -- pair barrier: exchange complete; warp 1 re-arms the raw buffer for the group's next tile --
//     stage B   warp 0: rows (0, 8), 1-4       warp 1: row pairs 5-7
//     -- pair barrier: power bins complete --
//     epilogue  warp 0: mel -> log -> DCT      warp 1: already in stage A of the next tile
// (shares chosen from the measured phase costs so that both warps finish a tile together).  A group needs nobody
// else: raw rows by per-lane bulk copies into its own shared-memory buffer, exchange in its own TMEM lanes, power
// bins in its own [bin][lane] buffer, statics straight to HBM (one 128-byte store per coefficient).
// Serves the configurations with a specialised epilogue plan, int16 PCM and the rectangular window (what the
// reference runs); everything else stays on k_frames_to_statics.
// ---------------------------------------------------------------------------
constexpr int kK1TGroups = 4;
constexpr int kK1TThreads = 2 * kK1TGroups * 32;
constexpr int kK1TRawBytes = 2 * kTRawBytes;                               // one group's raw samples, double-buffered
constexpr int kK1TPRows = 132;                                             // bins 0 .. 128 + 3 pad rows
constexpr int kK1TPbufBytes = kK1TPRows * kPStride * 4;                    // one group's power buffer: 19 008 bytes
constexpr int kK1TOffEnergy = kK1TGroups * (kK1TRawBytes + kK1TPbufBytes);
constexpr int kK1TOffSs = kK1TOffEnergy + kK1TGroups * kTileFrames * 4;
constexpr int kK1TOffBar = kK1TOffSs + kK1TGroups * kTileFrames * 4;
constexpr int kK1TSmem = kK1TOffBar + 128;
constexpr int kK1TSplitA = 1;            // stage A: warp 0 takes column quads [0, kK1TSplitA), warp 1 the rest
constexpr unsigned kK1TPairs0 = 0xD5u;   // stage B: warp 0 takes the row pairs whose bit is set (0, 2, 4, 6, 7), warp 1 the rest (1, 3, 5)

__constant__ float4 c_tw256p[128];       // packed stage-B twiddles of the 8 row pairs (fe_k1t.cuh: TTwiddles)
__constant__ float4 c_tw512p[64];        // packed real-FFT-split twiddles of pairs 1 .. 7
__constant__ float2 c_tw512[132];        // (cos, sin)(2 pi k / 512), k = 0 .. 128 (rows 0 and 8)

struct K1TParams {
    float epi_w[kEpiWCap];               // mel CSR weights (pre-scaled) + folded DCT rows: constant-bank operands
    float pscale;
    int fbank_log, dc_elim;
};

struct TmemExchange {
    uint32_t base;                       // tensor-memory address of this group's column 0 (lane field = 32 (warp % 4))
    __device__ __forceinline__ void st2(int col, float a, float b) const {
        asm volatile("tcgen05.st.sync.aligned.32x32b.x2.b32 [%0], {%1, %2};" :: "r"(base + (uint32_t)col), "f"(a), "f"(b) : "memory");
    }
    __device__ __forceinline__ void ld32(int col, float* v) const { tmem_ld32(base + (uint32_t)col, v); }
    __device__ __forceinline__ void wait_ld() const { tmem_wait_ld(); }
    __device__ __forceinline__ void wait_st() const { tmem_wait_st(); }
};

// the two warps of a lane group meet on a hardware named barrier (ids 1 .. 4); tensor-memory traffic is ordered
// across it by the tcgen05 fences
__device__ __forceinline__ void k1t_pair_sync(int group) {
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    asm volatile("bar.sync %0, 64;" :: "r"(1 + group) : "memory");
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}

#ifdef FE_K1_PROF
#define FE_TPROF_DECL long long prof[8] = {0}, tick = clock64()
#define FE_TTICK(i) do { const long long now__ = clock64(); prof[i] += now__ - tick; tick = now__; } while (0)
#else
#define FE_TPROF_DECL
#define FE_TTICK(i) do { } while (0)
#endif

template <int EPI>
__global__ void __launch_bounds__(kK1TThreads, 1)
k_frames_to_statics_t(const short* __restrict__ pcm, const short* __restrict__ scratch,
                      const TileDesc* __restrict__ tiles, int n_tiles,
                      const __grid_constant__ K1TParams P, float* __restrict__ statics) {
    extern __shared__ __align__(16) unsigned char smem_t[];
    __shared__ uint32_t s_tmem_base;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int group = warp & 3, half = warp >> 2;
    unsigned char* raw_w = smem_t + group * kK1TRawBytes;
    float* pbuf_w = reinterpret_cast<float*>(smem_t + kK1TGroups * kK1TRawBytes + group * kK1TPbufBytes);
    float* energy_w = reinterpret_cast<float*>(smem_t + kK1TOffEnergy) + group * kTileFrames;
    float* ss_w = reinterpret_cast<float*>(smem_t + kK1TOffSs) + group * kTileFrames;
    const uint32_t bar = smem_u32(smem_t + kK1TOffBar) + 16 * group;      // two mbarriers per group (raw double buffer)

    for (int i = tid; i < kK1TOffBar / 4; i += kK1TThreads) reinterpret_cast<uint32_t*>(smem_t)[i] = 0u;   // partial tiles read stale rows / bins
    if (half == 0 && lane == 0) { mbar_init(bar, 1); mbar_init(bar + 8, 1); }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(&s_tmem_base)), "r"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const TmemExchange ex{s_tmem_base + ((uint32_t)(group * 32) << 16)};
    const TTwiddles tw{c_tw256p, c_tw512p, c_tw512};
    float* pcol = pbuf_w + lane;

    const int stride = gridDim.x * kK1TGroups;
    int t = blockIdx.x * kK1TGroups + group;
    // raw samples of a tile: ONE bulk copy (the tile's frames overlap: (n - 1) * 160 + 400 samples), issued by lane 0
    // of warp 1 one tile ahead into the other buffer
    auto fetch = [&](int tile, int buf) {
        if (tile >= n_tiles || lane != 0) return;
        const TileDesc* td = tiles + tile;
        const long long off = td->pcm_off;
        const int2 ns = *reinterpret_cast<const int2*>(&td->n_frames);        // n_frames, src_sel
        const uint32_t bytes = (uint32_t)((ns.x - 1) * 160 + 400) * 2u;
        mbar_expect_tx(bar + 8 * buf, bytes);
        bulk_g2s(smem_u32(raw_w + buf * kTRawBytes), (ns.y ? scratch : pcm) + off, bytes, bar + 8 * buf);
    };
    if (half == 1) { fetch(t, 0); fetch(t + stride, 1); }
    uint32_t phase = 0;          // bit b = parity to wait for on buffer b
    int buf = 0;
    FE_TPROF_DECL;
    for (; t < n_tiles; t += stride) {
        float* out_t = statics + tiles[t].stat_off;
        FE_TTICK(7);
        mbar_wait<32>(bar + 8 * buf, (phase >> buf) & 1u);
        phase ^= 1u << buf;
        FE_TTICK(0);
        const uint4* raw4 = reinterpret_cast<const uint4*>(raw_w + buf * kTRawBytes) + lane * kTFrameVecs;
        const float ss = k1t_stage_a(raw4, ex, half == 0 ? 0 : kK1TSplitA, half == 0 ? kK1TSplitA : 4);
        if (half == 1) ss_w[lane] = ss;
        ex.wait_st();
        FE_TTICK(1);
        k1t_pair_sync(group);                                   // exchange complete, this buffer's samples consumed
        FE_TTICK(2);
        // the row-pair loop is the same for both warps (p is a uniform loop counter: the twiddles arrive through uniform
        // loads and are uniform-register operands of the packed FMAs); a warp skips the pairs of the other one
        const unsigned mine = half == 0 ? kK1TPairs0 : (~kK1TPairs0 & 0xffu);
        float x0 = 0.f, x256 = 0.f;
        if (half == 1) fetch(t + 2 * stride, buf);              // refill the drained buffer for the tile after the next one
#pragma unroll 1
        for (int p = 0; p < 8; ++p) {
            if (!((mine >> p) & 1u)) continue;
            float d0, d1;
            k1t_pair(ex, p, tw, pcol, d0, d1);
            if (p == 0) { x0 = d0; x256 = d1; }
        }
        if (half == 0) energy_w[lane] = frame_energy(ss + ss_w[lane], x0, x256, P.pscale);
        FE_TTICK(4);
        k1t_pair_sync(group);                                   // power bins complete
        FE_TTICK(5);
        if (half == 0) {
            // mel -> log -> DCT for this lane's frame (compile-time filterbank plan, weights in the constant bank)
            if (EPI == 1) epi_tile_spec<PlanMfcc40, 13, true, true>(pbuf_w, energy_w, out_t, P.epi_w, P.dc_elim != 0, lane);
            else if (P.fbank_log) epi_tile_spec<PlanFbank80, 80, false, true>(pbuf_w, energy_w, out_t, P.epi_w, false, lane);
            else epi_tile_spec<PlanFbank80, 80, false, false>(pbuf_w, energy_w, out_t, P.epi_w, false, lane);
            FE_TTICK(6);
        }
        buf ^= 1;
    }
#ifdef FE_K1_PROF
    if (lane == 0) for (int i = 0; i < 8; ++i) atomicAdd(&g_k1_prof[8 * half + i], (unsigned long long)prof[i]);
#endif
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(s_tmem_base), "r"(512) : "memory");
}

// ---------------------------------------------------------------------------
// K1U: K1T's lane = frame phases on 16 FFT warps + 4 epilogue warps per SM, written so that ptxas keeps every
// warp-uniform quantity in the UNIFORM datapath.
//
// Lane group g (= SM sub-partition g = tensor-memory lanes 32 g .. 32 g + 31) owns one tile at a time; its FOUR FFT
// warps (g, g + 4, g + 8, g + 12) split the tile's work evenly -- stage A: one column quad each, stage B: two row pairs
// each -- and warp 16 + g runs the mel -> log -> DCT epilogue of tile i while the FFT warps are already in tile i + 1
// (power bins double-buffered).  Five resident warps per scheduler instead of K1T's two.
//     FFT warps:  wait raw[b], "exchange drained" -> stage A (quad s) -> group barrier -> refill raw[b] (tile i + 2)
//                 -> wait pbuf[b] empty -> stage B (pairs 2 s, 2 s + 1; arrive "exchange drained" after the last load)
//                 -> arrive "pbuf[b] full"
//     epilogue:   wait "pbuf[b] full" -> energy, mel, log, DCT -> statics -> arrive "pbuf[b] empty"
//
// What ptxas needs for uniform-register operands (found with small probe kernels and cuobjdump; profiles/r02b_k1u.md):
//   * a value derived from threadIdx is never uniform to it; loop counters with constant or kernel-parameter bounds
//     are.  So the roles are entered through `for (ug) for (us) if (vote(ug == group && us == sub))`: inside, (ug, us)
//     are uniform loop counters and the tile index, buffer addresses, tensor-memory addresses, barrier ids and twiddle
//     indices derived from them are uniform-register arithmetic;
//   * a branch or loop exit whose predicate is not provably uniform makes everything it encloses "divergent" as soon as
//     it contains an `.aligned` instruction (tcgen05.ld / .st) or a loop: role tests and mbarrier spins therefore go
//     through `vote.all`, which yields a uniform predicate.
// With that, stage B's twiddles are `FFMA2 R, R, UR.F32x2, R` (2 issue cycles instead of 3 with three register pairs)
// and tcgen05 addresses need no per-instruction R2UR.
// ---------------------------------------------------------------------------
constexpr int kUGroups = 4;
constexpr int kUSub = 4;
constexpr int kUFftWarps = kUGroups * kUSub;
constexpr int kUThreads = (kUFftWarps + kUGroups) * 32;                    // 640
constexpr int kUPS = 32;                                                   // power-buffer row stride: [bin][lane]
constexpr int kUPBin0 = 5, kUPRows = 123;                                  // bins 5 .. 127: all the reference's filterbanks touch (fe_plans_gen.h)
constexpr int kUPbufBytes = kUPRows * kUPS * 4;                            // 15 744
// Raw samples of a tile: FOUR chunks of 8 frames each (190 sixteen-byte vectors: 7 * 160 + 400 samples), chunk c at
// vector 191 c of the buffer, and lane 8 i + j works on frame 8 (j & 3) + 2 i + (j >> 2).  The 16-byte loads of a
// quarter warp (fixed i) then fall into the 8 different bank groups (191 c + 20 t + w = -(j & 3) + 4 (j >> 2) + w mod 8):
// conflict-free, where one contiguous copy (lane stride 20 vectors) costs 4 wavefronts per quarter warp.
constexpr int kUChunkVecs = 191;
constexpr int kURawBytes = 4 * kUChunkVecs * 16;                           // 12 224 per buffer
#ifndef FE_K1U_RAW_BUFS
#define FE_K1U_RAW_BUFS 2          // raw buffers per group (copies issued this many tiles ahead)
#endif
#ifndef FE_K1U_PBUF_BUFS
#define FE_K1U_PBUF_BUFS 2         // power buffers per group (2: the epilogue of tile i may still run during stage B of i + 1)
#endif
constexpr int kURawBufs = FE_K1U_RAW_BUFS, kUPbufBufs = FE_K1U_PBUF_BUFS;
constexpr int kUOffPbuf = kUGroups * kURawBufs * kURawBytes;               // raw samples: [group][raw buf][4 chunks]
constexpr int kUOffSs = kUOffPbuf + kUGroups * kUPbufBufs * kUPbufBytes;   // power bins:  [group][pbuf][rows][32]
constexpr int kUOffX = kUOffSs + kUGroups * kUPbufBufs * kUSub * 32 * 4;   // sum of squares per quad: [group][pbuf][sub][32]
constexpr int kUOffBar = kUOffX + kUGroups * kUPbufBufs * 2 * 32 * 4;      // X[0], X[256]: [group][pbuf][2][32]
constexpr int kUSmem = kUOffBar + 192;                                     // mbarriers: raw full [g][2], pbuf empty [g][2], exchange drained [g]
constexpr int kUFftRegs = 104, kUEpiRegs = 64;                             // 16 x 32 x 104 + 4 x 32 x 64 = 640 x 96
static_assert(kUSmem <= 232448 - 16, "K1U shared memory");
static_assert(kURawBufs >= 1 && kURawBufs <= 2 && kUPbufBufs >= 1 && kUPbufBufs <= 2, "K1U buffer counts");

__device__ __forceinline__ void mbar_wait_vote(uint32_t bar, uint32_t parity) {       // uniform loop exit
    while (!__all_sync(0xffffffffu, mbar_try(bar, parity))) __nanosleep(20);
}
__device__ __forceinline__ void ubar_sync(int id, int threads) {
    asm volatile("bar.sync %0, %1;" :: "r"(id), "r"(threads) : "memory");
}
__device__ __forceinline__ void ubar_arrive(int id, int threads) {
    asm volatile("bar.arrive %0, %1;" :: "r"(id), "r"(threads) : "memory");
}
// the four FFT warps of a lane group meet on named barrier 1 + g; tensor-memory traffic is ordered across it by the fences
__device__ __forceinline__ void k1u_group_sync(int g) {
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    ubar_sync(1 + g, kUSub * 32);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}
// lane 8 i + j of a K1U warp owns this frame of the tile (see the raw layout above)
__device__ __forceinline__ int k1u_frame_of_lane(int lane) { return 8 * (lane & 3) + 2 * (lane >> 3) + ((lane >> 2) & 1); }

template <int EPI>
__global__ void __launch_bounds__(kUThreads, 1)
k_frames_to_statics_u(const short* __restrict__ pcm, const short* __restrict__ scratch,
                      const TileDesc* __restrict__ tiles, int n_tiles,
                      const __grid_constant__ K1TParams P, float* __restrict__ statics) {
    extern __shared__ __align__(16) unsigned char smem_u[];
    __shared__ uint32_t s_tmem_base;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int group = warp & 3, sub = warp >> 2;                 // sub == kUSub: the group's epilogue warp
    const uint32_t bar0 = smem_u32(smem_u + kUOffBar);

    for (int i = tid; i < kUOffBar / 4; i += kUThreads) reinterpret_cast<uint32_t*>(smem_u)[i] = 0u;   // partial tiles read stale rows / bins
    if (tid == 0) {
        for (int i = 0; i < 2 * kUGroups; ++i) { mbar_init(bar0 + 8 * i, 1); mbar_init(bar0 + 64 + 8 * i, 1); }
        for (int i = 0; i < kUGroups; ++i) mbar_init(bar0 + 128 + 8 * i, kUSub);
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(&s_tmem_base)), "r"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    if (s_tmem_base != 0u) __trap();                             // all 512 columns are ours: the allocation starts at 0
    const int tile_stride = gridDim.x * kUGroups;
    const int n_iter = (n_tiles + tile_stride - 1) / tile_stride;
    const int frame = k1u_frame_of_lane(lane);

    if (__all_sync(0xffffffffu, sub < kUSub)) {
        // =============================== FFT warps ===============================
        asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;\n" :: "n"(kUFftRegs));
        const TTwiddles tw{c_tw256p, c_tw512p, c_tw512};
#pragma unroll 1
        for (int ug = 0; ug < kUGroups; ++ug) {
#pragma unroll 1
        for (int us = 0; us < kUSub; ++us) {
        if (!__all_sync(0xffffffffu, ug == group && us == sub)) continue;
        const TmemExchange ex{(uint32_t)(ug * 32) << 16};
        unsigned char* raw_g = smem_u + ug * (kURawBufs * kURawBytes);
        float* pbuf_g = reinterpret_cast<float*>(smem_u + kUOffPbuf + ug * (kUPbufBufs * kUPbufBytes));
        float* ss_g = reinterpret_cast<float*>(smem_u + kUOffSs) + ug * (kUPbufBufs * kUSub * 32);
        float* x_g = reinterpret_cast<float*>(smem_u + kUOffX) + ug * (kUPbufBufs * 2 * 32);
        const uint32_t bar_raw = bar0 + 16 * ug, bar_empty = bar0 + 64 + 16 * ug, bar_drain = bar0 + 128 + 8 * ug;
        // raw samples: lane 0 of sub-warp 3 copies a tile as four chunks (frames 8 c .. 8 c + 7), kURawBufs tiles ahead; the
        // descriptor it needs was prefetched into L1 at the top of the previous tile
        auto desc_of = [&](int it) { return tiles + ((blockIdx.x + it * gridDim.x) * kUGroups + ug); };
        auto prefetch_desc = [&](int it) {
            if ((blockIdx.x + it * gridDim.x) * kUGroups + ug < n_tiles && lane == 0)
                asm volatile("prefetch.global.L1 [%0];" :: "l"(desc_of(it)));
        };
        auto fetch = [&](int it, int rb) {
            if ((blockIdx.x + it * gridDim.x) * kUGroups + ug >= n_tiles || lane != 0) return;
            const TileDesc* td = desc_of(it);
            const long long off = td->pcm_off;
            const int2 ns = *reinterpret_cast<const int2*>(&td->n_frames);        // n_frames, src_sel
            mbar_expect_tx(bar_raw + 8 * rb, (uint32_t)((ns.x - 1) * 160 + 400 + ((ns.x - 1) >> 3) * 240) * 2u);
            const short* src = (ns.y ? scratch : pcm) + off;
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                const int cnt = min(ns.x - 8 * c, 8);
                if (cnt > 0) bulk_g2s(smem_u32(raw_g + rb * kURawBytes + c * (kUChunkVecs * 16)), src + c * (8 * 160),
                                      (uint32_t)((cnt - 1) * 160 + 400) * 2u, bar_raw + 8 * rb);
            }
        };
        if (us == kUSub - 1) { fetch(0, 0); if (kURawBufs == 2) fetch(1, 1); prefetch_desc(kURawBufs); }
        const int lane_vec = (frame >> 3) * kUChunkVecs + (frame & 7) * kTFrameVecs;      // this lane's first vector in a raw buffer
#pragma unroll 1
        for (int it = 0; it < n_iter; ++it) {
            const int t = (blockIdx.x + it * gridDim.x) * kUGroups + ug;
            if (t >= n_tiles) break;
            const int rb = kURawBufs == 2 ? (it & 1) : 0, pb = kUPbufBufs == 2 ? (it & 1) : 0;
            mbar_wait_vote(bar_raw + 8 * rb, (uint32_t)(kURawBufs == 2 ? it >> 1 : it) & 1u);
            if (it > 0) {                                           // every warp of the group has loaded its last pair of tile it - 1
                mbar_wait_vote(bar_drain, (uint32_t)(it - 1) & 1u);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            }
            const uint4* raw4 = reinterpret_cast<const uint4*>(raw_g + rb * kURawBytes) + lane_vec;
            const float ss = k1t_stage_a<(EPI == 1 || EPI == 3)>(raw4, ex, us, us + 1);      // frame energy: cepstra only (c0)
            ex.wait_st();
            k1u_group_sync(ug);                                      // exchange complete, this buffer's samples consumed
            if (us == kUSub - 1) { fetch(it + kURawBufs, rb); prefetch_desc(it + kURawBufs + 1); }
            // the epilogue of the previous user of pbuf[pb] has drained it
            if (it >= kUPbufBufs) mbar_wait_vote(bar_empty + 8 * pb, (uint32_t)((kUPbufBufs == 2 ? it >> 1 : it) - 1) & 1u);
            if (EPI == 1 || EPI == 3) ss_g[(pb * kUSub + us) * 32 + lane] = ss;
            float* pcol = pbuf_g + pb * (kUPRows * kUPS) + lane - kUPBin0 * kUPS;      // row = bin - kUPBin0
#pragma unroll 1
            for (int pp = 0; pp < 2; ++pp) {
                float x0, x256;
                // "exchange drained" is signalled as soon as this warp's last pair sits in registers (split-phase: the
                // group does not meet again before the next tile's stage A stores)
                k1t_pair<kUPS, kUPBin0, 127>(ex, 2 * us + pp, tw, pcol, x0, x256, [&] {
                    if (pp == 1) {
                        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                        if (lane == 0) mbar_arrive(bar_drain);
                    }
                });
                if (us == 0 && pp == 0) { x_g[(pb * 2 + 0) * 32 + lane] = x0; x_g[(pb * 2 + 1) * 32 + lane] = x256; }
            }
            ubar_arrive(5 + 2 * ug + pb, 160);                      // power bins of this tile complete
        }
        }}
    } else {
        // ============================= epilogue warps =============================
        asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;\n" :: "n"(kUEpiRegs));
#pragma unroll 1
        for (int ug = 0; ug < kUGroups; ++ug) {
        if (!__all_sync(0xffffffffu, ug == group)) continue;
        const float* pbuf_g = reinterpret_cast<const float*>(smem_u + kUOffPbuf + ug * (kUPbufBufs * kUPbufBytes));
        const float* ss_g = reinterpret_cast<const float*>(smem_u + kUOffSs) + ug * (kUPbufBufs * kUSub * 32);
        const float* x_g = reinterpret_cast<const float*>(smem_u + kUOffX) + ug * (kUPbufBufs * 2 * 32);
        const uint32_t bar_empty = bar0 + 64 + 16 * ug;
#pragma unroll 1
        for (int it = 0; it < n_iter; ++it) {
            const int t = (blockIdx.x + it * gridDim.x) * kUGroups + ug;
            if (t >= n_tiles) break;
            const int pb = kUPbufBufs == 2 ? (it & 1) : 0;
            float* out_t = statics + tiles[t].stat_off;
            ubar_sync(5 + 2 * ug + pb, 160);
            float energy = 1.f;
            if (EPI == 1 || EPI == 3) {
                const float* ssb = ss_g + pb * kUSub * 32 + lane;
                const float ss = (ssb[0] + ssb[32]) + (ssb[64] + ssb[96]);
                energy = frame_energy(ss, x_g[(pb * 2 + 0) * 32 + lane], x_g[(pb * 2 + 1) * 32 + lane], P.pscale);
            }
            const float* pbp = pbuf_g + pb * (kUPRows * kUPS) - kUPBin0 * kUPS;         // row = bin - kUPBin0
            // mel -> log -> DCT for this lane's frame (compile-time filterbank plan, weights in the constant bank)
            // (EPI: 1 = 40 filters -> 13 cepstra, 3 = 40 -> 39 cepstra, 2 = fbank-80, 4 = fbank-40; as in k_frames_to_statics)
            if (EPI == 1) epi_tile_spec_e<PlanMfcc40, 13, true, true, kUPS>(pbp, energy, out_t, P.epi_w, P.dc_elim != 0, lane, frame);
            else if (EPI == 3) epi_tile_spec_e<PlanMfcc40, 39, true, true, kUPS>(pbp, energy, out_t, P.epi_w, P.dc_elim != 0, lane, frame);
            else if (EPI == 2 && P.fbank_log) epi_tile_spec_e<PlanFbank80, 80, false, true, kUPS>(pbp, energy, out_t, P.epi_w, false, lane, frame);
            else if (EPI == 2) epi_tile_spec_e<PlanFbank80, 80, false, false, kUPS>(pbp, energy, out_t, P.epi_w, false, lane, frame);
            else if (P.fbank_log) epi_tile_spec_e<PlanMfcc40, 40, false, true, kUPS>(pbp, energy, out_t, P.epi_w, false, lane, frame);
            else epi_tile_spec_e<PlanMfcc40, 40, false, false, kUPS>(pbp, energy, out_t, P.epi_w, false, lane, frame);
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_empty + 8 * pb);
        }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(s_tmem_base), "r"(512) : "memory");
}

// ---------------------------------------------------------------------------
// K2a k_utt_stats: per-utterance mean and 1 / (population std + 2^-30) of every statics column
// (speechpy.processing.cmvn, preprocess.py:85).  One 128-thread CTA per utterance (grid-stride),
// two passes, fixed-order block reductions (deterministic).  mean = x[0] + mean(x - x[0]): the
// shift keeps a constant column (digital silence) exactly constant.
// flags: bit0 subtract the mean, bit1 divide by the std (stats[u] = {mean[D], inv[D]}).
// ---------------------------------------------------------------------------
constexpr int kStatThreads = 128;

__global__ void __launch_bounds__(kStatThreads)
k_utt_stats(const UttDesc* __restrict__ utts, int n_utts, const float* __restrict__ statics,
            float* __restrict__ stats, int D, int flags) {
    extern __shared__ float sm_s[];           // red[kStatThreads] | mean[D]
    float* red = sm_s;
    float* mean = sm_s + kStatThreads;
    const bool f_mean = flags & 1, f_var = flags & 2;
    const int tid = threadIdx.x;
    const int G = D <= kStatThreads ? kStatThreads / D : 1;      // row groups
    for (int ui = blockIdx.x; ui < n_utts; ui += gridDim.x) {
        const int L = utts[ui].n_frames;
        float* st = stats + (long long)ui * 2 * D;
        if (L <= 0) continue;
        const float* x = statics + utts[ui].stat_off;
        for (int cb = 0; cb < D; cb += kStatThreads) {           // column blocks (D > 128 only loops)
            const int c = cb + tid % (D < kStatThreads ? D : kStatThreads);
            const int r = D < kStatThreads ? tid / D : 0;
            const bool act = r < G && c < D;
            const float shift = (act && f_mean) ? x[c] : 0.f;
            float s = 0.f;
            if (act && f_mean) {
                float a[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};      // eight row streams in flight per thread
                const float* xc = x + c;
                int t = r;
                for (; t + 7 * G < L; t += 8 * G) {
                    float v[8];
#pragma unroll
                    for (int k = 0; k < 8; ++k) v[k] = xc[(long long)(t + k * G) * D];
#pragma unroll
                    for (int k = 0; k < 8; ++k) a[k] += v[k] - shift;
                }
                for (; t < L; t += G) a[0] += xc[(long long)t * D] - shift;
                s = ((a[0] + a[1]) + (a[2] + a[3])) + ((a[4] + a[5]) + (a[6] + a[7]));
            }
            red[tid] = s;
            __syncthreads();
            if (tid < D - cb && tid < kStatThreads) {
                float m = 0.f;
                if (f_mean) {
                    const int w = D < kStatThreads ? D : kStatThreads;
                    for (int rr = 0; rr < G; ++rr) m += red[rr * w + tid];
                    m = x[cb + tid] + m / (float)L;
                }
                mean[tid] = m;
                st[cb + tid] = m;
            }
            __syncthreads();
            const float mu = act ? mean[c - cb] : 0.f;
            float q = 0.f;
            if (act && f_var) {
                float a[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
                const float* xc = x + c;
                int t = r;
                for (; t + 7 * G < L; t += 8 * G) {
                    float v[8];
#pragma unroll
                    for (int k = 0; k < 8; ++k) v[k] = xc[(long long)(t + k * G) * D] - mu;
#pragma unroll
                    for (int k = 0; k < 8; ++k) a[k] = fmaf(v[k], v[k], a[k]);
                }
                for (; t < L; t += G) { float d = xc[(long long)t * D] - mu; a[0] = fmaf(d, d, a[0]); }
                q = ((a[0] + a[1]) + (a[2] + a[3])) + ((a[4] + a[5]) + (a[6] + a[7]));
            }
            __syncthreads();
            red[tid] = q;
            __syncthreads();
            if (tid < D - cb && tid < kStatThreads) {
                float iv = 1.f;
                if (f_var) {
                    float v = 0.f;
                    const int w = D < kStatThreads ? D : kStatThreads;
                    for (int rr = 0; rr < G; ++rr) v += red[rr * w + tid];
                    iv = 1.0f / (sqrtf(v / (float)L) + 9.313225746154785e-10f);   // 2^-30
                }
                st[D + cb + tid] = iv;
            }
            __syncthreads();
        }
    }
}

// K2a for K1's tile-major statics ([D][32] blocks): one warp per (utterance, coefficient), lane = frame
// within the block, so every load is one 128-byte line and the reduction is a fixed-order shuffle tree
// (deterministic).  mean = x[0] + mean(x - x[0]); population variance from the same single pass.
constexpr int kStatTWarps = 4;

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

__global__ void __launch_bounds__(kStatTWarps * 32)
k_utt_stats_tiled(const UttDesc* __restrict__ utts, int n_utts, const float* __restrict__ statics,
                  float* __restrict__ stats, int D) {
    const int lane = threadIdx.x & 31;
    const long long n_items = (long long)n_utts * D;
    const long long w0 = (long long)blockIdx.x * kStatTWarps + (threadIdx.x >> 5);
    for (long long item = w0; item < n_items; item += (long long)gridDim.x * kStatTWarps) {
        const int ui = (int)(item / D), c = (int)(item % D);
        const int L = utts[ui].n_frames;
        if (L <= 0) continue;
        const float* x = statics + utts[ui].stat_off + c * 32 + lane;       // block j at x[j * 32 * D]
        const long long bs = 32LL * D;
        const int nb = L >> 5, tail = L & 31;
        // ONE pass over the column: sums of d = x - x[0] and d^2 (shifted data: a constant column -- digital silence --
        // stays exactly constant, and the cancellation in sum d^2 - (sum d)^2 / L is that of data centred on a sample)
        const float shift = __shfl_sync(0xffffffffu, x[0], 0);
        float a[4] = {0.f, 0.f, 0.f, 0.f}, q[4] = {0.f, 0.f, 0.f, 0.f};
        int j = 0;
        for (; j + 3 < nb; j += 4) {
            float v[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) v[k] = x[(j + k) * bs] - shift;
#pragma unroll
            for (int k = 0; k < 4; ++k) { a[k] += v[k]; q[k] = fmaf(v[k], v[k], q[k]); }
        }
        for (; j < nb; ++j) { const float d = x[j * bs] - shift; a[0] += d; q[0] = fmaf(d, d, q[0]); }
        if (lane < tail) { const float d = x[nb * bs] - shift; a[1] += d; q[1] = fmaf(d, d, q[1]); }
        const float sd = warp_sum((a[0] + a[1]) + (a[2] + a[3]));
        const float sq = warp_sum((q[0] + q[1]) + (q[2] + q[3]));
        if (lane == 0) {
            const float invL = 1.0f / (float)L;
            const float var = fmaxf((sq - sd * sd * invL) * invL, 0.f);
            float* st = stats + (long long)ui * 2 * D;
            st[c] = shift + sd * invL;
            st[D + c] = 1.0f / (sqrtf(var) + 9.313225746154785e-10f);       // 2^-30
        }
    }
}

// ---------------------------------------------------------------------------
// K2b k_norm_delta_pack: normalise, delta, delta-delta, pack the cube (L, D, 3)
// (speechpy.feature.extract_derivative_feature, preprocess.py:86).  Tile-parallel (same tile
// table as K1), streaming: 4 D bytes in, 12 D bytes out per frame, written with coalesced
// 16-byte stores from a shared-memory image of the tile.
// flags bit2: append deltas (else the output is the normalised (L, D) matrix); bit3: no statistics
// (mean 0, scale 1).  TR: the input is K1's tile-major [D][32] blocks, else a row-major (L, D) matrix
// (fe_postprocess).
// ---------------------------------------------------------------------------
constexpr int kPackThreads = 128;

__host__ __device__ inline int k2_smem_floats(int D, int tile_frames) {
    return 2 * D + (tile_frames + 8) * D + (tile_frames + 4) * D + tile_frames * 3 * D + 16;
}

// DT > 0: feature width known at compile time (index math by multiply-shift instead of integer division)
template <int DT, bool TR>
__global__ void __launch_bounds__(kPackThreads)
k_norm_delta_pack(const TileDesc* __restrict__ tiles, int n_tiles, const float* __restrict__ statics,
                  const float* __restrict__ stats, float* __restrict__ out, int D_rt, int tile_frames,
                  int delta_mode, int flags) {
    const int D = DT > 0 ? DT : D_rt;
    extern __shared__ __align__(16) float sm_p[];
    const bool f_delta = flags & 4;
    const int W = f_delta ? 3 : 1;
    float* cube = sm_p;                                            // [tile_frames][D][W], 16-byte aligned
    float* mi = cube + ((tile_frames * 3 * D + 3) & ~3);           // mean[D], inv[D]
    float* vt = mi + 2 * D;                                        // [(rows + 8)][D]
    float* d1 = vt + (tile_frames + 8) * D;                        // [(rows + 4)][D]
    const int tid = threadIdx.x;
    for (int ti = blockIdx.x; ti < n_tiles; ti += gridDim.x) {
        const TileDesc td = tiles[ti];
        const int nrow = td.n_frames, L = td.utt_frames, t0 = td.first_frame;
        const float* st = stats + (long long)td.utt * 2 * D;
        const float* x0 = statics + td.stat_off - (long long)t0 * D;          // frame 0 of the utterance (t0 % 32 == 0)
        if (flags & 8) { for (int i = tid; i < 2 * D; i += kPackThreads) mi[i] = i < D ? 0.f : 1.f; }
        else { for (int i = tid; i < 2 * D; i += kPackThreads) mi[i] = st[i]; }
        __syncthreads();
        if (!f_delta || delta_mode == 0) {
            // as shipped: everything is local to the frame
            const int n = nrow * D;
            const float* x = statics + td.stat_off;
            // vt image: element (row r, coefficient c) at r * rs + c * cs
            const int rs = TR ? 1 : D, cs = TR ? 33 : 1;
            if (TR) {
                for (int j = tid; j < 32 * D; j += kPackThreads) {            // coalesced over the [D][32] block
                    const int c = j >> 5, r = j & 31;
                    if (r < nrow) vt[c * 33 + r] = (x[j] - mi[c]) * mi[D + c];
                }
            } else {
                for (int i = tid; i < n; i += kPackThreads) {
                    const int c = i % D;
                    vt[i] = (x[i] - mi[c]) * mi[D + c];
                }
            }
            __syncthreads();
            for (int i = tid; i < n; i += kPackThreads) {
                const int c = i % D, r = i / D;
                const float* vr = vt + r * rs;
                const float v = vr[c * cs];
                if (!f_delta) { cube[i] = v; continue; }
                // d1[k] = (v[k+1] + 2 v[k+2]) / 10, d2[k] = (d1[k+1] + 2 d1[k+2]) / 10, indices clamped to D-1;
                // with ci = min(c+i, D-1): d1[c1] = (v[c2] + 2 v[c3]) / 10 and d1[c2] = (v[c3] + 2 v[c4]) / 10
                const float v1 = vr[min(c + 1, D - 1) * cs], v2 = vr[min(c + 2, D - 1) * cs];
                const float v3 = vr[min(c + 3, D - 1) * cs], v4 = vr[min(c + 4, D - 1) * cs];
                const float da = (v1 + 2.f * v2) * 0.1f;
                const float db = (v2 + 2.f * v3) * 0.1f;
                const float dc = (v3 + 2.f * v4) * 0.1f;
                float* q3 = cube + i * 3;
                q3[0] = v; q3[1] = da; q3[2] = (db + 2.f * dc) * 0.1f;
            }
        } else {
            // regression along time, edge replication; vt row i <-> frame clamp(t0-4+i), d1 row i <-> clamp(t0-2+i)
            for (int i = tid; i < (nrow + 8) * D; i += kPackThreads) {
                const int c = i % D, rr = i / D;
                const int a = min(max(t0 - 4 + rr, 0), L - 1);
                const float xv = TR ? x0[(long long)(a >> 5) * (32 * D) + c * 32 + (a & 31)] : x0[(long long)a * D + c];
                vt[i] = (xv - mi[c]) * mi[D + c];
            }
            __syncthreads();
            for (int i = tid; i < (nrow + 4) * D; i += kPackThreads) {
                const int c = i % D, rr = i / D;
                const int sfr = min(max(t0 - 2 + rr, 0), L - 1);
                float acc = 0.f;
#pragma unroll
                for (int k = 1; k <= 2; ++k) {
                    const int ap = min(sfr + k, L - 1) - (t0 - 4), am = max(sfr - k, 0) - (t0 - 4);
                    acc += (float)k * (vt[ap * D + c] - vt[am * D + c]);
                }
                d1[i] = acc * 0.1f;
            }
            __syncthreads();
            for (int i = tid; i < nrow * D; i += kPackThreads) {
                const int c = i % D, rr = i / D, tt = t0 + rr;
                float acc = 0.f;
#pragma unroll
                for (int k = 1; k <= 2; ++k) {
                    const int ap = min(tt + k, L - 1) - (t0 - 2), am = max(tt - k, 0) - (t0 - 2);
                    acc += (float)k * (d1[ap * D + c] - d1[am * D + c]);
                }
                float* q3 = cube + i * 3;
                q3[0] = vt[(rr + 4) * D + c]; q3[1] = d1[(rr + 2) * D + c]; q3[2] = acc * 0.1f;
            }
        }
        __syncthreads();
        const int total = nrow * W * D;
        float* dst = out + td.out_off + (long long)t0 * W * D;       // 16-byte aligned: t0 % 4 == 0
        const int n4 = total >> 2;
        for (int i = tid; i < n4; i += kPackThreads)
            reinterpret_cast<float4*>(dst)[i] = reinterpret_cast<const float4*>(cube)[i];
        for (int i = (n4 << 2) + tid; i < total; i += kPackThreads) dst[i] = cube[i];
        // the 0..3 pad floats that round the utterance's run up to 16 bytes are zeroed (deterministic buffers)
        if (t0 + nrow == L && tid < ((4 - (total & 3)) & 3)) dst[total + tid] = 0.f;
        __syncthreads();
    }
}

// ---------------------------------------------------------------------------
// K2b, fast path: CMVN + as-shipped deltas + cube for K1's tile-major statics.  The as-shipped deltas run
// along the COEFFICIENT axis, so with lane = frame everything is per-lane register arithmetic: one warp per
// tile, no CTA barrier.  A tile's [D][32] block is read with D coalesced 128-byte loads, the (32, D, 3) cube
// is staged through the warp's own shared-memory window (row stride 3 D is odd or padded -> conflict-free)
// and leaves as 16-byte coalesced stores.  flags as k_norm_delta_pack (bit2 deltas, bit3 no statistics).
// ---------------------------------------------------------------------------
template <int D>
struct CubeLocal {
    static constexpr int CH = D <= 16 ? D : (D == 39 ? 13 : (D % 20 == 0 && D <= 40 ? 20 : 16));    // coefficients staged per round
    static_assert(D % CH == 0, "feature width must be a multiple of the staging chunk");
    static constexpr int RS = (3 * CH) | 1;                              // staging row stride (odd: no bank conflicts)
    static constexpr int kWarps = 8;
    static constexpr int kSmemBytes = kWarps * 32 * RS * 4;
};

template <int D, bool DELTA>
__device__ __forceinline__ void cube_local_body(const TileDesc* __restrict__ tiles, int n_tiles, const float* __restrict__ statics,
                                                const float* __restrict__ stats, float* __restrict__ out, int flags, float* sm_c) {
    using C = CubeLocal<D>;
    constexpr int W = DELTA ? 3 : 1;
    constexpr int ROWLEN = W * D;                  // floats per output row
    constexpr int SEG = W * C::CH;                 // floats of a row written per round
    constexpr int RS = C::RS;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float* stage = sm_c + warp * 32 * RS;
    const long long w0 = (long long)blockIdx.x * C::kWarps + warp;
    for (long long ti = w0; ti < n_tiles; ti += (long long)gridDim.x * C::kWarps) {
        const TileDesc td = tiles[ti];
        const int nrow = td.n_frames;
        const float* x = statics + td.stat_off + lane;
        float v[D];
#pragma unroll
        for (int c = 0; c < D; ++c) v[c] = x[c * 32];                     // 128-byte lines; pad lanes read finite garbage
        if (!(flags & 8)) {
            const float* st = stats + (long long)td.utt * 2 * D;
#pragma unroll
            for (int c = 0; c < D; ++c) v[c] = (v[c] - __ldg(st + c)) * __ldg(st + D + c);   // warp-uniform, cached
        }
        float* dst = out + td.out_off + (long long)td.first_frame * ROWLEN;     // 16-byte aligned: first_frame % 4 == 0
        float* row = stage + lane * RS;
#pragma unroll
        for (int c0 = 0; c0 < D; c0 += C::CH) {
            if (DELTA) {
                // d1[k] = (v[k+1] + 2 v[k+2]) / 10, d2[k] = (d1[k+1] + 2 d1[k+2]) / 10, indices clamped to D-1
#pragma unroll
                for (int j = 0; j < C::CH; ++j) {
                    const int c = c0 + j;
                    auto cl = [](int i) { return i < D ? i : D - 1; };
                    const float d1c = (v[cl(c + 1)] + 2.f * v[cl(c + 2)]) * 0.1f;
                    const float d1a = (v[cl(cl(c + 1) + 1)] + 2.f * v[cl(cl(c + 1) + 2)]) * 0.1f;       // d1[min(c+1, D-1)]
                    const float d1b = (v[cl(cl(c + 2) + 1)] + 2.f * v[cl(cl(c + 2) + 2)]) * 0.1f;       // d1[min(c+2, D-1)]
                    row[3 * j] = v[c]; row[3 * j + 1] = d1c; row[3 * j + 2] = (d1a + 2.f * d1b) * 0.1f;
                }
            } else {
#pragma unroll
                for (int j = 0; j < C::CH; ++j) row[j] = v[c0 + j];
            }
            __syncwarp();
            if (SEG == ROWLEN && RS == ROWLEN) {
                // single round, contiguous staging (D = 13 with deltas): straight 16-byte copies
                const int total = nrow * ROWLEN, n4 = total >> 2;
                for (int i = lane; i < n4; i += 32) reinterpret_cast<float4*>(dst)[i] = reinterpret_cast<const float4*>(stage)[i];
                for (int i = (n4 << 2) + lane; i < total; i += 32) dst[i] = stage[i];
            } else if (SEG % 4 == 0 && ROWLEN % 4 == 0) {
                // every row segment is a whole number of aligned 16-byte pieces
                constexpr int Q = SEG / 4;
                for (int i = lane; i < nrow * Q; i += 32) {
                    const int r = i / Q, q = i - r * Q;
                    const float* sp = stage + r * RS + 4 * q;
                    reinterpret_cast<float4*>(dst + (long long)r * ROWLEN + W * c0)[q] = make_float4(sp[0], sp[1], sp[2], sp[3]);
                }
            } else {
                for (int i = lane; i < nrow * SEG; i += 32) {
                    const int r = i / SEG, e = i - r * SEG;
                    dst[(long long)r * ROWLEN + W * c0 + e] = stage[r * RS + e];
                }
            }
            __syncwarp();
        }
        // the 0..3 pad floats that round the utterance's run up to 16 bytes are zeroed (deterministic buffers)
        const int total = nrow * ROWLEN;
        if (td.first_frame + nrow == td.utt_frames && lane < ((4 - (total & 3)) & 3)) dst[total + lane] = 0.f;
    }
}

template <int D>
__global__ void __launch_bounds__(CubeLocal<D>::kWarps * 32)
k_cube_local(const TileDesc* __restrict__ tiles, int n_tiles, const float* __restrict__ statics,
             const float* __restrict__ stats, float* __restrict__ out, int flags) {
    extern __shared__ __align__(16) float sm_c[];
    if (flags & 4) cube_local_body<D, true>(tiles, n_tiles, statics, stats, out, flags, sm_c);
    else cube_local_body<D, false>(tiles, n_tiles, statics, stats, out, flags, sm_c);
}

// ---------------------------------------------------------------------------
// K2 fused: per-utterance CMVN statistics AND the cube in one kernel, one CTA per utterance (persistent).  The statics
// of an utterance (52 B x L for MFCC-13, 320 B x L for fbank-80: 64 KB .. 1.1 MB) are read from HBM ONCE for the
// statistics; the cube pass reads them again a few microseconds later out of the L2.  The two-kernel form (K2a + K2b)
// read them three times from HBM (mean pass, variance pass, cube pass): 2 240 instead of 1 280 B / frame for fbank-80.
//   phase 1  warp w takes the blocks w', w' + nw, ... of coefficient chunk k (CH coefficients per lane in registers,
//            lane = frame): sums of d = x - x[0] and d^2, fixed-order shuffle tree, per-warp partials in shared
//            memory, combined in warp order (deterministic).  mean = x[0] + sum d / L,
//            var = (sum d^2 - (sum d)^2 / L) / L (shifted data: a constant column stays exactly constant).
//   phase 2  warp per tile: normalise, as-shipped deltas (coefficient axis: per-lane register arithmetic), cube through
//            the warp's staging window, 16-byte coalesced stores -- the body of k_cube_local.
// flags as k_norm_delta_pack (bit2 deltas).  As-shipped delta mode only (time regression keeps K2a + k_norm_delta_pack).
// ---------------------------------------------------------------------------
template <int D>
struct UttCube {
    static constexpr int kWarps = 8;
    static constexpr int SCH = D <= 20 ? D : 20;                       // coefficients per statistics chunk
    static_assert(D % SCH == 0, "feature width must be a multiple of the statistics chunk");
    static constexpr int NCH = D / SCH;                                // chunks; kWarps / NCH warps share one
    static_assert(kWarps % NCH == 0, "warps must split evenly over the chunks");
    static constexpr int WPC = kWarps / NCH;
    static constexpr int kStageFloats = kWarps * 32 * CubeLocal<D>::RS;
    static constexpr int kSmemBytes = (kStageFloats + kWarps * 2 * SCH + 2 * D) * 4;
};

template <int D, bool DELTA>
__device__ __forceinline__ void cube_tile(const float* __restrict__ x, const float* st_mean, const float* st_inv, bool normalise,
                                          float* __restrict__ dst, int nrow, bool last_tile, float* stage, int lane) {
    using C = CubeLocal<D>;
    constexpr int W = DELTA ? 3 : 1;
    constexpr int ROWLEN = W * D, SEG = W * C::CH, RS = C::RS;
    float v[D];
#pragma unroll
    for (int c = 0; c < D; ++c) v[c] = x[c * 32 + lane];                   // 128-byte lines; pad lanes read finite garbage
    if (normalise) {
#pragma unroll
        for (int c = 0; c < D; ++c) v[c] = (v[c] - st_mean[c]) * st_inv[c];   // warp-uniform (shared memory broadcast)
    }
    float* row = stage + lane * RS;
#pragma unroll
    for (int c0 = 0; c0 < D; c0 += C::CH) {
        if (DELTA) {
#pragma unroll
            for (int j = 0; j < C::CH; ++j) {
                const int c = c0 + j;
                auto cl = [](int i) { return i < D ? i : D - 1; };
                const float d1c = (v[cl(c + 1)] + 2.f * v[cl(c + 2)]) * 0.1f;
                const float d1a = (v[cl(cl(c + 1) + 1)] + 2.f * v[cl(cl(c + 1) + 2)]) * 0.1f;
                const float d1b = (v[cl(cl(c + 2) + 1)] + 2.f * v[cl(cl(c + 2) + 2)]) * 0.1f;
                row[3 * j] = v[c]; row[3 * j + 1] = d1c; row[3 * j + 2] = (d1a + 2.f * d1b) * 0.1f;
            }
        } else {
#pragma unroll
            for (int j = 0; j < C::CH; ++j) row[j] = v[c0 + j];
        }
        __syncwarp();
        if (SEG == ROWLEN && RS == ROWLEN) {
            const int total = nrow * ROWLEN, n4 = total >> 2;
            // streaming stores: the cube is never read again on the device, it should not push the statics out of the L2
            for (int i = lane; i < n4; i += 32) __stcs(reinterpret_cast<float4*>(dst) + i, reinterpret_cast<const float4*>(stage)[i]);
            for (int i = (n4 << 2) + lane; i < total; i += 32) dst[i] = stage[i];
        } else if (SEG % 4 == 0 && ROWLEN % 4 == 0) {
            constexpr int Q = SEG / 4;
            for (int i = lane; i < nrow * Q; i += 32) {
                const int r = i / Q, q = i - r * Q;
                const float* sp = stage + r * RS + 4 * q;
                reinterpret_cast<float4*>(dst + (long long)r * ROWLEN + W * c0)[q] = make_float4(sp[0], sp[1], sp[2], sp[3]);
            }
        } else {
            for (int i = lane; i < nrow * SEG; i += 32) {
                const int r = i / SEG, e = i - r * SEG;
                dst[(long long)r * ROWLEN + W * c0 + e] = stage[r * RS + e];
            }
        }
        __syncwarp();
    }
    const int total = nrow * ROWLEN;
    if (last_tile && lane < ((4 - (total & 3)) & 3)) dst[total + lane] = 0.f;    // pad floats of the utterance's run
}

template <int D>
__global__ void __launch_bounds__(UttCube<D>::kWarps * 32)
k_utt_cmvn_cube(const UttDesc* __restrict__ utts, int n_utts, const float* __restrict__ statics, float* __restrict__ stats,
                float* __restrict__ out, int flags) {
    using U = UttCube<D>;
    extern __shared__ __align__(16) float sm_u[];
    float* stage_all = sm_u;
    float* part = sm_u + U::kStageFloats;                  // [kWarps][2 * SCH]
    float* st = part + U::kWarps * 2 * U::SCH;             // mean[D], inv[D]
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int chunk = warp / U::WPC, wsub = warp % U::WPC;
    const bool deltas = flags & 4;
    unsigned long long pol_keep;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol_keep));
    for (int ui = blockIdx.x; ui < n_utts; ui += gridDim.x) {
        const UttDesc u = utts[ui];
        const int L = u.n_frames;
        if (L <= 0) continue;
        const float* x = statics + u.stat_off;
        const int nb = (L + 31) >> 5;
        // ---- phase 1: statistics ----
        float shift[U::SCH], s[U::SCH], q[U::SCH];
#pragma unroll
        for (int c = 0; c < U::SCH; ++c) { shift[c] = __ldg(x + (chunk * U::SCH + c) * 32); s[c] = 0.f; q[c] = 0.f; }
        for (int b = wsub; b < nb; b += U::WPC) {
            const float* xb = x + (long long)b * 32 * D + chunk * U::SCH * 32 + lane;
            const bool live = b * 32 + lane < L;
#pragma unroll
            for (int c = 0; c < U::SCH; ++c) {
                float xv;        // evict-last: keep the utterance's statics in the L2 until phase 2 has read them again
                asm volatile("ld.global.L2::cache_hint.f32 %0, [%1], %2;" : "=f"(xv) : "l"(xb + c * 32), "l"(pol_keep));
                const float d = live ? xv - shift[c] : 0.f;
                s[c] += d; q[c] = fmaf(d, d, q[c]);
            }
        }
#pragma unroll
        for (int c = 0; c < U::SCH; ++c) { s[c] = warp_sum(s[c]); q[c] = warp_sum(q[c]); }
        if (lane == 0) {
#pragma unroll
            for (int c = 0; c < U::SCH; ++c) { part[warp * 2 * U::SCH + c] = s[c]; part[warp * 2 * U::SCH + U::SCH + c] = q[c]; }
        }
        __syncthreads();
        if (threadIdx.x < D) {
            const int c = threadIdx.x, ch = c / U::SCH, cc = c % U::SCH;
            float ss = 0.f, qq = 0.f;
            for (int w = 0; w < U::WPC; ++w) { ss += part[(ch * U::WPC + w) * 2 * U::SCH + cc]; qq += part[(ch * U::WPC + w) * 2 * U::SCH + U::SCH + cc]; }
            const float invL = 1.0f / (float)L;
            const float mean = __ldg(x + c * 32) + ss * invL;
            const float var = fmaxf((qq - ss * ss * invL) * invL, 0.f);
            const float inv = 1.0f / (sqrtf(var) + 9.313225746154785e-10f);       // 2^-30
            st[c] = mean; st[D + c] = inv;
            stats[(long long)ui * 2 * D + c] = mean; stats[(long long)ui * 2 * D + D + c] = inv;
        }
        __syncthreads();
        // ---- phase 2: cube, one warp per tile (statics come out of the L2 now) ----
        float* stage = stage_all + warp * 32 * CubeLocal<D>::RS;
        const int W = deltas ? 3 : 1;
        for (int b = warp; b < nb; b += U::kWarps) {
            const int nrow = min(32, L - b * 32);
            float* dst = out + u.out_off + (long long)b * 32 * W * D;
            if (deltas) cube_tile<D, true>(x + (long long)b * 32 * D, st, st + D, true, dst, nrow, b == nb - 1, stage, lane);
            else cube_tile<D, false>(x + (long long)b * 32 * D, st, st + D, true, dst, nrow, b == nb - 1, stage, lane);
        }
        __syncthreads();                                   // st / part are rewritten for the next utterance
    }
}

// ---------------------------------------------------------------------------
// K0: speed perturbation.  y[j] = sum_t taps[(j*down) % up][t] * x[floor(j*down/up) - (T/2 - 1) + t],
// x = 0 off the ends, int16 in -> (* gain) -> round-half-even, saturate -> int16 out.  T = taps per phase.
// speed_idx < 0: gain only.
// ---------------------------------------------------------------------------
constexpr int kK0Outputs = 2880;        // outputs per K0 tile: 288 groups of UP = 10 consecutive outputs
constexpr int kK0Taps = 128;            // taps per phase the fast kernels are compiled for (tables.RESAMPLE_TAPS)

// generic kernel: any ratio, any (even) number of taps; also serves gain-only utterances
__global__ void __launch_bounds__(256)
k_resample(const short* __restrict__ pcm, const UttDesc* __restrict__ utts,
           const int2* __restrict__ atiles, int n_atiles,
           const int* __restrict__ sp_up, const int* __restrict__ sp_down, const int* __restrict__ sp_tap_off,
           const float* __restrict__ taps_all, int ntaps, short* __restrict__ dst, int use_dst_off) {
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
            const int base = (int)(pos / up) - (ntaps / 2 - 1);
            const int ph = (int)(pos % up);
            const float* tp = taps + ph * ntaps;
            float acc = 0.f;
#pragma unroll 8
            for (int k = 0; k < ntaps; ++k) {
                int i = base + k;
                float xv = (i >= 0 && i < u.n_src) ? (float)__ldg(x + i) : 0.f;
                acc = fmaf(__ldg(tp + k), xv, acc);
            }
            if (u.gain != 1.f) acc *= u.gain;
            y[j] = (short)fminf(fmaxf(rintf(acc), -32768.f), 32767.f);
        }
    }
}

// K0, fast path for a compile-time ratio (the reference's speeds 0.9 / 1.1 -> 10/9, 10/11) and T = kK0Taps.
// A thread computes the UP consecutive outputs j = UP g .. UP g + UP - 1 of its group g: at every instruction all
// threads of the CTA are at the same (phase, tap), so the tap is a CONSTANT-BANK operand of the FFMA (the polyphase
// table travels in the kernel parameters: compile-time offset, no register, no load), and one shared-memory load
// of an input sample feeds ~UP FFMAs (windows of the UP phases overlap almost completely): T + DOWN - 1 loads for
// UP * T FFMAs = 0.11 loads per FMA.  A thread takes G = 2 consecutive groups: every tap then serves two FMAs, i.e. half
// the uniform loads per FMA (313 per 2 560 instead of per 1 280), and the two windows share all but DOWN samples.
// Same summation order per output as k_resample and the oracle (t ascending).
// Staging: the tile's input span is fetched as 16-byte vectors (8 int16 samples, at most two vectors per thread), one
// tile AHEAD of the arithmetic -- the vectors for tile i + 1 are in flight in registers while tile i is computed.
// Tiles start at multiples of kK0Outputs, i.e. at phase 0 and at a whole number of input samples: all index math
// is 32-bit, no division.
template <int VPT>
struct K0Stage {
    uint4 v[VPT];       // vector q holds 8 samples: utterance indices [n0 + 8 q THREADS, + 8)
    int n0;
    int n_src;
};

template <int UP, int T>
struct __align__(16) K0Taps { float w[UP * T]; };     // [phase][tap]

// float -> int16, round half to even, saturating: one F2I instead of rint + min + max + cast (same result; NaN -> 0)
__device__ __forceinline__ short f2s16_rn_sat(float a) {
    short r;
    asm("cvt.rni.sat.s16.f32 %0, %1;" : "=h"(r) : "f"(a));
    return r;
}

#ifndef FE_K0_GROUPS
#define FE_K0_GROUPS 3             // groups of UP outputs per thread: every tap (a uniform load) then serves three FMAs
#endif
constexpr int kK0Groups = FE_K0_GROUPS;

template <int UP, int DOWN, int T, int G = kK0Groups>
__global__ void __launch_bounds__(kK0Outputs / (UP * G))
k_resample_fast(const short* __restrict__ pcm, const UttDesc* __restrict__ utts,
                const int2* __restrict__ atiles, int n_atiles,
                const __grid_constant__ K0Taps<UP, T> W, short* __restrict__ dst, int use_dst_off) {
    constexpr int THREADS = kK0Outputs / (UP * G);                           // G consecutive groups of UP outputs per thread
    constexpr int HW = T / 2;
    constexpr int TILE_IN = kK0Outputs * DOWN / UP;                           // input samples a tile advances by (tiles start at
    static_assert(TILE_IN * UP == kK0Outputs * DOWN, "tile starts are phase 0");   // multiples of kK0Outputs: 32-bit index math)
    static_assert(kK0Outputs % 8 == 0 && THREADS * UP * G == kK0Outputs && (DOWN & 1) == 1, "tile geometry");
    constexpr int OFF_LAST = ((UP - 1) * DOWN) / UP;                          // window start of the last phase relative to phase 0
    constexpr int NX = OFF_LAST + T;                                          // input samples one group touches
    constexpr int NXG = NX + (G - 1) * DOWN;                                  // ... and a thread's G consecutive groups
    constexpr int SPAN = (THREADS * G - 1) * DOWN + NX;                       // input samples a tile touches
    constexpr int NV = (SPAN + 7 + 7) / 8;                                    // 16-byte vectors covering it from an aligned start
    constexpr int VPT = (NV + THREADS - 1) / THREADS;
    __shared__ __align__(16) float xs[VPT * THREADS * 8];
    __shared__ __align__(16) short ys[kK0Outputs];
    const int tid = threadIdx.x;

    // issue the staging loads of a tile (nothing is waited for here)
    auto fetch = [&](int tile, K0Stage<VPT>& sg, int2& te, UttDesc& u) {
        if (tile >= n_atiles) return;
        te = atiles[tile];
        u = utts[te.x];
        const int first = (te.y / kK0Outputs) * TILE_IN - (HW - 1);         // first input sample of the tile
        const int a0 = first & ~7;                                           // aligned down (two's complement: also for first < 0)
        sg.n0 = a0 + 8 * tid;
        sg.n_src = u.n_src;
#pragma unroll
        for (int q = 0; q < VPT; ++q) {
            const int n = sg.n0 + q * (8 * THREADS);
            sg.v[q] = make_uint4(0u, 0u, 0u, 0u);
            if (tid + q * THREADS < NV && n >= 0 && n < u.n_src)             // >= 1 valid sample: the vector lies inside the buffer
                sg.v[q] = __ldg(reinterpret_cast<const uint4*>(pcm + u.src_off + n));
        }
    };

    K0Stage<VPT> sg;
    int2 te_n = make_int2(0, 0);
    UttDesc u_n;
    fetch(blockIdx.x, sg, te_n, u_n);
    for (int tile = blockIdx.x; tile < n_atiles; tile += gridDim.x) {
        const int2 te = te_n;
        const UttDesc u = u_n;
        short* y = dst + (use_dst_off ? u.out_off : u.pcm_off) + te.y;
        const int nout = min(kK0Outputs, u.n_samples - te.y);
        const int first = (te.y / kK0Outputs) * TILE_IN - (HW - 1);
        const int a0 = first & ~7;
        __syncthreads();                                                     // previous tile's readers of xs / ys are done
#pragma unroll
        for (int qv = 0; qv < VPT; ++qv) {                                   // int16 -> float, samples off the ends are zero
            const int vi = tid + qv * THREADS;
            if (vi < NV) {
                const unsigned w[4] = {sg.v[qv].x, sg.v[qv].y, sg.v[qv].z, sg.v[qv].w};
                const int nb = sg.n0 + qv * (8 * THREADS);
                float f[8];
#pragma unroll
                for (int q = 0; q < 4; ++q) {                                 // zeros were loaded for dead vectors (n < 0 or n >= n_src)
                    f[2 * q] = (float)(short)(w[q] & 0xffffu);
                    f[2 * q + 1] = (float)(short)(w[q] >> 16);
                }
                if (nb + 8 > sg.n_src) {                                      // the vector straddles the utterance's end
#pragma unroll
                    for (int q = 0; q < 8; ++q) f[q] = (nb + q < sg.n_src) ? f[q] : 0.f;
                }
                if (nb < 0) {                                                 // ... or its start (first tile: first < 0)
#pragma unroll
                    for (int q = 0; q < 8; ++q) f[q] = (nb + q >= 0) ? f[q] : 0.f;
                }
                reinterpret_cast<float4*>(xs)[2 * vi] = make_float4(f[0], f[1], f[2], f[3]);
                reinterpret_cast<float4*>(xs)[2 * vi + 1] = make_float4(f[4], f[5], f[6], f[7]);
            }
        }
        __syncthreads();
        fetch(tile + gridDim.x, sg, te_n, u_n);                              // next tile's samples fly during the arithmetic
        const float* xw = xs + tid * (G * DOWN) + (first - a0);              // window of phase 0 of this thread's first group
        float acc[G][UP];
#pragma unroll
        for (int i = 0; i < G; ++i)
#pragma unroll
            for (int p = 0; p < UP; ++p) acc[i][p] = 0.f;
        // A ROLLED loop over chunks of TC taps: the straight-line form of G = 2 (3 000 instructions, 48 KB) no longer fits
        // the instruction cache and runs at half the speed.  The chunk counter is a uniform loop counter, so the taps still
        // reach the FMAs as uniform-register operands (LDCU c[0x0][UR + imm] from the kernel parameters).
#ifndef FE_K0_TC
#define FE_K0_TC 16
#endif
        constexpr int TC = FE_K0_TC, NW = TC + OFF_LAST + (G - 1) * DOWN;     // taps per chunk, samples a chunk touches
        static_assert(T % TC == 0, "taps per chunk");
#pragma unroll 1
        for (int c = 0; c < T / TC; ++c) {
            float xv[NW];
#pragma unroll
            for (int m = 0; m < NW; ++m) xv[m] = xw[c * TC + m];
            const float* wc = W.w + c * TC;
#pragma unroll
            for (int t4 = 0; t4 < TC; t4 += 4) {
#pragma unroll
                for (int p = 0; p < UP; ++p) {
                    // four consecutive taps of output p's phase: one 16-byte uniform load
                    const float4 w4 = *reinterpret_cast<const float4*>(wc + ((p * DOWN) % UP) * T + t4);
                    const float wv[4] = {w4.x, w4.y, w4.z, w4.w};
#pragma unroll
                    for (int e = 0; e < 4; ++e)
#pragma unroll
                        for (int i = 0; i < G; ++i)
                            acc[i][p] = fmaf(wv[e], xv[t4 + e + (p * DOWN) / UP + i * DOWN], acc[i][p]);
                }
            }
        }
#pragma unroll
        for (int i = 0; i < G; ++i)
#pragma unroll
            for (int p = 0; p < UP; ++p) ys[(tid * G + i) * UP + p] = f2s16_rn_sat(acc[i][p] * u.gain);     // gain 1 = exact identity
        __syncthreads();
        const int n8 = nout >> 3;                                            // y is 16-byte aligned (offsets % 8 == 0, tiles % 8 == 0)
        for (int i = tid; i < n8; i += THREADS) reinterpret_cast<int4*>(y)[i] = reinterpret_cast<const int4*>(ys)[i];
        for (int i = (n8 << 3) + tid; i < nout; i += THREADS) y[i] = ys[i];
    }
}

// K0': pre-emphasis y[n] = x[n] - a x[n-1], circular over the utterance (np.roll), to float scratch
template <int PCM_F32>
__global__ void __launch_bounds__(256)
k_preemph(const void* __restrict__ pcm, const UttDesc* __restrict__ utts, const int2* __restrict__ atiles,
          int n_atiles, float coef, float* __restrict__ dst) {
    for (int tile = blockIdx.x; tile < n_atiles; tile += gridDim.x) {
        const int2 te = atiles[tile];
        const UttDesc u = utts[te.x];
        const int j1 = min(te.y + kK0Outputs, u.n_samples);
        for (int j = te.y + threadIdx.x; j < j1; j += blockDim.x) {
            const int jp = j > 0 ? j - 1 : u.n_samples - 1;
            float a, b;
            if (PCM_F32) {
                const float* x = reinterpret_cast<const float*>(pcm) + u.src_off;
                a = x[j]; b = x[jp];
            } else {
                const short* x = reinterpret_cast<const short*>(pcm) + u.src_off;
                a = (float)x[j] * (1.0f / 32768.0f); b = (float)x[jp] * (1.0f / 32768.0f);
            }
            dst[u.pcm_off + j] = a - coef * b;
        }
    }
}

// ---------------------------------------------------------------------------
// K3: bucketed batches.  Slot = one utterance inside a dense [B, T_pad, D, planes] batch tensor: the
// utterance's n_valid floats are copied, the rest of the slot (padding up to the bucket boundary,
// tfrecord_data_loader.py:88-92 pad_to_bucket_boundary=True) is zero-filled.  Pure HBM traffic:
// 4 (n_valid + n_slot) bytes per slot.  One CTA per slot (grid-stride), 16-byte stores on the
// destination's alignment; loads are 16-byte when source and destination are congruent mod 16 bytes
// (always for 80-dim fbank rows, every other slot for MFCC-39 rows of 39 floats), else scalar but
// still fully coalesced.
// ---------------------------------------------------------------------------
struct __align__(8) PadSlot {
    long long src_off;      // float offset of the utterance's cube in the feature buffer
    long long dst_off;      // float offset of the slot in the batch buffer
    int n_valid;            // floats to copy (n_frames * D * planes)
    int n_slot;             // floats in the slot (T_pad * D * planes)
    int chunk_base;         // index of the slot's first work item (chunks of kPadChunk floats)
    int pad;
};

constexpr int kPadChunk = 8192;     // floats per work item (32 KB written): fine-grained enough to balance 148 SMs

__global__ void __launch_bounds__(256)
k_pad_slots(const float* __restrict__ src, const PadSlot* __restrict__ slots, int n_slots, int n_chunks,
            float* __restrict__ dst) {
    const int tid = threadIdx.x;
    for (int w = blockIdx.x; w < n_chunks; w += gridDim.x) {
        int lo = 0, hi = n_slots - 1;                                         // last slot with chunk_base <= w (block-uniform)
        while (lo < hi) {
            const int mid = (lo + hi + 1) >> 1;
            if (slots[mid].chunk_base <= w) lo = mid; else hi = mid - 1;
        }
        const PadSlot p = slots[lo];
        const float* s = src + p.src_off;
        float* d = dst + p.dst_off;
        const int head = min((int)((4 - (p.dst_off & 3)) & 3), p.n_slot);    // scalars up to the first 16-byte boundary of dst
        const int c = w - p.chunk_base;                                       // chunk c covers body vectors [c, c + 1) * kPadChunk / 4
        if (c == 0 && tid < head) d[tid] = tid < p.n_valid ? s[tid] : 0.f;
        const int nbody = (p.n_slot - head) >> 2;
        const bool congruent = ((p.src_off + head) & 3) == 0;
        const int full = p.n_valid >= head ? (p.n_valid - head) >> 2 : 0;     // body vectors made of valid floats only
        const int j1 = min(nbody, (c + 1) * (kPadChunk / 4));
#pragma unroll 4
        for (int j = c * (kPadChunk / 4) + tid; j < j1; j += 256) {
            const int e = head + 4 * j;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (j < full) {
                if (congruent) v = __ldcs(reinterpret_cast<const float4*>(s + e));
                else v = make_float4(__ldcs(s + e), __ldcs(s + e + 1), __ldcs(s + e + 2), __ldcs(s + e + 3));
            } else if (e < p.n_valid) {                                       // the vector straddles the end of the data
                v.x = s[e];
                if (e + 1 < p.n_valid) v.y = s[e + 1];
                if (e + 2 < p.n_valid) v.z = s[e + 2];
            }
            __stcs(reinterpret_cast<float4*>(d + e), v);
        }
        if (j1 == nbody) {                                                    // the slot's last chunk also writes the tail scalars
            const int e = head + 4 * nbody + tid;
            if (e < p.n_slot) d[e] = e < p.n_valid ? s[e] : 0.f;
        }
    }
}

// FP32 roofline denominator measured on the spot, no memory traffic.  Four instruction forms (VERDICT r1 item 3 asked
// for a cross-check of the packed probe): what the FP32 pipe sustains depends on where the operands come from -- the
// register file delivers about two operand words per lane and cycle (tools/ubench_issue2.cu):
//   MODE 0  scalar FFMA, warp-uniform multiplier and addend (the form that reaches the nominal 128 FMA / clk / SM)
//   MODE 1  scalar FFMA, three distinct register operands
//   MODE 2  packed FFMA2, uniform multiplier and addend (the round-1 probe)
//   MODE 3  packed FFMA2, three distinct register pairs
// 8 independent chains per thread, body unrolled 8x (loop overhead < 2 %).
template <int MODE>
__global__ void __launch_bounds__(256) k_fp32_peak(float* __restrict__ out, int iters, float ua, float ub) {
    float2 x[8], a[8], b[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        x[i] = make_float2(threadIdx.x * 1e-3f + i, threadIdx.x * 2e-3f - i);
        a[i] = make_float2(ua + i * 1e-7f + threadIdx.x * 1e-9f, ua * 0.999f + threadIdx.x * 1e-9f);
        b[i] = make_float2(ub + threadIdx.x * 1e-6f, ub * 0.5f + i);
    }
    const float2 u2a = make_float2(ua, ua), u2b = make_float2(ub, ub);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                if (MODE == 0) { x[i].x = fmaf(x[i].x, ua, ub); x[i].y = fmaf(x[i].y, ua, ub); }
                if (MODE == 1) { x[i].x = fmaf(x[i].x, a[i].x, b[i].x); x[i].y = fmaf(x[i].y, a[i].y, b[i].y); }
                if (MODE == 2) x[i] = __ffma2_rn(x[i], u2a, u2b);
                if (MODE == 3) x[i] = __ffma2_rn(x[i], a[i], b[i]);
            }
        }
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += x[i].x + x[i].y;
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

}  // namespace fe
