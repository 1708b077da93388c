// TEST INFRASTRUCTURE ONLY.  Replays the lane-level dataflow of K1
// (automatic-speech-recognition_b200/csrc/fe_core.cuh) on the CPU: lanes are a loop,
// shared memory is an array, __syncwarp is a phase boundary, packed f32x2 ops are
// emulated component-wise.  It lets the "not gpu" test-suite check the exact index
// math / swizzles / FFT codelets of the kernel against the float64 oracle without a GPU.
// Never linked into the product library.
#include <vector>
#include <cstring>
#include <cstdint>
#include <cstdlib>
#include <cmath>
#include "../../automatic-speech-recognition_b200/csrc/fe_tables.h"
#include "../../automatic-speech-recognition_b200/csrc/fe_k1t.cuh"

using namespace fe;

namespace {
void fill_tables(const fe_config& c, const HostTables& ht, bool in_f32, SmemTables& tb) {
    tb.tw256 = reinterpret_cast<const float4*>(ht.tw256.data());
    tb.tw512 = reinterpret_cast<const float4*>(ht.tw512.data());
    tb.window = c.window ? reinterpret_cast<const float2*>(ht.window.data()) : nullptr;
    tb.mel_desc = ht.mel_desc.data();
    tb.mel_w4 = reinterpret_cast<const float4*>(ht.mel_w.data() + (in_f32 ? (size_t)ht.mel_groups * 4 : 0));
    tb.dctf = ht.dctf.data();
    tb.nf = c.num_filters; tb.D = c.feat_dim; tb.dct_stride = ht.dct_stride; tb.nh = ht.nh;
    int max_bin = 0;
    for (int m = 0; m < c.num_filters; ++m) {
        int w = c.fb_row_start[m + 1] - c.fb_row_start[m];
        if (w > 0 && c.fb_first_bin[m] + w - 1 > max_bin) max_bin = c.fb_first_bin[m] + w - 1;
    }
    tb.full_spectrum = max_bin > 128;
    tb.is_mfcc = c.feat_type == FE_FEAT_MFCC; tb.fbank_log = c.fbank_log; tb.dc_elim = c.dc_elimination;
    tb.pscale = in_f32 ? 1.0f : 1.0f / 1073741824.0f;
}

template <int IN_F32>
int run(const fe_config& c, const void* pcm, int n_samples, float* statics) {
    constexpr int FL = 400, HOP = 160, ESZ = IN_F32 ? 4 : 2;
    if (c.frame_len != FL || c.hop != HOP) return -1;
    const int L = n_samples < FL ? 0 : (n_samples - FL) / HOP;
    HostTables ht; build_host_tables(c, ht);
    SmemTables tb; fill_tables(c, ht, IN_F32, tb);
    const int D = c.feat_dim;
    if (!ht.ok) return -2;
    float* e_w = static_cast<float*>(aligned_alloc(2048, kWarpFrames * kERegion * sizeof(float)));
    unsigned char* raw = static_cast<unsigned char*>(aligned_alloc(16, (3 * HOP + 13 * 32) * 4));
    std::vector<float> pbuf((size_t)ht.p_rows * kPStride, 0.f), lm((size_t)(c.num_filters + 4) * 32, 0.f), out_t((size_t)D * 32), energy(32, 1.f);
    float scr_w[32];
    static LaneZ z[32];
    for (int t0 = 0; t0 < L; t0 += kTileFrames) {                       // one CTA tile
        const int ntile = (L - t0) < kTileFrames ? (L - t0) : kTileFrames;
        for (int r = 0; r < ht.p_rows - 3; ++r) for (int f = 0; f < 32; ++f) pbuf[(size_t)r * kPStride + f] = 1e30f;   // poison
        for (int warp = 0; warp < kTileGroups; ++warp) {
            const int f0 = t0 + warp * kWarpFrames;
            const int nfw = (ntile - warp * kWarpFrames) < kWarpFrames ? (ntile - warp * kWarpFrames) : kWarpFrames;
            if (nfw <= 0) continue;
            for (int i = 0; i < kWarpFrames * kERegion; ++i) e_w[i] = 1e30f;      // poison: catches reads of unwritten words
            memset(raw, 0x7f, (3 * HOP + 13 * 32) * 4);
            memcpy(raw, static_cast<const unsigned char*>(pcm) + (size_t)f0 * HOP * ESZ, (size_t)((nfw - 1) * HOP + FL) * ESZ);
            for (int lane = 0; lane < 32; ++lane) {
                int fs = lane >> 3, t = lane & 7;
                if (fs < nfw) scr_w[lane] = c.window ? stage_a<FL, IN_F32, 1>(raw + fs * HOP * ESZ, e_w + fs * kERegion, tb, t, fs)
                                                     : stage_a<FL, IN_F32, 0>(raw + fs * HOP * ESZ, e_w + fs * kERegion, tb, t, fs);
            }
            for (int lane = 0; lane < 32; ++lane) {
                int fs = lane >> 3, t = lane & 7;
                if (fs < nfw) stage_b(e_w + fs * kERegion, z[lane], tb, t, fs);
            }
            for (int lane = 0; lane < 32; ++lane) {
                int fs = lane >> 3, t = lane & 7;
                if (fs >= nfw) continue;
                float x0, x256;
                post_pass(z[lane], pbuf.data() + warp * kWarpFrames + fs, tb, t, fs, x0, x256);
                if (t == 0) {
                    float s = 0.f;
                    for (int i = 0; i < 8; ++i) s += scr_w[fs * 8 + i];
                    energy[warp * kWarpFrames + fs] = frame_energy(s, x0, x256, tb.pscale);
                }
            }
        }
        // CTA-wide epilogue: lane = frame (only the tile's real frames are replayed), warp = filter / coefficient group
        for (int warp = 0; warp < kEpiWarps; ++warp)
            for (int lane = 0; lane < ntile; ++lane) epi_mel(pbuf.data(), tb.is_mfcc ? lm.data() : out_t.data(), tb, ht.epi_off[warp], ht.epi_cnt[warp], lane);
        if (tb.is_mfcc)
            for (int warp = 0; warp < kEpiWarps; ++warp)
                for (int lane = 0; lane < ntile; ++lane) epi_dct(lm.data(), energy.data(), out_t.data(), tb, warp, lane);
        for (int f = 0; f < ntile; ++f) for (int m = 0; m < D; ++m) statics[(long long)(t0 + f) * D + m] = out_t[(size_t)m * 32 + f];
    }
    free(e_w); free(raw);
    return L;
}
}  // namespace

// ---- K1T (lane = frame, exchange in tensor memory): the per-lane phases with the exchange as a plain array ----
namespace {
struct HostExchange {
    float m[512];
    bool written[512];
    HostExchange() { for (int i = 0; i < 512; ++i) { m[i] = 1e30f; written[i] = false; } }
    void st2(int col, float a, float b) { m[col] = a; m[col + 1] = b; written[col] = written[col + 1] = true; }
    void ld32(int col, float* v) { for (int i = 0; i < 32; ++i) v[i] = written[col + i] ? m[col + i] : NAN; }
    void wait_ld() {}
    void wait_st() {}
};

int run_t(const fe_config& c, const short* pcm, int n_samples, float* statics) {
    constexpr int FL = 400, HOP = 160;
    if (c.frame_len != FL || c.hop != HOP || c.window || c.pcm_dtype != FE_PCM_INT16) return -1;
    const int L = n_samples < FL ? 0 : (n_samples - FL) / HOP;
    HostTables ht; build_host_tables(c, ht);
    SmemTables tb; fill_tables(c, ht, false, tb);
    if (!ht.ok || tb.full_spectrum) return -2;
    std::vector<float4> t256p(128), t512p(64);
    std::vector<float2> t512(129);
    k1t_build_twiddles(t256p.data(), t512p.data(), t512.data());
    TTwiddles tw{t256p.data(), t512p.data(), t512.data()};
    const int D = c.feat_dim;
    std::vector<float> pbuf((size_t)ht.p_rows * kPStride, 0.f), lm((size_t)(c.num_filters + 4) * 32, 0.f), out_t((size_t)D * 32), energy(32, 1.f);
    alignas(16) uint32_t raw[204];
    for (int t0 = 0; t0 < L; t0 += kTileFrames) {
        const int ntile = (L - t0) < kTileFrames ? (L - t0) : kTileFrames;
        for (int r = 0; r < ht.p_rows - 3; ++r) for (int f = 0; f < 32; ++f) pbuf[(size_t)r * kPStride + f] = 1e30f;   // poison
        for (int lane = 0; lane < ntile; ++lane) {
            memset(raw, 0x7f, sizeof(raw));
            memcpy(raw, pcm + (size_t)(t0 + lane) * HOP, FL * 2);
            HostExchange ex;
            energy[lane] = k1t_frame(reinterpret_cast<const uint4*>(raw), ex, tw, pbuf.data() + lane, tb.pscale);
        }
        for (int warp = 0; warp < kEpiWarps; ++warp)
            for (int lane = 0; lane < ntile; ++lane) epi_mel(pbuf.data(), tb.is_mfcc ? lm.data() : out_t.data(), tb, ht.epi_off[warp], ht.epi_cnt[warp], lane);
        if (tb.is_mfcc)
            for (int warp = 0; warp < kEpiWarps; ++warp)
                for (int lane = 0; lane < ntile; ++lane) epi_dct(lm.data(), energy.data(), out_t.data(), tb, warp, lane);
        for (int f = 0; f < ntile; ++f) for (int m = 0; m < D; ++m) statics[(long long)(t0 + f) * D + m] = out_t[(size_t)m * 32 + f];
    }
    return L;
}
}  // namespace

// statics (L, D) of one utterance through the simulated K1T (int16 PCM, rectangular window); returns L or < 0
extern "C" int sim_statics_t(const fe_config* c, const void* pcm, int n_samples, float* statics) {
    return run_t(*c, static_cast<const short*>(pcm), n_samples, statics);
}

// statics (L, D) of one utterance through the simulated K1; returns L or < 0
extern "C" int sim_statics(const fe_config* c, const void* pcm, int n_samples, float* statics) {
    return c->pcm_dtype == FE_PCM_FLOAT32 ? run<1>(*c, pcm, n_samples, statics) : run<0>(*c, pcm, n_samples, statics);
}
