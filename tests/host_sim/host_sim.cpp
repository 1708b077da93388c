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
#include "../../automatic-speech-recognition_b200/csrc/fe_tables.h"

using namespace fe;

namespace {
void fill_tables(const fe_config& c, const HostTables& ht, bool in_f32, SmemTables& tb) {
    tb.tw256 = reinterpret_cast<const float4*>(ht.tw256.data());
    tb.tw512 = reinterpret_cast<const float4*>(ht.tw512.data());
    tb.window = c.window ? reinterpret_cast<const float2*>(ht.window.data()) : nullptr;
    tb.mel_n4 = ht.mel_n4; tb.mel_bi = ht.mel_bi.data();
    tb.mel_w = ht.mel_w.data() + (in_f32 ? (size_t)ht.mel_entries * 8 : 0);
    tb.dctf = ht.dctf.data();
    tb.mel_slots = ht.mel_slots; tb.nf = c.num_filters; tb.D = c.feat_dim; tb.dct_stride = ht.dct_stride; tb.nh = ht.nh;
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
    float* e_w = static_cast<float*>(aligned_alloc(2048, kWarpFrames * kERegion * sizeof(float)));
    unsigned char* raw = static_cast<unsigned char*>(aligned_alloc(16, (3 * HOP + 13 * 32) * 4));
    float scr_w[64];
    static LaneZ z[32];
    for (int f0 = 0; f0 < L; f0 += kWarpFrames) {
        const int nfw = (L - f0) < kWarpFrames ? (L - f0) : kWarpFrames;
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
            if (fs < nfw) stage_b(e_w + fs * kERegion, z[lane], t, fs);
        }
        for (int lane = 0; lane < 32; ++lane) {
            int fs = lane >> 3, t = lane & 7;
            if (fs >= nfw) continue;
            float x0, x256;
            post_pass(z[lane], power_row(e_w, fs), tb, t, fs, x0, x256);
            if (t == 0) {
                float s = 0.f;
                for (int i = 0; i < 8; ++i) s += scr_w[fs * 8 + i];
                scr_w[32 + fs] = frame_energy(s, x0, x256, tb.pscale);
            }
        }
        for (int lane = 0; lane < 32; ++lane) if ((lane >> 3) < nfw) mel_phase(e_w, tb, lane & 7, lane >> 3);
        float* dst = statics + (long long)f0 * D;
        if (tb.is_mfcc && (tb.nf & 7) == 0) {
            for (int lane = 0; lane < 32; ++lane)
                if ((lane >> 3) < nfw) dct_phase_fused(e_w, scr_w + 32, tb, lane & 7, lane >> 3, dst + (lane >> 3) * D, true);
        } else if (tb.is_mfcc) {
            for (int lane = 0; lane < 32; ++lane) if ((lane >> 3) < nfw) fold_phase(e_w, tb, lane & 7, lane >> 3);
            for (int lane = 0; lane < 32; ++lane)
                if ((lane >> 3) < nfw) dct_phase(e_w, scr_w + 32, tb, lane & 7, lane >> 3, dst + (lane >> 3) * D, true);
        } else {
            for (int f = 0; f < nfw; ++f) for (int m = 0; m < D; ++m) dst[f * D + m] = logmel_row(e_w, f)[m];
        }
    }
    free(e_w); free(raw);
    return L;
}
}  // namespace

// statics (L, D) of one utterance through the simulated K1; returns L or < 0
extern "C" int sim_statics(const fe_config* c, const void* pcm, int n_samples, float* statics) {
    return c->pcm_dtype == FE_PCM_FLOAT32 ? run<1>(*c, pcm, n_samples, statics) : run<0>(*c, pcm, n_samples, statics);
}
