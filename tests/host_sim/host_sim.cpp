// TEST INFRASTRUCTURE ONLY.  Replays the lane-level dataflow of K1
// (automatic-speech-recognition_b200/csrc/fe_core.cuh) on the CPU: lanes are a loop,
// shared memory is an array, __syncwarp is a phase boundary.  It lets the "not gpu"
// test-suite check the exact index math / swizzles / FFT codelets of the kernel
// against the float64 oracle without a GPU.  Never linked into the product library.
#include <vector>
#include <cstring>
#include <cstdint>
#include "../../automatic-speech-recognition_b200/csrc/fe_core.cuh"

using namespace fe;

extern "C" int sim_statics_i16(const int16_t* pcm, int n_samples, int frame_len, int hop,
                               int nf, int D, int is_mfcc, int fbank_log, int dc_elim,
                               const int* fb_start, const int* fb_bin0, const float* fb_w /*unscaled*/,
                               const float* dct /*[D*nf]*/, const float* window /*[frame_len] or null*/,
                               const float* tw256 /*[256*2]*/, const float* tw512 /*[257*2]*/,
                               float* statics /*[L*D]*/) {
    if (frame_len != 400 || hop != 160) return -1;
    constexpr int FL = 400, HOP = 160;
    const int L = n_samples < FL ? 0 : (n_samples - FL) / HOP;
    // tables as the kernel sees them in shared memory
    std::vector<float2> s_tw256(16 * kTw256Stride), s_tw512(kBins);
    for (int i = 0; i < 256; ++i) s_tw256[(i >> 4) * kTw256Stride + (i & 15)] = make_float2(tw256[2 * i], tw256[2 * i + 1]);
    for (int i = 0; i < kBins; ++i) s_tw512[i] = make_float2(tw512[2 * i], tw512[2 * i + 1]);
    const int nnz = fb_start[nf];
    std::vector<float> s_fbw(nnz > 0 ? nnz : 1);
    int max_bin = 0;
    for (int m = 0; m < nf; ++m) { int w = fb_start[m + 1] - fb_start[m]; if (w > 0 && fb_bin0[m] + w - 1 > max_bin) max_bin = fb_bin0[m] + w - 1; }
    for (int i = 0; i < nnz; ++i) s_fbw[i] = fb_w[i] * (1.0f / 2048.0f);
    const int nf4 = (nf + 3) & ~3, dct_stride = nf4 + 4;
    std::vector<float> s_dct((size_t)D * dct_stride + 4, 0.f);
    if (is_mfcc) for (int k = 0; k < D; ++k) for (int m = 0; m < nf; ++m) s_dct[(size_t)k * dct_stride + m] = dct[k * nf + m];
    std::vector<float> s_win;
    if (window) { s_win.assign(13 * 32, 0.f); for (int n = 0; n < FL; ++n) s_win[(n / 32) * 32 + pcm_pos(n % 32)] = window[n]; }
    SmemTables tb;
    tb.tw256 = s_tw256.data(); tb.tw512 = s_tw512.data(); tb.window = window ? s_win.data() : nullptr;
    tb.fb_start = fb_start; tb.fb_bin0 = fb_bin0; tb.fb_w = s_fbw.data(); tb.dct = s_dct.data();
    tb.nf = nf; tb.D = D; tb.dct_stride = dct_stride; tb.full_spectrum = max_bin > 128;
    tb.is_mfcc = is_mfcc; tb.fbank_log = fbank_log; tb.dc_elim = dc_elim;

    alignas(16) static float pcm_w[3 * HOP + 13 * 32];
    alignas(16) static float e_w[kWarpFrames * kERegion];
    float scr_w[64];
    for (int f0 = 0; f0 < L; f0 += kWarpFrames) {
        const int nfw = (L - f0) < kWarpFrames ? (L - f0) : kWarpFrames;
        // poison shared memory so that reads of never-written words show up
        for (auto& v : pcm_w) v = 1e30f;
        for (auto& v : e_w) v = 1e30f;
        const int n_samp = (nfw - 1) * HOP + FL;
        const int items = ((n_samp + 31) >> 5) << 3;
        for (int lane = 0; lane < 32; ++lane)
            for (int id = lane; id < items; id += 32) stage_item_i16(pcm + (long long)f0 * HOP, n_samp, id, pcm_w);
        // phase 1
        for (int lane = 0; lane < 32; ++lane) {
            int fs = lane >> 3, t = lane & 7;
            if (fs < nfw) scr_w[lane] = stage_a<FL>(pcm_w + fs * HOP, e_w + fs * kERegion, tb, t, fs);
        }
        // phase 2 (all lanes load before anybody overwrites: registers per lane kept in z[])
        static LaneZ z[32];
        for (int lane = 0; lane < 32; ++lane) {
            int fs = lane >> 3, t = lane & 7;
            if (fs < nfw) stage_b(e_w + fs * kERegion, z[lane], t, fs);
        }
        // phase 3
        for (int lane = 0; lane < 32; ++lane) {
            int fs = lane >> 3, t = lane & 7;
            if (fs >= nfw) continue;
            float x0, x256;
            post_pass(z[lane], power_row(e_w, fs), tb, t, x0, x256);
            if (t == 0) {
                float s = 0.f;
                for (int i = 0; i < 8; ++i) s += scr_w[fs * 8 + i];
                scr_w[32 + fs] = frame_energy(s, x0, x256);
            }
        }
        // phase 4
        for (int lane = 0; lane < 32; ++lane)
            for (int id = lane; id < kWarpFrames * nf4; id += 32) mel_phase(e_w, tb, id, nfw);
        // phase 5
        float* dst = statics + (long long)f0 * D;
        for (int lane = 0; lane < 32; ++lane)
            for (int id = lane; id < nfw * D; id += 32) dst[id] = emit_phase(e_w, scr_w + 32, tb, id);
    }
    return L;
}

// raw 512-point power spectrum of one 400-sample frame through the same phases (debug aid)
extern "C" void sim_power_frame(const float* frame400, const float* tw256, const float* tw512, float* p257 /* |X|^2/512 */) {
    std::vector<float2> s_tw256(16 * kTw256Stride), s_tw512(kBins);
    for (int i = 0; i < 256; ++i) s_tw256[(i >> 4) * kTw256Stride + (i & 15)] = make_float2(tw256[2 * i], tw256[2 * i + 1]);
    for (int i = 0; i < kBins; ++i) s_tw512[i] = make_float2(tw512[2 * i], tw512[2 * i + 1]);
    SmemTables tb; memset(&tb, 0, sizeof(tb));
    tb.tw256 = s_tw256.data(); tb.tw512 = s_tw512.data(); tb.full_spectrum = 1;
    alignas(16) float pcm_w[13 * 32];
    alignas(16) float e_f[kERegion];
    for (int n = 0; n < 416; ++n) pcm_w[(n / 32) * 32 + pcm_pos(n % 32)] = n < 400 ? frame400[n] : 0.f;
    for (int t = 0; t < 8; ++t) stage_a<400>(pcm_w, e_f, tb, t, 0);
    LaneZ z[8];
    for (int t = 0; t < 8; ++t) stage_b(e_f, z[t], t, 0);
    float x0, x256;
    for (int t = 0; t < 8; ++t) post_pass(z[t], e_f, tb, t, x0, x256);
    for (int k = 0; k < kBins; ++k) p257[k] = e_f[k] * (1.0f / 2048.0f);
}
