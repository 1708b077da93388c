// Host-side re-layout of the configuration tables into the shapes the K1 phases read
// (shared by the library's fe_configure and by tests/host_sim, so the CPU replay of the
// kernel dataflow sees byte-identical tables).
#pragma once
#include <algorithm>
#include <vector>
#include "../../include/asr_frontend.h"
#include "fe_core.cuh"

namespace fe {

struct HostTables {
    std::vector<float> tw256;       // [15][16][4] stage-B input twiddles, row j - 1, cfg = flip*8 + t
    std::vector<float> tw512;       // [8][16][4]  post-pass twiddles, row k2, cfg = flip*8 + t
    std::vector<float> window;      // [rows*32]    (w[2m], w[2m+1]) per complex point
    std::vector<int> mel_desc;      // [mel_groups] flat group list per epilogue warp: row | last << 10 | filter << 16
    std::vector<float> mel_w;       // [2][mel_groups*4] weights in the same order: int16-count scale, then float scale
    int epi_off[kEpiWarps] = {0}, epi_cnt[kEpiWarps] = {0};     // each warp's slice of the list (count % 4 == 0)
    std::vector<float> dctf;        // [D][dct_stride]
    int mel_groups = 0, nh = 0, dct_stride = 0;
    int p_rows = 0;                 // rows of the power buffer: 129 (bins 0..128) or 257, + 3 zero pad rows
    bool ok = true;                 // false: a run does not fit the descriptor fields
    int epi_plan = 0;               // 0 generic epilogue; specialised: 1 mfcc 40 -> 13, 2 fbank-80, 3 mfcc 40 -> 39, 4 fbank-40
    std::vector<float> epi_w;       // [2][epi_w_n] specialised epilogue's weights: mel CSR (pre-scaled) + folded DCT rows
    int epi_w_n = 0;
};

inline void build_host_tables(const fe_config& c, HostTables& t) {
    // stage-B input twiddles W_256^(j row), j = 1..15: (wr(rx), wr(ry), wi(rx), wi(ry)) for the lane's two rows
    t.tw256.assign(15 * 16 * 4, 0.f);
    for (int j = 1; j < 16; ++j)
        for (int cfg = 0; cfg < 16; ++cfg) {
            const int tt = cfg & 7, fs = (cfg >> 3) << 1;
            const int rx = row_x(tt, fs), ry = row_y(tt, fs);
            float* o = &t.tw256[((j - 1) * 16 + cfg) * 4];
            o[0] = c.tw256[(j * 16 + rx) * 2];     o[1] = c.tw256[(j * 16 + ry) * 2];
            o[2] = c.tw256[(j * 16 + rx) * 2 + 1]; o[3] = c.tw256[(j * 16 + ry) * 2 + 1];
        }
    // post-pass twiddles: (cos, cos, sin, sin) of 2 pi k / 512 for k = rx + 16 k2 and ry + 16 k2
    t.tw512.assign(8 * 16 * 4, 0.f);
    for (int k2 = 0; k2 < 8; ++k2)
        for (int cfg = 0; cfg < 16; ++cfg) {
            const int tt = cfg & 7, fs = (cfg >> 3) << 1;
            const int kx = row_x(tt, fs) + 16 * k2, ky = row_y(tt, fs) + 16 * k2;
            float* o = &t.tw512[(k2 * 16 + cfg) * 4];
            o[0] = c.tw512[kx * 2];     o[1] = c.tw512[ky * 2];
            o[2] = c.tw512[kx * 2 + 1]; o[3] = c.tw512[ky * 2 + 1];
        }
    // mel plan: runs are padded to a multiple of 4 weights (zeros); reads past bin 128 / 256 land in the
    // three zero pad rows of the power buffer
    const int nf = c.num_filters;
    int max_bin = 0;
    for (int m = 0; m < nf; ++m) {
        const int n = c.fb_row_start[m + 1] - c.fb_row_start[m];
        if (n > 0) max_bin = std::max(max_bin, c.fb_first_bin[m] + n - 1);
    }
    t.p_rows = (max_bin > 128 ? 257 : 129) + 3;
    // Epilogue warp w owns the filter pairs (p, nf-1-p), p = w, w + 4, ... (a narrow low filter with a wide
    // high one: balanced work).  Its filters are flattened into one list of 4-bin weight groups, the last
    // group of every filter flagged; the list is padded with zero groups to a multiple of 4 and followed by
    // 8 more zero groups so that the software pipeline can prefetch past the end without bounds checks.
    t.nh = (nf + 1) / 2;
    t.mel_desc.clear();
    std::vector<float> w16, wf;
    for (int w = 0; w < kEpiWarps; ++w) {
        t.epi_off[w] = (int)t.mel_desc.size();
        int cnt = 0;
        for (int p = w; p < t.nh; p += kEpiWarps) {
            const int pair[2] = {p, nf - 1 - p};
            for (int k = 0; k < (pair[1] != pair[0] ? 2 : 1); ++k) {
                const int m = pair[k];
                const int n = c.fb_row_start[m + 1] - c.fb_row_start[m];
                const int n4 = std::max(1, (n + 3) / 4);                    // an empty filter still emits its (zero) value
                const int b0 = n > 0 ? c.fb_first_bin[m] : 0;
                if (b0 + 4 * n4 > t.p_rows || m > 32767) t.ok = false;
                for (int q = 0; q < n4; ++q) {
                    t.mel_desc.push_back((b0 + 4 * q) | ((q == n4 - 1) ? 1 << 10 : 0) | (m << 16));
                    for (int i = 0; i < 4; ++i) {
                        const int idx = 4 * q + i;
                        const float v = idx < n ? c.fb_weights[c.fb_row_start[m] + idx] * (1.0f / 2048.0f) : 0.f;   // columns hold |2X|^2
                        w16.push_back(v * (1.0f / 1073741824.0f)); wf.push_back(v);
                    }
                    ++cnt;
                }
            }
        }
        const int padded = (cnt + 3) & ~3;
        t.epi_cnt[w] = padded;
        for (int q = cnt; q < padded + 8; ++q) { t.mel_desc.push_back(0); for (int i = 0; i < 4; ++i) { w16.push_back(0.f); wf.push_back(0.f); } }
    }
    t.mel_groups = (int)t.mel_desc.size();
    t.mel_w = w16;
    t.mel_w.insert(t.mel_w.end(), wf.begin(), wf.end());
    // folded DCT: y_c = sum_{n < nh} C[c][n] * (x[n] + (-1)^c x[nf-1-n])
    t.dct_stride = 0;
    t.dctf.clear();
    if (c.feat_type == FE_FEAT_MFCC) {
        t.dct_stride = (t.nh + 3) & ~3;                  // 16-byte groups, read warp-uniformly
        t.dctf.assign((size_t)c.feat_dim * t.dct_stride, 0.f);
        for (int k = 0; k < c.feat_dim; ++k)
            for (int m = 0; m < t.nh; ++m) t.dctf[(size_t)k * t.dct_stride + m] = c.dct[k * nf + m];
    }
    // specialised epilogue: only when the caller's filterbank has exactly a baked structure
    t.epi_plan = 0; t.epi_w.clear(); t.epi_w_n = 0;
    const bool is40 = plan_matches<PlanMfcc40>(c.fb_row_start, c.fb_first_bin, nf);
    if (c.feat_type == FE_FEAT_MFCC && c.feat_dim == 13 && is40) t.epi_plan = 1;
    if (c.feat_type == FE_FEAT_FBANK && plan_matches<PlanFbank80>(c.fb_row_start, c.fb_first_bin, nf)) t.epi_plan = 2;
    if (c.feat_type == FE_FEAT_MFCC && c.feat_dim == 39 && is40) t.epi_plan = 3;       // run.sh:41-50 default feat_dim
    if (c.feat_type == FE_FEAT_FBANK && is40) t.epi_plan = 4;
    if (t.epi_plan) {
        const bool mfcc = c.feat_type == FE_FEAT_MFCC;
        const int nnz = c.fb_nnz, nd = mfcc ? c.feat_dim * t.nh : 0;
        t.epi_w_n = nnz + nd;
        t.epi_w.assign((size_t)2 * t.epi_w_n, 0.f);
        for (int i = 0; i < nnz; ++i) {
            const float v = c.fb_weights[i] * (1.0f / 2048.0f);
            t.epi_w[i] = v * (1.0f / 1073741824.0f);
            t.epi_w[(size_t)t.epi_w_n + i] = v;
        }
        for (int k = 0; k < (mfcc ? c.feat_dim : 0); ++k)
            for (int m = 0; m < t.nh; ++m)
                t.epi_w[nnz + k * t.nh + m] = t.epi_w[(size_t)t.epi_w_n + nnz + k * t.nh + m] = c.dct[k * nf + m];
    }
    t.window.clear();
    if (c.window) {
        const int rows = (c.frame_len + 31) / 32;
        t.window.assign((size_t)rows * 32, 0.f);
        for (int n = 0; n < c.frame_len; ++n) t.window[n] = c.window[n];
    }
}

}  // namespace fe
