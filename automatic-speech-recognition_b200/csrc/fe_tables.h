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
    std::vector<float> tw256;       // [6][16][4]  twiddle bases m = 1,2,3,4,8,12; cfg = swap*8 + t
    std::vector<float> tw512;       // [16][4]     post-pass bases; cfg = flip*8 + t
    std::vector<float> window;      // [rows*32]    (w[2m], w[2m+1]) per complex point
    std::vector<int> mel_slot_off, mel_b0, mel_id, mel_bi;    // mel_bi = id << 16 | first bin
    std::vector<float> mel_w;       // [2][entries*8]: int16-count scale, then float scale
    std::vector<float> dctf;        // [D][dct_stride]
    int mel_slots = 0, mel_entries = 0, nh = 0, dct_stride = 0;
    int mel_n4[16] = {0}, mel_e4[16] = {0};
};

inline void build_host_tables(const fe_config& c, HostTables& t) {
    // stage-A twiddle bases: (wr(jx m), wr(jy m), wi(jx m), wi(jy m)), jx = t + 8 swap, jy = t + 8 (1 - swap),
    // for m = 1, 2, 3 (W^(j b)) and m = 4, 8, 12 (W^(4 j a))
    t.tw256.assign(6 * 16 * 4, 0.f);
    const int ms[6] = {1, 2, 3, 4, 8, 12};
    for (int mi = 0; mi < 6; ++mi)
        for (int cfg = 0; cfg < 16; ++cfg) {
            const int swap = cfg >> 3, tt = cfg & 7, m = ms[mi];
            const int jx = tt + 8 * swap, jy = tt + 8 * (1 - swap);
            float* o = &t.tw256[(mi * 16 + cfg) * 4];
            o[0] = c.tw256[(jx * 16 + m) * 2];     o[1] = c.tw256[(jy * 16 + m) * 2];
            o[2] = c.tw256[(jx * 16 + m) * 2 + 1]; o[3] = c.tw256[(jy * 16 + m) * 2 + 1];
        }
    // post-pass twiddle bases: (cos rx, cos ry, sin rx, sin ry) (angles 2 pi r / 512); bins r + 16 k2 by rotation
    t.tw512.assign(16 * 4, 0.f);
    for (int cfg = 0; cfg < 16; ++cfg) {
        const int tt = cfg & 7, fs = (cfg >> 3) << 1;
        const int kx = row_x(tt, fs), ky = row_y(tt, fs);
        float* o = &t.tw512[cfg * 4];
        o[0] = c.tw512[kx * 2];     o[1] = c.tw512[ky * 2];
        o[2] = c.tw512[kx * 2 + 1]; o[3] = c.tw512[ky * 2 + 1];
    }
    // mel plan: filters sorted by run length, 8 per slot (one per lane of a frame); every slot
    // is padded to the longest run in it (multiple of 4) so trip counts are lane-uniform
    const int nf = c.num_filters;
    std::vector<int> order(nf);
    for (int m = 0; m < nf; ++m) order[m] = m;
    std::stable_sort(order.begin(), order.end(), [&](int a, int b) {
        return c.fb_row_start[a + 1] - c.fb_row_start[a] < c.fb_row_start[b + 1] - c.fb_row_start[b]; });
    const int S = (nf + 7) / 8;
    t.mel_slot_off.assign(S + 1, 0); t.mel_b0.assign(S * 8, 0); t.mel_id.assign(S * 8, -1);
    // power-row reads are 16-byte loads: every run starts at its first bin rounded down to a
    // multiple of 4 (weights shifted right by the remainder, zero filled)
    for (int s = 0; s < S; ++s) {
        int e = 0;
        for (int g = 0; g < 8 && s * 8 + g < nf; ++g) {
            const int m = order[s * 8 + g];
            const int n = c.fb_row_start[m + 1] - c.fb_row_start[m];
            e = std::max(e, n > 0 ? (c.fb_first_bin[m] & 3) + n : 0);
        }
        e = std::max(4, (e + 3) & ~3);
        t.mel_slot_off[s + 1] = t.mel_slot_off[s] + e;
    }
    const int entries = t.mel_slot_off[S];
    t.mel_w.assign((size_t)entries * 8 * 2, 0.f);
    for (int s = 0; s < S; ++s) {
        const int e = t.mel_slot_off[s + 1] - t.mel_slot_off[s];
        for (int g = 0; g < 8 && s * 8 + g < nf; ++g) {
            const int m = order[s * 8 + g];
            const int n = c.fb_row_start[m + 1] - c.fb_row_start[m];
            int bin0 = n > 0 ? (c.fb_first_bin[m] & ~3) : 0;
            int shift = n > 0 ? (c.fb_first_bin[m] & 3) : 0;
            // keep padded reads inside the 260-float row (257 bins + 3 of slack after it)
            while (bin0 + e > 260) { bin0 -= 4; shift += 4; }
            t.mel_id[s * 8 + g] = m; t.mel_b0[s * 8 + g] = bin0;
            for (int i = 0; i < n; ++i) {
                const float v = c.fb_weights[c.fb_row_start[m] + i] * (1.0f / 2048.0f);   // rows hold |2X|^2
                const int ee = t.mel_slot_off[s] + shift + i;       // float4 groups: [entry/4][lane][entry%4]
                const size_t at = ((size_t)(ee >> 2) * 8 + g) * 4 + (ee & 3);
                t.mel_w[at] = v * (1.0f / 1073741824.0f);
                t.mel_w[(size_t)entries * 8 + at] = v;
            }
        }
    }
    t.mel_slots = S; t.mel_entries = entries;
    for (int s = 0; s < S && s < 16; ++s) { t.mel_e4[s] = t.mel_slot_off[s] >> 2; t.mel_n4[s] = (t.mel_slot_off[s + 1] - t.mel_slot_off[s]) >> 2; }
    t.mel_bi.assign(S * 8, 0);
    for (int i = 0; i < S * 8; ++i) t.mel_bi[i] = ((t.mel_id[i] < 0 ? 0xffff : t.mel_id[i]) << 16) | t.mel_b0[i];
    // folded DCT: y_c = sum_{n < nh} C[c][n] * (x[n] + (-1)^c x[nf-1-n])
    t.nh = (nf + 1) / 2;
    t.dct_stride = 0;
    t.dctf.clear();
    if (c.feat_type == FE_FEAT_MFCC) {
        t.dct_stride = (t.nh + 3) & ~3;                  // 16-byte rows; stride/4 odd -> the 8 rows read by a
        if (((t.dct_stride >> 2) & 1) == 0) t.dct_stride += 4;   // frame's lanes fall in distinct bank groups
        t.dctf.assign((size_t)c.feat_dim * t.dct_stride, 0.f);
        for (int k = 0; k < c.feat_dim; ++k)
            for (int m = 0; m < t.nh; ++m) t.dctf[(size_t)k * t.dct_stride + m] = c.dct[k * nf + m];
    }
    t.window.clear();
    if (c.window) {
        const int rows = (c.frame_len + 31) / 32;
        t.window.assign((size_t)rows * 32, 0.f);
        for (int n = 0; n < c.frame_len; ++n) t.window[n] = c.window[n];
    }
}

}  // namespace fe
