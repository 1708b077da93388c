// C-ABI implementation (see include/asr_frontend.h for the contract and the
// reference call sites each entry point replaces).
#include <cuda_runtime.h>
#include <stdint.h>
#include <string.h>
#include <stdlib.h>
#include <math.h>
#include <string>
#include <vector>
#include <algorithm>

#include "../../include/asr_frontend.h"
#include "fe_kernels.cuh"
#include "flac_gpu.cuh"
#include "fe_tables.h"

using namespace fe;

namespace {

thread_local std::string g_create_error;

struct DevBuf {
    void* p = nullptr;
    size_t cap = 0;
};

struct HostPinned {
    void* p = nullptr;
    size_t cap = 0;
};

}  // namespace

// per-run device/host scratch and the stream it is used on.  Lane 0 serves every call; lanes 1..2
// exist only for the chunked host pipeline of fe_run (H2D / kernels / D2H of neighbouring chunks overlap).
struct Lane {
    cudaStream_t stream = nullptr;
    DevBuf d_utts, d_tile_prefix, d_tiles, d_atile_prefix, d_atiles, d_statics, d_stats, d_pcm, d_out, d_scratch;
    HostPinned h_stage;
    std::vector<UttDesc> utts;
    std::vector<long long> tile_prefix, atile_prefix;
    int stats_utts = 0;               // utterances of the current batch (size of d_stats when a batch runs in L2 groups)
    cudaEvent_t done = nullptr;       // end of the kernels of the last run on this lane
    bool done_valid = false;
};

struct fe_handle {
    int device = 0;
    int num_sms = 0;
    cudaStream_t stream = nullptr;
    bool configured = false;
    fe_config cfg{};
    std::string err;

    // device tables
    DevBuf tw256, tw512, window, mel_desc, mel_w, dct;
    int dct_stride = 0, full_spectrum = 0, mel_groups = 0, p_rows = 0, nh = 0;
    int epi_off[kEpiWarps] = {0}, epi_cnt[kEpiWarps] = {0};
    int epi_plan = 0, epi_w_n = 0;        // specialised epilogue (fe_plans_gen.h) and its weights
    std::vector<float> epi_w;
    bool scratch_f32 = false;     // pre-emphasis materialises float PCM in the scratch buffer
    // resampler
    std::vector<int> sp_up, sp_down, sp_tap_off;
    std::vector<float> sp_taps;       // host copy: the fast kernels take their table as a kernel parameter
    int sp_ntaps = 0;                 // taps per phase
    DevBuf d_sp_up, d_sp_down, d_sp_tap_off, d_taps;
    // bucketed batches (fe_pad_batches): slot table, time of the last profiled launch
    DevBuf d_pad;
    float pad_ms = 0.f;
    // FLAC decode on the device (fe_decode_flac)
    DevBuf d_flac_bytes, d_flac_files, d_flac_frames, d_flac_cands, d_flac_ctr, d_flac_pcm;
    float flac_ms[3] = {0.f, 0.f, 0.f};

    Lane lane[3];

    bool k2_split = false;            // FE_K2_SPLIT=1: statistics and cube as two kernels (K2a + K2b), the round-1 form
    int k1t = 2;                      // 2 (default): K1U where it applies; FE_K1T=0: always K1; FE_K1T=1: the lane = frame / tensor-memory kernel (fe_k1t.cuh) where it applies; measured
                                      // slower than K1 (profiles/r02_k1t.md), kept selectable for A/B runs
    int profiling = 0;
    // profiled runs since fe_set_profiling(1): one event set per run (no sync inside the timed region)
    struct ProfSet { cudaEvent_t e[5]; bool k0, k2; std::vector<cudaEvent_t> ge; int groups = 0; };   // ge: (K1 end, K2 end) per L2 group
    std::vector<ProfSet> prof;
    size_t prof_used = 0;
    int64_t launches = 0;
    size_t k1_smem[2] = {0, 0};      // [raw int16 input, float input]
    long long l2_group_bytes = 0;              // statics per K1 -> K2 group (FE_L2_GROUP_MB; 0 = one launch per kernel, the default:
                                               // grouping measured slower at every size, profiles/r02_l2_groups.jsonl)
    long long pipe_chunk_bytes = 192LL << 20;   // PCM bytes per chunk of the host pipeline (FE_PIPE_CHUNK_MB overrides)
};

namespace {

int fail(fe_handle* h, int code, const std::string& msg) {
    if (h) h->err = msg; else g_create_error = msg;
    return code;
}

#define FE_CUDA(h, expr)                                                                     \
    do {                                                                                     \
        cudaError_t e__ = (expr);                                                            \
        if (e__ != cudaSuccess)                                                              \
            return fail(h, FE_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(e__)); \
    } while (0)

int ensure(fe_handle* h, DevBuf& b, size_t bytes) {
    if (bytes <= b.cap) return FE_OK;
    if (b.p) { FE_CUDA(h, cudaFree(b.p)); b.p = nullptr; b.cap = 0; }
    size_t want = bytes + bytes / 8 + 256;
    FE_CUDA(h, cudaMalloc(&b.p, want));
    b.cap = want;
    return FE_OK;
}

int ensure_pinned(fe_handle* h, HostPinned& b, size_t bytes) {
    if (bytes <= b.cap) return FE_OK;
    if (b.p) { FE_CUDA(h, cudaFreeHost(b.p)); b.p = nullptr; b.cap = 0; }
    size_t want = bytes + bytes / 4 + 4096;
    FE_CUDA(h, cudaMallocHost(&b.p, want));
    b.cap = want;
    return FE_OK;
}

int upload(fe_handle* h, DevBuf& b, const void* src, size_t bytes) {
    int rc = ensure(h, b, bytes ? bytes : 16);
    if (rc) return rc;
    if (bytes) FE_CUDA(h, cudaMemcpy(b.p, src, bytes, cudaMemcpyHostToDevice));
    return FE_OK;
}

void release(DevBuf& b) { if (b.p) cudaFree(b.p); b.p = nullptr; b.cap = 0; }

bool is_device_ptr(const void* p) {
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return false; }
    return a.type == cudaMemoryTypeDevice || a.type == cudaMemoryTypeManaged;
}

inline long long round_up(long long x, long long m) { return (x + m - 1) / m * m; }

// per-utterance host plan shared by fe_plan / fe_run
struct Plan {
    long long total_frames = 0, total_out = 0, total_tiles = 0, total_scratch = 0, total_atiles = 0;
    long long pcm_span = 0;
    bool any_scratch = false;
};

int make_plan(fe_handle* h, Lane& ln, const int64_t* pcm_offsets, const int64_t* pcm_lengths, int32_t n,
              const int32_t* speed_idx, const float* gain, int64_t* out_offsets, int32_t* n_frames,
              Plan& pl, bool fill_desc) {
    const fe_config& c = h->cfg;
    const int width = c.feat_dim * (c.cmvn ? 3 : 1);
    const long long align = c.pcm_dtype == FE_PCM_INT16 ? 8 : 4;
    if (fill_desc) { ln.utts.resize(n); ln.tile_prefix.resize(n + 1); ln.atile_prefix.resize(n + 1); }
    long long out_off = 0;
    for (int i = 0; i < n; ++i) {
        const long long len = pcm_lengths[i];
        if (len < 0 || len > 0x7fffffffLL) return fail(h, FE_ERR_INVALID, "utterance length out of range");
        int sidx = speed_idx ? speed_idx[i] : -1;
        if (sidx >= (int)h->sp_up.size()) return fail(h, FE_ERR_INVALID, "speed_idx not configured");
        if (sidx >= 0 && h->sp_up[sidx] == h->sp_down[sidx]) sidx = -1;
        if (sidx >= 0 && c.pcm_dtype != FE_PCM_INT16)
            return fail(h, FE_ERR_INVALID, "speed perturbation needs int16 PCM");
        long long n_eff = sidx >= 0 ? fe_resampled_length(len, h->sp_up[sidx], h->sp_down[sidx]) : len;
        long long L = fe_num_frames(n_eff, c.frame_len, c.hop);
        if (L > 0x7fffffffLL / (width > 0 ? width : 1)) return fail(h, FE_ERR_INVALID, "utterance too long");
        if (out_offsets) out_offsets[i] = out_off;
        if (n_frames) n_frames[i] = (int32_t)L;
        const float g = gain ? gain[i] : 1.f;
        const bool preemph = c.preemph != 0.f;
        if (preemph && (sidx >= 0 || g != 1.f))
            return fail(h, FE_ERR_INVALID, "pre-emphasis cannot be combined with speed / gain perturbation");
        if (g != 1.f && c.pcm_dtype != FE_PCM_INT16)
            return fail(h, FE_ERR_INVALID, "gain perturbation needs int16 PCM");
        const bool via_scratch = sidx >= 0 || g != 1.f || preemph;
        if (fill_desc) {
            UttDesc& u = ln.utts[i];
            const long long off = pcm_offsets[i];
            if (off < 0 || off % align) return fail(h, FE_ERR_INVALID, "pcm_offsets must be multiples of 16 bytes");
            u.src_off = off;
            u.n_src = (int)len;
            u.n_samples = (int)n_eff;
            u.n_frames = (int)L;
            u.speed_idx = sidx;
            u.gain = g;
            u.src_sel = via_scratch ? 1 : 0;
            u.pcm_off = via_scratch ? pl.total_scratch : off;
            u.out_off = out_off;
            u.stat_off = pl.total_tiles * kTileFrames * c.feat_dim;     // K1 -> K2 statics: [D][32] blocks
            ln.tile_prefix[i] = pl.total_tiles;
            ln.atile_prefix[i] = pl.total_atiles;
            pl.pcm_span = std::max(pl.pcm_span, off + len);
        }
        if (via_scratch) {
            pl.any_scratch = true;
            pl.total_scratch += round_up(n_eff, 8);
            pl.total_atiles += (n_eff + kK0Outputs - 1) / kK0Outputs;
        }
        pl.total_frames += L;
        pl.total_tiles += statics_tiles(L);
        out_off += round_up(L * width, 4);
    }
    if (out_offsets) out_offsets[n] = out_off;
    if (fill_desc) { ln.tile_prefix[n] = pl.total_tiles; ln.atile_prefix[n] = pl.total_atiles; }
    pl.total_out = out_off;
    return FE_OK;
}

template <typename K>
int set_smem(fe_handle* h, K kernel, size_t bytes) {
    FE_CUDA(h, cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
    return FE_OK;
}

int launch_k1(fe_handle* h, cudaStream_t st, const void* pcm, const void* scratch, bool in_f32,
              const TileDesc* tiles, int n_tiles, const DevTables& dt, float* statics) {
    const fe_config& c = h->cfg;
    int grid = std::min<long long>(n_tiles, (long long)h->num_sms);        // persistent: one 16-warp CTA per SM
    if (grid <= 0) return FE_OK;
    const bool geo_main = c.frame_len == 400 && c.hop == 160;
    if (!geo_main && !fe_geometry_supported(c.frame_len, c.hop)) return fail(h, FE_ERR_INVALID, "unsupported frame geometry");
    // int16 PCM, rectangular window, a specialised filterbank plan (everything the reference runs): K1U, lane = frame
    // with the FFT exchange in tensor memory (fe_k1t.cuh, k_frames_to_statics_u).  FE_K1T=0 keeps K1 for A/B runs,
    // FE_K1T=1 selects the first tensor-memory kernel (two warps per tile; plans 1 and 2 only).
    if (geo_main && !in_f32 && c.window == nullptr && h->epi_plan >= 1 && h->epi_plan <= 4 && h->k1t != 0 &&
        !(h->k1t == 1 && h->epi_plan > 2)) {
        K1TParams T;
        memset(T.epi_w, 0, sizeof(T.epi_w));
        memcpy(T.epi_w, h->epi_w.data(), sizeof(float) * (size_t)h->epi_w_n);
        T.pscale = dt.pscale; T.fbank_log = dt.fbank_log; T.dc_elim = dt.dc_elim;
#define FE_LAUNCH_K1U(EPI)                                                                                               \
        do {                                                                                                             \
            FE_CUDA(h, cudaFuncSetAttribute(k_frames_to_statics_u<EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize, kUSmem)); \
            k_frames_to_statics_u<EPI><<<gu, kUThreads, kUSmem, st>>>((const short*)pcm, (const short*)scratch, tiles, n_tiles, T, statics); \
        } while (0)
        if (h->k1t == 2) {
            const int gu = (int)std::min<long long>((n_tiles + kUGroups - 1) / kUGroups, (long long)h->num_sms);
            if (h->epi_plan == 1) FE_LAUNCH_K1U(1);
            else if (h->epi_plan == 2) FE_LAUNCH_K1U(2);
            else if (h->epi_plan == 3) FE_LAUNCH_K1U(3);
            else FE_LAUNCH_K1U(4);
            h->launches++;
            FE_CUDA(h, cudaGetLastError());
            return FE_OK;
        }
#undef FE_LAUNCH_K1U
        const int gt = (int)std::min<long long>((n_tiles + kK1TGroups - 1) / kK1TGroups, (long long)h->num_sms);
        if (h->epi_plan == 1) {
            FE_CUDA(h, cudaFuncSetAttribute(k_frames_to_statics_t<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, kK1TSmem));
            k_frames_to_statics_t<1><<<gt, kK1TThreads, kK1TSmem, st>>>((const short*)pcm, (const short*)scratch, tiles, n_tiles, T, statics);
        } else {
            FE_CUDA(h, cudaFuncSetAttribute(k_frames_to_statics_t<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, kK1TSmem));
            k_frames_to_statics_t<2><<<gt, kK1TThreads, kK1TSmem, st>>>((const short*)pcm, (const short*)scratch, tiles, n_tiles, T, statics);
        }
        h->launches++;
        FE_CUDA(h, cudaGetLastError());
        return FE_OK;
    }
    K1Params P;
    P.dt = dt;
    P.L = k1_smem_layout(c.num_filters, h->mel_groups, h->p_rows, c.feat_dim, h->dct_stride, c.window != nullptr,
                         c.frame_len, c.hop, c.feat_type == FE_FEAT_MFCC, in_f32, h->epi_plan);
    P.dbg = getenv("FE_K1_DBG") ? atoi(getenv("FE_K1_DBG")) : 0;
    const bool win = c.window != nullptr;
    memset(P.epi_w, 0, sizeof(P.epi_w));
    if (h->epi_plan) memcpy(P.epi_w, h->epi_w.data() + (in_f32 ? (size_t)h->epi_w_n : 0), sizeof(float) * (size_t)h->epi_w_n);
#define FE_LAUNCH_K1G(FL, HOP, F32, WIN, EPI)                                                              \
    do {                                                                                                   \
        FE_CUDA(h, cudaFuncSetAttribute(k_frames_to_statics<FL, HOP, F32, WIN, EPI>,                       \
                                        cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->k1_smem[F32])); \
        k_frames_to_statics<FL, HOP, F32, WIN, EPI><<<grid, kK1Threads, h->k1_smem[F32], st>>>(            \
            pcm, scratch, tiles, n_tiles, P, statics);                                                     \
    } while (0)
#define FE_LAUNCH_K1E(F32, WIN, EPI) FE_LAUNCH_K1G(400, 160, F32, WIN, EPI)
    // other frame geometries (the reference takes frame_length / frame_step from its arguments, las/arguments.py:33-40,
    // and fs from every file): generic epilogue only
#define FE_LAUNCH_K1O(FL, HOP)                                                                             \
    do {                                                                                                   \
        if (!in_f32 && !c.window) FE_LAUNCH_K1G(FL, HOP, 0, 0, 0);                                         \
        else if (!in_f32) FE_LAUNCH_K1G(FL, HOP, 0, 1, 0);                                                 \
        else if (!c.window) FE_LAUNCH_K1G(FL, HOP, 1, 0, 0);                                               \
        else FE_LAUNCH_K1G(FL, HOP, 1, 1, 0);                                                              \
    } while (0)
    if (!geo_main) {
        if (c.frame_len == 200 && c.hop == 80) FE_LAUNCH_K1O(200, 80);            // 25 / 10 ms at 8 kHz
        else if (c.frame_len == 320 && c.hop == 160) FE_LAUNCH_K1O(320, 160);     // 20 / 10 ms at 16 kHz (speechpy's own default length)
        else if (c.frame_len == 480 && c.hop == 160) FE_LAUNCH_K1O(480, 160);     // 30 / 10 ms at 16 kHz
        else return fail(h, FE_ERR_INVALID, "unsupported frame geometry");
        h->launches++;
        FE_CUDA(h, cudaGetLastError());
        return FE_OK;
    }
#define FE_LAUNCH_K1(F32, WIN)                                                                             \
    do {                                                                                                   \
        if (h->epi_plan == 1) FE_LAUNCH_K1E(F32, WIN, 1);                                                  \
        else if (h->epi_plan == 2) FE_LAUNCH_K1E(F32, WIN, 2);                                             \
        else if (h->epi_plan == 3) FE_LAUNCH_K1E(F32, WIN, 3);                                             \
        else if (h->epi_plan == 4) FE_LAUNCH_K1E(F32, WIN, 4);                                             \
        else FE_LAUNCH_K1E(F32, WIN, 0);                                                                   \
    } while (0)
    if (!in_f32 && !win) FE_LAUNCH_K1(0, 0);
    else if (!in_f32) FE_LAUNCH_K1(0, 1);
    else if (!win) FE_LAUNCH_K1(1, 0);
    else FE_LAUNCH_K1(1, 1);
#undef FE_LAUNCH_K1E
#undef FE_LAUNCH_K1G
#undef FE_LAUNCH_K1O
#undef FE_LAUNCH_K1
    h->launches++;
    FE_CUDA(h, cudaGetLastError());
    return FE_OK;
}

// K0: speed / gain perturbation.  Ratios 10/9 and 10/11 (the reference's 0.9 / 1.1) go through the register-tiled
// kernels, everything else (other ratios, gain only) through the generic one.  The tile table is grouped by
// class on the host (class_atiles) so that no kernel walks tiles it has to skip.
int atile_class(const fe_handle* h, const UttDesc& u) {
    if (u.speed_idx < 0 || h->sp_ntaps != kK0Taps) return 2;
    const int up = h->sp_up[u.speed_idx], down = h->sp_down[u.speed_idx];
    if (up != 10 || (down != 9 && down != 11)) return 2;
    // two entries with the same ratio could carry different taps: only the first one rides the fast kernel
    for (int i = 0; i < u.speed_idx; ++i) if (h->sp_up[i] == up && h->sp_down[i] == down) return 2;
    return down == 9 ? 0 : 1;
}

void class_atiles(const fe_handle* h, const UttDesc* utts, std::vector<int2>& at, int counts[3]) {
    std::vector<int2> cls[3];
    for (const int2& t : at) cls[atile_class(h, utts[t.x])].push_back(t);
    size_t o = 0;
    for (int k = 0; k < 3; ++k) { counts[k] = (int)cls[k].size(); std::copy(cls[k].begin(), cls[k].end(), at.begin() + o); o += cls[k].size(); }
}

// first configured speed with this ratio (-1: none)
int speed_with_ratio(const fe_handle* h, int up, int down) {
    for (size_t i = 0; i < h->sp_up.size(); ++i) if (h->sp_up[i] == up && h->sp_down[i] == down) return (int)i;
    return -1;
}

int launch_k0(fe_handle* h, cudaStream_t st, const short* src, const UttDesc* utts, const int2* atiles, const int counts[3],
              short* dst, int use_dst_off) {
    const int* up = (const int*)h->d_sp_up.p; const int* down = (const int*)h->d_sp_down.p;
    const int* toff = (const int*)h->d_sp_tap_off.p; const float* taps = (const float*)h->d_taps.p;
    auto grid = [&](int n) { return (int)std::min<long long>(n, 16LL * h->num_sms); };
    if (counts[0] > 0) {
        K0Taps<10, kK0Taps> W;
        memcpy(W.w, h->sp_taps.data() + h->sp_tap_off[speed_with_ratio(h, 10, 9)], sizeof(W.w));
        k_resample_fast<10, 9, kK0Taps><<<grid(counts[0]), kK0Outputs / (10 * kK0Groups), 0, st>>>(src, utts, atiles, counts[0], W, dst, use_dst_off);
        h->launches++;
    }
    if (counts[1] > 0) {
        K0Taps<10, kK0Taps> W;
        memcpy(W.w, h->sp_taps.data() + h->sp_tap_off[speed_with_ratio(h, 10, 11)], sizeof(W.w));
        k_resample_fast<10, 11, kK0Taps><<<grid(counts[1]), kK0Outputs / (10 * kK0Groups), 0, st>>>(src, utts, atiles + counts[0], counts[1], W, dst, use_dst_off);
        h->launches++;
    }
    if (counts[2] > 0) {
        k_resample<<<grid(counts[2]), 256, 0, st>>>(src, utts, atiles + counts[0] + counts[1], counts[2], up, down, toff, taps,
                                                    h->sp_ntaps, dst, use_dst_off);
        h->launches++;
    }
    FE_CUDA(h, cudaGetLastError());
    return FE_OK;
}

// K2a + K2b: per-utterance statistics, then the tile-parallel normalise / delta / pack pass.
// tiled: `statics` are K1's [D][32] blocks (else a row-major (L, D) matrix per utterance: fe_postprocess).
// flags: bit0 subtract the mean, bit1 divide by the std, bit2 append deltas.
// utt0: index of utts[0] in the batch (a group of a larger batch): its statistics live at stats[utt0 * 2 D], which is
// where the tiles (TileDesc::utt is batch-wide) look them up.
int launch_k2(fe_handle* h, Lane& L, cudaStream_t st, const UttDesc* utts, int n_utts, const TileDesc* tiles, long long n_tiles,
              const float* statics, float* out, int D, int delta_mode, int flags, bool tiled, int utt0) {
    int rc;
    if (utt0 == 0 && (rc = ensure(h, L.d_stats, sizeof(float) * 2 * (size_t)D * (size_t)std::max(n_utts, L.stats_utts)))) return rc;
    float* d_stats_g = (float*)L.d_stats.p + (size_t)utt0 * 2 * D;
    // fused path: statistics + cube in one kernel, statics read from HBM once (K1's tile-major statics, full CMVN,
    // as-shipped deltas, a baked feature width)
    if (tiled && (flags & 3) == 3 && (delta_mode == 0 || !(flags & 4)) && D == 13 && !h->k2_split && n_tiles > 0) {
#define FE_LAUNCH_UTT(DT)                                                                                           \
        do {                                                                                                        \
            using U = UttCube<DT>;                                                                                  \
            FE_CUDA(h, cudaFuncSetAttribute(k_utt_cmvn_cube<DT>, cudaFuncAttributeMaxDynamicSharedMemorySize, U::kSmemBytes)); \
            const int per_sm = std::max(1, std::min(4, (200 * 1024) / U::kSmemBytes));                              \
            const int grid = (int)std::min<long long>(n_utts, (long long)per_sm * h->num_sms);                      \
            k_utt_cmvn_cube<DT><<<grid, U::kWarps * 32, U::kSmemBytes, st>>>(utts, n_utts, statics, d_stats_g, out, flags); \
        } while (0)
        FE_LAUNCH_UTT(13);       // wider rows (fbank-40 / -80) measured slower fused: one CTA per utterance leaves too few bytes in flight
#undef FE_LAUNCH_UTT
        h->launches++;
        FE_CUDA(h, cudaGetLastError());
        return FE_OK;
    }
    if (flags & 3) {
        if (tiled && (flags & 3) == 3) {
            const long long items = (long long)n_utts * D;
            const int grid = (int)std::min<long long>((items + kStatTWarps - 1) / kStatTWarps, 32LL * h->num_sms);
            k_utt_stats_tiled<<<grid, kStatTWarps * 32, 0, st>>>(utts, n_utts, statics, d_stats_g, D);
        } else {
            if (tiled) return fail(h, FE_ERR_INVALID, "internal: partial normalisation on tiled statics");
            const int grid = (int)std::min<long long>(n_utts, 32LL * h->num_sms);
            k_utt_stats<<<grid, kStatThreads, (kStatThreads + std::min(D, kStatThreads)) * sizeof(float), st>>>(
                utts, n_utts, statics, d_stats_g, D, flags);
        }
        h->launches++;
    } else {
        flags |= 8;              // no statistics: mean 0, scale 1 inside the pack kernel
    }
    // fast path: K1's tile-major statics with the as-shipped (per-frame) deltas -- one warp per tile, no CTA barrier
    const bool local = tiled && (delta_mode == 0 || !(flags & 4)) && (D == 13 || D == 39 || D == 40 || D == 80);
    if (n_tiles > 0 && local) {
#define FE_LAUNCH_CUBE(DT)                                                                                          \
        do {                                                                                                        \
            using C = CubeLocal<DT>;                                                                                \
            FE_CUDA(h, cudaFuncSetAttribute(k_cube_local<DT>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::kSmemBytes)); \
            const int per_sm = std::max(1, std::min(8, (220 * 1024) / C::kSmemBytes));                              \
            const int grid = (int)std::min<long long>((n_tiles + C::kWarps - 1) / C::kWarps, (long long)per_sm * h->num_sms); \
            k_cube_local<DT><<<grid, C::kWarps * 32, C::kSmemBytes, st>>>(tiles, (int)n_tiles, statics,             \
                (const float*)L.d_stats.p, out, flags);                                                             \
        } while (0)
        if (D == 13) FE_LAUNCH_CUBE(13);
        else if (D == 39) FE_LAUNCH_CUBE(39);
        else if (D == 40) FE_LAUNCH_CUBE(40);
        else FE_LAUNCH_CUBE(80);
#undef FE_LAUNCH_CUBE
        h->launches++;
    } else if (n_tiles > 0) {
        const size_t smem = k2_smem_floats(D, kTileFrames) * sizeof(float);
        const int grid = (int)std::min<long long>(n_tiles, 12LL * h->num_sms);
#define FE_LAUNCH_PACK(DT, TR)                                                                                      \
        do {                                                                                                        \
            FE_CUDA(h, cudaFuncSetAttribute(k_norm_delta_pack<DT, TR>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
            k_norm_delta_pack<DT, TR><<<grid, kPackThreads, smem, st>>>(tiles, (int)n_tiles, statics,               \
                (const float*)L.d_stats.p, out, D, kTileFrames, delta_mode, flags);                                 \
        } while (0)
        if (tiled) {
            if (D == 13) FE_LAUNCH_PACK(13, true);
            else if (D == 40) FE_LAUNCH_PACK(40, true);
            else if (D == 80) FE_LAUNCH_PACK(80, true);
            else FE_LAUNCH_PACK(0, true);
        } else {
            if (D == 13) FE_LAUNCH_PACK(13, false);
            else if (D == 39) FE_LAUNCH_PACK(39, false);
            else FE_LAUNCH_PACK(0, false);
        }
#undef FE_LAUNCH_PACK
        h->launches++;
    }
    FE_CUDA(h, cudaGetLastError());
    return FE_OK;
}

DevTables dev_tables(const fe_handle* h, bool in_f32) {
    const fe_config& c = h->cfg;
    DevTables dt;
    dt.tw256 = (const float4*)h->tw256.p;
    dt.tw512 = (const float4*)h->tw512.p;
    dt.window = c.window ? (const float2*)h->window.p : nullptr;
    dt.mel_desc = (const int*)h->mel_desc.p;
    dt.mel_w = (const float*)h->mel_w.p + (in_f32 ? 4 * (size_t)h->mel_groups : 0);
    dt.dctf = (const float*)h->dct.p;
    dt.mel_groups = h->mel_groups; dt.p_rows = h->p_rows;
    for (int i = 0; i < kEpiWarps; ++i) { dt.epi_off[i] = h->epi_off[i]; dt.epi_cnt[i] = h->epi_cnt[i]; }
    dt.nf = c.num_filters; dt.D = c.feat_dim; dt.dct_stride = h->dct_stride; dt.nh = h->nh;
    dt.full_spectrum = h->full_spectrum;
    dt.is_mfcc = c.feat_type == FE_FEAT_MFCC;
    dt.fbank_log = c.fbank_log; dt.dc_elim = c.dc_elimination;
    dt.pscale = in_f32 ? 1.0f : (1.0f / 1073741824.0f);      // raw int16 counts: (1/32768)^2
    return dt;
}

}  // namespace

// ---------------------------------------------------------------------------
extern "C" {

int fe_abi_version(void) { return FE_ABI_VERSION; }

int64_t fe_num_frames(int64_t n_samples, int32_t frame_len, int32_t hop) {
    if (hop <= 0 || n_samples < frame_len) return 0;
    return (n_samples - frame_len) / hop;      // floor for non-negative operands
}

int fe_geometry_supported(int32_t frame_len, int32_t hop) {
    return (frame_len == 400 && hop == 160) || (frame_len == 320 && hop == 160) || (frame_len == 480 && hop == 160) ||
           (frame_len == 200 && hop == 80);
}

int64_t fe_resampled_length(int64_t n_samples, int32_t up, int32_t down) {
    if (up <= 0 || down <= 0) return n_samples;
    return (n_samples * up + down - 1) / down;
}

const char* fe_last_error(fe_handle* h) { return h ? h->err.c_str() : g_create_error.c_str(); }

int fe_create(int device, fe_handle** out) {
    if (!out) return fail(nullptr, FE_ERR_INVALID, "out is NULL");
    *out = nullptr;
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0)
        return fail(nullptr, FE_ERR_CUDA, std::string("no CUDA device (there is no CPU fallback): ") +
                                              cudaGetErrorString(e));
    if (device < 0 || device >= n) return fail(nullptr, FE_ERR_INVALID, "device index out of range");
    FE_CUDA(nullptr, cudaSetDevice(device));
    cudaDeviceProp prop;
    FE_CUDA(nullptr, cudaGetDeviceProperties(&prop, device));
    if (prop.major < 10)
        return fail(nullptr, FE_ERR_INVALID, "this library is built for sm_100a (Blackwell) only");
    fe_handle* h = new fe_handle();
    h->device = device;
    h->num_sms = prop.multiProcessorCount;
    FE_CUDA(nullptr, cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
    for (int i = 0; i < 3; ++i) {
        if (i == 0) h->lane[i].stream = h->stream;
        else FE_CUDA(nullptr, cudaStreamCreateWithFlags(&h->lane[i].stream, cudaStreamNonBlocking));
        FE_CUDA(nullptr, cudaEventCreateWithFlags(&h->lane[i].done, cudaEventDisableTiming));
    }
    h->k1t = getenv("FE_K1T") ? atoi(getenv("FE_K1T")) : 2;
    h->k2_split = getenv("FE_K2_SPLIT") != nullptr;
    {   // K1T's twiddles: universal constants, float64 on the host, rounded once (idempotent across handles)
        std::vector<float4> t256p(128), t512p(64);
        std::vector<float2> t512(132, make_float2(0.f, 0.f));
        k1t_build_twiddles(t256p.data(), t512p.data(), t512.data());
        FE_CUDA(nullptr, cudaMemcpyToSymbol(c_tw256p, t256p.data(), sizeof(float4) * 128));
        FE_CUDA(nullptr, cudaMemcpyToSymbol(c_tw512p, t512p.data(), sizeof(float4) * 64));
        FE_CUDA(nullptr, cudaMemcpyToSymbol(c_tw512, t512.data(), sizeof(float2) * 132));
    }
    if (const char* e = getenv("FE_L2_GROUP_MB")) h->l2_group_bytes = atoll(e) << 20;
    if (const char* e = getenv("FE_PIPE_CHUNK_MB")) { long long mb = atoll(e); if (mb > 0) h->pipe_chunk_bytes = mb << 20; }
    *out = h;
    return FE_OK;
}

int fe_destroy(fe_handle* h) {
    if (!h) return FE_OK;
    cudaSetDevice(h->device);
    cudaDeviceSynchronize();
    for (DevBuf* b : {&h->tw256, &h->tw512, &h->window, &h->mel_desc, &h->mel_w, &h->dct,
                      &h->d_sp_up, &h->d_sp_down, &h->d_sp_tap_off, &h->d_taps, &h->d_pad,
                      &h->d_flac_bytes, &h->d_flac_files, &h->d_flac_frames, &h->d_flac_cands, &h->d_flac_ctr, &h->d_flac_pcm})
        release(*b);
    for (Lane& L : h->lane) {
        for (DevBuf* b : {&L.d_utts, &L.d_tile_prefix, &L.d_tiles, &L.d_atile_prefix, &L.d_atiles, &L.d_statics,
                          &L.d_stats, &L.d_pcm, &L.d_out, &L.d_scratch})
            release(*b);
        if (L.h_stage.p) cudaFreeHost(L.h_stage.p);
        if (L.done) cudaEventDestroy(L.done);
        if (L.stream && L.stream != h->stream) cudaStreamDestroy(L.stream);
    }
    for (auto& p : h->prof) { for (auto& e : p.e) if (e) cudaEventDestroy(e); for (auto& e : p.ge) cudaEventDestroy(e); }
    if (h->stream) cudaStreamDestroy(h->stream);
    delete h;
    return FE_OK;
}

int fe_configure(fe_handle* h, const fe_config* c) {
    if (!h || !c) return FE_ERR_INVALID;
    if (c->abi_version != FE_ABI_VERSION) return fail(h, FE_ERR_INVALID, "fe_config.abi_version mismatch");
    if (c->nfft != kNfft) return fail(h, FE_ERR_INVALID, "only nfft = 512 is supported");
    if (!fe_geometry_supported(c->frame_len, c->hop))
        return fail(h, FE_ERR_INVALID, "unsupported frame geometry: kernels are built for 400/160 (25 / 10 ms at 16 kHz), 320/160, 480/160 and "
                                       "200/80 samples (25 / 10 ms at 8 kHz)");
    if (c->num_filters < 1 || c->num_filters > kMaxFilters) return fail(h, FE_ERR_INVALID, "num_filters out of range [1,128]");
    if (c->feat_dim < 1 || c->feat_dim > kMaxFilters) return fail(h, FE_ERR_INVALID, "feat_dim out of range [1,128]");
    if (c->feat_type == FE_FEAT_MFCC && c->feat_dim > c->num_filters)
        return fail(h, FE_ERR_INVALID, "mfcc: feat_dim must be <= num_filters");
    if (c->feat_type == FE_FEAT_FBANK && c->feat_dim != c->num_filters)
        return fail(h, FE_ERR_INVALID, "fbank: feat_dim must equal num_filters");
    if (c->feat_type != FE_FEAT_MFCC && c->feat_type != FE_FEAT_FBANK) return fail(h, FE_ERR_INVALID, "bad feat_type");
    if (!c->fb_row_start || !c->fb_first_bin || !c->fb_weights || !c->tw256 || !c->tw512)
        return fail(h, FE_ERR_INVALID, "missing table pointer");
    if (c->feat_type == FE_FEAT_MFCC && !c->dct) return fail(h, FE_ERR_INVALID, "mfcc needs the DCT table");
    if (c->fb_row_start[c->num_filters] != c->fb_nnz) return fail(h, FE_ERR_INVALID, "filterbank CSR inconsistent");
    FE_CUDA(h, cudaSetDevice(h->device));
    FE_CUDA(h, cudaStreamSynchronize(h->stream));

    int max_bin = 0;
    for (int m = 0; m < c->num_filters; ++m) {
        int w = c->fb_row_start[m + 1] - c->fb_row_start[m];
        if (w < 0 || c->fb_first_bin[m] < 0 || c->fb_first_bin[m] + w > kBins)
            return fail(h, FE_ERR_INVALID, "filterbank row outside the 257 FFT bins");
        if (w > 0) max_bin = std::max(max_bin, c->fb_first_bin[m] + w - 1);
    }
    h->full_spectrum = max_bin > 128;

    int rc;
    HostTables ht;
    build_host_tables(*c, ht);
    if (!ht.ok) return fail(h, FE_ERR_INVALID, "filterbank does not fit the mel plan (run too long)");
    h->mel_groups = ht.mel_groups; h->p_rows = ht.p_rows; h->nh = ht.nh; h->dct_stride = ht.dct_stride;
    memcpy(h->epi_off, ht.epi_off, sizeof(h->epi_off)); memcpy(h->epi_cnt, ht.epi_cnt, sizeof(h->epi_cnt));
    h->epi_plan = ht.epi_w_n <= kEpiWCap ? ht.epi_plan : 0; h->epi_w_n = ht.epi_w_n; h->epi_w = ht.epi_w;
    if (getenv("FE_K1_GENERIC")) h->epi_plan = 0;         // force the run-time mel plan (tests, comparisons)
    if (!(c->frame_len == 400 && c->hop == 160)) h->epi_plan = 0;     // the specialised epilogues are built for the main geometry
    if ((rc = upload(h, h->tw256, ht.tw256.data(), ht.tw256.size() * sizeof(float)))) return rc;
    if ((rc = upload(h, h->tw512, ht.tw512.data(), ht.tw512.size() * sizeof(float)))) return rc;
    if ((rc = upload(h, h->mel_desc, ht.mel_desc.data(), ht.mel_desc.size() * sizeof(int)))) return rc;
    if ((rc = upload(h, h->mel_w, ht.mel_w.data(), ht.mel_w.size() * sizeof(float)))) return rc;
    if (c->feat_type == FE_FEAT_MFCC && (rc = upload(h, h->dct, ht.dctf.data(), ht.dctf.size() * sizeof(float)))) return rc;
    if (c->window && (rc = upload(h, h->window, ht.window.data(), ht.window.size() * sizeof(float)))) return rc;
    h->sp_up.clear(); h->sp_down.clear(); h->sp_tap_off.clear(); h->sp_taps.clear(); h->sp_ntaps = 0;
    if (c->n_speeds > 0) {
        if (!c->speed_up || !c->speed_down || !c->speed_taps) return fail(h, FE_ERR_INVALID, "missing resampler tables");
        if (c->speed_ntaps < 2 || c->speed_ntaps > 256 || (c->speed_ntaps & 1))
            return fail(h, FE_ERR_INVALID, "speed_ntaps must be even, 2..256");
        h->sp_ntaps = c->speed_ntaps;
        int off = 0;
        for (int i = 0; i < c->n_speeds; ++i) {
            if (c->speed_up[i] < 1 || c->speed_down[i] < 1 || c->speed_up[i] > 4096)
                return fail(h, FE_ERR_INVALID, "bad speed ratio");
            h->sp_up.push_back(c->speed_up[i]); h->sp_down.push_back(c->speed_down[i]); h->sp_tap_off.push_back(off);
            off += c->speed_up[i] * c->speed_ntaps;
        }
        h->sp_taps.assign(c->speed_taps, c->speed_taps + off);
        if ((rc = upload(h, h->d_sp_up, h->sp_up.data(), h->sp_up.size() * sizeof(int)))) return rc;
        if ((rc = upload(h, h->d_sp_down, h->sp_down.data(), h->sp_down.size() * sizeof(int)))) return rc;
        if ((rc = upload(h, h->d_sp_tap_off, h->sp_tap_off.data(), h->sp_tap_off.size() * sizeof(int)))) return rc;
        if ((rc = upload(h, h->d_taps, c->speed_taps, (size_t)off * sizeof(float)))) return rc;
    }

    h->cfg = *c;       // pointer members are only used as "present" flags from here on
    for (int f32 = 0; f32 < 2; ++f32) {
        const bool used = (f32 != 0) == (c->preemph != 0.f || c->pcm_dtype == FE_PCM_FLOAT32);
        if (!used) continue;
        K1Smem L = k1_smem_layout(c->num_filters, h->mel_groups, h->p_rows, c->feat_dim, h->dct_stride, c->window != nullptr,
                                  c->frame_len, c->hop, c->feat_type == FE_FEAT_MFCC, f32, h->epi_plan);
        if (h->epi_plan && L.total > kK1SmemMax) {           // the 4-slot layout does not fit: generic epilogue
            h->epi_plan = 0;
            L = k1_smem_layout(c->num_filters, h->mel_groups, h->p_rows, c->feat_dim, h->dct_stride, c->window != nullptr,
                               c->frame_len, c->hop, c->feat_type == FE_FEAT_MFCC, f32, 0);
        }
        h->k1_smem[f32] = L.total;
        if (L.total > kK1SmemMax) return fail(h, FE_ERR_INVALID, "configuration needs too much shared memory");
    }
    h->configured = true;
    return FE_OK;
}

int fe_plan(fe_handle* h, const int64_t* pcm_lengths, int32_t n_utts, const int32_t* speed_idx,
            int64_t* out_offsets, int32_t* n_frames) {
    if (!h) return FE_ERR_INVALID;
    if (!h->configured) return fail(h, FE_ERR_STATE, "fe_configure first");
    if (n_utts < 0 || (n_utts > 0 && !pcm_lengths)) return fail(h, FE_ERR_INVALID, "bad arguments");
    Plan pl;
    return make_plan(h, h->lane[0], nullptr, pcm_lengths, n_utts, speed_idx, nullptr, out_offsets, n_frames, pl, false);
}

// One batch on one lane: plan, descriptor upload, (H2D), tile build, K0, K1, K2, (D2H).  Asynchronous on
// `st` unless `final_sync`.  `pcm` / `out` may be host or device pointers.
static int run_core(fe_handle* h, Lane& L, cudaStream_t st, const void* pcm, const int64_t* pcm_offsets,
                    const int64_t* pcm_lengths, int32_t n_utts, const int32_t* speed_idx, const float* gain,
                    float* out, int64_t out_capacity, int64_t* out_offsets, int32_t* n_frames,
                    bool allow_prof, bool final_sync) {
    const fe_config& c = h->cfg;

    Plan pl;
    int rc = make_plan(h, L, pcm_offsets, pcm_lengths, n_utts, speed_idx, gain, out_offsets, n_frames, pl, true);
    if (rc) return rc;
    if (pl.total_out > out_capacity) return fail(h, FE_ERR_CAPACITY, "out buffer too small for this batch");
    if (pl.total_tiles > 0x7fffffffLL || pl.total_atiles > 0x7fffffffLL) return fail(h, FE_ERR_INVALID, "batch too large");

    const size_t esz = c.pcm_dtype == FE_PCM_INT16 ? 2 : 4;
    const bool pcm_on_dev = is_device_ptr(pcm), out_on_dev = is_device_ptr(out);

    // descriptors -> pinned staging -> device (async on st)
    const size_t b_utts = sizeof(UttDesc) * (size_t)n_utts, b_pref = sizeof(long long) * (size_t)(n_utts + 1);
    if ((rc = ensure_pinned(h, L.h_stage, b_utts + 2 * b_pref))) return rc;
    if ((rc = ensure(h, L.d_utts, b_utts))) return rc;
    if ((rc = ensure(h, L.d_tile_prefix, b_pref))) return rc;
    if ((rc = ensure(h, L.d_atile_prefix, b_pref))) return rc;
    if ((rc = ensure(h, L.d_tiles, sizeof(TileDesc) * (size_t)std::max<long long>(pl.total_tiles, 1)))) return rc;
    if ((rc = ensure(h, L.d_statics, sizeof(float) * (size_t)std::max<long long>(pl.total_tiles * kTileFrames * c.feat_dim, 1)))) return rc;
    const bool preemph = c.preemph != 0.f;
    const bool k1_f32 = preemph || c.pcm_dtype == FE_PCM_FLOAT32;      // what K1's stage A reads
    if (pl.any_scratch) {
        if ((rc = ensure(h, L.d_scratch, (preemph ? 4 : 2) * (size_t)pl.total_scratch + 64))) return rc;
        if ((rc = ensure(h, L.d_atiles, sizeof(int2) * (size_t)pl.total_atiles))) return rc;
    }
    const void* d_pcm = pcm;
    float* d_out = out;
    if (!pcm_on_dev) { if ((rc = ensure(h, L.d_pcm, esz * (size_t)pl.pcm_span))) return rc; d_pcm = L.d_pcm.p; }
    if (!out_on_dev) { if ((rc = ensure(h, L.d_out, sizeof(float) * (size_t)std::max<long long>(pl.total_out, 1)))) return rc; d_out = (float*)L.d_out.p; }

    // K0 tile table: per-utterance start index, utterances grouped by resampler class (atile_prefix is reused
    // for it: entry i = first tile of utterance i in the grouped table, -1 = none)
    int k0_counts[3] = {0, 0, 0};
    if (pl.any_scratch) {
        long long cnt[3] = {0, 0, 0};
        for (int i = 0; i < n_utts; ++i) {
            const long long nt = L.atile_prefix[i + 1] - L.atile_prefix[i];
            if (nt > 0) cnt[preemph ? 2 : atile_class(h, L.utts[i])] += nt;
        }
        long long base[3] = {0, cnt[0], cnt[0] + cnt[1]};
        std::vector<long long>& as = L.atile_prefix;
        long long prev = as[0];
        for (int i = 0; i < n_utts; ++i) {
            const long long nxt = as[i + 1], nt = nxt - prev;
            prev = nxt;
            if (nt > 0) { long long& b = base[preemph ? 2 : atile_class(h, L.utts[i])]; as[i] = b; b += nt; }
            else as[i] = -1;
        }
        for (int k = 0; k < 3; ++k) k0_counts[k] = (int)cnt[k];
    }
    // the previous run on this lane may still be reading the staging area
    if (L.done_valid) FE_CUDA(h, cudaEventSynchronize(L.done));
    unsigned char* hs = (unsigned char*)L.h_stage.p;
    memcpy(hs, L.utts.data(), b_utts);
    memcpy(hs + b_utts, L.tile_prefix.data(), b_pref);
    memcpy(hs + b_utts + b_pref, L.atile_prefix.data(), b_pref);
    FE_CUDA(h, cudaMemcpyAsync(L.d_utts.p, hs, b_utts, cudaMemcpyHostToDevice, st));
    FE_CUDA(h, cudaMemcpyAsync(L.d_tile_prefix.p, hs + b_utts, b_pref, cudaMemcpyHostToDevice, st));
    if (pl.any_scratch)
        FE_CUDA(h, cudaMemcpyAsync(L.d_atile_prefix.p, hs + b_utts + b_pref, b_pref, cudaMemcpyHostToDevice, st));
    if (((uintptr_t)d_pcm & 15) != 0) return fail(h, FE_ERR_INVALID, "pcm buffer must be 16-byte aligned");
    if (!pcm_on_dev) FE_CUDA(h, cudaMemcpyAsync(L.d_pcm.p, pcm, esz * (size_t)pl.pcm_span, cudaMemcpyHostToDevice, st));

    const bool prof = allow_prof && h->profiling != 0 && h->prof_used < 4096;
    fe_handle::ProfSet* ps = nullptr;
    if (prof) {
        if (h->prof_used == h->prof.size()) {
            fe_handle::ProfSet n{};
            for (auto& e : n.e) FE_CUDA(h, cudaEventCreate(&e));
            h->prof.push_back(n);
        }
        ps = &h->prof[h->prof_used++];
        ps->k0 = ps->k2 = false;
        FE_CUDA(h, cudaEventRecord(ps->e[0], st));
    }
    const int tb = 256, gb = (n_utts + tb - 1) / tb;
    if (pl.total_tiles > 0) {
        k_build_tiles<<<gb, tb, 0, st>>>((const UttDesc*)L.d_utts.p, (const long long*)L.d_tile_prefix.p, n_utts,
                                         c.hop, c.feat_dim, (TileDesc*)L.d_tiles.p);
        h->launches++;
    }
    if (pl.any_scratch) {
        // K0's tile table is expanded on the device from the per-utterance start indices (grouped by class)
        k_build_atiles<<<gb, tb, 0, st>>>((const UttDesc*)L.d_utts.p, (const long long*)L.d_atile_prefix.p, n_utts,
                                          kK0Outputs, (int2*)L.d_atiles.p);
        h->launches++;
        if (prof) FE_CUDA(h, cudaEventRecord(ps->e[1], st));
        int grid = (int)std::min<long long>(pl.total_atiles, 16LL * h->num_sms);
        if (preemph) {
            if (c.pcm_dtype == FE_PCM_INT16)
                k_preemph<0><<<grid, 256, 0, st>>>(d_pcm, (const UttDesc*)L.d_utts.p, (const int2*)L.d_atiles.p,
                                                   (int)pl.total_atiles, c.preemph, (float*)L.d_scratch.p);
            else
                k_preemph<1><<<grid, 256, 0, st>>>(d_pcm, (const UttDesc*)L.d_utts.p, (const int2*)L.d_atiles.p,
                                                   (int)pl.total_atiles, c.preemph, (float*)L.d_scratch.p);
        } else {
            if ((rc = launch_k0(h, st, (const short*)d_pcm, (const UttDesc*)L.d_utts.p, (const int2*)L.d_atiles.p,
                                k0_counts, (short*)L.d_scratch.p, 0))) return rc;
        }
        if (prof) ps->k0 = true;
    }
    if (prof) FE_CUDA(h, cudaEventRecord(ps->e[2], st));
    DevTables dt = dev_tables(h, k1_f32);
    const int k2_flags = c.cmvn ? 7 : 0;
    L.stats_utts = n_utts;
    if ((rc = ensure(h, L.d_stats, sizeof(float) * 2 * (size_t)c.feat_dim * (size_t)n_utts))) return rc;
    // K1 -> K2a -> K2b run over groups of utterances whose statics fit the L2 (126 MB on the B200): K2 reads them
    // three times (mean pass, variance pass, cube pass) and finds them where K1 has just left them instead of in
    // HBM.  Measured on the fbank-80 pass (statics 320 B / frame): K2 3.56 -> see profiles/r02_l2_groups.json.
    // Profiling records one event pair per group.
    const long long group_bytes = pl.total_frames == 0 ? 0 : h->l2_group_bytes;
    if (prof) ps->groups = 0;
    if (group_bytes <= 0) {
        if ((rc = launch_k1(h, st, d_pcm, L.d_scratch.p, k1_f32, (const TileDesc*)L.d_tiles.p,
                            (int)pl.total_tiles, dt, (float*)L.d_statics.p))) return rc;
        if (prof) FE_CUDA(h, cudaEventRecord(ps->e[3], st));
        if (pl.total_frames > 0) {
            // cmvn: statistics + normalise + deltas + cube; no cmvn: the same pack kernel only re-lays the blocks out as (L, D)
            if ((rc = launch_k2(h, L, st, (const UttDesc*)L.d_utts.p, n_utts, (const TileDesc*)L.d_tiles.p, pl.total_tiles,
                                (const float*)L.d_statics.p, d_out, c.feat_dim, c.delta_mode, k2_flags, true, 0))) return rc;
            if (prof) ps->k2 = true;
        }
    } else {
        const long long tiles_per_group = std::max<long long>(1, group_bytes / ((long long)kTileFrames * c.feat_dim * 4));
        int u0 = 0;
        while (u0 < n_utts) {
            int u1 = u0 + 1;
            while (u1 < n_utts && L.tile_prefix[(size_t)u1 + 1] - L.tile_prefix[(size_t)u0] <= tiles_per_group) ++u1;
            const long long t0 = L.tile_prefix[(size_t)u0], nt = L.tile_prefix[(size_t)u1] - t0;
            if (nt > 0) {
                if (prof && ps->ge.size() < (size_t)(2 * ps->groups + 2)) {
                    cudaEvent_t a, b;
                    FE_CUDA(h, cudaEventCreate(&a)); FE_CUDA(h, cudaEventCreate(&b));
                    ps->ge.push_back(a); ps->ge.push_back(b);
                }
                if ((rc = launch_k1(h, st, d_pcm, L.d_scratch.p, k1_f32, (const TileDesc*)L.d_tiles.p + t0, (int)nt, dt,
                                    (float*)L.d_statics.p))) return rc;
                if (prof) FE_CUDA(h, cudaEventRecord(ps->ge[(size_t)2 * ps->groups], st));
                if ((rc = launch_k2(h, L, st, (const UttDesc*)L.d_utts.p + u0, u1 - u0, (const TileDesc*)L.d_tiles.p + t0, nt,
                                    (const float*)L.d_statics.p, d_out, c.feat_dim, c.delta_mode, k2_flags, true, u0))) return rc;
                if (prof) { FE_CUDA(h, cudaEventRecord(ps->ge[(size_t)2 * ps->groups + 1], st)); ps->groups++; ps->k2 = true; }
            }
            u0 = u1;
        }
    }
    FE_CUDA(h, cudaGetLastError());
    if (prof) FE_CUDA(h, cudaEventRecord(ps->e[4], st));
    if (!out_on_dev)
        FE_CUDA(h, cudaMemcpyAsync(out, d_out, sizeof(float) * (size_t)pl.total_out, cudaMemcpyDeviceToHost, st));
    FE_CUDA(h, cudaEventRecord(L.done, st));
    L.done_valid = true;
    if (final_sync) FE_CUDA(h, cudaStreamSynchronize(st));
    return FE_OK;
}



int fe_run(fe_handle* h, const void* pcm, const int64_t* pcm_offsets, const int64_t* pcm_lengths,
           int32_t n_utts, const int32_t* speed_idx, const float* gain,
           float* out, int64_t out_capacity, int64_t* out_offsets, int32_t* n_frames, void* stream) {
    if (!h) return FE_ERR_INVALID;
    if (!h->configured) return fail(h, FE_ERR_STATE, "fe_configure first");
    if (n_utts < 0) return fail(h, FE_ERR_INVALID, "n_utts < 0");
    if (n_utts == 0) { if (out_offsets) out_offsets[0] = 0; return FE_OK; }
    if (!pcm || !pcm_offsets || !pcm_lengths || !out) return fail(h, FE_ERR_INVALID, "NULL buffer");
    FE_CUDA(h, cudaSetDevice(h->device));
    const fe_config& c = h->cfg;
    const bool pcm_on_dev = is_device_ptr(pcm), out_on_dev = is_device_ptr(out);
    const long long esz = c.pcm_dtype == FE_PCM_INT16 ? 2 : 4;
    long long span = 0;
    bool monotonic = true;
    for (int i = 0; i < n_utts; ++i) {
        if (i && pcm_offsets[i] < pcm_offsets[i - 1] + pcm_lengths[i - 1]) monotonic = false;
        span = std::max<long long>(span, pcm_offsets[i] + pcm_lengths[i]);
    }
    // ---- host in, host out, large batch: chunked pipeline over lanes 1 and 2 so that the H2D copy of
    // chunk i+1, the kernels of chunk i and the D2H copy of chunk i-1 overlap (PCIe is full duplex) ----
    if (!pcm_on_dev && !out_on_dev && monotonic && span * esz > 2 * h->pipe_chunk_bytes && out_offsets && n_frames) {
        Plan pl;
        int rc = make_plan(h, h->lane[0], nullptr, pcm_lengths, n_utts, speed_idx, gain, out_offsets, n_frames, pl, false);
        if (rc) return rc;
        if (pl.total_out > out_capacity) return fail(h, FE_ERR_CAPACITY, "out buffer too small for this batch");
        std::vector<int64_t> loc_off, loc_out((size_t)n_utts + 1);
        std::vector<int32_t> loc_nf((size_t)n_utts);
        int first = 0, k = 0;
        while (first < n_utts) {
            int last = first;                                   // chunk = utterances [first, last]
            const long long base = pcm_offsets[first];
            while (last + 1 < n_utts && (pcm_offsets[last + 1] + pcm_lengths[last + 1] - base) * esz <= h->pipe_chunk_bytes) ++last;
            const int n = last - first + 1;
            loc_off.assign((size_t)n, 0);
            for (int i = 0; i < n; ++i) loc_off[(size_t)i] = pcm_offsets[first + i] - base;
            Lane& L = h->lane[1 + (k & 1)];
            rc = run_core(h, L, L.stream, (const char*)pcm + base * esz, loc_off.data(), pcm_lengths + first, n,
                          speed_idx ? speed_idx + first : nullptr, gain ? gain + first : nullptr,
                          out + out_offsets[first], out_offsets[last + 1] - out_offsets[first],
                          loc_out.data(), loc_nf.data(), false, false);
            if (rc) return rc;
            first = last + 1; ++k;
        }
        FE_CUDA(h, cudaStreamSynchronize(h->lane[1].stream));
        FE_CUDA(h, cudaStreamSynchronize(h->lane[2].stream));
        return FE_OK;
    }
    cudaStream_t st = stream ? (cudaStream_t)stream : h->stream;
    return run_core(h, h->lane[0], st, pcm, pcm_offsets, pcm_lengths, n_utts, speed_idx, gain, out, out_capacity,
                    out_offsets, n_frames, true, !out_on_dev);
}

int fe_perturb(fe_handle* h, const int16_t* pcm, const int64_t* pcm_offsets, const int64_t* pcm_lengths,
               int32_t n_utts, const int32_t* speed_idx, const float* gain, int16_t* dst,
               int64_t dst_capacity, int64_t* dst_offsets, int64_t* dst_lengths, void* stream) {
    if (!h) return FE_ERR_INVALID;
    if (!h->configured) return fail(h, FE_ERR_STATE, "fe_configure first");
    if (n_utts <= 0) { if (dst_offsets) dst_offsets[0] = 0; return n_utts == 0 ? FE_OK : FE_ERR_INVALID; }
    if (!pcm || !pcm_offsets || !pcm_lengths || !dst || !dst_offsets || !dst_lengths)
        return fail(h, FE_ERR_INVALID, "NULL buffer");
    FE_CUDA(h, cudaSetDevice(h->device));
    cudaStream_t st = stream ? (cudaStream_t)stream : h->stream;
    std::vector<UttDesc> ut((size_t)n_utts);
    std::vector<int2> at;
    long long off = 0, span = 0;
    for (int i = 0; i < n_utts; ++i) {
        int sidx = speed_idx ? speed_idx[i] : -1;
        if (sidx >= (int)h->sp_up.size()) return fail(h, FE_ERR_INVALID, "speed_idx not configured");
        if (sidx >= 0 && h->sp_up[sidx] == h->sp_down[sidx]) sidx = -1;
        long long len = pcm_lengths[i];
        long long n_eff = sidx >= 0 ? fe_resampled_length(len, h->sp_up[sidx], h->sp_down[sidx]) : len;
        UttDesc& u = ut[(size_t)i];
        memset(&u, 0, sizeof(u));
        u.src_off = pcm_offsets[i]; u.n_src = (int)len; u.n_samples = (int)n_eff; u.speed_idx = sidx;
        u.gain = gain ? gain[i] : 1.f; u.out_off = off;
        dst_offsets[i] = off; dst_lengths[i] = n_eff;
        for (long long j = 0; j < n_eff; j += kK0Outputs) at.push_back(make_int2(i, (int)j));
        off += round_up(n_eff, 8);
        span = std::max(span, pcm_offsets[i] + len);
    }
    dst_offsets[n_utts] = off;
    if (off > dst_capacity) return fail(h, FE_ERR_CAPACITY, "dst buffer too small");
    if (at.empty()) return FE_OK;
    const bool src_dev = is_device_ptr(pcm), dst_dev = is_device_ptr(dst);
    int rc;
    if ((rc = ensure(h, h->lane[0].d_utts, sizeof(UttDesc) * ut.size()))) return rc;
    if ((rc = ensure(h, h->lane[0].d_atiles, sizeof(int2) * at.size()))) return rc;
    const short* d_src = pcm; short* d_dst = dst;
    if (!src_dev) { if ((rc = ensure(h, h->lane[0].d_pcm, 2 * (size_t)span))) return rc; d_src = (const short*)h->lane[0].d_pcm.p; }
    if (!dst_dev) { if ((rc = ensure(h, h->lane[0].d_scratch, 2 * (size_t)off))) return rc; d_dst = (short*)h->lane[0].d_scratch.p; }
    FE_CUDA(h, cudaMemcpyAsync(h->lane[0].d_utts.p, ut.data(), sizeof(UttDesc) * ut.size(), cudaMemcpyHostToDevice, st));
    int k0_counts[3];
    class_atiles(h, ut.data(), at, k0_counts);
    FE_CUDA(h, cudaMemcpyAsync(h->lane[0].d_atiles.p, at.data(), sizeof(int2) * at.size(), cudaMemcpyHostToDevice, st));
    if (!src_dev) FE_CUDA(h, cudaMemcpyAsync(h->lane[0].d_pcm.p, pcm, 2 * (size_t)span, cudaMemcpyHostToDevice, st));
    FE_CUDA(h, cudaStreamSynchronize(st));
    int grid = (int)std::min<long long>((long long)at.size(), 16LL * h->num_sms);
    (void)grid;
    if ((rc = launch_k0(h, st, d_src, (const UttDesc*)h->lane[0].d_utts.p, (const int2*)h->lane[0].d_atiles.p, k0_counts, d_dst, 1))) return rc;
    FE_CUDA(h, cudaGetLastError());
    if (!dst_dev) {
        FE_CUDA(h, cudaMemcpyAsync(dst, d_dst, 2 * (size_t)off, cudaMemcpyDeviceToHost, st));
        FE_CUDA(h, cudaStreamSynchronize(st));
    }
    return FE_OK;
}

int fe_postprocess(fe_handle* h, const float* feats, const int64_t* feat_offsets, const int32_t* n_frames,
                   int32_t n_utts, int32_t D, int32_t mode, int32_t delta_mode, float* out, int64_t out_capacity,
                   int64_t* out_offsets, void* stream) {
    if (!h) return FE_ERR_INVALID;
    if (n_utts < 0 || D < 1 || D > 1024) return fail(h, FE_ERR_INVALID, "bad n_utts / D");
    if (n_utts == 0) { if (out_offsets) out_offsets[0] = 0; return FE_OK; }
    if (!feats || !feat_offsets || !n_frames || !out || !out_offsets) return fail(h, FE_ERR_INVALID, "NULL buffer");
    FE_CUDA(h, cudaSetDevice(h->device));
    cudaStream_t st = stream ? (cudaStream_t)stream : h->stream;
    const int W = (mode & 4) ? 3 : 1;
    std::vector<UttDesc> ut((size_t)n_utts);
    long long off = 0, span = 0;
    for (int i = 0; i < n_utts; ++i) {
        UttDesc& u = ut[(size_t)i];
        memset(&u, 0, sizeof(u));
        if (feat_offsets[i] < 0 || n_frames[i] < 0) return fail(h, FE_ERR_INVALID, "negative offset / length");
        u.stat_off = feat_offsets[i]; u.out_off = off; u.n_frames = n_frames[i];
        out_offsets[i] = off;
        off += round_up((long long)n_frames[i] * D * W, 4);
        span = std::max<long long>(span, feat_offsets[i] + (long long)n_frames[i] * D);
    }
    out_offsets[n_utts] = off;
    if (off > out_capacity) return fail(h, FE_ERR_CAPACITY, "out buffer too small");
    const bool in_dev = is_device_ptr(feats), out_dev = is_device_ptr(out);
    int rc;
    std::vector<long long> tpref((size_t)n_utts + 1, 0);
    for (int i = 0; i < n_utts; ++i) tpref[(size_t)i + 1] = tpref[(size_t)i] + statics_tiles(n_frames[i]);
    const long long n_tiles = tpref[(size_t)n_utts];
    if (n_tiles > 0x7fffffffLL) return fail(h, FE_ERR_INVALID, "batch too large");
    if ((rc = ensure(h, h->lane[0].d_utts, sizeof(UttDesc) * ut.size()))) return rc;
    if ((rc = ensure(h, h->lane[0].d_tile_prefix, sizeof(long long) * tpref.size()))) return rc;
    if ((rc = ensure(h, h->lane[0].d_tiles, sizeof(TileDesc) * (size_t)std::max<long long>(n_tiles, 1)))) return rc;
    const float* d_in = feats; float* d_out = out;
    if (!in_dev) { if ((rc = ensure(h, h->lane[0].d_statics, sizeof(float) * (size_t)std::max<long long>(span, 1)))) return rc; d_in = (const float*)h->lane[0].d_statics.p; }
    if (!out_dev) { if ((rc = ensure(h, h->lane[0].d_out, sizeof(float) * (size_t)std::max<long long>(off, 1)))) return rc; d_out = (float*)h->lane[0].d_out.p; }
    FE_CUDA(h, cudaMemcpyAsync(h->lane[0].d_utts.p, ut.data(), sizeof(UttDesc) * ut.size(), cudaMemcpyHostToDevice, st));
    FE_CUDA(h, cudaMemcpyAsync(h->lane[0].d_tile_prefix.p, tpref.data(), sizeof(long long) * tpref.size(), cudaMemcpyHostToDevice, st));
    if (!in_dev) FE_CUDA(h, cudaMemcpyAsync(h->lane[0].d_statics.p, feats, sizeof(float) * (size_t)span, cudaMemcpyHostToDevice, st));
    FE_CUDA(h, cudaStreamSynchronize(st));
    if (n_tiles > 0) {
        k_build_tiles<<<(n_utts + 255) / 256, 256, 0, st>>>((const UttDesc*)h->lane[0].d_utts.p, (const long long*)h->lane[0].d_tile_prefix.p,
                                                            n_utts, 160, D, (TileDesc*)h->lane[0].d_tiles.p);
        h->launches++;
    }
    if ((rc = launch_k2(h, h->lane[0], st, (const UttDesc*)h->lane[0].d_utts.p, n_utts, (const TileDesc*)h->lane[0].d_tiles.p, n_tiles, d_in, d_out,
                        D, delta_mode, mode, false, 0))) return rc;
    if (!out_dev) {
        FE_CUDA(h, cudaMemcpyAsync(out, d_out, sizeof(float) * (size_t)off, cudaMemcpyDeviceToHost, st));
    }
    FE_CUDA(h, cudaStreamSynchronize(st));
    return FE_OK;
}

int fe_pad_batches(fe_handle* h, const float* feats, const int64_t* src_offsets, const int32_t* valid_floats,
                   const int64_t* dst_offsets, const int32_t* slot_floats, int32_t n_slots, float* dst,
                   int64_t dst_capacity, void* stream) {
    if (!h) return FE_ERR_INVALID;
    if (n_slots < 0) return fail(h, FE_ERR_INVALID, "bad n_slots");
    if (n_slots == 0) return FE_OK;
    if (!feats || !src_offsets || !valid_floats || !dst_offsets || !slot_floats || !dst) return fail(h, FE_ERR_INVALID, "NULL buffer");
    FE_CUDA(h, cudaSetDevice(h->device));
    cudaStream_t st = stream ? (cudaStream_t)stream : h->stream;
    std::vector<PadSlot> sl((size_t)n_slots);
    long long span_src = 0, span_dst = 0, n_chunks = 0;
    for (int i = 0; i < n_slots; ++i) {
        if (src_offsets[i] < 0 || dst_offsets[i] < 0 || valid_floats[i] < 0 || slot_floats[i] < 0)
            return fail(h, FE_ERR_INVALID, "negative offset / length");
        if (valid_floats[i] > slot_floats[i]) return fail(h, FE_ERR_INVALID, "utterance longer than its slot (bucket boundary)");
        sl[(size_t)i] = PadSlot{src_offsets[i], dst_offsets[i], valid_floats[i], slot_floats[i], (int)n_chunks, 0};
        n_chunks += std::max(1, (slot_floats[i] / 4 + kPadChunk / 4 - 1) / (kPadChunk / 4));      // >= 1: head / tail scalars
        if (n_chunks > 0x7fffffffLL) return fail(h, FE_ERR_INVALID, "batch too large");
        span_src = std::max<long long>(span_src, src_offsets[i] + valid_floats[i]);
        span_dst = std::max<long long>(span_dst, dst_offsets[i] + slot_floats[i]);
    }
    if (span_dst > dst_capacity) return fail(h, FE_ERR_CAPACITY, "batch buffer too small");
    const bool in_dev = is_device_ptr(feats), out_dev = is_device_ptr(dst);
    int rc;
    Lane& L = h->lane[0];
    if ((rc = ensure(h, h->d_pad, sizeof(PadSlot) * sl.size()))) return rc;
    const float* d_in = feats; float* d_out = dst;
    if (!in_dev) { if ((rc = ensure(h, L.d_statics, sizeof(float) * (size_t)std::max<long long>(span_src, 1)))) return rc; d_in = (const float*)L.d_statics.p; }
    if (!out_dev) { if ((rc = ensure(h, L.d_out, sizeof(float) * (size_t)std::max<long long>(span_dst, 1)))) return rc; d_out = (float*)L.d_out.p; }
    FE_CUDA(h, cudaMemcpyAsync(h->d_pad.p, sl.data(), sizeof(PadSlot) * sl.size(), cudaMemcpyHostToDevice, st));
    if (!in_dev) FE_CUDA(h, cudaMemcpyAsync(L.d_statics.p, feats, sizeof(float) * (size_t)span_src, cudaMemcpyHostToDevice, st));
    FE_CUDA(h, cudaStreamSynchronize(st));                   // sl / feats may be freed by the caller after return
    const int grid = (int)std::min<long long>(n_chunks, 8LL * h->num_sms);
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    if (h->profiling) { FE_CUDA(h, cudaEventCreate(&e0)); FE_CUDA(h, cudaEventCreate(&e1)); FE_CUDA(h, cudaEventRecord(e0, st)); }
    k_pad_slots<<<grid, 256, 0, st>>>(d_in, (const PadSlot*)h->d_pad.p, n_slots, (int)n_chunks, d_out);
    h->launches++;
    FE_CUDA(h, cudaGetLastError());
    if (h->profiling) {
        FE_CUDA(h, cudaEventRecord(e1, st));
        FE_CUDA(h, cudaEventSynchronize(e1));
        FE_CUDA(h, cudaEventElapsedTime(&h->pad_ms, e0, e1));
        cudaEventDestroy(e0); cudaEventDestroy(e1);
    }
    if (!out_dev) FE_CUDA(h, cudaMemcpyAsync(dst, d_out, sizeof(float) * (size_t)span_dst, cudaMemcpyDeviceToHost, st));
    if (!out_dev || !in_dev) FE_CUDA(h, cudaStreamSynchronize(st));
    return FE_OK;
}

int fe_decode_flac(fe_handle* h, const uint8_t* bytes, int64_t total_bytes, const fe_flac_file* files, int32_t n_files,
                   int16_t* pcm, int64_t pcm_capacity, int32_t* status, void* stream) {
    if (!h) return FE_ERR_INVALID;
    if (n_files < 0 || total_bytes < 0) return fail(h, FE_ERR_INVALID, "bad n_files / total_bytes");
    if (n_files == 0) return FE_OK;
    if (!bytes || !files || !pcm || !status) return fail(h, FE_ERR_INVALID, "NULL buffer");
    FE_CUDA(h, cudaSetDevice(h->device));
    cudaStream_t st = stream ? (cudaStream_t)stream : h->stream;
    std::vector<FlacFile> ff((size_t)n_files);
    long long frames_total = 0, pcm_span = 0;
    for (int i = 0; i < n_files; ++i) {
        const fe_flac_file& s = files[i];
        if (s.byte_offset < 0 || (s.byte_offset & 15) || s.n_bytes < 0 || s.byte_offset + s.n_bytes > total_bytes)
            return fail(h, FE_ERR_INVALID, "file byte ranges must be 16-byte aligned and inside the buffer");
        if (i > 0 && s.byte_offset < files[i - 1].byte_offset + files[i - 1].n_bytes)
            return fail(h, FE_ERR_INVALID, "files must be laid out in ascending, non-overlapping byte ranges");
        if (s.pcm_offset < 0 || (s.pcm_offset & 7)) return fail(h, FE_ERR_INVALID, "pcm_offset must be a multiple of 8 samples");
        if (s.n_samples < 0 || s.first_frame < 0 || s.first_frame > s.n_bytes) return fail(h, FE_ERR_INVALID, "bad stream layout");
        if (s.block_size < 16 || (s.block_size & 7) || s.block_size > 65535)
            return fail(h, FE_ERR_INVALID, "unsupported FLAC block size (fixed block size, multiple of 8, expected)");
        if (s.bits_per_sample < 4 || s.bits_per_sample > 16) return fail(h, FE_ERR_INVALID, "unsupported bits per sample (4..16)");
        FlacFile& f = ff[(size_t)i];
        f.byte_off = s.byte_offset; f.pcm_off = s.pcm_offset; f.n_bytes = s.n_bytes; f.first_frame = s.first_frame;
        f.n_samples = s.n_samples; f.block_size = s.block_size; f.bps = s.bits_per_sample;
        f.frame_base = (int)frames_total;
        f.n_frames = (s.n_samples + s.block_size - 1) / s.block_size;
        f.pad = 0;
        frames_total += f.n_frames;
        if (frames_total > 0x7fffffffLL) return fail(h, FE_ERR_INVALID, "batch too large");
        pcm_span = std::max<long long>(pcm_span, s.pcm_offset + s.n_samples);
    }
    if (pcm_span > pcm_capacity) return fail(h, FE_ERR_CAPACITY, "pcm buffer too small");
    const bool in_dev = is_device_ptr(bytes), out_dev = is_device_ptr(pcm);
    const int cap = (int)std::min<long long>(frames_total + frames_total / 64 + 1024, 0x7fffffffLL);
    const size_t ctr_ints = 1 + 2 * (size_t)n_files;                      // n_cands, cand_per_file[n], status[n]
    int rc;
    if ((rc = ensure(h, h->d_flac_files, sizeof(FlacFile) * ff.size()))) return rc;
    if ((rc = ensure(h, h->d_flac_frames, sizeof(FlacFrame) * (size_t)std::max<long long>(frames_total, 1)))) return rc;
    if ((rc = ensure(h, h->d_flac_cands, sizeof(int4) * (size_t)cap))) return rc;
    if ((rc = ensure(h, h->d_flac_ctr, sizeof(int) * ctr_ints))) return rc;
    const uint8_t* d_bytes = bytes;
    if (!in_dev) {                                                        // the kernels read up to 4 KB past the last file
        if ((rc = ensure(h, h->d_flac_bytes, (size_t)total_bytes + kFlacPadBytes))) return rc;
        FE_CUDA(h, cudaMemcpyAsync(h->d_flac_bytes.p, bytes, (size_t)total_bytes, cudaMemcpyHostToDevice, st));
        FE_CUDA(h, cudaMemsetAsync((char*)h->d_flac_bytes.p + total_bytes, 0, kFlacPadBytes, st));
        d_bytes = (const uint8_t*)h->d_flac_bytes.p;
    }
    short* d_pcm = pcm;
    if (!out_dev) {
        if ((rc = ensure(h, h->d_flac_pcm, sizeof(short) * (size_t)std::max<long long>(pcm_span, 8)))) return rc;
        d_pcm = (short*)h->d_flac_pcm.p;
    }
    int* d_ncand = (int*)h->d_flac_ctr.p;
    int* d_cpf = d_ncand + 1;
    int* d_status = d_cpf + n_files;
    FE_CUDA(h, cudaMemcpyAsync(h->d_flac_files.p, ff.data(), sizeof(FlacFile) * ff.size(), cudaMemcpyHostToDevice, st));
    FE_CUDA(h, cudaMemsetAsync(h->d_flac_frames.p, 0, sizeof(FlacFrame) * (size_t)std::max<long long>(frames_total, 1), st));
    FE_CUDA(h, cudaMemsetAsync(h->d_flac_ctr.p, 0, sizeof(int) * ctr_ints, st));
    cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};
    if (h->profiling) { for (auto& e : ev) FE_CUDA(h, cudaEventCreate(&e)); FE_CUDA(h, cudaEventRecord(ev[0], st)); }
    const long long nvec = (total_bytes + 15) >> 4;
    const int g_scan = (int)std::min<long long>((nvec + 255) / 256, 32LL * h->num_sms);
    if (nvec > 0) {
        k_flac_scan<<<std::max(g_scan, 1), 256, 0, st>>>(d_bytes, total_bytes, (const FlacFile*)h->d_flac_files.p, n_files,
                                                         (int4*)h->d_flac_cands.p, d_ncand, cap, d_cpf);
        h->launches++;
    }
    if (h->profiling) FE_CUDA(h, cudaEventRecord(ev[1], st));
    const int g_dec = (int)std::min<long long>((frames_total + 127) / 128 + 1, 64LL * h->num_sms);
    k_flac_decode<<<g_dec, 128, 0, st>>>(d_bytes, (const FlacFile*)h->d_flac_files.p, (const int4*)h->d_flac_cands.p, d_ncand, cap,
                                         (FlacFrame*)h->d_flac_frames.p, d_pcm, d_status, 0);
    h->launches++;
    if (h->profiling) FE_CUDA(h, cudaEventRecord(ev[2], st));
    k_flac_validate<<<(n_files + 255) / 256, 256, 0, st>>>((const FlacFile*)h->d_flac_files.p, n_files, (const FlacFrame*)h->d_flac_frames.p,
                                                           d_cpf, d_status);
    h->launches++;
    if (h->profiling) FE_CUDA(h, cudaEventRecord(ev[3], st));
    FE_CUDA(h, cudaGetLastError());
    std::vector<int> hs((size_t)n_files);
    FE_CUDA(h, cudaMemcpyAsync(hs.data(), d_status, sizeof(int) * (size_t)n_files, cudaMemcpyDeviceToHost, st));
    FE_CUDA(h, cudaStreamSynchronize(st));
    if (h->profiling) {
        for (int k = 0; k < 3; ++k) FE_CUDA(h, cudaEventElapsedTime(&h->flac_ms[k], ev[k], ev[k + 1]));
        for (auto& e : ev) cudaEventDestroy(e);
    }
    bool redo = false;
    for (int v : hs) redo |= (v == kFlacRedo);
    if (redo) {                                                           // a stray header-like byte run: decode chain positions again
        k_flac_decode<<<g_dec, 128, 0, st>>>(d_bytes, (const FlacFile*)h->d_flac_files.p, (const int4*)h->d_flac_cands.p, d_ncand, cap,
                                             (FlacFrame*)h->d_flac_frames.p, d_pcm, d_status, 1);
        h->launches++;
        FE_CUDA(h, cudaGetLastError());
        FE_CUDA(h, cudaStreamSynchronize(st));
    }
    int n_bad = 0;
    for (int i = 0; i < n_files; ++i) {
        status[i] = (hs[(size_t)i] == kFlacCorrupt) ? FE_ERR_INVALID : FE_OK;
        n_bad += status[i] != 0;
    }
    if (!out_dev) {
        FE_CUDA(h, cudaMemcpyAsync(pcm, d_pcm, sizeof(short) * (size_t)pcm_span, cudaMemcpyDeviceToHost, st));
        FE_CUDA(h, cudaStreamSynchronize(st));
    }
    if (n_bad) return fail(h, FE_ERR_INVALID, "FLAC decode: " + std::to_string(n_bad) + " file(s) corrupt, truncated or not fixed-block-size mono (see status[])");
    return FE_OK;
}

int fe_get_flac_ms(fe_handle* h, float ms[3]) {
    if (!h || !ms) return FE_ERR_INVALID;
    for (int k = 0; k < 3; ++k) ms[k] = h->flac_ms[k];
    return FE_OK;
}

int fe_get_pad_ms(fe_handle* h, float* ms) {
    if (!h || !ms) return FE_ERR_INVALID;
    *ms = h->pad_ms;
    return FE_OK;
}

int fe_host_alloc(fe_handle* h, int64_t n_bytes, void** out) {
    if (!h || !out || n_bytes <= 0) return h ? fail(h, FE_ERR_INVALID, "bad size / NULL out") : FE_ERR_INVALID;
    FE_CUDA(h, cudaSetDevice(h->device));
    FE_CUDA(h, cudaHostAlloc(out, (size_t)n_bytes, cudaHostAllocDefault));
    return FE_OK;
}

int fe_host_free(fe_handle* h, void* p) {
    if (!h) return FE_ERR_INVALID;
    if (p) FE_CUDA(h, cudaFreeHost(p));
    return FE_OK;
}

int fe_sync(fe_handle* h) {
    if (!h) return FE_ERR_INVALID;
    FE_CUDA(h, cudaSetDevice(h->device));
    for (Lane& L : h->lane) if (L.done_valid) FE_CUDA(h, cudaEventSynchronize(L.done));
    FE_CUDA(h, cudaStreamSynchronize(h->stream));
    return FE_OK;
}

int fe_measure_fp32_peaks(fe_handle* h, float tflops[4]) {
    if (!h || !tflops) return FE_ERR_INVALID;
    FE_CUDA(h, cudaSetDevice(h->device));
    const int grid = h->num_sms * 8, block = 256, iters = 2500;
    int rc;
    if ((rc = ensure(h, h->lane[0].d_out, sizeof(float) * (size_t)grid * block))) return rc;
    cudaEvent_t e0, e1;
    FE_CUDA(h, cudaEventCreate(&e0)); FE_CUDA(h, cudaEventCreate(&e1));
    float* o = (float*)h->lane[0].d_out.p;
    for (int mode = 0; mode < 4; ++mode) {
        float best = 0.f;
        for (int rep = 0; rep < 4; ++rep) {                                      // rep 0 = warm-up
            FE_CUDA(h, cudaEventRecord(e0, h->stream));
            if (mode == 0) k_fp32_peak<0><<<grid, block, 0, h->stream>>>(o, iters, 1.0001f, 0.5f);
            else if (mode == 1) k_fp32_peak<1><<<grid, block, 0, h->stream>>>(o, iters, 1.0001f, 0.5f);
            else if (mode == 2) k_fp32_peak<2><<<grid, block, 0, h->stream>>>(o, iters, 1.0001f, 0.5f);
            else k_fp32_peak<3><<<grid, block, 0, h->stream>>>(o, iters, 1.0001f, 0.5f);
            FE_CUDA(h, cudaEventRecord(e1, h->stream));
            FE_CUDA(h, cudaEventSynchronize(e1));
            float ms = 0.f;
            FE_CUDA(h, cudaEventElapsedTime(&ms, e0, e1));
            const double flops = (double)grid * block * iters * 8.0 * 8.0 * 2.0 * 2.0;   // 64 FMA pairs per iteration
            if (rep > 0) best = std::max(best, (float)(flops / (ms * 1e-3) / 1e12));
        }
        tflops[mode] = best;
        h->launches += 4;
    }
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    FE_CUDA(h, cudaGetLastError());
    return FE_OK;
}

int fe_measure_fp32_peak(fe_handle* h, float* tflops) {
    if (!tflops) return FE_ERR_INVALID;
    float v[4];
    const int rc = fe_measure_fp32_peaks(h, v);
    if (rc) return rc;
    *tflops = std::max(std::max(v[0], v[1]), std::max(v[2], v[3]));
    return FE_OK;
}

int fe_set_profiling(fe_handle* h, int on) {
    if (!h) return FE_ERR_INVALID;
    h->profiling = on;
    h->prof_used = 0;            // a new measurement window
    return FE_OK;
}

int fe_get_kernel_ms(fe_handle* h, float ms[4]) {
    if (!h || !ms) return FE_ERR_INVALID;
    ms[0] = ms[1] = ms[2] = ms[3] = 0.f;
    if (h->prof_used == 0) return fail(h, FE_ERR_STATE, "no profiled run");
    FE_CUDA(h, cudaEventSynchronize(h->prof[h->prof_used - 1].e[4]));
    double acc[4] = {0, 0, 0, 0};
    for (size_t i = 0; i < h->prof_used; ++i) {
        const fe_handle::ProfSet& p = h->prof[i];
        float v = 0.f;
        if (p.k0) { FE_CUDA(h, cudaEventElapsedTime(&v, p.e[1], p.e[2])); acc[0] += v; }
        if (p.groups > 0) {                       // K1 / K2 alternate over the L2 groups
            cudaEvent_t prev = p.e[2];
            for (int g = 0; g < p.groups; ++g) {
                FE_CUDA(h, cudaEventElapsedTime(&v, prev, p.ge[(size_t)2 * g])); acc[1] += v;
                FE_CUDA(h, cudaEventElapsedTime(&v, p.ge[(size_t)2 * g], p.ge[(size_t)2 * g + 1])); acc[2] += v;
                prev = p.ge[(size_t)2 * g + 1];
            }
        } else {
            FE_CUDA(h, cudaEventElapsedTime(&v, p.e[2], p.e[3])); acc[1] += v;
            if (p.k2) { FE_CUDA(h, cudaEventElapsedTime(&v, p.e[3], p.e[4])); acc[2] += v; }
        }
        FE_CUDA(h, cudaEventElapsedTime(&v, p.e[0], p.e[4])); acc[3] += v;
    }
    for (int k = 0; k < 4; ++k) ms[k] = (float)(acc[k] / (double)h->prof_used);   // mean over the window
    return FE_OK;
}

int64_t fe_launch_count(fe_handle* h) { return h ? h->launches : 0; }

// profiling aid (FE_K1_DBG & 8): read and clear the K1 per-phase cycle counters
int fe_debug_counters(fe_handle* h, uint64_t out[16]) {
    if (!h || !out) return FE_ERR_INVALID;
    FE_CUDA(h, cudaSetDevice(h->device));
    FE_CUDA(h, cudaDeviceSynchronize());
    FE_CUDA(h, cudaMemcpyFromSymbol(out, g_k1_prof, sizeof(uint64_t) * 16));
    uint64_t z[16] = {0};
    FE_CUDA(h, cudaMemcpyToSymbol(g_k1_prof, z, sizeof(z)));
    return FE_OK;
}

int64_t fe_device_bytes(fe_handle* h) {
    if (!h) return 0;
    size_t t = 0;
    for (const Lane& L : h->lane)
        for (const DevBuf* b : {&L.d_utts, &L.d_tile_prefix, &L.d_tiles, &L.d_atile_prefix, &L.d_atiles,
                                &L.d_statics, &L.d_stats, &L.d_pcm, &L.d_out, &L.d_scratch})
            t += b->cap;
    return (int64_t)t;
}

}  // extern "C"
