/* C-ABI of the B200 acoustic front-end (libasr_frontend.so).
 *
 * The reference has no FFI for this path; its de-facto boundary is one Python
 * function plus the library calls it makes (all in /root/reference):
 *
 *   process_audios(audio_path, args)            preprocess.py:50-91
 *     speechpy.feature.mfcc(...)                preprocess.py:72-76
 *     speechpy.feature.mfe(...)                 preprocess.py:78-82
 *     speechpy.processing.cmvn(feat, True)      preprocess.py:85
 *     speechpy.feature.extract_derivative_feature(feat)   preprocess.py:86
 *     feat.astype(np.float32)                   preprocess.py:88
 *   SpeedAugmentation / VolumeAugmentation      utils/augmentation.py:6-31, 33-56
 *
 * These entry points are what a ctypes binding behind that function binds
 * (INTEGRATION.md shows the stub).  Plain pointers and sizes only; every call
 * returns 0 on success and a negative code on failure (never throws, never
 * aborts); fe_last_error() gives the message.  A handle is NOT thread-safe: use
 * one handle per (host thread, GPU).  There is no CPU fallback: without a CUDA
 * device fe_create fails.
 */
#ifndef ASR_FRONTEND_H_
#define ASR_FRONTEND_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FE_ABI_VERSION 2

#define FE_OK 0
#define FE_ERR_INVALID (-1)      /* bad argument / unsupported configuration */
#define FE_ERR_CUDA (-2)         /* CUDA runtime error (message has the detail) */
#define FE_ERR_CAPACITY (-3)     /* caller's output buffer is too small */
#define FE_ERR_STATE (-4)        /* call order (e.g. run before configure) */

#define FE_FEAT_MFCC 0           /* args.feat_type == 'mfcc'   preprocess.py:71 */
#define FE_FEAT_FBANK 1          /* args.feat_type == 'fbank'  preprocess.py:77 */
#define FE_DELTA_SPEECHPY 0      /* speechpy.processing.derivative_extraction as shipped */
#define FE_DELTA_TIME_REGRESSION 1
#define FE_PCM_INT16 0
#define FE_PCM_FLOAT32 1

typedef struct fe_handle fe_handle;

/* Configuration + host-built constant tables (copied during fe_configure; the
 * caller may free them afterwards).  Tables are computed in float64 on the host
 * and rounded to float32 so host and device cannot disagree on a filter edge. */
typedef struct fe_config {
    int32_t abi_version;      /* FE_ABI_VERSION */
    int32_t sample_rate;      /* informational (fs comes from the file, preprocess.py:69) */
    int32_t frame_len;        /* samples; args.frame_length ms -> 400 at 16 kHz */
    int32_t hop;              /* samples; args.frame_step ms   -> 160 at 16 kHz */
    int32_t nfft;             /* 512 (speechpy default fft_length) */
    int32_t num_filters;      /* 40 for mfcc; args.feat_dim for fbank */
    int32_t feat_dim;         /* D: num_cepstral (mfcc) or num_filters (fbank) */
    int32_t feat_type;        /* FE_FEAT_* */
    int32_t cmvn;             /* args.cmvn: 1 -> CMVN + deltas, cube (L, D, 3); 0 -> (L, D) */
    int32_t delta_mode;       /* FE_DELTA_* */
    int32_t fbank_log;        /* 0 = linear mel energies (what mfe returns), 1 = log */
    int32_t dc_elimination;   /* 1: c0 <- log(frame energy) (speechpy mfcc default) */
    int32_t pcm_dtype;        /* FE_PCM_* */
    float   preemph;          /* 0 = off (reference); else y[n] = x[n] - preemph * x[n-1], circular */
    int32_t fb_nnz;
    const int32_t* fb_row_start;   /* [num_filters + 1] CSR row starts */
    const int32_t* fb_first_bin;   /* [num_filters] first FFT bin of each filter's run */
    const float*   fb_weights;     /* [fb_nnz] triangle weights (unscaled) */
    const float*   dct;            /* [feat_dim * num_filters] DCT-II ortho rows (mfcc only) */
    const float*   window;         /* [frame_len] or NULL = rectangular (reference) */
    const float*   tw256;          /* [16*16*2] W_256^(j k) as (cos, -sin) */
    const float*   tw512;          /* [257*2]  (cos, sin)(2 pi k / 512) */
    int32_t n_speeds;              /* resampler variants (speed perturbation), may be 0 */
    const int32_t* speed_up;       /* [n_speeds] speed = down / up */
    const int32_t* speed_down;     /* [n_speeds] */
    const float*   speed_taps;     /* concatenated [up_i][speed_ntaps] polyphase taps: row p = (j * down) % up of output j,
                                      column t weighs input floor(j * down / up) - (speed_ntaps / 2 - 1) + t */
    int32_t speed_ntaps;           /* taps per phase, even, 2..256 (ABI 2; the round-2 filter has 128) */
} fe_config;

/* Pure host helper: frame-count rule of speechpy.processing.stack_frames with
 * zero_padding=False, floor((n - frame_len) / hop), clamped at 0.
 * (tfrecord_data_loader.py:78-79 pins it: 522320 -> 3262, 559280 -> 3493.) */
int64_t fe_num_frames(int64_t n_samples, int32_t frame_len, int32_t hop);

/* 1 if the kernels are built for this frame geometry (samples): 400/160 = 25 / 10 ms at 16 kHz (the reference's
 * run.sh), 320/160, 480/160, 200/80 = 25 / 10 ms at 8 kHz.  The reference takes frame_length / frame_step from its
 * arguments (las/arguments.py:33-40) and fs from every file (preprocess.py:69). */
int fe_geometry_supported(int32_t frame_len, int32_t hop);

/* ceil(n * up / down): utterance length after speed perturbation. */
int64_t fe_resampled_length(int64_t n_samples, int32_t up, int32_t down);

int fe_create(int device, fe_handle** out);
int fe_destroy(fe_handle* h);
int fe_configure(fe_handle* h, const fe_config* cfg);

/* Host-only planning: per-utterance frame counts and output offsets (in floats,
 * each utterance 16-byte aligned), so the caller can size `out`.
 *   pcm_lengths[n]  samples per utterance
 *   speed_idx[n]    index into the configured speeds, -1 = none; NULL = none
 *   out_offsets[n+1], n_frames[n]  filled on return; out_offsets[n] = total floats */
int fe_plan(fe_handle* h, const int64_t* pcm_lengths, int32_t n_utts, const int32_t* speed_idx,
            int64_t* out_offsets, int32_t* n_frames);

/* The hot path.  Replaces the loop body of process_audios (preprocess.py:67-89)
 * for a whole batch of utterances.
 *   pcm           packed PCM, host OR device pointer (detected); int16 or float32 per config
 *   pcm_offsets[n] element offset of each utterance (multiple of 8 for int16, 4 for float32)
 *   pcm_lengths[n] samples per utterance
 *   speed_idx/gain per-utterance perturbation (NULL = none); gain g: clip(round(g * x))
 *   out           host OR device pointer with room for out_capacity floats; utterance i
 *                 is the C-contiguous (L_i, D, 3) float32 cube at out + out_offsets[i]
 *                 ((L_i, D) when cmvn == 0)
 *   out_offsets[n+1], n_frames[n]  filled on return (host arrays)
 *   stream        cudaStream_t to run on, NULL = the handle's own stream
 * Asynchronous for device `out` (fe_sync to wait); synchronous when `out` is host memory. */
int fe_run(fe_handle* h, const void* pcm, const int64_t* pcm_offsets, const int64_t* pcm_lengths,
           int32_t n_utts, const int32_t* speed_idx, const float* gain,
           float* out, int64_t out_capacity, int64_t* out_offsets, int32_t* n_frames,
           void* stream);

/* Speed / volume perturbation only (utils/augmentation.py:6-56 without the
 * file round-trip): int16 in -> int16 out, host or device pointers.
 *   dst_offsets[n+1] filled on return (element offsets, multiples of 8). */
int fe_perturb(fe_handle* h, const int16_t* pcm, const int64_t* pcm_offsets,
               const int64_t* pcm_lengths, int32_t n_utts, const int32_t* speed_idx,
               const float* gain, int16_t* dst, int64_t dst_capacity, int64_t* dst_offsets,
               int64_t* dst_lengths, void* stream);

/* speechpy.processing.cmvn(vec, variance_normalization) and
 * speechpy.feature.extract_derivative_feature(feature) on their own (preprocess.py:85-86),
 * for a batch of (L_i, D) float32 matrices, host or device pointers.  Needs no fe_configure.
 *   feat_offsets[n] float offsets of each matrix, n_frames[n] rows
 *   mode  bit0 subtract the per-column mean, bit1 divide by (std + 2^-30), bit2 append delta and
 *         delta-delta -> (L, D, 3); without bit2 the result stays (L, D)
 *   out_offsets[n+1] filled on return (floats, 16-byte aligned).  Synchronous. */
#define FE_POST_MEAN 1
#define FE_POST_VAR 2
#define FE_POST_DELTAS 4
int fe_postprocess(fe_handle* h, const float* feats, const int64_t* feat_offsets, const int32_t* n_frames,
                   int32_t n_utts, int32_t feat_dim, int32_t mode, int32_t delta_mode,
                   float* out, int64_t out_capacity, int64_t* out_offsets, void* stream);

/* Bucketed, padded batches: what tf.data's bucket_by_sequence_length(..., pad_to_bucket_boundary=True)
 * hands the model (tfrecord_data_loader.py:75-94), built on the device from the cubes fe_run left in HBM.
 * Slot i of the batch buffer (dst + dst_offsets[i], slot_floats[i] = T_pad * D * planes floats) receives the
 * valid_floats[i] = L_i * D * planes floats at feats + src_offsets[i] followed by zeros.  Which utterance goes
 * to which slot (bucket by length, batch size per bucket, drop L >= 1710) is the caller's plan -- the host
 * mirror builds it with the reference's boundaries.  One launch for any number of batches; host or device
 * pointers; asynchronous on the stream when both are device pointers.  Needs no fe_configure. */
int fe_pad_batches(fe_handle* h, const float* feats, const int64_t* src_offsets, const int32_t* valid_floats,
                   const int64_t* dst_offsets, const int32_t* slot_floats, int32_t n_slots,
                   float* dst, int64_t dst_capacity, void* stream);
int fe_get_pad_ms(fe_handle* h, float* ms);    /* duration of the last fe_pad_batches launch (profiling on) */

/* FLAC decode on the device: sf.read (preprocess.py:69) behind the PCIe link.  The caller uploads (or
 * passes host memory holding) the raw bytes of n FLAC files, each starting at a 16-byte aligned offset
 * of one buffer, with the stream layout the host probe reports (aio_flac_layout, asr_audio_io.h); the
 * int16 samples of file i appear at pcm + pcm_offset (host or device memory), ready for fe_run.
 * Streams must be mono, <= 16 bits, fixed block size (a multiple of 8) -- what libFLAC, SoX and this
 * library's encoder write.  Every frame's CRC-8 / CRC-16 and the frame chain of every file are
 * verified on the device; status[i] = 0 or FE_ERR_INVALID (corrupt, truncated, unsupported layout) and
 * the call returns FE_ERR_INVALID if any file failed (the others are decoded).  A device `bytes` buffer
 * must be readable for 4096 bytes past total_bytes.  Synchronous. */
typedef struct fe_flac_file {
    int64_t byte_offset;      /* of the file's first byte in `bytes`, multiple of 16 */
    int64_t pcm_offset;       /* int16 element offset of its first sample in `pcm`, multiple of 8 */
    int32_t n_bytes;
    int32_t first_frame;      /* byte offset of the first audio frame inside the file */
    int32_t n_samples;        /* STREAMINFO total samples */
    int32_t block_size;       /* STREAMINFO min == max block size */
    int32_t bits_per_sample;
    int32_t reserved;
} fe_flac_file;
int fe_decode_flac(fe_handle* h, const uint8_t* bytes, int64_t total_bytes, const fe_flac_file* files, int32_t n_files,
                   int16_t* pcm, int64_t pcm_capacity, int32_t* status, void* stream);
int fe_get_flac_ms(fe_handle* h, float ms[3]);   /* scan, decode, validate of the last call (profiling on) */

/* Page-locked host memory for the caller's staging buffers (file bytes, PCM, cubes): transfers from / to such
 * buffers run at the PCIe rate and asynchronously, pageable ones at a fraction of it.  Freed with fe_host_free
 * (or by fe_destroy of a still-live handle is NOT implied: the caller owns these). */
int fe_host_alloc(fe_handle* h, int64_t n_bytes, void** out);
int fe_host_free(fe_handle* h, void* p);

int fe_sync(fe_handle* h);

/* Measurement hooks.  With profiling on, every kernel of every fe_run is bracketed by CUDA
 * events on the launching stream (no synchronisation); fe_get_kernel_ms returns the MEAN over the
 * runs since fe_set_profiling(h, 1) was last called.
 *   ms[0] resample  ms[1] frames->statics  ms[2] cmvn+delta+pack  ms[3] whole device pass */
int fe_set_profiling(fe_handle* h, int on);
/* FP32 CUDA-core peak of this GPU (TFLOP/s), the measured denominator of the kernels' FP32 roofline: the best of four
 * independent-chain probes -- [0] scalar FFMA with warp-uniform multiplier / addend, [1] scalar FFMA with three
 * register operands, [2] packed FFMA2 with uniform operands, [3] packed FFMA2 with three register pairs.
 * fe_measure_fp32_peaks returns all four (what the pipe sustains depends on where the operands come from). */
int fe_measure_fp32_peak(fe_handle* h, float* tflops);
int fe_measure_fp32_peaks(fe_handle* h, float tflops[4]);
int fe_get_kernel_ms(fe_handle* h, float ms[4]);
int64_t fe_launch_count(fe_handle* h);       /* kernels launched since fe_create */
/* profiling aid: with FE_K1_DBG=8 in the environment K1 accumulates clock64 totals per phase
 * (lane 0 of every warp); this reads and clears them.  Layout: see g_k1_prof in fe_kernels.cuh. */
int fe_debug_counters(fe_handle* h, uint64_t out[16]);
int64_t fe_device_bytes(fe_handle* h);       /* scratch currently held on the device */

const char* fe_last_error(fe_handle* h);     /* h may be NULL: last fe_create error */
int fe_abi_version(void);

#ifdef __cplusplus
}
#endif
#endif  /* ASR_FRONTEND_H_ */
