/* Host-side audio ingest / egress of libasr_frontend.so: the file boundary either
 * side of the CUDA front-end.
 *
 * Replaces, in /root/reference:
 *   sf.read(p)                                  preprocess.py:69   (libsndfile FLAC decode)
 *   tfm.build(src, dst) output files (16-bit)   utils/augmentation.py:28,53 (SoX FLAC encode)
 *
 * The decoder writes int16 PCM straight into the packed, 16-byte-aligned batch
 * buffer that fe_run() consumes (asr_frontend.h), from a pool of host threads,
 * so no per-utterance array is created between the file and the GPU.
 *
 * Formats: native FLAC (all subframe types: constant, verbatim, fixed, LPC; Rice
 * and Rice2 partitions with escapes; wasted bits; all stereo decorrelation
 * modes; frame CRC-8 / CRC-16 and the STREAMINFO MD5 are verified) at <= 16 bits
 * per sample, and RIFF/WAVE 16-bit PCM.  Plain pointers and sizes; every call
 * returns 0 or a negative AIO_ERR_* code, never throws.  All functions are
 * thread-safe (no global state).  This is host code: no CUDA device is needed.
 */
#ifndef ASR_AUDIO_IO_H_
#define ASR_AUDIO_IO_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define AIO_OK 0
#define AIO_ERR_INVALID (-1)       /* bad argument */
#define AIO_ERR_IO (-2)            /* open / read / write failed */
#define AIO_ERR_FORMAT (-3)        /* not a FLAC / WAV stream, or corrupt (CRC, MD5, truncated) */
#define AIO_ERR_UNSUPPORTED (-4)   /* valid stream outside what the front-end takes (> 16 bit, ...) */
#define AIO_ERR_CAPACITY (-5)      /* output buffer too small */

#define AIO_FMT_FLAC 1
#define AIO_FMT_WAV 2

typedef struct aio_info {
    int32_t format;            /* AIO_FMT_* */
    int32_t sample_rate;
    int32_t channels;
    int32_t bits_per_sample;
    int64_t n_samples;         /* per channel; -1 if the stream does not say (FLAC total_samples == 0) */
} aio_info;

/* Header parse only (no sample is decoded). */
int aio_probe_memory(const uint8_t* data, int64_t n_bytes, aio_info* info);
int aio_probe_file(const char* path, aio_info* info);

/* Whole-stream decode to interleaved int16 (samples with fewer than 16 bits are
 * scaled up, which keeps x / 2^(bits-1) -- what sf.read returns -- unchanged).
 * n_decoded receives samples per channel.  check_md5 != 0 also verifies the
 * STREAMINFO MD5 of the decoded audio (FLAC; skipped when the stream has none). */
int aio_decode_memory(const uint8_t* data, int64_t n_bytes, int16_t* out, int64_t capacity,
                      int64_t* n_decoded, int check_md5);
int aio_decode_file(const char* path, int16_t* out, int64_t capacity, int64_t* n_decoded, int check_md5);

/* Batch ingest with a pool of n_threads host threads (<= 0: one per core).
 * aio_probe_files fills info[n]; status[n] gets the per-file code.
 * aio_decode_files decodes file i (mono only) to out + offsets[i] (caller-planned,
 * capacity lengths[i] samples as probed).  The return value is the first
 * failing file's code (0 if none). */
int aio_probe_files(const char* const* paths, int32_t n, int32_t n_threads, aio_info* info, int32_t* status);
int aio_decode_files(const char* const* paths, int32_t n, int32_t n_threads, int16_t* out,
                     const int64_t* offsets, const int64_t* lengths, int check_md5, int32_t* status);

/* Device-side FLAC decode support (fe_decode_flac, asr_frontend.h): the stream layout the GPU kernels
 * need from the metadata -- byte offset of the first audio frame and the stream's block sizes -- and a
 * thread-pooled raw file read into one buffer (file i at buf + offsets[i], sizes[i] bytes; no decoding). */
typedef struct aio_flac_layout_t {
    int64_t n_samples;         /* per channel (0 if unknown) */
    int32_t first_frame;       /* byte offset of the first audio frame */
    int32_t min_block, max_block;   /* STREAMINFO values; min_block = 0 when the first frame uses variable blocking (0xFFF9) */
    int32_t sample_rate, channels, bits_per_sample;
} aio_flac_layout_t;
int aio_flac_layout(const uint8_t* data, int64_t n_bytes, aio_flac_layout_t* out);
int aio_flac_layouts(const uint8_t* buf, const int64_t* offsets, const int64_t* sizes, int32_t n, int32_t n_threads,
                     aio_flac_layout_t* out, int32_t* status);     /* batch form: file i at buf + offsets[i] */
int aio_file_sizes(const char* const* paths, int32_t n, int64_t* sizes);
int aio_read_files(const char* const* paths, int32_t n, int32_t n_threads, uint8_t* buf, const int64_t* offsets,
                   const int64_t* sizes, int32_t* status);

/* Encoders (mono or interleaved multi-channel int16).  FLAC: fixed block size
 * 4096; constant / verbatim / fixed-predictor (order 0..4) subframes with
 * partitioned Rice coding, frame CRCs and the STREAMINFO MD5 -- a subset of the
 * format every conforming decoder reads (no LPC search: lossless either way). */
int64_t aio_flac_bound(int64_t n_samples, int32_t channels);     /* worst-case encoded bytes */
int aio_encode_flac(const int16_t* pcm, int64_t n_samples, int32_t channels, int32_t sample_rate,
                    uint8_t* out, int64_t capacity, int64_t* n_bytes);
int aio_write_file(const char* path, const int16_t* pcm, int64_t n_samples, int32_t channels,
                   int32_t sample_rate, int32_t format);
int aio_write_files(const char* const* paths, int32_t n, int32_t n_threads, const int16_t* pcm,
                    const int64_t* offsets, const int64_t* lengths, int32_t sample_rate, int32_t format,
                    int32_t* status);

const char* aio_strerror(int code);

#ifdef __cplusplus
}
#endif
#endif  /* ASR_AUDIO_IO_H_ */
