/* Host-side record writer / reader of libasr_frontend.so: the consumer side of the
 * CUDA front-end's arrays.
 *
 * Replaces, in /root/reference:
 *   create_tfrecords(X, y, filename, ...)       create_tfrecord.py:43-97
 *       tf.train.Example{feat: FloatList(feat.flatten()), shape: Int64List(feat.shape),
 *                        token: Int64List(token)}.SerializeToString()   create_tfrecord.py:83-90
 *       tf.python_io.TFRecordWriter.write        create_tfrecord.py:69,90
 *   data_parser / tf.parse_single_example        tfrecord_data_loader.py:24-52   (reader, for verification
 *                                                                                  and TF-free consumers)
 *
 * The feature cubes are serialised straight from the flat float32 output buffer of
 * fe_run() (asr_frontend.h: utterance i = out + out_offsets[i], n_frames[i] x row floats),
 * so no per-utterance array or Python object is touched between the D2H copy and the file.
 *
 * File format (TFRecord): per record  u64 length | u32 masked_crc32c(length) | bytes |
 * u32 masked_crc32c(bytes), little endian, mask(c) = ((c >> 15) | (c << 17)) + 0xa282ead8,
 * CRC-32C (Castagnoli).  Payload: protobuf tensorflow.Example with the map entries in
 * key order (feat, shape, token) -- what deterministic protobuf serialisation emits.
 * Plain pointers and sizes; 0 or a negative RIO_ERR_* code; thread-safe; host only.
 */
#ifndef ASR_RECORD_IO_H_
#define ASR_RECORD_IO_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RIO_OK 0
#define RIO_ERR_INVALID (-1)
#define RIO_ERR_IO (-2)
#define RIO_ERR_FORMAT (-3)      /* bad length / CRC / protobuf */
#define RIO_ERR_CAPACITY (-4)

uint32_t rio_crc32c(const void* data, int64_t n_bytes);           /* plain CRC-32C (check: "123456789" -> 0xE3069283) */
uint32_t rio_masked_crc32c(const void* data, int64_t n_bytes);    /* as stored in TFRecord files */

/* Serialised size / bytes of one Example.  feat: n_feat floats; shape: n_shape int64; token: n_token int64. */
int64_t rio_example_size(int64_t n_feat, const int64_t* shape, int32_t n_shape, const int64_t* token, int64_t n_token);
int rio_example_serialize(const float* feat, int64_t n_feat, const int64_t* shape, int32_t n_shape,
                          const int64_t* token, int64_t n_token, uint8_t* out, int64_t capacity, int64_t* n_bytes);

/* Write records [0, n) to one TFRecord file.  Record r takes its cube from
 * feats + feat_offsets[r] (n_frames[r] * feat_dim * planes floats, shape = (n_frames[r], feat_dim, planes))
 * and its tokens from tokens + token_offsets[r] (token_lens[r] int64 values). */
int rio_write_tfrecord(const char* path, int32_t n, const float* feats, const int64_t* feat_offsets,
                       const int32_t* n_frames, int32_t feat_dim, int32_t planes,
                       const int64_t* tokens, const int64_t* token_offsets, const int32_t* token_lens);

/* Several files at once from a pool of host threads: file f holds records [file_start[f], file_start[f+1]). */
int rio_write_tfrecords(const char* const* paths, int32_t n_files, const int32_t* file_start, int32_t n_threads,
                        const float* feats, const int64_t* feat_offsets, const int32_t* n_frames, int32_t feat_dim,
                        int32_t planes, const int64_t* tokens, const int64_t* token_offsets,
                        const int32_t* token_lens, int32_t* status);

/* Reader.  rio_index_tfrecord verifies every CRC and returns the record count; with non-NULL arrays
 * (capacity entries) it also reports per record the number of feat floats, the shape (3 values,
 * zero-filled if shorter) and the number of tokens.  rio_read_tfrecord then fills caller-planned
 * buffers: feats + feat_offsets[r], tokens + token_offsets[r]. */
int64_t rio_index_tfrecord(const char* path, int64_t capacity, int64_t* n_feat, int64_t* shapes3, int64_t* n_token);
int rio_read_tfrecord(const char* path, int64_t n, float* feats, const int64_t* feat_offsets,
                      int64_t* tokens, const int64_t* token_offsets);

const char* rio_strerror(int code);

#ifdef __cplusplus
}
#endif
#endif  /* ASR_RECORD_IO_H_ */
