"""ctypes binding of libasr_frontend.so (declarations mirror include/asr_frontend.h).

There is deliberately no fallback: if the shared library is missing or cannot be
loaded the import of the product API fails with an explicit error."""
import ctypes as C
import os

from . import build as _build

FE_ABI_VERSION = 2
FE_OK, FE_ERR_INVALID, FE_ERR_CUDA, FE_ERR_CAPACITY, FE_ERR_STATE = 0, -1, -2, -3, -4
FE_FEAT_MFCC, FE_FEAT_FBANK = 0, 1
FE_DELTA_SPEECHPY, FE_DELTA_TIME_REGRESSION = 0, 1
FE_PCM_INT16, FE_PCM_FLOAT32 = 0, 1
FE_POST_MEAN, FE_POST_VAR, FE_POST_DELTAS = 1, 2, 4

_i32p = C.POINTER(C.c_int32)
_i64p = C.POINTER(C.c_int64)
_f32p = C.POINTER(C.c_float)


class FeConfig(C.Structure):
    _fields_ = [
        ("abi_version", C.c_int32), ("sample_rate", C.c_int32), ("frame_len", C.c_int32),
        ("hop", C.c_int32), ("nfft", C.c_int32), ("num_filters", C.c_int32),
        ("feat_dim", C.c_int32), ("feat_type", C.c_int32), ("cmvn", C.c_int32),
        ("delta_mode", C.c_int32), ("fbank_log", C.c_int32), ("dc_elimination", C.c_int32),
        ("pcm_dtype", C.c_int32), ("preemph", C.c_float), ("fb_nnz", C.c_int32),
        ("fb_row_start", _i32p), ("fb_first_bin", _i32p), ("fb_weights", _f32p),
        ("dct", _f32p), ("window", _f32p), ("tw256", _f32p), ("tw512", _f32p),
        ("n_speeds", C.c_int32), ("speed_up", _i32p), ("speed_down", _i32p), ("speed_taps", _f32p),
        ("speed_ntaps", C.c_int32),
    ]


class FeFlacFile(C.Structure):
    _fields_ = [("byte_offset", C.c_int64), ("pcm_offset", C.c_int64), ("n_bytes", C.c_int32), ("first_frame", C.c_int32),
                ("n_samples", C.c_int32), ("block_size", C.c_int32), ("bits_per_sample", C.c_int32), ("reserved", C.c_int32)]


# every symbol include/asr_frontend.h declares: name -> (restype, argtypes)
SYMBOLS = {
    "fe_abi_version": (C.c_int, []),
    "fe_num_frames": (C.c_int64, [C.c_int64, C.c_int32, C.c_int32]),
    "fe_resampled_length": (C.c_int64, [C.c_int64, C.c_int32, C.c_int32]),
    "fe_geometry_supported": (C.c_int, [C.c_int32, C.c_int32]),
    "fe_create": (C.c_int, [C.c_int, C.POINTER(C.c_void_p)]),
    "fe_destroy": (C.c_int, [C.c_void_p]),
    "fe_configure": (C.c_int, [C.c_void_p, C.POINTER(FeConfig)]),
    "fe_plan": (C.c_int, [C.c_void_p, _i64p, C.c_int32, _i32p, _i64p, _i32p]),
    "fe_run": (C.c_int, [C.c_void_p, C.c_void_p, _i64p, _i64p, C.c_int32, _i32p, _f32p,
                         C.c_void_p, C.c_int64, _i64p, _i32p, C.c_void_p]),
    "fe_perturb": (C.c_int, [C.c_void_p, C.c_void_p, _i64p, _i64p, C.c_int32, _i32p, _f32p,
                             C.c_void_p, C.c_int64, _i64p, _i64p, C.c_void_p]),
    "fe_postprocess": (C.c_int, [C.c_void_p, C.c_void_p, _i64p, _i32p, C.c_int32, C.c_int32, C.c_int32, C.c_int32,
                                 C.c_void_p, C.c_int64, _i64p, C.c_void_p]),
    "fe_pad_batches": (C.c_int, [C.c_void_p, C.c_void_p, _i64p, _i32p, _i64p, _i32p, C.c_int32, C.c_void_p, C.c_int64,
                                 C.c_void_p]),
    "fe_get_pad_ms": (C.c_int, [C.c_void_p, _f32p]),
    "fe_decode_flac": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.POINTER(FeFlacFile), C.c_int32, C.c_void_p, C.c_int64,
                                 _i32p, C.c_void_p]),
    "fe_get_flac_ms": (C.c_int, [C.c_void_p, _f32p]),
    "fe_host_alloc": (C.c_int, [C.c_void_p, C.c_int64, C.POINTER(C.c_void_p)]),
    "fe_host_free": (C.c_int, [C.c_void_p, C.c_void_p]),
    "fe_sync": (C.c_int, [C.c_void_p]),
    "fe_set_profiling": (C.c_int, [C.c_void_p, C.c_int]),
    "fe_measure_fp32_peak": (C.c_int, [C.c_void_p, _f32p]),
    "fe_measure_fp32_peaks": (C.c_int, [C.c_void_p, _f32p]),
    "fe_get_kernel_ms": (C.c_int, [C.c_void_p, _f32p]),
    "fe_launch_count": (C.c_int64, [C.c_void_p]),
    "fe_debug_counters": (C.c_int, [C.c_void_p, C.POINTER(C.c_uint64)]),
    "fe_device_bytes": (C.c_int64, [C.c_void_p]),
    "fe_last_error": (C.c_char_p, [C.c_void_p]),
}

AIO_OK, AIO_ERR_INVALID, AIO_ERR_IO, AIO_ERR_FORMAT, AIO_ERR_UNSUPPORTED, AIO_ERR_CAPACITY = 0, -1, -2, -3, -4, -5
AIO_FMT_FLAC, AIO_FMT_WAV = 1, 2


class AioInfo(C.Structure):
    _fields_ = [("format", C.c_int32), ("sample_rate", C.c_int32), ("channels", C.c_int32),
                ("bits_per_sample", C.c_int32), ("n_samples", C.c_int64)]


class AioFlacLayout(C.Structure):
    _fields_ = [("n_samples", C.c_int64), ("first_frame", C.c_int32), ("min_block", C.c_int32), ("max_block", C.c_int32),
                ("sample_rate", C.c_int32), ("channels", C.c_int32), ("bits_per_sample", C.c_int32)]


_cpp = C.POINTER(C.c_char_p)
_infop = C.POINTER(AioInfo)
# every symbol include/asr_audio_io.h declares (host-side FLAC / WAV ingest and egress)
AIO_SYMBOLS = {
    "aio_probe_memory": (C.c_int, [C.c_void_p, C.c_int64, _infop]),
    "aio_probe_file": (C.c_int, [C.c_char_p, _infop]),
    "aio_decode_memory": (C.c_int, [C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, _i64p, C.c_int]),
    "aio_decode_file": (C.c_int, [C.c_char_p, C.c_void_p, C.c_int64, _i64p, C.c_int]),
    "aio_probe_files": (C.c_int, [_cpp, C.c_int32, C.c_int32, _infop, _i32p]),
    "aio_decode_files": (C.c_int, [_cpp, C.c_int32, C.c_int32, C.c_void_p, _i64p, _i64p, C.c_int, _i32p]),
    "aio_flac_layout": (C.c_int, [C.c_void_p, C.c_int64, C.POINTER(AioFlacLayout)]),
    "aio_flac_layouts": (C.c_int, [C.c_void_p, _i64p, _i64p, C.c_int32, C.c_int32, C.POINTER(AioFlacLayout), _i32p]),
    "aio_file_sizes": (C.c_int, [_cpp, C.c_int32, _i64p]),
    "aio_read_files": (C.c_int, [_cpp, C.c_int32, C.c_int32, C.c_void_p, _i64p, _i64p, _i32p]),
    "aio_flac_bound": (C.c_int64, [C.c_int64, C.c_int32]),
    "aio_encode_flac": (C.c_int, [C.c_void_p, C.c_int64, C.c_int32, C.c_int32, C.c_void_p, C.c_int64, _i64p]),
    "aio_write_file": (C.c_int, [C.c_char_p, C.c_void_p, C.c_int64, C.c_int32, C.c_int32, C.c_int32]),
    "aio_write_files": (C.c_int, [_cpp, C.c_int32, C.c_int32, C.c_void_p, _i64p, _i64p, C.c_int32, C.c_int32, _i32p]),
    "aio_strerror": (C.c_char_p, [C.c_int]),
}

RIO_OK, RIO_ERR_INVALID, RIO_ERR_IO, RIO_ERR_FORMAT, RIO_ERR_CAPACITY = 0, -1, -2, -3, -4
# every symbol include/asr_record_io.h declares (TFRecord / tf.train.Example writer and reader)
RIO_SYMBOLS = {
    "rio_crc32c": (C.c_uint32, [C.c_void_p, C.c_int64]),
    "rio_masked_crc32c": (C.c_uint32, [C.c_void_p, C.c_int64]),
    "rio_example_size": (C.c_int64, [C.c_int64, _i64p, C.c_int32, _i64p, C.c_int64]),
    "rio_example_serialize": (C.c_int, [C.c_void_p, C.c_int64, _i64p, C.c_int32, _i64p, C.c_int64, C.c_void_p, C.c_int64, _i64p]),
    "rio_write_tfrecord": (C.c_int, [C.c_char_p, C.c_int32, C.c_void_p, _i64p, _i32p, C.c_int32, C.c_int32,
                                     _i64p, _i64p, _i32p]),
    "rio_write_tfrecords": (C.c_int, [_cpp, C.c_int32, _i32p, C.c_int32, C.c_void_p, _i64p, _i32p, C.c_int32, C.c_int32,
                                      _i64p, _i64p, _i32p, _i32p]),
    "rio_index_tfrecord": (C.c_int64, [C.c_char_p, C.c_int64, _i64p, _i64p, _i64p]),
    "rio_read_tfrecord": (C.c_int, [C.c_char_p, C.c_int64, C.c_void_p, _i64p, _i64p, _i64p]),
    "rio_strerror": (C.c_char_p, [C.c_int]),
}

_lib = None


class FrontendLibraryError(RuntimeError):
    pass


def library_path():
    """The in-tree library; FE_LIB points profiling / experiment builds (build.build_library(out=...)) at another file."""
    return os.environ.get("FE_LIB") or _build.LIB_PATH


def load():
    """dlopen the in-tree CUDA library (build it with ``python -m <pkg>.build`` or
    ``__graft_entry__.build()``).  No CPU path exists behind this call."""
    global _lib
    if _lib is not None:
        return _lib
    path = library_path()
    if not os.path.exists(path):
        raise FrontendLibraryError(
            "%s is missing: build the CUDA extension first (__graft_entry__.build()); "
            "this package has no CPU fallback" % path)
    try:
        lib = C.CDLL(path)
    except OSError as e:
        raise FrontendLibraryError("cannot load %s: %s (no CPU fallback)" % (path, e))
    for name, (res, args) in list(SYMBOLS.items()) + list(AIO_SYMBOLS.items()) + list(RIO_SYMBOLS.items()):
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    if lib.fe_abi_version() != FE_ABI_VERSION:
        raise FrontendLibraryError("ABI version mismatch: rebuild %s" % path)
    _lib = lib
    return lib
