"""Audio ingest / egress for the file-based boundary: replaces ``sf.read`` at
/root/reference/preprocess.py:69 and the files SoX writes at utils/augmentation.py:28,53.

FLAC (what LibriSpeech ships, ``prepare_libri_data.sh``) and 16-bit WAV are decoded and
encoded by the native codec in libasr_frontend.so (include/asr_audio_io.h, ``aio_*``);
``.npy`` holds raw int16 / float samples.  ``read_audio_batch`` decodes a file list with a
pool of host threads straight into the packed, 16-byte-aligned int16 buffer ``fe_run``
consumes, so nothing is copied between the decoder and the H2D transfer.  No third-party
audio package is used (the reference needs ``soundfile`` + libsndfile for the same job)."""
import ctypes as C
import os

import numpy as np

from . import _lib

DEFAULT_FS = 16000
_MAX_SAMPLES_PER_BYTE = 8192      # a constant subframe holds 65 535 samples in ~9 bytes; more than that is a damaged header
_FORMATS = {".flac": _lib.AIO_FMT_FLAC, ".wav": _lib.AIO_FMT_WAV}


class AudioFormatError(RuntimeError):
    pass


class MixedSampleRates(ValueError):
    """Files of different sample rates in one batch call (process_audios then runs one pass per rate)."""


class UnsupportedStreamError(AudioFormatError):
    """A valid stream that the DEVICE decoder does not take (stereo, > 16 bit, variable block size, no sample
    count); the host decoder reads it."""


def _check(rc, what):
    if rc != 0:
        msg = _lib.load().aio_strerror(int(rc))
        raise AudioFormatError("%s: %s (%d)" % (what, msg.decode() if msg else "error", rc))


def _paths_array(paths):
    arr = (C.c_char_p * len(paths))()
    arr[:] = [os.fsencode(p) for p in paths]
    return arr


def probe(path):
    """-> dict(format, sample_rate, channels, bits_per_sample, n_samples) from the header only."""
    info = _lib.AioInfo()
    _check(_lib.load().aio_probe_file(os.fsencode(path), C.byref(info)), path)
    return {"format": {1: "flac", 2: "wav"}[info.format], "sample_rate": info.sample_rate, "channels": info.channels,
            "bits_per_sample": info.bits_per_sample, "n_samples": info.n_samples}


def probe_batch(paths, n_threads=0):
    """Header probe of a file list with the native thread pool -> list of dicts (see probe)."""
    n = len(paths)
    info = (_lib.AioInfo * max(n, 1))()
    status = np.zeros(max(n, 1), dtype=np.int32)
    rc = _lib.load().aio_probe_files(_paths_array(paths), n, int(n_threads), info,
                                     status.ctypes.data_as(C.POINTER(C.c_int32)))
    if rc != 0:
        bad = int(np.flatnonzero(status)[0])
        _check(int(status[bad]), paths[bad])
    return [{"format": {1: "flac", 2: "wav"}[info[i].format], "sample_rate": info[i].sample_rate,
             "channels": info[i].channels, "bits_per_sample": info[i].bits_per_sample, "n_samples": info[i].n_samples}
            for i in range(n)]


def decode_bytes(data, check_md5=True):
    """FLAC / WAV bytes -> (int16 samples [n] or [n, channels], fs)."""
    lib = _lib.load()
    buf = np.frombuffer(data, dtype=np.uint8)
    info = _lib.AioInfo()
    _check(lib.aio_probe_memory(C.c_void_p(buf.ctypes.data), buf.size, C.byref(info)), "probe")
    claimed, ch = info.n_samples, info.channels
    # The sample count is a header field: never trust it for the allocation.  Start from what the bytes
    # plausibly hold and grow while the decoder reports a full buffer, up to the claimed count.
    n = max(buf.size * 16 // max(ch, 1), 65536)
    if claimed >= 0:
        n = min(n, claimed)
    while True:
        out = np.empty(max(n * ch, 1), dtype=np.int16)
        got = C.c_int64()
        rc = lib.aio_decode_memory(C.c_void_p(buf.ctypes.data), buf.size, C.c_void_p(out.ctypes.data), n * ch,
                                   C.byref(got), int(bool(check_md5)))
        if rc == _lib.AIO_ERR_CAPACITY and (claimed < 0 or n < claimed) and n < (1 << 36):
            n = n * 4 if claimed < 0 else min(n * 4, claimed)
            continue
        _check(rc, "decode")
        break
    out = out[:got.value * ch]
    return (out if ch == 1 else out.reshape(-1, ch)), info.sample_rate


def read_audio(path, check_md5=True):
    """-> (samples, fs).  int16 for 16-bit (or narrower) PCM sources -- the kernels scale by
    1/32768, which equals the float64 ``sf.read`` hands the reference -- float32 for .npy floats.
    Multi-channel files come back as (n, channels), like ``sf.read``."""
    ext = os.path.splitext(path)[1].lower()
    if ext == ".npy":
        data = np.load(path)
        if data.ndim != 1:
            raise ValueError("%s: 1-D samples expected" % path)
        if data.dtype != np.int16:
            data = data.astype(np.float32)
        return data, DEFAULT_FS
    with open(path, "rb") as f:
        raw = f.read()
    try:
        return decode_bytes(raw, check_md5)
    except AudioFormatError as e:
        raise AudioFormatError("%s: %s" % (path, e))


def plan_batch(lengths, align=8):
    """Sample counts -> (offsets[n], total) with every utterance starting on a 16-byte boundary."""
    lengths = np.asarray(lengths, dtype=np.int64)
    padded = (lengths + align - 1) // align * align
    offsets = np.zeros(lengths.size, dtype=np.int64)
    if lengths.size > 1:
        np.cumsum(padded[:-1], out=offsets[1:])
    return offsets, int(padded.sum())


def read_audio_batch(paths, n_threads=0, check_md5=True, out=None):
    """Decode mono FLAC / WAV files into one packed int16 buffer.

    Returns (packed, offsets[n], lengths[n], fs).  ``out`` may be a caller-owned (e.g. pinned)
    int16 array to decode into.  All files must share one sample rate."""
    lib = _lib.load()
    n = len(paths)
    if n == 0:
        return np.zeros(8, np.int16), np.zeros(0, np.int64), np.zeros(0, np.int64), DEFAULT_FS
    arr = _paths_array(paths)
    info = (_lib.AioInfo * n)()
    status = np.zeros(n, dtype=np.int32)
    rc = lib.aio_probe_files(arr, n, int(n_threads), info, status.ctypes.data_as(C.POINTER(C.c_int32)))
    if rc != 0:
        bad = int(np.flatnonzero(status)[0])
        _check(int(status[bad]), paths[bad])
    fs = info[0].sample_rate
    lengths = np.empty(n, dtype=np.int64)
    for i in range(n):
        if info[i].channels != 1:
            raise ValueError("%s: mono audio expected" % paths[i])
        if info[i].sample_rate != fs:
            raise MixedSampleRates("mixed sample rates in one call: %d vs %d (%s)" % (fs, info[i].sample_rate, paths[i]))
        if info[i].n_samples < 0:
            raise AudioFormatError("%s: FLAC stream without a sample count; use read_audio" % paths[i])
        if info[i].n_samples > _MAX_SAMPLES_PER_BYTE * max(os.path.getsize(paths[i]), 1):
            raise AudioFormatError("%s: implausible sample count %d in the header" % (paths[i], info[i].n_samples))
        lengths[i] = info[i].n_samples
    offsets, total = plan_batch(lengths)
    if out is None:
        out = np.zeros(max(total, 8), dtype=np.int16)
    elif out.dtype != np.int16 or out.size < total or not out.flags.c_contiguous:
        raise ValueError("out must be a C-contiguous int16 array of at least %d samples" % total)
    rc = lib.aio_decode_files(arr, n, int(n_threads), C.c_void_p(out.ctypes.data),
                              offsets.ctypes.data_as(C.POINTER(C.c_int64)), lengths.ctypes.data_as(C.POINTER(C.c_int64)),
                              int(bool(check_md5)), status.ctypes.data_as(C.POINTER(C.c_int32)))
    if rc != 0:
        bad = int(np.flatnonzero(status)[0])
        _check(int(status[bad]), paths[bad])
    return out, offsets, lengths, fs


def flac_layout(data):
    """Stream layout of FLAC bytes (for the device decoder) -> dict."""
    buf = np.frombuffer(data, dtype=np.uint8)
    lay = _lib.AioFlacLayout()
    _check(_lib.load().aio_flac_layout(C.c_void_p(buf.ctypes.data), buf.size, C.byref(lay)), "flac_layout")
    return {k: getattr(lay, k) for k, _ in _lib.AioFlacLayout._fields_}


def load_flac_batch(paths, n_threads=0, out=None):
    """Raw bytes of mono FLAC files in ONE buffer for ``Frontend.decode_flac`` (the GPU decoder): file i
    starts at a 16-byte aligned offset.  No sample is decoded on the host -- only the metadata blocks are
    parsed.  Returns (buf uint8, files: ctypes array of fe_flac_file with pcm offsets planned, pcm_offsets[n],
    lengths[n], fs, total_bytes).  ``out`` may be a caller-owned (e.g. pinned) uint8 array."""
    lib = _lib.load()
    n = len(paths)
    arr = _paths_array(paths)
    sizes = np.zeros(max(n, 1), dtype=np.int64)
    if lib.aio_file_sizes(arr, n, sizes.ctypes.data_as(C.POINTER(C.c_int64))) != 0:
        bad = int(np.flatnonzero(sizes[:n] < 0)[0])
        _check(_lib.AIO_ERR_IO, paths[bad])
    sizes = sizes[:n]
    if np.any(sizes >= 2 ** 31):
        raise AudioFormatError("file larger than 2 GB")
    if np.any(sizes >= 2 ** 28):
        # the device decoder addresses bits with 32-bit integers: files of 256 MB or more go to the host decoder
        raise UnsupportedStreamError("%s: file of %d MB; the device decoder takes files below 256 MB; use read_audio_batch"
                                     % (paths[int(np.argmax(sizes))], int(sizes.max()) >> 20))
    offsets = np.zeros(n, dtype=np.int64)
    if n > 1:
        np.cumsum((sizes[:-1] + 15) // 16 * 16, out=offsets[1:])
    total = int(offsets[-1] + sizes[-1]) if n else 0
    if out is None:
        out = np.zeros(total + 16, dtype=np.uint8)
    elif out.dtype != np.uint8 or out.size < total or not out.flags.c_contiguous:
        raise ValueError("out must be a C-contiguous uint8 array of at least %d bytes" % total)
    status = np.zeros(max(n, 1), dtype=np.int32)
    rc = lib.aio_read_files(arr, n, int(n_threads), C.c_void_p(out.ctypes.data), offsets.ctypes.data_as(C.POINTER(C.c_int64)),
                            sizes.ctypes.data_as(C.POINTER(C.c_int64)), status.ctypes.data_as(C.POINTER(C.c_int32)))
    if rc != 0:
        bad = int(np.flatnonzero(status)[0])
        _check(int(status[bad]), paths[bad])
    files = (_lib.FeFlacFile * max(n, 1))()
    lays = (_lib.AioFlacLayout * max(n, 1))()
    rc = lib.aio_flac_layouts(C.c_void_p(out.ctypes.data), offsets.ctypes.data_as(C.POINTER(C.c_int64)),
                              sizes.ctypes.data_as(C.POINTER(C.c_int64)), n, int(n_threads), lays,
                              status.ctypes.data_as(C.POINTER(C.c_int32)))
    if rc != 0:
        bad = int(np.flatnonzero(status)[0])
        _check(int(status[bad]), paths[bad])
    L = np.ctypeslib.as_array(lays)[:n]                                   # structured views: no per-file Python work
    F = np.ctypeslib.as_array(files)[:n]
    fs = int(L["sample_rate"][0]) if n else DEFAULT_FS
    if np.any(L["channels"] != 1):
        raise ValueError("%s: mono audio expected" % paths[int(np.flatnonzero(L["channels"] != 1)[0])])
    if np.any(L["sample_rate"] != fs):
        i = int(np.flatnonzero(L["sample_rate"] != fs)[0])
        raise MixedSampleRates("mixed sample rates in one call: %d vs %d (%s)" % (fs, L["sample_rate"][i], paths[i]))
    unsup = ((L["min_block"] != L["max_block"]) | (L["n_samples"] <= 0) | (L["max_block"] % 8 != 0) | (L["max_block"] < 16) |
             (L["bits_per_sample"] > 16) | (L["n_samples"] >= 2 ** 31) | (L["n_samples"] > _MAX_SAMPLES_PER_BYTE * np.maximum(sizes, 1)))
    if np.any(unsup):
        raise UnsupportedStreamError("%s: the device decoder takes fixed-block-size streams (multiple of 8) of at most "
                                     "16 bits with a sample count; use read_audio_batch" % paths[int(np.flatnonzero(unsup)[0])])
    lengths = L["n_samples"].astype(np.int64)
    pcm_offsets, _ = plan_batch(lengths)
    F["byte_offset"], F["n_bytes"], F["first_frame"] = offsets, sizes, L["first_frame"]
    F["n_samples"], F["block_size"], F["bits_per_sample"] = lengths, L["max_block"], L["bits_per_sample"]
    F["pcm_offset"] = pcm_offsets
    return out, files, pcm_offsets, lengths, fs, total


def encode_flac(pcm16, fs=DEFAULT_FS, channels=1):
    """int16 samples (interleaved if channels > 1) -> FLAC bytes."""
    lib = _lib.load()
    pcm16 = np.ascontiguousarray(pcm16, dtype=np.int16).reshape(-1)
    n = pcm16.size // channels
    cap = int(lib.aio_flac_bound(n, channels))
    buf = np.empty(cap, dtype=np.uint8)
    nb = C.c_int64()
    _check(lib.aio_encode_flac(C.c_void_p(pcm16.ctypes.data), n, channels, int(fs), C.c_void_p(buf.ctypes.data), cap,
                               C.byref(nb)), "encode_flac")
    return buf[:nb.value].tobytes()


def write_flac(path, pcm16, fs=DEFAULT_FS, channels=1):
    pcm16 = np.ascontiguousarray(pcm16, dtype=np.int16).reshape(-1)
    _check(_lib.load().aio_write_file(os.fsencode(path), C.c_void_p(pcm16.ctypes.data), pcm16.size // channels,
                                      channels, int(fs), _lib.AIO_FMT_FLAC), path)


def write_audio(path, pcm16, fs=DEFAULT_FS):
    """16-bit mono output, container chosen by extension (SoX does the same)."""
    pcm16 = np.ascontiguousarray(pcm16, dtype=np.int16)
    ext = os.path.splitext(path)[1].lower()
    if ext == ".npy":
        with open(path, "wb") as f:
            np.save(f, pcm16)
        return
    if ext not in _FORMATS:
        raise ValueError("%s: unsupported output format %r (flac, wav, npy)" % (path, ext))
    _check(_lib.load().aio_write_file(os.fsencode(path), C.c_void_p(pcm16.ctypes.data), pcm16.size, 1, int(fs),
                                      _FORMATS[ext]), path)


def write_audio_batch(paths, packed, offsets, lengths, fs=DEFAULT_FS, n_threads=0):
    """Encode utterance i = packed[offsets[i] : offsets[i] + lengths[i]] to paths[i] (flac / wav by
    extension, one format per call) with a pool of host threads."""
    n = len(paths)
    if n == 0:
        return
    exts = {os.path.splitext(p)[1].lower() for p in paths}
    if len(exts) != 1 or next(iter(exts)) not in _FORMATS:
        raise ValueError("one of .flac / .wav per call, got %r" % sorted(exts))
    packed = np.ascontiguousarray(packed, dtype=np.int16)
    offsets = np.ascontiguousarray(offsets, dtype=np.int64)
    lengths = np.ascontiguousarray(lengths, dtype=np.int64)
    status = np.zeros(n, dtype=np.int32)
    rc = _lib.load().aio_write_files(_paths_array(paths), n, int(n_threads), C.c_void_p(packed.ctypes.data),
                                     offsets.ctypes.data_as(C.POINTER(C.c_int64)),
                                     lengths.ctypes.data_as(C.POINTER(C.c_int64)), int(fs), _FORMATS[next(iter(exts))],
                                     status.ctypes.data_as(C.POINTER(C.c_int32)))
    if rc != 0:
        bad = int(np.flatnonzero(status)[0])
        _check(int(status[bad]), paths[bad])
