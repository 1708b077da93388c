"""TF-free drop-in for the consumer of the feature arrays: create_tfrecord.py of the reference.

``create_tfrecords(X, y, filename, num_files, file_start_index)`` keeps the reference's
signature, file naming and record split (/root/reference/create_tfrecord.py:43-97) and writes
the same ``tf.train.Example{feat, shape, token}`` records in TFRecord framing, but through the
native serializer in libasr_frontend.so (include/asr_record_io.h): cubes are read in place
(the views ``process_audios`` returns into the flat ``fe_run`` output buffer are serialised
without a copy or a Python-level ``flatten()``), files are written by a pool of host threads.
``read_tfrecord`` / ``data_parser`` are the parser side (tfrecord_data_loader.py:24-52).
TensorFlow is not needed (and not installed here); the byte format is checked against
``google.protobuf`` in tests/test_tfrecord.py."""
import ctypes as C
import os

import numpy as np

from . import _lib

MAXLEN = 1710                      # create_tfrecord.py:28
NUM_FILE_PER_TFRECORD = 5000       # create_tfrecord.py:29


class RecordError(RuntimeError):
    pass


def _check(rc, what):
    if rc != 0:
        msg = _lib.load().rio_strerror(int(rc))
        raise RecordError("%s: %s (%d)" % (what, msg.decode() if msg else "error", rc))


def _i64p(a):
    return a.ctypes.data_as(C.POINTER(C.c_int64))


def _i32p(a):
    return a.ctypes.data_as(C.POINTER(C.c_int32))


def crc32c(data):
    buf = np.frombuffer(data, dtype=np.uint8)
    return int(_lib.load().rio_crc32c(C.c_void_p(buf.ctypes.data if buf.size else 0), buf.size))


def masked_crc32c(data):
    buf = np.frombuffer(data, dtype=np.uint8)
    return int(_lib.load().rio_masked_crc32c(C.c_void_p(buf.ctypes.data if buf.size else 0), buf.size))


def serialize_example(feat, token):
    """One ``tf.train.Example`` as create_tfrecord.py:83-87 builds it -> bytes."""
    feat = np.ascontiguousarray(feat, dtype=np.float32)
    shape = np.asarray(feat.shape, dtype=np.int64)
    token = np.ascontiguousarray(token, dtype=np.int64).reshape(-1)
    lib = _lib.load()
    size = int(lib.rio_example_size(feat.size, _i64p(shape), shape.size, _i64p(token), token.size))
    out = np.empty(max(size, 1), dtype=np.uint8)
    nb = C.c_int64()
    _check(lib.rio_example_serialize(C.c_void_p(feat.ctypes.data), feat.size, _i64p(shape), shape.size, _i64p(token),
                                     token.size, C.c_void_p(out.ctypes.data), size, C.byref(nb)), "serialize_example")
    return out[:nb.value].tobytes()


def _gather(feats, tokens):
    """Object arrays -> (base pointer, element offsets, n_frames, D, planes, token buffer, offsets, lens).
    Cubes are addressed in place (element offsets relative to the lowest address)."""
    n = len(feats)
    keep = []
    addrs = np.empty(n, dtype=np.int64)
    nfr = np.empty(n, dtype=np.int32)
    D = planes = None
    for i, f in enumerate(feats):
        if not (isinstance(f, np.ndarray) and f.dtype == np.float32 and f.flags.c_contiguous):
            f = np.ascontiguousarray(f, dtype=np.float32)
        if f.ndim not in (2, 3):
            raise ValueError("feature %d: (L, D, 3) or (L, D) expected, got shape %r" % (i, f.shape))
        d, p = int(f.shape[1]), (int(f.shape[2]) if f.ndim == 3 else 0)
        if D is None:
            D, planes = d, p
        elif (d, p) != (D, planes):
            raise ValueError("feature %d: trailing shape %r differs from %r" % (i, f.shape[1:], (D, planes)))
        keep.append(f)
        addrs[i] = f.ctypes.data
        nfr[i] = f.shape[0]
    base = int(addrs.min()) if n else 0
    rel = addrs - base
    if n and np.any(rel % 4):
        raise ValueError("float32 cubes must be 4-byte aligned relative to each other")
    tok_lens = np.fromiter((len(t) for t in tokens), dtype=np.int32, count=n)
    tok_off = np.zeros(n, dtype=np.int64)
    if n > 1:
        np.cumsum(tok_lens[:-1], out=tok_off[1:])
    tok = np.zeros(max(int(tok_lens.sum()), 1), dtype=np.int64)
    for t, o, m in zip(tokens, tok_off, tok_lens):
        tok[o:o + m] = np.asarray(t, dtype=np.int64)
    return keep, base, rel // 4, nfr, (D or 1), (planes or 0), tok, tok_off, tok_lens


def write_packed(paths, file_start, feats_ptr, feat_offsets, n_frames, feat_dim, planes, tokens, token_offsets, token_lens,
                 n_threads=0):
    """Lowest level: records [file_start[f], file_start[f+1]) of a flat float32 buffer -> paths[f]."""
    lib = _lib.load()
    nf = len(paths)
    arr = (C.c_char_p * nf)()
    arr[:] = [os.fsencode(p) for p in paths]
    fs = np.ascontiguousarray(file_start, dtype=np.int32)
    status = np.zeros(max(nf, 1), dtype=np.int32)
    rc = lib.rio_write_tfrecords(arr, nf, _i32p(fs), int(n_threads), C.c_void_p(int(feats_ptr)),
                                 _i64p(np.ascontiguousarray(feat_offsets, dtype=np.int64)),
                                 _i32p(np.ascontiguousarray(n_frames, dtype=np.int32)), int(feat_dim), int(planes),
                                 _i64p(tokens), _i64p(np.ascontiguousarray(token_offsets, dtype=np.int64)),
                                 _i32p(np.ascontiguousarray(token_lens, dtype=np.int32)), _i32p(status))
    if rc != 0:
        bad = np.flatnonzero(status)
        if bad.size:                                  # a per-file failure: name the file
            _check(int(status[bad[0]]), paths[int(bad[0])])
        _check(rc, "rio_write_tfrecords")             # rejected before any file was touched


def create_tfrecords(X, y, filename, num_files=5, file_start_index=1, n_threads=0):
    """Create tfrecords for dataset (create_tfrecord.py:43-97): ``num_files`` files
    ``{filename}-{i + file_start_index}.tfrecord`` with ``len(X) // num_files`` records each, the
    remainder in the last one.  Returns the list of written paths."""
    feats, tokens = X, y
    assert len(feats) == len(tokens)               # "Check if the number of sample points matches."
    n = len(feats)
    if num_files <= 0:
        return []
    per = n // num_files
    starts = [i * per for i in range(num_files)] + [n]
    paths = ["%s-%d.tfrecord" % (filename, i + file_start_index) for i in range(num_files)]
    keep, base, off, nfr, D, planes, tok, tok_off, tok_lens = _gather(feats, tokens)
    write_packed(paths, starts, base, off, nfr, D, planes, tok, tok_off, tok_lens, n_threads)
    del keep
    for i, p in enumerate(paths):
        print("create {}-{}.tfrecord -- contains {} records".format(filename, str(i + file_start_index),
                                                                    starts[i + 1] - starts[i]))
    print("Total records: {}".format(n))
    return paths


def read_tfrecord(path):
    """-> list of (feat ndarray reshaped to its stored shape, token int64 array); every CRC is verified."""
    lib = _lib.load()
    n = int(lib.rio_index_tfrecord(os.fsencode(path), 0, None, None, None))
    if n < 0:
        _check(n, path)
    nfeat = np.zeros(max(n, 1), dtype=np.int64)
    shapes = np.zeros((max(n, 1), 3), dtype=np.int64)
    ntok = np.zeros(max(n, 1), dtype=np.int64)
    rc = int(lib.rio_index_tfrecord(os.fsencode(path), n, _i64p(nfeat), _i64p(shapes), _i64p(ntok)))
    if rc < 0:
        _check(rc, path)
    foff = np.zeros(max(n, 1), dtype=np.int64)
    toff = np.zeros(max(n, 1), dtype=np.int64)
    if n > 1:
        np.cumsum(nfeat[:n - 1], out=foff[1:n])
        np.cumsum(ntok[:n - 1], out=toff[1:n])
    feats = np.empty(max(int(nfeat[:n].sum()), 1), dtype=np.float32)
    toks = np.empty(max(int(ntok[:n].sum()), 1), dtype=np.int64)
    _check(lib.rio_read_tfrecord(os.fsencode(path), n, C.c_void_p(feats.ctypes.data), _i64p(foff), _i64p(toks),
                                 _i64p(toff)), path)
    out = []
    for i in range(n):
        shp = tuple(int(v) for v in shapes[i] if v > 0) if nfeat[i] else tuple(int(v) for v in shapes[i])
        f = feats[foff[i]:foff[i] + nfeat[i]]
        if int(np.prod(shp)) == f.size:
            f = f.reshape(shp)
        out.append((f, toks[toff[i]:toff[i] + ntok[i]]))
    return out


def data_parser(record):
    """tfrecord_data_loader.py:24-52 on one (feat, token) pair from ``read_tfrecord``:
    -> ((feat (L, D, 3) float32, featlen), (token int32, tokenlen))."""
    feat, token = record
    feat = np.asarray(feat, dtype=np.float32).reshape(feat.shape[0], feat.shape[1], 3)
    token = np.asarray(token).astype(np.int32)
    return (feat, int(feat.shape[0])), (token, int(token.shape[0]))
