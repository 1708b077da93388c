"""TF-free TFRecord / tf.train.Example writer behind the reference's create_tfrecords signature
(/root/reference/create_tfrecord.py:43-97) and the parser side (tfrecord_data_loader.py:24-52).

Independent checks: CRC-32C against the RFC 3720 B.4 vectors, the Example bytes against
google.protobuf (message types rebuilt from tensorflow's example.proto / feature.proto field
numbers), the file framing against a pure-Python reader written here.  Byte-exact work."""
import importlib
import os
import struct

import numpy as np
import pytest

from conftest import ROOT, PKG


@pytest.fixture(scope="module")
def tfr(pkg):
    return importlib.import_module(PKG + ".tfrecord")


def _crc32c_py(data):
    c = 0xFFFFFFFF
    for b in data:
        c ^= b
        for _ in range(8):
            c = (c >> 1) ^ 0x82F63B78 if c & 1 else c >> 1
    return c ^ 0xFFFFFFFF


def _mask(c):
    return ((((c >> 15) | (c << 17)) & 0xFFFFFFFF) + 0xA282EAD8) & 0xFFFFFFFF


def test_crc32c_known_answers(tfr):
    assert tfr.crc32c(b"123456789") == 0xE3069283
    assert tfr.crc32c(b"") == 0
    assert tfr.crc32c(bytes(32)) == 0x8A9136AA                       # RFC 3720 B.4
    assert tfr.crc32c(b"\xff" * 32) == 0x62A8AB43
    assert tfr.crc32c(bytes(range(32))) == 0x46DD794E
    assert tfr.crc32c(bytes(range(31, -1, -1))) == 0x113FDB5C
    rng = np.random.default_rng(0)
    for n in (1, 7, 8, 9, 63, 1000):
        d = rng.integers(0, 256, n, dtype=np.uint8).tobytes()
        assert tfr.crc32c(d) == _crc32c_py(d)
        assert tfr.masked_crc32c(d) == _mask(_crc32c_py(d))


def _example_class():
    """tensorflow.Example rebuilt from the field numbers of tensorflow/core/example/{example,feature}.proto."""
    from google.protobuf import descriptor_pb2, descriptor_pool, message_factory
    fd = descriptor_pb2.FileDescriptorProto(name="asr_b200_test_example.proto", package="tensorflow", syntax="proto3")
    F = descriptor_pb2.FieldDescriptorProto

    def msg(name):
        m = fd.message_type.add()
        m.name = name
        return m
    m = msg("BytesList"); m.field.add(name="value", number=1, type=F.TYPE_BYTES, label=F.LABEL_REPEATED)
    m = msg("FloatList"); f = m.field.add(name="value", number=1, type=F.TYPE_FLOAT, label=F.LABEL_REPEATED); f.options.packed = True
    m = msg("Int64List"); f = m.field.add(name="value", number=1, type=F.TYPE_INT64, label=F.LABEL_REPEATED); f.options.packed = True
    m = msg("Feature")
    m.oneof_decl.add(name="kind")
    for i, (n, t) in enumerate((("bytes_list", "BytesList"), ("float_list", "FloatList"), ("int64_list", "Int64List"))):
        m.field.add(name=n, number=i + 1, type=F.TYPE_MESSAGE, label=F.LABEL_OPTIONAL, type_name=".tensorflow." + t, oneof_index=0)
    m = msg("Features")
    e = m.nested_type.add(name="FeatureEntry")
    e.options.map_entry = True
    e.field.add(name="key", number=1, type=F.TYPE_STRING, label=F.LABEL_OPTIONAL)
    e.field.add(name="value", number=2, type=F.TYPE_MESSAGE, label=F.LABEL_OPTIONAL, type_name=".tensorflow.Feature")
    m.field.add(name="feature", number=1, type=F.TYPE_MESSAGE, label=F.LABEL_REPEATED, type_name=".tensorflow.Features.FeatureEntry")
    m = msg("Example")
    m.field.add(name="features", number=1, type=F.TYPE_MESSAGE, label=F.LABEL_OPTIONAL, type_name=".tensorflow.Features")
    pool = descriptor_pool.DescriptorPool()
    pool.Add(fd)
    return message_factory.GetMessageClass(pool.FindMessageTypeByName("tensorflow.Example"))


@pytest.mark.parametrize("shape,ntok", [((7, 13, 3), 5), ((1, 80, 3), 1), ((0, 13, 3), 0), ((4, 13), 3), ((300, 13, 3), 219)])
def test_example_bytes_match_google_protobuf(tfr, shape, ntok):
    Example = _example_class()
    rng = np.random.default_rng(sum(shape))
    feat = rng.standard_normal(shape).astype(np.float32)
    token = rng.integers(0, 5000, ntok).astype(np.int64)
    if ntok > 2:
        token[1] = -3                                                    # negative int64 -> 10-byte varint
        token[2] = 2 ** 40
    mine = tfr.serialize_example(feat, token)
    ex = Example()
    ex.ParseFromString(mine)                                             # a real protobuf parser reads our bytes
    f = ex.features.feature
    assert sorted(f.keys()) == ["feat", "shape", "token"]
    assert np.array_equal(np.asarray(f["feat"].float_list.value, np.float32), feat.reshape(-1))
    assert list(f["shape"].int64_list.value) == list(shape)
    assert list(f["token"].int64_list.value) == token.tolist()
    # ... and builds the same bytes from the same content (create_tfrecord.py:83-87, deterministic map order)
    ex2 = Example()
    ex2.features.feature["feat"].float_list.value.extend(feat.reshape(-1).tolist())
    ex2.features.feature["shape"].int64_list.value.extend(list(shape))
    ex2.features.feature["token"].int64_list.value.extend(token.tolist())
    if feat.size == 0:
        ex2.features.feature["feat"].float_list.SetInParent()
    if ntok == 0:
        ex2.features.feature["token"].int64_list.SetInParent()
    assert ex2.SerializeToString(deterministic=True) == mine


def _read_records_py(path):
    out = []
    with open(path, "rb") as f:
        while True:
            hdr = f.read(12)
            if not hdr:
                return out
            n, lcrc = struct.unpack("<QI", hdr)
            assert _mask(_crc32c_py(hdr[:8])) == lcrc
            data = f.read(n)
            (dcrc,) = struct.unpack("<I", f.read(4))
            assert _mask(_crc32c_py(data)) == dcrc
            out.append(data)


def _cubes(n, D=13, planes=3, seed=0, as_views=True):
    rng = np.random.default_rng(seed)
    lens = rng.integers(1, 40, n)
    if as_views:                                                          # like process_audios: views into one flat buffer
        flat = rng.standard_normal(int(lens.sum()) * D * planes + 8 * n).astype(np.float32)
        cubes, o = [], 0
        for L in lens:
            cubes.append(flat[o:o + L * D * planes].reshape(L, D, planes))
            o += (L * D * planes + 3) // 4 * 4
    else:
        cubes = [rng.standard_normal((L, D, planes)).astype(np.float32) for L in lens]
    X = np.empty(n, dtype=object)
    for i, c in enumerate(cubes):
        X[i] = c
    y = np.empty(n, dtype=object)
    for i in range(n):
        y[i] = rng.integers(0, 5000, rng.integers(1, 30)).tolist()
    return X, y


@pytest.mark.parametrize("as_views", [True, False])
def test_create_tfrecords_signature_split_and_round_trip(tfr, tmp_path, as_views):
    X, y = _cubes(11, as_views=as_views)
    paths = tfr.create_tfrecords(X, y, str(tmp_path / "train-100"), num_files=3, file_start_index=4, n_threads=2)
    assert paths == [str(tmp_path / ("train-100-%d.tfrecord" % i)) for i in (4, 5, 6)]   # create_tfrecord.py:69
    counts, k = [], 0
    Example = _example_class()
    for p in paths:
        recs = tfr.read_tfrecord(p)
        raw = _read_records_py(p)                                        # framing checked by an independent reader
        assert len(raw) == len(recs)
        counts.append(len(recs))
        for (feat, token), data in zip(recs, raw):
            assert feat.shape == X[k].shape and np.array_equal(feat, X[k])
            assert token.tolist() == list(y[k])
            ex = Example(); ex.ParseFromString(data)
            assert list(ex.features.feature["shape"].int64_list.value) == list(X[k].shape)
            (f2, L), (t2, tl) = tfr.data_parser((feat, token))           # tfrecord_data_loader.py:24-52
            assert L == X[k].shape[0] and tl == len(y[k]) and t2.dtype == np.int32 and f2.shape[2] == 3
            k += 1
    assert counts == [3, 3, 5]                                           # remainder in the last file (:72-74)
    with pytest.raises(AssertionError):
        tfr.create_tfrecords(X, y[:5], str(tmp_path / "bad"), 1)         # :58


def test_two_dimensional_features_and_corruption(tfr, tmp_path):
    X, y = _cubes(4, D=80, planes=3, seed=3)
    X2 = np.empty(4, dtype=object)
    for i in range(4):
        X2[i] = np.ascontiguousarray(X[i][:, :, 0])                      # args.cmvn false: (L, D)
    (p,) = tfr.create_tfrecords(X2, y, str(tmp_path / "dev"), 1)
    for (feat, token), want in zip(tfr.read_tfrecord(p), X2):
        assert feat.shape == want.shape and np.array_equal(feat, want)
    raw = bytearray(open(p, "rb").read())
    raw[len(raw) // 2] ^= 1
    open(p, "wb").write(bytes(raw))
    with pytest.raises(tfr.RecordError):
        tfr.read_tfrecord(p)
    open(p, "wb").write(bytes(raw[:len(raw) // 3]))
    with pytest.raises(tfr.RecordError):
        tfr.read_tfrecord(p)
    with pytest.raises(tfr.RecordError):
        tfr.read_tfrecord(str(tmp_path / "missing.tfrecord"))


def test_empty_partition_raises_record_error_not_index_error(tfr, tmp_path):
    """rio_write_tfrecords rejects the call before touching status[] -> RecordError, or empty files."""
    empty = np.empty(0, dtype=object)
    try:
        out = tfr.create_tfrecords(empty, empty, str(tmp_path / "e"), 1)
    except tfr.RecordError:
        return
    assert out is None or all(os.path.exists(p) for p in out)
