"""The C-ABI shared library loads here (no GPU) and exports exactly what include/asr_frontend.h declares."""
import ctypes
import importlib
import os
import re

import pytest

from conftest import ROOT, PKG, _has_gpu


def header_symbols(header="asr_frontend.h", prefix="fe_"):
    txt = open(os.path.join(ROOT, "include", header)).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(%s[a-z_0-9]+)\s*\(" % prefix, txt)))


def test_library_exports_every_declared_symbol(pkg):
    _lib = importlib.import_module(PKG + "._lib")
    lib = ctypes.CDLL(pkg.library_path())
    names = header_symbols()
    assert len(names) >= 14
    for n in names:
        assert hasattr(lib, n), "missing export %s" % n
    assert sorted(_lib.SYMBOLS) == names, "ctypes table and header disagree"
    assert _lib.load().fe_abi_version() == _lib.FE_ABI_VERSION
    aio = header_symbols("asr_audio_io.h", "aio_")
    assert len(aio) >= 10
    for n in aio:
        assert hasattr(lib, n), "missing export %s" % n
    assert sorted(_lib.AIO_SYMBOLS) == aio, "ctypes table and asr_audio_io.h disagree"
    assert ctypes.sizeof(_lib.AioInfo) == 24
    rio = header_symbols("asr_record_io.h", "rio_")
    assert len(rio) >= 9
    for n in rio:
        assert hasattr(lib, n), "missing export %s" % n
    assert sorted(_lib.RIO_SYMBOLS) == rio, "ctypes table and asr_record_io.h disagree"


def test_pure_host_entry_points(pkg, golden):
    lib = importlib.import_module(PKG + "._lib").load()
    for n, L in zip(golden["frame_count_n"], golden["frame_count_L"]):
        assert lib.fe_num_frames(int(n), 400, 160) == int(L)
    assert lib.fe_num_frames(399, 400, 160) == 0
    assert lib.fe_resampled_length(16000, 10, 9) == 17778 and lib.fe_resampled_length(16000, 10, 11) == 14546
    assert pkg.num_frames(559280) == 3493


def test_config_struct_matches_header(pkg):
    _lib = importlib.import_module(PKG + "._lib")
    txt = open(os.path.join(ROOT, "include", "asr_frontend.h")).read()
    body = txt[txt.index("typedef struct fe_config {"):txt.index("} fe_config;")]
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
    fields = re.findall(r"\b(?:int32_t|float|const\s+\w+\s*\*)\s*(\w+)\s*;", body)
    assert fields == [f[0] for f in _lib.FeConfig._fields_]


@pytest.mark.skipif(_has_gpu(), reason="checks the no-GPU failure mode")
def test_no_cpu_fallback(pkg):
    with pytest.raises(RuntimeError, match="no CUDA device"):
        pkg.Frontend()
    from conftest import make_args
    import numpy as np
    with pytest.raises(RuntimeError):
        pkg.process_pcm([np.zeros(1000, np.int16)], make_args())


def _struct_fields(header, name):
    txt = open(os.path.join(ROOT, "include", header)).read()
    body = txt[txt.index("typedef struct %s {" % name):]
    body = body[:body.index("}")]
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
    out = []
    for decl in body.split(";"):
        m = re.match(r"\s*(?:typedef struct \w+ \{)?\s*(int32_t|int64_t|float)\s+([\w\s,]+)$", decl.strip())
        if m:
            out += [(m.group(1), n.strip()) for n in m.group(2).split(",")]
    return out


def test_plain_structs_match_headers(pkg):
    """fe_flac_file / aio_info / aio_flac_layout_t: field order, types and sizes of the ctypes mirrors."""
    _lib = importlib.import_module(PKG + "._lib")
    ctype = {"int32_t": ctypes.c_int32, "int64_t": ctypes.c_int64, "float": ctypes.c_float}
    for header, cname, mirror, size in (("asr_frontend.h", "fe_flac_file", _lib.FeFlacFile, 40),
                                        ("asr_audio_io.h", "aio_info", _lib.AioInfo, 24),
                                        ("asr_audio_io.h", "aio_flac_layout_t", _lib.AioFlacLayout, 32)):
        fields = _struct_fields(header, cname)
        assert [n for _, n in fields] == [f[0] for f in mirror._fields_], cname
        assert [ctype[t] for t, _ in fields] == [f[1] for f in mirror._fields_], cname
        assert ctypes.sizeof(mirror) == size, cname
