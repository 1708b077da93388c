import importlib
import os
import subprocess
import sys
import types

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

PKG = "automatic-speech-recognition_b200"


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def _has_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if _has_gpu():
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def pkg():
    b = importlib.import_module(PKG + ".build")
    b.build_library()
    return importlib.import_module(PKG)


@pytest.fixture(scope="session")
def ref():
    from oracle import speechpy_ref
    return speechpy_ref


@pytest.fixture(scope="session")
def sox():
    from oracle import sox_ref
    return sox_ref


def make_args(**kw):
    d = dict(frame_step=10, frame_length=25, feat_dim=13, feat_type="mfcc", cmvn=True)
    d.update(kw)
    return types.SimpleNamespace(**d)


@pytest.fixture()
def args():
    return make_args()


@pytest.fixture(scope="session")
def golden():
    path = os.path.join(ROOT, "tests", "golden", "golden_v1.npz")
    return np.load(path)


@pytest.fixture(scope="session")
def host_sim(pkg):
    """CPU replay of the K1 lane-level dataflow (tests/host_sim/host_sim.cpp), built on demand."""
    import ctypes
    src = os.path.join(ROOT, "tests", "host_sim", "host_sim.cpp")
    so = os.path.join(ROOT, "tests", "host_sim", "libhost_sim.so")
    deps = [src] + [os.path.join(ROOT, PKG, "csrc", f) for f in ("fe_core.cuh", "fe_tables.h", "fe_k1t.cuh")]
    if not os.path.exists(so) or any(os.path.getmtime(d) > os.path.getmtime(so) for d in deps):
        subprocess.run(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-o", so, src], check=True)
    lib = ctypes.CDLL(so)
    _lib = importlib.import_module(PKG + "._lib")
    lib.sim_statics.restype = ctypes.c_int
    lib.sim_statics.argtypes = [ctypes.POINTER(_lib.FeConfig), ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p]
    lib.sim_statics_t.restype = ctypes.c_int
    lib.sim_statics_t.argtypes = lib.sim_statics.argtypes
    fr = importlib.import_module(PKG + ".frontend")

    def run(pcm, kernel="k1", **cfg_kw):
        c = fr.FrontendConfig(**cfg_kw)
        cfg, keep, _ = fr.make_fe_config(c)
        pcm = np.ascontiguousarray(pcm)
        L = max((len(pcm) - c.frame_len) // c.hop, 0) if len(pcm) >= c.frame_len else 0
        out = np.zeros((L, c.feat_dim), np.float32)
        fn = lib.sim_statics_t if kernel == "k1t" else lib.sim_statics
        rc = fn(ctypes.byref(cfg), pcm.ctypes.data, len(pcm), out.ctypes.data)
        assert rc == L, (rc, L)
        return out
    return run


def assert_close(got, want, abs_tol=1e-3, rel_tol=1e-4, what=""):
    """north_star tolerance: |d| <= 1e-3 or |d| <= 1e-4 |ref|, element-wise."""
    got = np.asarray(got, dtype=np.float64)
    want = np.asarray(want, dtype=np.float64)
    assert got.shape == want.shape, (what, got.shape, want.shape)
    if got.size == 0:
        return 0.0
    err = np.abs(got - want)
    ok = (err <= abs_tol) | (err <= rel_tol * np.abs(want))
    assert ok.all(), "%s: max abs err %.3g (worst rel %.3g)" % (what, err.max(), (err / (np.abs(want) + 1e-300)).max())
    return float(err.max())
