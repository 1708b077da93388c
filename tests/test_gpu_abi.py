"""The C-ABI itself on the GPU: device-resident buffers, error codes, planning, measurement hooks."""
import ctypes as C
import importlib

import numpy as np
import pytest

from conftest import PKG, assert_close

pytestmark = pytest.mark.gpu


def test_device_resident_run_matches_host_run(pkg, ref):
    import torch
    pcm = pkg.synth.corpus(12, 1.0, 6.0, seed=21)
    fe = pkg.Frontend(pkg.FrontendConfig())
    packed, off, lens = pkg.pack_pcm(pcm)
    host_out, host_off, host_n = fe.run_packed(packed, off, lens)
    d_pcm = torch.from_numpy(packed).cuda()
    launches0 = fe.launch_count()
    d_out, d_off, d_n = fe.run_packed(d_pcm, off, lens, stream=torch.cuda.current_stream().cuda_stream)
    fe.sync()
    torch.cuda.synchronize()
    assert d_out.is_cuda and np.array_equal(d_off, host_off) and np.array_equal(d_n, host_n)
    assert np.array_equal(d_out.cpu().numpy()[:int(host_off[-1])], host_out[:int(host_off[-1])])
    assert fe.launch_count() - launches0 == 3                         # build tiles, K1, K2 (per-utterance CMVN + cube)
    fe.set_profiling(True)
    fe.run_packed(d_pcm, off, lens, out=d_out)
    ms = fe.kernel_ms()
    assert ms["frames_to_statics"] > 0 and ms["cmvn_delta_pack"] > 0 and ms["resample"] == 0
    assert fe.device_bytes() > 0
    for a, p in zip(fe.split(d_out, d_off, d_n), pcm):
        assert_close(a, ref.features_one(p), what="device resident")
    fe.close()


def test_error_codes(pkg):
    _lib = importlib.import_module(PKG + "._lib")
    lib = _lib.load()
    h = C.c_void_p()
    assert lib.fe_create(0, C.byref(h)) == 0
    lens = np.array([16000], np.int64); off = np.array([0], np.int64)
    oo = np.zeros(2, np.int64); nf = np.zeros(1, np.int32)
    p64 = lambda a: a.ctypes.data_as(C.POINTER(C.c_int64))
    assert lib.fe_plan(h, p64(lens), 1, None, p64(oo), nf.ctypes.data_as(C.POINTER(C.c_int32))) == _lib.FE_ERR_STATE
    assert b"fe_configure" in lib.fe_last_error(h)
    assert lib.fe_create(9999, C.byref(C.c_void_p())) == _lib.FE_ERR_INVALID
    fr = importlib.import_module(PKG + ".frontend")
    cfg, keep, _ = fr.make_fe_config(pkg.FrontendConfig())
    assert lib.fe_configure(h, C.byref(cfg)) == 0
    assert lib.fe_plan(h, p64(lens), 1, None, p64(oo), nf.ctypes.data_as(C.POINTER(C.c_int32))) == 0
    assert nf[0] == 97 and oo[1] == (97 * 39 + 3) // 4 * 4
    pcm = np.zeros(16000, np.int16); out = np.zeros(100, np.float32)
    rc = lib.fe_run(h, pcm.ctypes.data, p64(off), p64(lens), 1, None, None, out.ctypes.data, out.size,
                    p64(oo), nf.ctypes.data_as(C.POINTER(C.c_int32)), None)
    assert rc == _lib.FE_ERR_CAPACITY
    bad_off = np.array([3], np.int64)
    out = np.zeros(int(oo[1]), np.float32)
    rc = lib.fe_run(h, pcm.ctypes.data, p64(bad_off), p64(lens), 1, None, None, out.ctypes.data, out.size,
                    p64(oo), nf.ctypes.data_as(C.POINTER(C.c_int32)), None)
    assert rc == _lib.FE_ERR_INVALID and b"16 bytes" in lib.fe_last_error(h)
    cfg.frame_len = 352
    assert lib.fe_configure(h, C.byref(cfg)) == _lib.FE_ERR_INVALID    # unsupported geometry fails loudly
    assert lib.fe_destroy(h) == 0


def test_one_bad_utterance_does_not_poison_the_batch(pkg, ref):
    rng = np.random.default_rng(5)
    good = pkg.synth.corpus(3, 1.0, 2.0, seed=6)
    tiny = (rng.normal(size=450) * 1000).astype(np.int16)              # 0 frames
    fe = pkg.Frontend(pkg.FrontendConfig())
    got = fe.extract([good[0], tiny, good[1], np.zeros(5000, np.int16), good[2]])
    assert got[1].shape == (0, 13, 3)
    for a, p in zip([got[0], got[2], got[4]], good):
        assert_close(a, ref.features_one(p), what="batch with empty utterance")
    assert np.isfinite(got[3]).all()
    fe.close()


def test_chunked_host_pipeline_equals_single_shot(pkg, monkeypatch):
    """host in / host out batches above 2 chunks go through the 2-lane H2D | kernels | D2H pipeline;
    results must be bit-identical to the single-shot path (FE_PIPE_CHUNK_MB shrinks the chunk for the test)."""
    rng = np.random.default_rng(31)
    lens = pkg.synth.durations(300, 2, 15, rng)
    pcm = pkg.synth.noise_corpus_fast(lens, seed=32)                  # ~80 MB of PCM
    packed, off, ln = pkg.pack_pcm(pcm)
    fe = pkg.Frontend(pkg.FrontendConfig())
    ref_out, ref_off, ref_n = fe.run_packed(packed, off, ln)
    fe.close()
    monkeypatch.setenv("FE_PIPE_CHUNK_MB", "4")
    fe2 = pkg.Frontend(pkg.FrontendConfig())
    out, o2, n2 = fe2.run_packed(packed, off, ln)
    assert np.array_equal(o2, ref_off) and np.array_equal(n2, ref_n)
    assert np.array_equal(out[:int(o2[-1])], ref_out[:int(o2[-1])])
    # with perturbation (scratch buffers per lane)
    sp = np.array([(-1, 0, 1)[i % 3] for i in range(len(pcm))], np.int32)
    a, _, _ = fe2.run_packed(packed, off, ln, speed_idx=sp)
    fe2.close()
    monkeypatch.delenv("FE_PIPE_CHUNK_MB")
    fe3 = pkg.Frontend(pkg.FrontendConfig())
    b, ob, _ = fe3.run_packed(packed, off, ln, speed_idx=sp)
    assert np.array_equal(a[:int(ob[-1])], b[:int(ob[-1])])
    fe3.close()
