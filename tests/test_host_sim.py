"""The exact lane-level dataflow of the frames->statics kernel (stage A / exchange /
stage B / real-FFT split / mel / log / DCT), replayed on the CPU by tests/host_sim and
checked against the float64 oracle.  Catches index-math, swizzle and codelet errors
without a GPU; the GPU parity tests (-m gpu) check the real thing."""
import numpy as np
import pytest


@pytest.fixture(scope="module")
def utts(pkg):
    u = pkg.synth.corpus(3, 1.0, 3.0, seed=99)
    u.append(u[0][:400 + 160 * 5 + 17])        # partial 4-frame group
    u.append(u[1][:400 + 160 * 2 + 3])
    return u


def test_mfcc_statics(host_sim, ref, utts):
    for u in utts:
        want = ref.mfcc(ref.pcm_to_float(u), 16000, 0.025, 0.010, 13)
        assert np.abs(host_sim(u) - want).max() < 2e-5


def test_fbank_statics_relative(host_sim, ref, utts):
    for u in utts:
        want, _ = ref.mfe(ref.pcm_to_float(u), 16000, 0.025, 0.010, 80)
        got = host_sim(u, feat_type="fbank", feat_dim=80)
        assert np.max(np.abs(got - want) / np.abs(want)) < 1e-4


@pytest.mark.parametrize("kw", [dict(bin_map="nfft_plus_one"), dict(feat_dim=40), dict(feat_dim=7, num_filters=26),
                                dict(dc_elimination=False), dict(window=np.hamming(400))])
def test_mfcc_variants(host_sim, ref, utts, kw):
    u = utts[0]
    okw = dict(num_cepstral=kw.get("feat_dim", 13), num_filters=kw.get("num_filters", 40),
               bin_map=kw.get("bin_map", "coefficients_plus_one"), window=kw.get("window"),
               dc_elimination=kw.get("dc_elimination", True))
    want = ref.mfcc(ref.pcm_to_float(u), 16000, 0.025, 0.010, **okw)
    assert np.abs(host_sim(u, **kw) - want).max() < 2e-5


def test_float_pcm_and_log_fbank(host_sim, ref, utts):
    u = utts[1]
    f = ref.pcm_to_float(u)
    assert np.abs(host_sim(f.astype(np.float32), pcm_dtype="float32") - ref.mfcc(f, 16000, 0.025, 0.010, 13)).max() < 2e-5
    want = np.log(ref.mfe(f, 16000, 0.025, 0.010, 23)[0])
    assert np.abs(host_sim(u, feat_type="fbank", feat_dim=23, fbank_log=True) - want).max() < 2e-5


def test_digital_silence_hits_eps_exactly(host_sim, ref):
    z = np.zeros(3000, np.int16)
    got = host_sim(z)
    want = ref.mfcc(ref.pcm_to_float(z), 16000, 0.025, 0.010, 13)
    assert np.abs(got - want).max() < 1e-5       # log(2.22e-16) on c0, ~0 elsewhere


def test_k1t_lane_per_frame_dataflow(host_sim, ref, utts):
    """fe_k1t.cuh (lane = frame, exchange in tensor memory) replayed with the exchange as an array: reads of
    unwritten exchange words come back NaN, so a wrong column / row index cannot hide."""
    for u in utts:
        want = ref.mfcc(ref.pcm_to_float(u), 16000, 0.025, 0.010, 13)
        got = host_sim(u, kernel="k1t")
        assert np.isfinite(got).all()
        assert np.abs(got - want).max() < 2e-5
        assert np.abs(got - host_sim(u)).max() < 2e-5                 # and it agrees with the K1 dataflow
    for u in utts[:2]:
        want, _ = ref.mfe(ref.pcm_to_float(u), 16000, 0.025, 0.010, 80)
        got = host_sim(u, kernel="k1t", feat_type="fbank", feat_dim=80)
        assert np.max(np.abs(got - want) / np.abs(want)) < 1e-4
    z = np.zeros(3000, np.int16)
    want = ref.mfcc(ref.pcm_to_float(z), 16000, 0.025, 0.010, 13)
    assert np.abs(host_sim(z, kernel="k1t") - want).max() < 1e-5       # digital silence: exact eps path
