"""The CPU oracle against every known answer the reference pins for this path
 and against the committed golden vectors."""
import hashlib

import numpy as np
import pytest
from scipy.fftpack import dct

from conftest import make_args


def test_frame_count_rule_matches_reference_comments(ref, golden):
    # tfrecord_data_loader.py:78-79: "max dev_featlen: 3262", "max test_featlen: 3493"
    for n, L in zip(golden["frame_count_n"], golden["frame_count_L"]):
        assert max(ref.num_frames(int(n)), 0) == int(L)
    assert ref.num_frames(522320) == 3262 and ref.num_frames(559280) == 3493


def test_filterbank_edges_and_sparsity(ref, golden):
    e40 = ref.filterbank_edges(40, 257, 16000)
    assert e40.tolist() == golden["edges40"].tolist()
    assert e40[0] == 4 and e40[-1] == 128          # 300 Hz floor, (257+1) bin map: top half unused
    fb = ref.filterbanks(40, 257, 16000, 0, 8000)
    assert (fb != 0).sum() == 200 and len(set(np.nonzero(fb)[1])) == 123
    fb80 = ref.filterbanks(80, 257, 16000, 0, 8000)
    assert (fb80 != 0).sum() == 184 and ((fb80 != 0).sum(1) > 0).all()
    assert ref.filterbank_edges(80, 257, 16000).tolist() == golden["edges80"].tolist()
    assert ref.filterbank_edges(40, 257, 16000, bin_map="nfft_plus_one")[-1] >= 255


def test_power_spectrum_parseval_and_energy(ref):
    rng = np.random.default_rng(0)
    fr = rng.normal(size=(7, 400))
    p = ref.power_spectrum(fr, 512)
    assert p.shape == (7, 257)
    full = (np.abs(np.fft.fft(fr, 512, axis=-1)) ** 2).sum(1)
    x0, xn = fr.sum(1), (fr * (-1.0) ** np.arange(400)).sum(1)
    np.testing.assert_allclose(p.sum(1), (full + x0 ** 2 + xn ** 2) / 2 / 512, rtol=1e-12)
    np.testing.assert_allclose(full, 512 * (fr ** 2).sum(1), rtol=1e-12)


def test_dct_is_scipy_ortho(ref, pkg):
    x = np.random.default_rng(1).normal(size=(5, 40))
    want = dct(x, type=2, axis=-1, norm="ortho")[:, :13]
    np.testing.assert_allclose(x @ pkg.tables.dct_ortho(40, 13).T, want, atol=1e-13)
    full = pkg.tables.dct_ortho(40, 40)
    np.testing.assert_allclose(full @ full.T, np.eye(40), atol=1e-13)


def test_cmvn_and_deltas(ref):
    rng = np.random.default_rng(2)
    x = rng.normal(3.0, 2.0, size=(50, 13))
    y = ref.cmvn(x, True)
    np.testing.assert_allclose(y.mean(0), 0, atol=1e-12)
    np.testing.assert_allclose(y.std(0), 1, atol=1e-6)
    d = ref.derivative_extraction(x, 2)                       # as shipped: along the coefficient axis
    k = np.minimum(np.arange(13)[None, :] + np.array([[1], [2]]), 12)
    np.testing.assert_allclose(d, (x[:, k[0]] + 2 * x[:, k[1]]) / 10)
    dt = ref.derivative_extraction(x, 2, "time_regression")
    xp = np.pad(x, ((2, 2), (0, 0)), "edge")
    np.testing.assert_allclose(dt, ((xp[3:-1] - xp[1:-3]) + 2 * (xp[4:] - xp[:-4])) / 10)
    cube = ref.extract_derivative_feature(x)
    assert cube.shape == (50, 13, 3) and np.array_equal(cube[:, :, 0], x)


def test_process_audios_contract(ref, pkg, args):
    pcm = pkg.synth.corpus(3, 0.5, 1.0, seed=5)
    feats, featlen = ref.process_audios(pcm, args)
    assert feats.dtype == object and feats.shape == (3,)
    for f, L, p in zip(feats, featlen, pcm):
        assert f.dtype == np.float32 and f.shape == (L, 13, 3) and L == (len(p) - 400) // 160
    f2, _ = ref.process_audios(pcm, make_args(cmvn=False))
    assert f2[0].shape == (featlen[0], 13)
    fb, _ = ref.process_audios(pcm, make_args(feat_type="fbank", feat_dim=80))
    assert fb[0].shape == (featlen[0], 80, 3)


def test_short_inputs(ref, args):
    rng = np.random.default_rng(3)
    x = (rng.normal(size=470) * 1000).astype(np.int16)          # 400 <= N < 560 -> zero frames
    assert ref.features_one(x).shape == (0, 13, 3)
    with pytest.raises(ValueError):                              # N < 400: speechpy's np.tile raises
        ref.features_one(x[:399])
    one = ref.features_one((rng.normal(size=560) * 1000).astype(np.int16))
    assert one.shape == (1, 13, 3) and np.all(one[:, :, 0] == 0)  # single frame: std = 0 -> 0 / eps


def test_zero_handling_on_silence(ref):
    f = ref.mfcc(np.zeros(2000), 16000, 0.025, 0.010, 13)
    assert np.isfinite(f).all()
    np.testing.assert_allclose(f[:, 0], np.log(np.finfo(float).eps))


def test_golden_vectors(ref, sox, golden):
    for i in range(3):
        p = golden["pcm_%d" % i]
        assert hashlib.sha256(p.tobytes()).digest() == golden["sha_%d" % i].tobytes()
        np.testing.assert_array_equal(ref.features_one(p), golden["mfcc13_%d" % i])
        np.testing.assert_array_equal(ref.features_one(p, cmvn_flag=False), golden["mfcc13_nocmvn_%d" % i])
        np.testing.assert_array_equal(ref.features_one(p, delta_mode="time_regression"), golden["mfcc13_timereg_%d" % i])
        np.testing.assert_array_equal(ref.features_one(p, feat_dim=80, feat_type="fbank"), golden["fbank80_%d" % i])
    np.testing.assert_array_equal(sox.speed_perturb(golden["pcm_0"], 0.9), golden["speed09_pcm_0"])
    np.testing.assert_array_equal(sox.volume_perturb(golden["pcm_0"], 1.23), golden["gain123_pcm_0"])


def test_resampler_frequency_response(sox):
    """the speed-perturbation filter must stand next to SoX `rate -h` (utils/augmentation.py:16-18,28):
    pass-band flat to +-0.1 dB up to 0.9 x min(Nyquist_in, Nyquist_out), >= 100 dB rejection from that Nyquist upward
    (no aliasing for speed 1.1, no imaging for speed 0.9).  Response computed from the polyphase taps themselves."""
    table = {}
    for speed in (0.9, 1.1):
        up, down = sox.speed_ratio(speed)
        taps = sox.polyphase_taps(speed)
        W = sox.HALF_WIDTH
        tau = (np.arange(up)[:, None] / up + (W - 1) - np.arange(2 * W)[None, :]).ravel()    # tap positions, input samples
        h = taps.ravel()
        nyq = min(1.0, up / down)                                                            # in units of the input Nyquist

        def gain_db(f):                                                                      # f relative to the input Nyquist
            return 20 * np.log10(np.abs(np.exp(-1j * np.pi * np.outer(f, tau)) @ h) / up + 1e-300)
        pb = gain_db(np.linspace(0.0, 0.9 * nyq, 500))
        sb = gain_db(np.concatenate((np.linspace(nyq, 3.0, 3000), np.linspace(3.0, float(up), 1500))))
        assert np.abs(pb).max() < 0.1, (speed, pb.min(), pb.max())
        assert sb.max() < -100.0, (speed, sb.max())
        table[speed] = [round(float(gain_db(np.array([f * nyq]))[0]), 2) for f in (0.5, 0.8, 0.9, 0.95, 1.0, 1.1)]
    # DESIGN.md quotes this table (dB at 0.5 / 0.8 / 0.9 / 0.95 / 1.0 / 1.1 of the lower Nyquist)
    assert table[0.9][4] < -100 and table[1.1][4] < -100 and abs(table[0.9][2]) < 0.1 and abs(table[1.1][2]) < 0.1
    # the polyphase evaluation (upfirdn) and the literal formula agree bit for bit on int16 output
    rng = np.random.default_rng(3)
    x = (rng.normal(size=6000) * 6000).astype(np.int16)
    for speed in (0.9, 1.1, 0.8):
        assert np.array_equal(sox.speed_perturb(x, speed), sox.speed_perturb_direct(x, speed))


def test_resampler_oracle_properties(sox):
    n = 16000
    t = np.arange(n) / 16000.0
    x = np.round(8000 * np.sin(2 * np.pi * 1000 * t)).astype(np.int16)
    assert np.array_equal(sox.speed_perturb(x, 1.0), x)
    for s, up, down in ((0.9, 10, 9), (1.1, 10, 11)):
        y = sox.speed_perturb(x, s)
        assert len(y) == -(-n * up // down) == sox.out_length(n, s)
        spec = np.abs(np.fft.rfft(y[2000:2000 + 8192] * np.hanning(8192)))
        peak = np.argmax(spec) * 16000.0 / 8192
        assert abs(peak - 1000 * s) < 4.0                        # pitch scales with speed
        assert abs(np.abs(y[1000:-1000]).max() - 8000) < 8       # unity passband gain (+-0.05 dB ripple)
    dc = np.full(4000, 1000, np.int16)
    assert np.all(np.abs(sox.speed_perturb(dc, 0.9)[100:-100].astype(int) - 1000) <= 1)
    assert sox.volume_perturb(np.array([30000, -30000, 100], np.int16), 1.5).tolist() == [32767, -32768, 150]
