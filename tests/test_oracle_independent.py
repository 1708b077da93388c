"""Independent cross-checks of the two CPU oracles.  The oracles restate third-party code that
cannot run here (speechpy, SoX), so nothing can pin them against the real thing; what CAN be done is to check every
stage against an implementation that shares no code with oracle/: a direct float64 DFT, torch.fft, torch.unfold,
scipy.fft / torchaudio's DCT matrix, torchaudio's mel-scale helpers, scipy.signal.resample_poly and
torchaudio.functional.resample, and 10-line numpy re-derivations written from SURVEY.md Appendix A / B."""
import numpy as np
import pytest


@pytest.fixture(scope="module")
def frames():
    rng = np.random.default_rng(11)
    return rng.normal(size=(7, 400))


def test_power_spectrum_vs_direct_dft_and_torch(ref, frames):
    """Appendix A.3: P = |rfft(frame, 512)|^2 / 512 -- against the DFT sum itself and against torch's FFT."""
    import torch
    n = np.arange(400)[None, :]
    k = np.arange(257)[:, None]
    dft = np.exp(-2j * np.pi * k * n / 512.0)                       # (257, 400): zero padding = the missing columns
    want = np.abs(frames @ dft.T) ** 2 / 512.0
    got = ref.power_spectrum(frames, 512)
    assert got.shape == (7, 257)
    np.testing.assert_allclose(got, want, rtol=1e-10, atol=1e-12)
    t = torch.fft.rfft(torch.from_numpy(frames), n=512, dim=-1)
    np.testing.assert_allclose(got, (t.abs() ** 2 / 512.0).numpy(), rtol=1e-10, atol=1e-12)
    # Parseval, the identity the kernel's frame energy uses: sum_257 P = sum x^2 / 2 + (X0^2 + X256^2) / 1024
    x0 = frames.sum(1)
    x256 = (frames * (-1.0) ** np.arange(400)).sum(1)
    np.testing.assert_allclose(got.sum(1), (frames ** 2).sum(1) / 2 + (x0 ** 2 + x256 ** 2) / 1024, rtol=1e-12)


def test_framing_vs_torch_unfold(ref):
    """Appendix A.2: frames i*160 .. i*160+400, L = floor((N-400)/160) -- torch.unfold yields the textbook L + 1 windows,
    the reference drops the last one (tfrecord_data_loader.py:78-79: 522320 -> 3262, 559280 -> 3493)."""
    import torch
    rng = np.random.default_rng(12)
    for n in (560, 720, 16000, 21923, 400 + 160 * 32 + 159):
        x = rng.normal(size=n)
        u = torch.from_numpy(x).unfold(0, 400, 160).numpy()
        got = ref.stack_frames(x, 16000, 0.025, 0.010)
        assert got.shape[0] == u.shape[0] - 1 == (n - 400) // 160
        assert np.array_equal(got, u[:-1])
    assert ref.num_frames(522320) == 3262 and ref.num_frames(559280) == 3493


def test_dct_vs_scipy_fft_and_torchaudio(ref, pkg):
    """Appendix A.5: scipy.fftpack.dct(type=2, norm='ortho')[:, :13] -- against scipy.fft (pocketfft, a different
    module), torchaudio's DCT matrix and the closed form."""
    import scipy.fft
    import torchaudio.functional as F
    rng = np.random.default_rng(13)
    x = rng.normal(size=(9, 40))
    from scipy.fftpack import dct as fp_dct
    want = fp_dct(x, type=2, axis=-1, norm="ortho")[:, :13]
    np.testing.assert_allclose(scipy.fft.dct(x, type=2, axis=-1, norm="ortho")[:, :13], want, atol=1e-12)
    m = F.create_dct(13, 40, norm="ortho").double().numpy()           # (n_mels, n_mfcc)
    np.testing.assert_allclose(x @ m, want, atol=1e-5)                 # torchaudio builds the matrix in float32
    k = np.arange(13)[:, None]
    n = np.arange(40)[None, :]
    closed = np.cos(np.pi * k * (2 * n + 1) / 80.0) * np.where(k == 0, np.sqrt(1 / 40.0), np.sqrt(2 / 40.0))
    np.testing.assert_allclose(x @ closed.T, want, atol=1e-12)
    np.testing.assert_allclose(pkg.tables.dct_ortho(40, 13), closed, atol=1e-15)


def test_mel_edges_vs_torchaudio_melscale(ref):
    """Appendix A.4: mel = 1127 ln(1 + f/700) (HTK), 42 points linear in mel between 300 Hz and 8 kHz, bin =
    floor(258 f / fs).  torchaudio's helpers use 2595 log10(1 + f/700): the constant cancels in hz -> mel -> hz."""
    from torchaudio.functional.functional import _hz_to_mel, _mel_to_hz
    import torch
    m = torch.linspace(_hz_to_mel(300.0, "htk"), _hz_to_mel(8000.0, "htk"), 42, dtype=torch.float64)
    hz = _mel_to_hz(m, "htk").numpy()
    edges = np.floor(258 * hz / 16000.0).astype(int)
    got = ref.filterbank_edges(40, 257, 16000)
    # float32 round-off inside torchaudio's linspace can move an edge that sits exactly on an integer: allow none here,
    # and fall back to the closed form if torchaudio ever disagrees by its own rounding
    closed = 700.0 * (np.exp(np.linspace(np.log(1 + 300 / 700.0), np.log(1 + 8000 / 700.0), 42)) - 1)
    assert np.array_equal(got, np.floor(258 * closed / 16000.0).astype(int))
    assert np.abs(hz - closed).max() < 1e-6 * 8000 and (edges != got).sum() <= 1
    assert got.tolist() == [4, 5, 6, 7, 8, 9, 10, 12, 13, 14, 16, 17, 19, 20, 22, 24, 26, 28, 30, 32, 35, 37, 40, 42, 45, 49,
                            52, 55, 59, 63, 67, 71, 75, 80, 85, 90, 96, 102, 108, 114, 121, 128]


def test_filterbank_triangles_rederived(ref):
    """Appendix A.4 triangle rule written out per bin: rising (k-l)/(m-l) on l < k <= m, falling (r-k)/(r-m) on
    m <= k < r (the falling branch wins at k == m), zero at the ends."""
    e = ref.filterbank_edges(40, 257, 16000)
    fb = ref.filterbanks(40, 257, 16000)
    mine = np.zeros((40, 257))
    for i in range(40):
        l, m, r = int(e[i]), int(e[i + 1]), int(e[i + 2])
        for k in range(l, r + 1):
            w = 0.0
            if l < k <= m:
                w = (k - l) / (m - l)
            if m <= k < r:
                w = (r - k) / (r - m)
            mine[i, k] = w
    assert np.array_equal(fb, mine)
    assert int((fb != 0).sum()) == 200 and len(np.unique(np.nonzero(fb)[1])) == 123
    fb80 = ref.filterbanks(80, 257, 16000)
    assert int((fb80 != 0).sum()) == 184 and (fb80.sum(1) > 0).all()


def test_cmvn_and_deltas_rederived(ref):
    """Appendix A.6 / A.7 in ten lines of numpy, no oracle code."""
    rng = np.random.default_rng(14)
    x = rng.normal(size=(57, 13)) * rng.uniform(0.1, 5, size=13) + rng.normal(size=13)
    mu, sd = x.mean(0), x.std(0)                                       # population std (ddof = 0)
    cm = (x - mu) / (sd + 2.0 ** -30)
    np.testing.assert_allclose(ref.cmvn(x, True), cm, rtol=1e-12, atol=1e-14)
    D = 13
    k = np.arange(D)
    shipped = lambda f: (f[:, np.minimum(k + 1, D - 1)] + 2 * f[:, np.minimum(k + 2, D - 1)]) / 10.0
    d1 = shipped(cm)
    cube = np.stack((cm, d1, shipped(d1)), axis=2)
    np.testing.assert_allclose(ref.extract_derivative_feature(cm), cube, rtol=1e-12, atol=1e-14)
    L = len(cm)
    t = np.arange(L)
    reg = lambda f: sum(n * (f[np.minimum(t + n, L - 1)] - f[np.maximum(t - n, 0)]) for n in (1, 2)) / 10.0
    r1 = reg(cm)
    np.testing.assert_allclose(ref.extract_derivative_feature(cm, "time_regression"), np.stack((cm, r1, reg(r1)), axis=2),
                               rtol=1e-12, atol=1e-14)


def test_mfcc_chain_rederived(ref):
    """The whole static chain once more, straight from Appendix A, on one utterance (float64, no oracle functions)."""
    from scipy.fft import dct, rfft
    rng = np.random.default_rng(15)
    x = np.round(rng.normal(size=16000 + 123) * 2000).astype(np.int16)
    sig = x.astype(np.float64) / 32768.0
    L = (len(sig) - 400) // 160
    fr = np.stack([sig[i * 160:i * 160 + 400] for i in range(L)])
    P = np.abs(rfft(fr, 512)) ** 2 / 512
    en = P.sum(1)
    en[en == 0] = np.finfo(float).eps
    hz = 700.0 * (np.exp(np.linspace(np.log(1 + 300 / 700.0), np.log(1 + 8000 / 700.0), 42)) - 1)
    e = np.floor(258 * hz / 16000).astype(int)
    fb = np.zeros((40, 257))
    for i in range(40):
        l, m, r = e[i:i + 3]
        for kk in range(l, r + 1):
            if l < kk <= m:
                fb[i, kk] = (kk - l) / (m - l)
            if m <= kk < r:
                fb[i, kk] = (r - kk) / (r - m)
    mel = P @ fb.T
    mel[mel == 0] = np.finfo(float).eps
    c = dct(np.log(mel), type=2, axis=-1, norm="ortho")[:, :13]
    c[:, 0] = np.log(en)
    np.testing.assert_allclose(ref.mfcc(sig, 16000, 0.025, 0.010, 13), c, rtol=1e-9, atol=1e-10)


def test_resampler_vs_scipy_resample_poly(sox):
    """Appendix B: the defined resampler is `rational polyphase FIR, zero-phase alignment as scipy.signal.resample_poly`.
    Fed the same prototype filter, scipy's implementation must give the same samples (its `window` argument accepts
    the FIR itself; scipy multiplies it by `up` -- interpolation gain -- which our per-phase taps already carry)."""
    from scipy.signal import resample_poly
    rng = np.random.default_rng(16)
    x = (rng.normal(size=9000) * 5000).astype(np.int16)
    for speed in (0.9, 1.1):
        up, down = sox.speed_ratio(speed)
        h = sox.prototype_filter(speed)                             # odd length, symmetric: scipy centres it the same way
        y = resample_poly(x.astype(np.float64), up, down, window=h / up)
        want = sox.requantize(y)
        got = sox.speed_perturb(x, speed)
        assert len(got) == len(want) == -(-len(x) * up // down)
        assert np.array_equal(got, want)


def test_resampler_vs_torchaudio_in_band(sox):
    """A different filter design altogether (torchaudio's windowed sinc): in the pass-band both must reproduce the same
    band-limited signal -- checks ratio direction, alignment (zero phase) and unity gain, not the transition band."""
    import torch
    import torchaudio.functional as F
    t = np.arange(32000) / 16000.0
    x = 6000 * (np.sin(2 * np.pi * 440 * t) + 0.5 * np.sin(2 * np.pi * 2500 * t + 1.0) + 0.25 * np.sin(2 * np.pi * 5200 * t + 2.0))
    x16 = np.round(x).astype(np.int16)
    for speed in (0.9, 1.1):
        up, down = sox.speed_ratio(speed)
        got = sox.speed_perturb(x16, speed).astype(np.float64)
        ta = F.resample(torch.from_numpy(x16.astype(np.float64)), down, up, lowpass_filter_width=64, rolloff=0.94,
                        resampling_method="sinc_interp_kaiser", beta=10.5).numpy()
        n = min(len(got), len(ta))
        d = np.abs(got[2000:n - 2000] - ta[2000:n - 2000])
        assert d.max() < 12.0, (speed, d.max())                       # < 0.2 % of the 6000-count amplitude (+ 1 LSB rounding)
