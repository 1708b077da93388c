"""Host-built constant tables uploaded once per configuration (``fe_configure``).

Everything here is computed in float64 on the host and rounded once to float32,
so the device never evaluates a ``floor``/``exp``/``cos`` whose last ulp could
move a filter edge.  The formulas are the ones speechpy applies on every call
from /root/reference/preprocess.py:72-82 (filterbank rebuilt per utterance
there; built once here):

* mel filterbank  -> CSR rows (first bin, run of weights) of the triangular
  filters; as shipped the bin map is floor((257 + 1) * f / fs) with a 300 Hz
  floor on the low edge, so only bins ~5..128 carry weight.
* DCT-II, norm='ortho', first ``D`` rows (scipy.fftpack.dct as used by
  speechpy.feature.mfcc).
* FFT twiddles for the 16x16 decomposition of the 256-point complex FFT that
  carries the 512-point real FFT, and the W_512^k post-pass factors.
* polyphase taps of the speed-perturbation resampler (spec in DESIGN.md).
"""
from fractions import Fraction

import numpy as np

NFFT = 512
NBINS = NFFT // 2 + 1
FLOAT64_EPS = float(np.finfo(np.float64).eps)

# Speed-perturbation resampler (DESIGN.md "K0"): Kaiser-windowed sinc, 128 taps per phase.  Measured response
# (tests/test_oracle.py::test_resampler_frequency_response): +-0.05 dB up to 0.9 x min(Nyquist_in, Nyquist_out),
# <= -104 dB from that Nyquist upward (no aliasing / imaging above it) -- below the 16-bit quantisation floor.
RESAMPLE_HALF_WIDTH = 64
RESAMPLE_TAPS = 2 * RESAMPLE_HALF_WIDTH
RESAMPLE_BETA = 10.5
RESAMPLE_ROLLOFF = 0.9425


def hz_to_mel(f):
    return 1127 * np.log(1 + f / 700.)


def mel_to_hz(m):
    return 700 * (np.exp(m / 1127.0) - 1)


def mel_edges(num_filters, fs, low_freq=0, high_freq=None, bin_map="coefficients_plus_one",
              coefficients=NBINS, fft_length=NFFT):
    """Integer FFT-bin edges (num_filters + 2) of the triangular mel filters."""
    high_freq = high_freq or fs / 2
    low_freq = low_freq or 300
    if high_freq > fs / 2:
        raise ValueError("High frequency cannot be greater than half of the sampling frequency!")
    if low_freq < 0:
        raise ValueError("low frequency cannot be less than zero!")
    hertz = mel_to_hz(np.linspace(hz_to_mel(low_freq), hz_to_mel(high_freq), num_filters + 2))
    if bin_map == "coefficients_plus_one":
        scale = coefficients + 1
    elif bin_map == "nfft_plus_one":
        scale = fft_length + 1
    else:
        raise ValueError("unknown bin_map %r" % (bin_map,))
    return np.floor(scale * hertz / fs).astype(np.int64)


def mel_filterbank_dense(num_filters, fs, low_freq=0, high_freq=None,
                         bin_map="coefficients_plus_one", coefficients=NBINS):
    """(num_filters, coefficients) float64 table, triangle by triangle."""
    edges = mel_edges(num_filters, fs, low_freq, high_freq, bin_map, coefficients)
    fb = np.zeros((num_filters, coefficients), dtype=np.float64)
    for m in range(num_filters):
        left, mid, right = (int(e) for e in edges[m:m + 3])
        for k in range(left, min(right, coefficients - 1) + 1):
            w = 0.0
            if left < k <= mid:
                w = (k - left) / (mid - left)
            if mid <= k < right:
                w = (right - k) / (right - mid)
            fb[m, k] = w
    return fb


def filterbank_csr(fb):
    """Dense table -> (row_start[nf+1] i32, first_bin[nf] i32, weights f64).
    Each filter's non-zeros are one contiguous run of bins, so a row is
    (first_bin, weights[row_start[m]:row_start[m+1]])."""
    nf = fb.shape[0]
    row_start = np.zeros(nf + 1, dtype=np.int32)
    first_bin = np.zeros(nf, dtype=np.int32)
    weights = []
    for m in range(nf):
        nz = np.nonzero(fb[m])[0]
        if nz.size:
            lo, hi = int(nz[0]), int(nz[-1])
            first_bin[m] = lo
            weights.extend(fb[m, lo:hi + 1].tolist())
        row_start[m + 1] = len(weights)
    return row_start, first_bin, np.asarray(weights, dtype=np.float64)


def dct_ortho(num_filters, num_cepstral):
    """First ``num_cepstral`` rows of the orthonormal DCT-II of size num_filters."""
    n = np.arange(num_filters, dtype=np.float64)[None, :]
    k = np.arange(num_cepstral, dtype=np.float64)[:, None]
    mat = np.cos(np.pi * k * (2 * n + 1) / (2 * num_filters))
    mat[0, :] *= np.sqrt(1.0 / num_filters)
    mat[1:, :] *= np.sqrt(2.0 / num_filters)
    return mat


def twiddles_256():
    """W_256^(j*k) for j,k in 0..15 as (cos, -sin) pairs, shape (16, 16, 2)."""
    j = np.arange(16)[:, None]
    k = np.arange(16)[None, :]
    ang = 2.0 * np.pi * ((j * k) % 256) / 256.0
    return np.stack((np.cos(ang), -np.sin(ang)), axis=-1)


def twiddles_512():
    """(cos, sin) of 2*pi*k/512 for k = 0..256, shape (257, 2)."""
    ang = 2.0 * np.pi * np.arange(NBINS) / NFFT
    return np.stack((np.cos(ang), np.sin(ang)), axis=-1)


def speed_ratio(speed):
    """speed -> (up, down), speed == down / up (0.9 -> (10, 9))."""
    fr = Fraction(str(speed)).limit_denominator(1000)
    return fr.denominator, fr.numerator


def resampled_length(n_in, speed):
    up, down = speed_ratio(speed)
    return -((-int(n_in) * up) // down)


def resampler_taps(speed):
    """(up, RESAMPLE_TAPS) float64 polyphase taps of the Kaiser-windowed sinc
    y[j] = sum_i x[i] g(j*down/up - i); row p serves outputs with (j*down)%up == p,
    column t weighs input floor(j*down/up) - (RESAMPLE_HALF_WIDTH - 1) + t."""
    up, down = speed_ratio(speed)
    fc = RESAMPLE_ROLLOFF * min(1.0, up / down)
    p = np.arange(up, dtype=np.float64)[:, None] / up
    t = np.arange(RESAMPLE_TAPS, dtype=np.float64)[None, :]
    tau = p + (RESAMPLE_HALF_WIDTH - 1) - t
    inside = np.abs(tau) < RESAMPLE_HALF_WIDTH
    arg = np.sqrt(np.clip(1.0 - (tau / RESAMPLE_HALF_WIDTH) ** 2, 0.0, None))
    win = np.i0(RESAMPLE_BETA * arg) / np.i0(RESAMPLE_BETA)
    return np.where(inside, fc * np.sinc(fc * tau) * win, 0.0)
