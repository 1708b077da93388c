"""CPU oracle for speed / volume perturbation  --  TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Reference path: /root/reference/utils/augmentation.py:6-31 (SpeedAugmentation:
``sox.Transformer().speed(s)`` -> ``sox -D -V2 in out speed s``) and :33-56
(VolumeAugmentation: ``vol(g)``, g = np.around(U(lo, hi), 2)).  The arithmetic is
inside the SoX binary (pysox + SoX 14.4.x, requirements.txt:7 UNPINNED, neither
installed here nor vendored), whose ``rate -h`` multi-stage resampler is not
restated.

SoX PARITY UNPINNED.  What is pinned is the *contract*: output length ~ N / speed,
same sample rate, band-limited, re-quantised to 16 bit (the reference writes
16-bit FLAC and re-reads it through soundfile, so the int16 rounding is part of
its path).  This file DEFINES the resampler the CUDA path is checked against:

    y[j] = sum_i x[i] * g(j * down / up - i)                      (x = 0 off the ends)
    g(t) = fc * sinc(fc * t) * I0(beta * sqrt(1 - (t/W)^2)) / I0(beta),   |t| < W
    speed = down / up  (0.9 -> 9/10, 1.1 -> 11/10),  n_out = ceil(N * up / down)
    fc = 0.9425 * min(1, up / down),  W = 64 input samples (128 taps / phase),
    beta = 10.5;   float64;  then rint(y) clipped to int16 (the reference's FLAC round trip).

Round 2 replaced the round-1 spec (32 taps: -6 dB at 0.95 Nyquist, -12 dB at Nyquist) with one that can stand
next to SoX ``rate -h`` (linear phase, no aliasing by default): measured response of THIS filter
(tests/test_oracle.py::test_resampler_frequency_response): flat to +-0.05 dB up to 0.9 x min(Nyquist_in,
Nyquist_out), <= -104 dB from that Nyquist upward.  SoX's own pass-band is 95 % and its rejection ~125 dB; the
16-bit output quantisation floor (-98 dB) is above both stop-bands.  ``speed_perturb`` evaluates the formula with
scipy.signal.upfirdn (same products, polyphase order); ``speed_perturb_direct`` is the literal double loop used to
pin it on small inputs.
"""
from fractions import Fraction
import math

import numpy as np

HALF_WIDTH = 64
TAPS = 2 * HALF_WIDTH
KAISER_BETA = 10.5
ROLLOFF = 0.9425


def speed_ratio(speed):
    """speed -> (up, down) with speed == down / up."""
    fr = Fraction(str(speed)).limit_denominator(1000)
    return fr.denominator, fr.numerator


def out_length(n_in, speed):
    up, down = speed_ratio(speed)
    return -((-n_in * up) // down)


def kernel(t, up, down):
    fc = ROLLOFF * min(1.0, up / down)
    t = np.asarray(t, dtype=np.float64)
    inside = np.abs(t) < HALF_WIDTH
    arg = np.sqrt(np.clip(1.0 - (t / HALF_WIDTH) ** 2, 0.0, None))
    win = np.i0(KAISER_BETA * arg) / np.i0(KAISER_BETA)
    return np.where(inside, fc * np.sinc(fc * t) * win, 0.0)


def polyphase_taps(speed):
    """taps[p, t] = g(p/up + (HALF_WIDTH-1) - t): weight of input sample
    floor(j*down/up) - (HALF_WIDTH-1) + t for an output whose phase is p."""
    up, down = speed_ratio(speed)
    p = np.arange(up, dtype=np.float64)[:, None] / up
    t = np.arange(TAPS, dtype=np.float64)[None, :]
    return kernel(p + (HALF_WIDTH - 1) - t, up, down)


def requantize(y_int_scale):
    """float (already in int16 units) -> int16, round-half-even, saturating."""
    return np.clip(np.rint(y_int_scale), -32768, 32767).astype(np.int16)


def speed_perturb_direct(pcm16, speed):
    """int16 in -> int16 out, the defining formula evaluated literally (float64 gather + dot per output)."""
    pcm16 = np.asarray(pcm16)
    assert pcm16.dtype == np.int16
    up, down = speed_ratio(speed)
    if up == down:
        return pcm16.copy()
    n = pcm16.shape[0]
    n_out = out_length(n, speed)
    taps = polyphase_taps(speed)
    x = np.concatenate((np.zeros(TAPS), pcm16.astype(np.float64), np.zeros(TAPS)))
    y = np.empty(n_out)
    for j0 in range(0, n_out, 1 << 15):                    # chunks bound the (outputs x taps) gather
        j = np.arange(j0, min(n_out, j0 + (1 << 15)), dtype=np.int64)
        pos = j * down
        base = pos // up - (HALF_WIDTH - 1) + TAPS        # index into padded x
        idx = base[:, None] + np.arange(TAPS)[None, :]
        y[j0:j0 + len(j)] = np.einsum("jt,jt->j", x[idx], taps[pos % up])
    return requantize(y)


def prototype_filter(speed):
    """The same kernel as one FIR at the up-sampled rate: h[n] = g((n - C) / up), C = W * up - 1 (odd length 2C + 1,
    symmetric).  y[j] = sum_i x[i] h[j * down - i * up + C]."""
    up, down = speed_ratio(speed)
    c = HALF_WIDTH * up - 1
    return kernel((np.arange(2 * c + 1, dtype=np.float64) - c) / up, up, down)


def speed_perturb(pcm16, speed):
    """int16 in -> int16 out, the defined resampler (float64), evaluated as a polyphase filter."""
    from scipy.signal import upfirdn
    pcm16 = np.asarray(pcm16)
    assert pcm16.dtype == np.int16
    up, down = speed_ratio(speed)
    if up == down:
        return pcm16.copy()
    n_out = out_length(pcm16.shape[0], speed)
    h = prototype_filter(speed)
    c = HALF_WIDTH * up - 1
    lead = (-c) % down                                     # zeros in front so that the centre lands on a kept output
    full = upfirdn(np.concatenate((np.zeros(lead), h)), pcm16.astype(np.float64), up, down)
    first = (c + lead) // down
    y = full[first:first + n_out]
    if y.shape[0] < n_out:
        y = np.concatenate((y, np.zeros(n_out - y.shape[0])))
    return requantize(y)


def volume_perturb(pcm16, gain):
    """SoX ``vol g``: amplitude multiply, clip at full scale, 16-bit output."""
    pcm16 = np.asarray(pcm16)
    assert pcm16.dtype == np.int16
    return requantize(pcm16.astype(np.float64) * float(gain))
