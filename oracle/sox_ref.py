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
    fc = 0.95 * min(1, up / down),  W = 16 input samples (32 taps / phase),
    beta = 14.769656459379492;   float64;  then rint(y * 32768) clipped to int16.
"""
from fractions import Fraction
import math

import numpy as np

HALF_WIDTH = 16
TAPS = 2 * HALF_WIDTH
KAISER_BETA = 14.769656459379492
ROLLOFF = 0.95


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


def speed_perturb(pcm16, speed):
    """int16 in -> int16 out, the defined resampler (direct form, float64)."""
    pcm16 = np.asarray(pcm16)
    assert pcm16.dtype == np.int16
    up, down = speed_ratio(speed)
    if up == down:
        return pcm16.copy()
    n = pcm16.shape[0]
    n_out = out_length(n, speed)
    taps = polyphase_taps(speed)
    x = np.concatenate((np.zeros(TAPS), pcm16.astype(np.float64), np.zeros(TAPS)))
    j = np.arange(n_out, dtype=np.int64)
    pos = j * down
    base = pos // up - (HALF_WIDTH - 1) + TAPS            # index into padded x
    phase = pos % up
    idx = base[:, None] + np.arange(TAPS)[None, :]
    y = np.einsum("jt,jt->j", x[idx], taps[phase])
    return requantize(y)


def volume_perturb(pcm16, gain):
    """SoX ``vol g``: amplitude multiply, clip at full scale, 16-bit output."""
    pcm16 = np.asarray(pcm16)
    assert pcm16.dtype == np.int16
    return requantize(pcm16.astype(np.float64) * float(gain))
