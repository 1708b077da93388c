"""Seeded synthetic 16 kHz int16 utterances (there is no corpus and no network).

Broadband on purpose: white + low-passed Gaussian noise + a few amplitude-modulated
harmonics, peak 0.5 full scale.  Narrowband or digitally silent inputs leave mel
bins at the FP32 FFT noise floor where ``log`` amplifies rounding; those are
covered by dedicated edge-case tests, not by the parity sets.
"""
import numpy as np

FS = 16000


def durations(n, lo, hi, rng, dist="uniform"):
    """Utterance lengths in samples.  ``uniform``: U(lo, hi) seconds;
    ``librispeech``: clip(N(12.3, 3.8^2), lo, hi) seconds."""
    if dist == "uniform":
        sec = rng.uniform(lo, hi, size=n)
    elif dist == "librispeech":
        sec = np.clip(rng.normal(12.3, 3.8, size=n), lo, hi)
    else:
        raise ValueError(dist)
    return np.round(sec * FS).astype(np.int64)


def utterance(n_samples, rng):
    """One broadband utterance, int16."""
    n = int(n_samples)
    white = rng.normal(0.0, 0.05, size=n)
    g = rng.normal(0.0, 0.1, size=n + 7)
    c = np.cumsum(np.concatenate(([0.0], g)))
    pinkish = (c[8:] - c[:-8]) / 8.0 * np.sqrt(8.0)
    t = np.arange(n, dtype=np.float64) / FS
    f0 = rng.uniform(80.0, 300.0)
    nh = int(rng.integers(3, 6))
    am = 0.6 + 0.4 * np.sin(2 * np.pi * rng.uniform(1.0, 4.0) * t + rng.uniform(0, 2 * np.pi))
    tone = np.zeros(n)
    for h in range(1, nh + 1):
        tone += (0.3 / h) * np.sin(2 * np.pi * f0 * h * t + rng.uniform(0, 2 * np.pi))
    x = white + pinkish + am * tone
    x *= 0.5 / max(np.max(np.abs(x)), 1e-9)
    x = np.clip(x, -1.0, 1.0 - 2.0 ** -15)
    return np.round(x * 32768.0).astype(np.int16)


def corpus(n, lo, hi, seed, dist="uniform"):
    """List of n int16 utterances, reproducible from ``seed``."""
    rng = np.random.default_rng(seed)
    lens = durations(n, lo, hi, rng, dist)
    return [utterance(int(m), rng) for m in lens]


def noise_corpus_fast(lengths, seed):
    """Cheap broadband int16 noise for throughput runs (content does not change the
    work done: no data-dependent branches apart from exact-zero handling)."""
    rng = np.random.default_rng(seed)
    out = []
    for m in lengths:
        a = rng.integers(-6000, 6000, size=int(m), dtype=np.int16)
        out.append(a)
    return out
