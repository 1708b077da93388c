"""BASELINE.json configs[0..2] at their FULL sizes, every utterance against the CPU oracle:
1 000 utterances U(2,15) s mfcc-39;  2 000 LibriSpeech-length utterances fbank-80 (linear as shipped, and log);
1 000 utterances x speeds 0.9 / 1.0 / 1.1.  The oracle runs on all host cores (tests/parity_util.py); the
per-config error summary is written to gpurun_out/r02_parity_*.json (committed copies: profiles/r02_parity.json)."""
import numpy as np
import pytest

import parity_util as pu
from conftest import make_args

pytestmark = pytest.mark.gpu


def _run(pkg, pcm, args, **sw):
    feats, featlen = pkg.process_pcm(pcm, args, **sw)
    return list(feats), featlen


def test_config1_every_utterance(pkg):
    rng = np.random.default_rng(1234)
    lens = pkg.synth.durations(1000, 2, 15, rng)
    pcm = pu.gen_corpus(lens, 1234)
    feats, featlen = _run(pkg, pcm, make_args())
    assert featlen == [int((n - 400) // 160) for n in lens]                   # frame counts bit-exact
    s = pu.oracle_errors(pcm, feats, dict(feat_dim=13, feat_type="mfcc"))
    s["config"] = "configs[0]: 1000 utterances U(2,15) s, MFCC-39 (13+d+dd) + per-utterance CMVN"
    pu.record("config1", s)
    assert s["shape_mismatches"] == 0 and s["elements_out_of_tolerance"] == 0, s
    assert s["max_abs_err"] < 1e-3, s


@pytest.mark.parametrize("log", [False, True])
def test_config2_every_utterance(pkg, log):
    rng = np.random.default_rng(2345)
    lens = pkg.synth.durations(2000, 2, 35, rng, "librispeech")
    pcm = pu.gen_corpus(lens, 2345)
    feats, featlen = _run(pkg, pcm, make_args(feat_type="fbank", feat_dim=80), fbank_log=log)
    assert featlen == [int((n - 400) // 160) for n in lens]
    s = pu.oracle_errors(pcm, feats, dict(feat_dim=80, feat_type="fbank", fbank_log=log))
    s["config"] = "configs[1]: 2000 utterances clip(N(12.3,3.8^2),2,35) s, fbank-80 %s + CMVN" % ("log" if log else "linear (mfe, as shipped)")
    pu.record("config2_%s" % ("log" if log else "linear"), s)
    assert s["shape_mismatches"] == 0 and s["elements_out_of_tolerance"] == 0, s


def test_config3_every_utterance_three_speeds(pkg):
    rng = np.random.default_rng(3456)
    base = pkg.synth.durations(1000, 2, 15, rng)
    one = pu.gen_corpus(base, 3456)
    pcm = [p for p in one for _ in range(3)]
    speeds = [0.9, 1.0, 1.1] * 1000
    fe = pkg.Frontend(pkg.FrontendConfig())
    feats = fe.extract(pcm, speeds=speeds)                                    # resampler in front of the framing, one batch
    gpu_pcm = fe.perturb(pcm, speeds=speeds)                                  # what K0 hands K1 (int16)
    fe.close()
    s = pu.oracle_errors(pcm, feats, dict(feat_dim=13, feat_type="mfcc"), speeds=speeds, gpu_pcm=gpu_pcm)
    s["config"] = "configs[2]: 1000 utterances U(2,15) s x speeds 0.9/1.0/1.1, MFCC-39 + CMVN (oracle: oracle/sox_ref.py resampler)"
    pu.record("config3", s)
    assert s["shape_mismatches"] == 0, s                                      # resampled lengths -> frame counts bit-exact
    # integer stage: the resampler re-quantises to int16 like the reference's FLAC round trip; FP32 accumulation of 128
    # taps against the FP64 oracle can land on the other side of a .5 boundary: +-1 LSB on < 0.2 % of the samples
    r = s["resampler_int16_vs_oracle"]
    assert r["length_mismatches"] == 0 and r["max_abs_lsb"] <= 1 and r["fraction_differing"] < 2e-3, r
    # feature chain given the SAME resampled samples: the north-star tolerance, every element
    g = s["given_the_kernels_resampled_pcm"]
    assert g["shape_mismatches"] == 0 and g["elements_out_of_tolerance"] == 0, g
    # end to end against the oracle's own resampled samples: a +-1 LSB difference in a low-energy mel band is amplified
    # by the log (the float64 oracle moves by the same amount when ITS input changes by one LSB): bounded, rare
    assert s["elements_out_of_tolerance"] < 1e-5 * s["elements"] and s["median_utterance_max_abs_err"] < 1e-3, s


def _tilted_noise(n, tilt_db, rng):
    """Gaussian noise whose power spectrum falls by ``tilt_db`` between 0 and 4 kHz (the band the as-shipped
    filterbank reads), flat above; float32 so that the in-frame range is not capped by 16-bit quantisation."""
    w = np.fft.rfft(rng.normal(size=n))
    f = np.fft.rfftfreq(n, 1.0 / 16000)
    w *= 10.0 ** (-(tilt_db * np.minimum(f, 4000.0) / 4000.0) / 20.0)
    x = np.fft.irfft(w, n)
    return (x * (0.3 / np.abs(x).max())).astype(np.float32)


def _sweep(pkg, make, steps, **sw):
    from oracle import speechpy_ref as R
    rows, first_fail = [], None
    args = make_args()
    for step in steps:
        pcm = [make(step) for _ in range(6)]
        feats, _ = pkg.process_pcm(pcm, args, **sw)
        worst, units, bad, total, rng_db = 0.0, 0.0, 0, 0, []
        for x, f in zip(pcm, feats):
            x64 = R.pcm_to_float(x) if x.dtype == np.int16 else x.astype(np.float64)
            want = R.features_one(x if x.dtype == np.int16 else x64, **sw).astype(np.float64)
            err = np.abs(f.astype(np.float64) - want)
            u = np.minimum(err / pu.ABS_TOL, err / (pu.REL_TOL * np.abs(want) + 1e-300))
            worst, units = max(worst, float(err.max())), max(units, float(u.max()))
            bad += int((u > 1).sum()); total += int(u.size)
            mel, _ = R.mfe(x64, 16000, 0.025, 0.010, 40, window=sw.get("window"))
            rng_db.append(float(np.median(10 * np.log10(mel.max(1) / mel.min(1)))))
        rows.append({"step_db": step, "median_in_frame_mel_range_db": float(np.median(rng_db)), "max_abs_err": worst,
                     "max_tolerance_units": units, "elements_out_of_tolerance": bad, "elements": total})
        if first_fail is None and bad:
            first_fail = step
    return rows, first_fail


def test_fp32_error_vs_in_frame_dynamic_range(pkg, ref):
    """the kernels are FP32 and the log is lg2.approx; the oracle is FP64.  The error is a function
    of how far below the frame's strongest bin a mel band sits (FP32 FFT round-off is relative to the strongest bin).
    Three sweeps, each recording the error per step and the first step that leaves the north-star tolerance
    (1e-3 abs / 1e-4 rel after CMVN):
      (a) tilted Gaussian noise, rectangular window (as shipped): leakage of the 400-sample rectangular window fills
          the weak bands, the measured in-frame mel range saturates near 34 dB whatever the tilt;
      (b) the same noise through a Hann window (`window` switch): the range follows the tilt;
      (c) a 1 kHz tone over a white floor at -R dB, rectangular window: the narrowband worst case.
    Float32 PCM so that the range is not capped by 16-bit quantisation.  Broadband material sits at 20-35 dB."""
    rng = np.random.default_rng(77)
    t = np.arange(48000) / 16000.0
    tone = lambda r: (0.3 * np.sin(2 * np.pi * 1000.0 * t + rng.uniform(0, 6.28))
                      + 0.3 * 10.0 ** (-r / 20.0) * rng.normal(size=t.size)).astype(np.float32)
    steps = (20, 30, 40, 50, 60, 70, 80, 90, 100)
    a, fa = _sweep(pkg, lambda r: _tilted_noise(48000, r, rng), steps)
    b, fb = _sweep(pkg, lambda r: _tilted_noise(48000, r, rng), steps, window=np.hanning(400))
    c, fc = _sweep(pkg, tone, steps)
    # (d) the same tone over a floor as int16 PCM: the path the reference's files take, i.e. the lane-per-frame kernel K1U
    # (float32 PCM runs on K1); 16-bit quantisation puts its own floor 98 dB below full scale
    tone16 = lambda r: np.clip(np.rint(tone(r).astype(np.float64) * 32768.0 / 0.35), -32768, 32767).astype(np.int16)
    d, fd = _sweep(pkg, tone16, (20, 30, 40, 50, 60, 70, 80))
    pu.record("dynamic_range", {
        "tilted_noise_rectangular_window": {"sweep": a, "first_step_out_of_tolerance_db": fa},
        "tilted_noise_hann_window": {"sweep": b, "first_step_out_of_tolerance_db": fb},
        "tone_over_white_floor_rectangular_window": {"sweep": c, "first_step_out_of_tolerance_db": fc},
        "tone_over_white_floor_int16_k1u": {"sweep": d, "first_step_out_of_tolerance_db": fd},
        "input": "float32 PCM, 6 x 3 s per step, MFCC-39 + CMVN; step = spectral tilt over 0-4 kHz (a, b) or floor below the tone (c)",
        "tolerance": {"abs": pu.ABS_TOL, "rel": pu.REL_TOL}})
    for rows in (a, b, c, d):
        by = {r["step_db"]: r for r in rows}
        for s_ in (20, 30, 40):                                        # the range broadband material and speech occupy
            assert by[s_]["elements_out_of_tolerance"] == 0, by[s_]
    assert all(r["elements_out_of_tolerance"] == 0 for r in a)         # as shipped: leakage bounds the range, FP32 is enough
