"""Full-size parity helpers (TEST INFRASTRUCTURE): the oracle is run over EVERY utterance of a BASELINE-sized
set on all host cores (fork pool; the arrays are inherited copy-on-write, only statistics travel back)."""
import json
import multiprocessing as mp
import os
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ABS_TOL, REL_TOL = 1e-3, 1e-4          # BASELINE.json north_star: max-abs 1e-3 / rel 1e-4 after CMVN

_G = {}


def _gen_one(i):
    from importlib import import_module
    synth = import_module("automatic-speech-recognition_b200.synth")
    rng = np.random.default_rng([_G["seed"], i])
    return synth.utterance(int(_G["lens"][i]), rng)


def gen_corpus(lens, seed, procs=None):
    """Broadband speech-like int16 utterances (synth.utterance), one independent stream per utterance so that the
    set can be generated on all cores: utterance i = synth.utterance(lens[i], default_rng([seed, i]))."""
    _G.update(seed=int(seed), lens=np.asarray(lens))
    procs = procs or os.cpu_count() or 1
    with mp.get_context("fork").Pool(procs) as pool:
        return pool.map(_gen_one, range(len(lens)), chunksize=max(1, len(lens) // (8 * procs)))


def _cmp(got, want):
    if got.shape != want.shape:
        return (-1.0, -1.0, -1, 0, 0.0, 0.0, 0.0)
    if want.size == 0:
        return (0.0, 0.0, 0, 0, 0.0, 0.0, 0.0)
    err = np.abs(got - want)
    bad = (err > ABS_TOL) & (err > REL_TOL * np.abs(want))
    units = np.minimum(err / ABS_TOL, err / (REL_TOL * np.abs(want) + 1e-300))        # <= 1 passes
    planes = [float(err[..., k].max()) for k in range(3)] if want.ndim == 3 else [float(err.max()), 0.0, 0.0]
    return (float(err.max()), float(units.max()), int(bad.sum()), int(want.size), planes[0], planes[1], planes[2])


def _err_one(i):
    """Row: index, then _cmp(cube, oracle features of the oracle's PCM); with a resampler in the path also
    _cmp(cube, oracle features of the GPU-resampled PCM) and the int16 difference GPU vs oracle resampler."""
    from oracle import speechpy_ref as ref
    pcm = _G["pcm"][i]
    got = np.asarray(_G["cubes"][i], dtype=np.float64)
    extra = (0.0, 0.0, 0, 0, 0.0, 0.0, 0.0, 0, 0, 0)
    if _G.get("speeds") is not None and _G["speeds"][i] != 1.0:
        from oracle import sox_ref
        pcm = sox_ref.speed_perturb(pcm, float(_G["speeds"][i]))
        if _G.get("gpu_pcm") is not None:
            g = np.asarray(_G["gpu_pcm"][i])
            if g.shape == pcm.shape:
                d = np.abs(g.astype(np.int32) - pcm.astype(np.int32))
                lsb = (int(d.max()) if d.size else 0, int((d > 0).sum()), int(d.size))
            else:
                lsb = (-1, 0, 0)
            extra = _cmp(got, ref.features_one(g, **_G["kw"]).astype(np.float64)) + lsb
    return (i,) + _cmp(got, ref.features_one(pcm, **_G["kw"]).astype(np.float64)) + extra


def oracle_errors(pcm, cubes, kw, speeds=None, gpu_pcm=None, procs=None):
    """Compare every cube with the oracle's; returns a summary dict (no assertion here).  ``gpu_pcm``: the kernels'
    resampled int16 utterances, to separate the resampler's +-1 LSB rounding from the feature chain."""
    _G.update(pcm=pcm, cubes=cubes, kw=dict(kw), speeds=speeds, gpu_pcm=gpu_pcm)
    procs = procs or os.cpu_count() or 1
    t0 = time.perf_counter()
    with mp.get_context("fork").Pool(procs) as pool:
        rows = pool.map(_err_one, range(len(pcm)), chunksize=max(1, len(pcm) // (8 * procs)))
    dt = time.perf_counter() - t0
    rows = np.array(rows, dtype=np.float64)
    worst = int(rows[np.argmax(rows[:, 2]), 0])
    out = _summary(rows, pcm, worst, dt, procs)
    if gpu_pcm is not None:
        x = rows[:, 8:]
        out["given_the_kernels_resampled_pcm"] = {
            "max_abs_err": float(x[:, 0].max()), "max_tolerance_units": float(x[:, 1].max()),
            "elements_out_of_tolerance": int(np.maximum(x[:, 2], 0).sum()), "shape_mismatches": int((x[:, 0] < 0).sum())}
        out["resampler_int16_vs_oracle"] = {
            "max_abs_lsb": int(x[:, 7].max()), "length_mismatches": int((x[:, 7] < 0).sum()),
            "samples_differing": int(x[:, 8].sum()), "samples": int(x[:, 9].sum()),
            "fraction_differing": float(x[:, 8].sum() / max(x[:, 9].sum(), 1))}
    return out


def _summary(rows, pcm, worst, dt, procs):
    return {
        "utterances": int(len(pcm)), "elements": int(rows[:, 4].sum()),
        "audio_hours": float(sum(len(p) for p in pcm)) / 16000 / 3600,
        "shape_mismatches": int((rows[:, 1] < 0).sum()),
        "max_abs_err": float(rows[:, 1].max()), "max_tolerance_units": float(rows[:, 2].max()),
        "elements_out_of_tolerance": int(np.maximum(rows[:, 3], 0).sum()),
        "max_abs_err_by_plane": {"static": float(rows[:, 5].max()), "delta": float(rows[:, 6].max()), "delta2": float(rows[:, 7].max())},
        "median_utterance_max_abs_err": float(np.median(rows[:, 1])),
        "worst_utterance": worst, "tolerance": {"abs": ABS_TOL, "rel": REL_TOL, "rule": "abs <= 1e-3 OR abs <= 1e-4 |ref|, element-wise"},
        "oracle_seconds": dt, "oracle_processes": procs,
    }


def record(name, summary):
    """Drop the summary where the driver's GPU run collects artefacts (copied into profiles/ by hand)."""
    d = os.path.join(ROOT, "gpurun_out")
    try:
        os.makedirs(d, exist_ok=True)
        with open(os.path.join(d, "r02_parity_%s.json" % name), "w") as f:
            json.dump(summary, f, indent=1)
    except OSError:
        pass
