"""bench.py's host-side contract (no GPU): the workload table behind --config, the reference arm's JSON line on a tiny
sample, and the shard generator's determinism.  The GPU legs are exercised by the driver's own run of bench.py."""
import importlib.util
import json
import os
import subprocess
import sys

import numpy as np

from conftest import ROOT


def _bench():
    spec = importlib.util.spec_from_file_location("bench_mod", os.path.join(ROOT, "bench.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


def test_workload_table_matches_survey_8d():
    b = _bench()
    w = b.WORKLOADS
    assert set(w) == {"configs1", "configs2", "configs4"}
    assert w["configs4"]["metric"] == b.METRIC == "audio-hours/sec MFCC-39+CMVN"
    # SURVEY.md 8d, algorithmic work per frame: rFFT 11 520 + power / energy 1 284 + mel 400 + log 41 + DCT 1 040 (- 1 rounding)
    assert w["configs4"]["flop_k1"] == 14284 and w["configs4"]["flop_all"] == 14440
    assert w["configs4"]["bytes_k1"] == 320 + 52 and w["configs4"]["bytes_all"] == 320 + 156
    assert w["configs1"]["bytes_all"] == 320 + 960 and w["configs1"]["planes"] == 240
    assert w["configs2"]["speeds"] == (0.9, 1.0, 1.1)
    lens = b.shard_lengths(2.0, 5678)
    assert np.array_equal(lens, b.shard_lengths(2.0, 5678))                 # seeded
    assert abs(lens.sum() / 16000 / 3600 - 2.0) < 0.02 and lens.min() >= 2 * 16000 and lens.max() <= 35 * 16000


def test_reference_arm_prints_one_json_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                          "--cpu-sample-hours", "0.02"], capture_output=True, text=True, timeout=300, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-500:]
    lines = [l for l in out.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "audio-hours/sec MFCC-39+CMVN" and d["unit"] == "audio-h/s"
    assert d["value"] > 0 and d["gpu_launches"] == 0 and d["cpu_baseline"]["kind"] == "port"
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
