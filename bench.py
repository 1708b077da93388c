#!/usr/bin/env python
"""Headline benchmark: audio-hours/second of MFCC-39 + per-utterance CMVN on N B200s.

    python bench.py --gpus N --steps K --warmup W            (N > 1: launched under torchrun)
    python bench.py --impl reference ...                     (the reference's CPU path, host cores)

A step = one pass of the hot path (tile build, frames->statics, cmvn+delta+pack) over this
rank's shard of BASELINE.json configs[4]: the synthetic 1000-hour LibriSpeech-length corpus,
sharded 8 ways (125 audio-hours per GPU, weak scaling: N = 8 is the whole corpus).  `value`
times the device-resident pass with CUDA events; `e2e` times the same call from pinned host
PCM to pinned host features (H2D + D2H inside).  One JSON line on stdout (rank 0)."""
import argparse
import importlib
import json
import os
import subprocess
import sys
import threading
import time
import types

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
PKG = "automatic-speech-recognition_b200"

METRIC = "audio-hours/sec MFCC-39+CMVN"
UNIT = "audio-h/s"
FS = 16000
FLOP_PER_FRAME_K1 = 14284        # rFFT 11520 + power/energy 1284 + mel 400 + log 41 + DCT 1040 (- rounding)
FLOP_PER_FRAME_ALL = 14440       # + CMVN 78 + deltas 78
BYTES_PER_FRAME_ALL = 476        # 320 B int16 in + 156 B cube out


# BASELINE.json configs[]: what one bench line can time.  flop / bytes per frame: SURVEY.md 8d (algorithmic).
WORKLOADS = {
    "configs4": dict(metric=METRIC, hours=125.0, cfg={}, planes=39, shape=(13, 3), flop_k1=FLOP_PER_FRAME_K1, flop_all=FLOP_PER_FRAME_ALL,
                     bytes_k1=372, bytes_all=BYTES_PER_FRAME_ALL, speeds=None, oracle=dict(feat_dim=13, feat_type="mfcc"),
                     text="BASELINE configs[4]: 1000-hour LibriSpeech-length corpus, MFCC-39 (13+d+dd) + per-utterance CMVN, "
                          "sharded 8 ways -> %.0f audio-h per GPU (weak scaling; N=8 is the whole corpus)"),
    "configs1": dict(metric="audio-hours/sec fbank-80+CMVN", hours=30.0, cfg=dict(feat_type="fbank", feat_dim=80), planes=240, shape=(80, 3),
                     flop_k1=11520 + 1284 + 368, flop_all=11520 + 1284 + 368 + 960, bytes_k1=320 + 320, bytes_all=320 + 960, speeds=None,
                     oracle=dict(feat_dim=80, feat_type="fbank"),
                     text="BASELINE configs[1]: 80 mel energies (linear, as speechpy.mfe ships them) + CMVN on a LibriSpeech-length "
                          "distribution, %.0f audio-h per GPU"),
    "configs2": dict(metric="audio-hours/sec MFCC-39+CMVN with speed perturbation 0.9/1.0/1.1", hours=30.0, cfg={}, planes=39, shape=(13, 3),
                     flop_k1=FLOP_PER_FRAME_K1, flop_all=FLOP_PER_FRAME_ALL, bytes_k1=372, bytes_all=BYTES_PER_FRAME_ALL,
                     speeds=(0.9, 1.0, 1.1), oracle=dict(feat_dim=13, feat_type="mfcc"),
                     text="BASELINE configs[2]: MFCC-39 + CMVN with the speed perturbation of utils/augmentation.py in front "
                          "(speeds 0.9 / 1.0 / 1.1 cycling over the utterances), %.0f audio-h of input per GPU"),
}


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--hours-per-gpu", type=float, default=None, help="default: 125 (configs4), 30 (configs1 / configs2)")
    ap.add_argument("--config", default="configs4", choices=sorted(WORKLOADS),
                    help="BASELINE.json configs[] entry; the driver's default run is configs4 (the headline metric)")
    ap.add_argument("--files-hours", type=float, default=16.0,
                    help="audio-hours of FLAC files per rank for the e2e_files leg (process_audios on paths; half of it per rank at N > 1); 0 = skip")
    ap.add_argument("--e2e-steps", type=int, default=2)
    ap.add_argument("--cpu-sample-hours", type=float, default=None, help="audio-hours the CPU baseline times")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    return ap.parse_args()


def shard_lengths(hours, seed):
    """LibriSpeech-like utterance lengths: clip(N(12.3, 3.8^2), 2, 35) s."""
    synth = importlib.import_module(PKG + ".synth")
    rng = np.random.default_rng(seed)
    n = int(hours * 3600.0 / 12.3) + 1
    lens = synth.durations(n, 2, 35, rng, "librispeech")
    cut = int(np.searchsorted(np.cumsum(lens), hours * 3600.0 * FS)) + 1
    return lens[:cut]


def ref_args():
    return types.SimpleNamespace(frame_step=10, frame_length=25, feat_dim=13, feat_type="mfcc", cmvn=True)


# ----------------------------------------------------------------------------------------
# CPU reference path (oracle port of preprocess.py:50-91 over speechpy), all host cores
# ----------------------------------------------------------------------------------------
def _cpu_worker(chunk):
    from oracle import speechpy_ref as ref
    a = ref_args()
    t = 0
    try:                                    # one BLAS thread per worker process: the pool owns the cores
        from threadpoolctl import threadpool_limits
        ctx = threadpool_limits(1)
    except Exception:
        import contextlib
        ctx = contextlib.nullcontext()
    with ctx:
        for p in chunk:
            f = ref.features_one(p, FS, a.frame_length, a.frame_step, a.feat_dim, a.feat_type, a.cmvn)
            t += len(f)
    return t


_CHK = {}


def _chk_worker(k):
    """Checker only (never timed, never shipped): oracle features of one picked utterance vs the kernels' cube."""
    from oracle import speechpy_ref as ref
    pcm = _CHK["pcm"][k]
    if _CHK.get("speed") is not None and _CHK["speed"][k] != 1.0:
        from oracle import sox_ref
        pcm = sox_ref.speed_perturb(pcm, float(_CHK["speed"][k]))
    want = ref.features_one(pcm, **_CHK.get("kw", {})).astype(np.float64)
    err = np.abs(_CHK["cube"][k].astype(np.float64) - want)
    return float(err.max()) if err.size else 0.0, int(((err > 1e-3) & (err > 1e-4 * np.abs(want))).sum())


def cpu_pass(pcm_list, pool, cores):
    chunks = [pcm_list[i::cores] for i in range(cores)]
    t0 = time.perf_counter()
    frames = sum(pool.map(_cpu_worker, chunks))
    return time.perf_counter() - t0, frames


def cpu_baseline(pcm_list, reps=1):
    import multiprocessing as mp
    cores = os.cpu_count() or 1
    hours = sum(len(p) for p in pcm_list) / FS / 3600.0
    with mp.get_context("fork").Pool(cores) as pool:
        cpu_pass(pcm_list[:cores], pool, cores)                      # warm the workers
        best = min(cpu_pass(pcm_list, pool, cores)[0] for _ in range(reps))
    return hours / best, cores, hours, best


def cpu_model():
    try:
        for line in open("/proc/cpuinfo"):
            if line.startswith("model name"):
                return line.split(":", 1)[1].strip()
    except OSError:
        pass
    return "unknown"


def host_sample(hours, seed):
    synth = importlib.import_module(PKG + ".synth")
    lens = shard_lengths(hours, seed)
    return synth.noise_corpus_fast(lens, seed + 1)


def run_reference(a):
    """--impl reference: the reference's CPU path on this box's host cores, same metric/config."""
    rank = int(os.environ.get("RANK", 0))
    if rank != 0:
        return
    import multiprocessing as mp
    cores = os.cpu_count() or 1
    hours = a.cpu_sample_hours
    if not hours:
        # size the per-step sample so that the whole --steps/--warmup run takes about 2.5 minutes:
        # probe the host's rate on a small sample first
        probe = host_sample(0.05 * cores, 991)
        with mp.get_context("fork").Pool(cores) as pool:
            cpu_pass(probe[:cores], pool, cores)
            dt, _ = cpu_pass(probe, pool, cores)
        rate = sum(len(p) for p in probe) / FS / 3600.0 / dt             # audio-h / s
        per_step_s = min(30.0, max(2.0, 150.0 / (a.warmup + a.steps)))
        hours = rate * per_step_s
    pcm = host_sample(hours, 5678)
    hours = sum(len(p) for p in pcm) / FS / 3600.0
    times = []
    with mp.get_context("fork").Pool(cores) as pool:
        for i in range(a.warmup + a.steps):
            dt, _ = cpu_pass(pcm, pool, cores)
            if i >= a.warmup:
                times.append(dt)
    ms = 1e3 * sum(times) / len(times)
    value = hours / (ms * 1e-3)
    sample = "%.2f audio-h (%d utterances) of the same LibriSpeech-length distribution per step" % (hours, len(pcm))
    emit({
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": a.gpus, "steps": a.steps,
        "warmup": a.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": "configs[4] 1000-hour LibriSpeech-length corpus, MFCC-39 (13+d+dd) + per-utterance CMVN; "
                               "bounded sample per step", "sample": sample, "cpu": cpu_model()},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample,
                         "note": "oracle = numpy restatement of preprocess.py:50-91 over speechpy (speechpy itself not "
                                 "installable here); multiprocessing over all host cores; decode excluded"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    })


# ----------------------------------------------------------------------------------------
class ClockSampler(threading.Thread):
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        super().__init__(daemon=True)
        self.idx = gpu_index
        self.stop_flag = threading.Event()
        self.samples = []

    def run(self):
        while not self.stop_flag.is_set():
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.idx), "--query-gpu=" + self.Q,
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
                f = [x.strip() for x in out.strip().split(",")]
                if len(f) >= 8:
                    self.samples.append(f)
            except Exception:
                pass
            self.stop_flag.wait(0.1)

    def summary(self):
        sm = [float(s[1]) for s in self.samples if s[1].replace(".", "").isdigit()]
        mx = [float(s[2]) for s in self.samples if s[2].replace(".", "").isdigit()]
        reasons = set()
        for s in self.samples:
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), s[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(self.samples)}


_REAL_STDOUT = None


def _guard_stdout():
    """The contract is ONE JSON line on stdout.  Libraries write there too (NCCL prints its version line to fd 1
    when the environment sets NCCL_DEBUG), so everything but the final line is routed to stderr at the fd level."""
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)


def emit(line):
    sys.stdout.flush()
    if _REAL_STDOUT is not None:
        os.dup2(_REAL_STDOUT, 1)
    print(json.dumps(line), flush=True)


def _gen_speechlike(i):
    synth = importlib.import_module(PKG + ".synth")
    return synth.utterance(int(_CHK["lens"][i]), np.random.default_rng([_CHK["seed"], i]))


def files_leg(pkg, hours, rank, local, world, dist, torch):
    """e2e_files: FLAC files on disk (page cache) -> process_audios(paths, args) -> object array of cubes, per rank.
    Speech-like synthetic utterances (they compress like speech, ~0.55 of the PCM), written once with the package's
    own FLAC encoder, read back through the path the reference's caller uses.  Timed: 1 warm-up + 2 passes."""
    import multiprocessing as mp
    import shutil
    import tempfile
    lens = shard_lengths(hours, 777 + rank)
    _CHK.update(lens=lens, seed=4242 + rank)
    procs = max(1, (os.cpu_count() or 1) // max(int(os.environ.get("LOCAL_WORLD_SIZE", world)), 1))
    with mp.get_context("fork").Pool(procs) as pool:
        pcm = pool.map(_gen_speechlike, range(len(lens)), chunksize=8)
    d = tempfile.mkdtemp(prefix="asr_b200_bench_r%d_" % rank)
    try:
        paths = [os.path.join(d, "%06d.flac" % i) for i in range(len(pcm))]
        packed, off, ln = pkg.pack_pcm(pcm)
        pkg.audio_io.write_audio_batch(paths, packed, off, ln)
        del packed
        nbytes = sum(os.path.getsize(p) for p in paths)
        hours = float(sum(len(p) for p in pcm)) / FS / 3600.0
        args = ref_args()
        feats, featlen = pkg.process_audios(paths, args, device=local)                 # warm: page cache, staging buffers
        assert featlen == [int((len(p) - 400) // 160) for p in pcm]
        out_bytes = int(sum(f.nbytes for f in feats))
        if dist is not None:
            dist.barrier()
        t0 = time.perf_counter()
        for _ in range(2):
            feats, featlen = pkg.process_audios(paths, args, device=local)
        dt = (time.perf_counter() - t0) / 2
        td = torch.tensor([dt], dtype=torch.float64, device="cuda")
        th = torch.tensor([hours, float(nbytes), float(out_bytes)], dtype=torch.float64, device="cuda")
        if dist is not None:
            dist.all_reduce(td, op=dist.ReduceOp.MAX)
            dist.all_reduce(th, op=dist.ReduceOp.SUM)
        return {"value": float(th[0]) / float(td[0]), "unit": UNIT, "ms_per_step": float(td[0]) * 1e3,
                "files": int(len(paths)) * world, "audio_hours": float(th[0]), "file_bytes_read_per_step": int(th[1]),
                "h2d_bytes_per_step": int(th[1]), "d2h_bytes_per_step": int(th[2]),
                "api": "process_audios(list of .flac paths, args) -> (object ndarray of (L, 13, 3) float32, featlen): "
                       "file bytes uploaded compressed, FLAC decoded on the GPU (fe_decode_flac), fe_run, cubes to host",
                "data": "speech-like synthetic utterances (synth.utterance), FLAC-compressed to %.2f of the PCM, page cache" %
                        (float(th[1]) / (float(th[0]) * 3600 * FS * 2))}
    finally:
        shutil.rmtree(d, ignore_errors=True)


def main():
    a = parse()
    _guard_stdout()
    if a.impl == "reference":
        return run_reference(a)

    import torch
    import torch.distributed as dist
    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    if world != a.gpus and world > 1:
        raise SystemExit("--gpus %d but WORLD_SIZE=%d" % (a.gpus, world))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    pkg = importlib.import_module(PKG)
    importlib.import_module(PKG + ".build").build_library()

    W = WORKLOADS[a.config]
    if a.hours_per_gpu is None:
        a.hours_per_gpu = W["hours"]
    # ---- this rank's shard, generated on the device (seeded) ----
    lens = shard_lengths(a.hours_per_gpu, 5678 + rank)
    pad = (lens + 7) // 8 * 8
    off = np.concatenate(([0], np.cumsum(pad)))[:-1].astype(np.int64)
    total = int(pad.sum())
    hours = float(lens.sum()) / FS / 3600.0
    g = torch.Generator(device="cuda")
    g.manual_seed(91 + rank)
    d_pcm = torch.empty(total, dtype=torch.int16, device="cuda")
    CH = 1 << 27
    for s in range(0, total, CH):
        e = min(total, s + CH)
        d_pcm[s:e] = (torch.randn(e - s, device="cuda", generator=g) * 3000.0).clamp_(-32768, 32767).to(torch.int16)
    cfg = pkg.FrontendConfig(**W["cfg"])            # default: mfcc, D=13, cmvn, as-shipped deltas: MFCC-39 cube
    fe = pkg.Frontend(cfg, device=local)
    speeds = None if W["speeds"] is None else [W["speeds"][i % len(W["speeds"])] for i in range(len(lens))]
    speed_idx = fe.speed_indices(speeds)
    out_off, nfr = fe.plan(lens, speed_idx)
    frames = int(nfr.sum())
    d_out = torch.empty(int(out_off[-1]), dtype=torch.float32, device="cuda")
    # an explicit (non-default) stream: the kernels, the timing events and the library's own
    # per-kernel events all live on it
    tstream = torch.cuda.Stream()
    torch.cuda.synchronize()
    torch.cuda.set_stream(tstream)
    stream = tstream.cuda_stream
    assert stream != 0

    def step():
        fe.run_packed(d_pcm, off, lens, speed_idx=speed_idx, out=d_out, stream=stream)

    peak_variants = fe.measure_fp32_peaks()
    fp32_peak = max(peak_variants.values())
    for _ in range(max(a.warmup, 3)):
        step()
    torch.cuda.synchronize()

    # ---- timed region: device-resident (no per-kernel events inside: they come from a second, profiled pass) ----
    sampler = ClockSampler(local)
    sampler.start()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    l0 = fe.launch_count()
    k1_ms, k2_ms = [], []
    ev[0].record()
    for _ in range(a.steps):
        step()
    ev[1].record()
    torch.cuda.synchronize()
    launches = fe.launch_count() - l0
    ms_total = ev[0].elapsed_time(ev[1])
    # the same K steps once more with the library's per-kernel events (on the launching stream) for the roofline
    fe.set_profiling(True)
    for _ in range(a.steps):
        step()
    torch.cuda.synchronize()
    km = fe.kernel_ms()                               # mean over those steps
    k1_ms.append(km["frames_to_statics"]); k2_ms.append(km["cmvn_delta_pack"])
    k0_ms = float(km["resample"])
    if world > 1:
        dist.barrier()
    sampler.stop_flag.set()
    sampler.join(timeout=3)
    t = torch.tensor([ms_total, hours, float(frames)], dtype=torch.float64, device="cuda")
    if world > 1:
        tmax = t.clone(); dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        tsum = t.clone(); dist.all_reduce(tsum, op=dist.ReduceOp.SUM)
        ms_total, hours_all, frames_all = float(tmax[0]), float(tsum[1]), float(tsum[2])
    else:
        hours_all, frames_all = hours, float(frames)
    ms_per_step = ms_total / a.steps
    value = hours_all / (ms_per_step * 1e-3)
    fe.set_profiling(False)

    # ---- end to end: pinned host PCM -> features in pinned host memory, through the public API ----
    # Every rank pins its shard (14.4 GB in + 7 GB out at the default size).  If the box cannot pin that for
    # all ranks, all ranks together fall back to the same leading fraction of their shards (said in "sample");
    # every decision below is collective, so that no rank waits at a barrier another one skipped.
    def all_min(x):
        t_ = torch.tensor([float(x)], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t_, op=dist.ReduceOp.MIN)
        return float(t_[0])

    e2e, chk = None, None
    n_all = len(lens)
    need = total * 2 + int(out_off[-1]) * 4
    try:
        import psutil
        avail = psutil.virtual_memory().available / max(int(os.environ.get("LOCAL_WORLD_SIZE", world)), 1)
    except Exception:
        avail = float("inf")
    frac = all_min(min(1.0, 0.6 * avail / need))
    n_e2e = n_all if frac >= 1.0 else max(int(n_all * frac), 1)
    tot_e = int(off[n_e2e - 1] + pad[n_e2e - 1])
    out_e = int(out_off[n_e2e])
    hours_e = float(lens[:n_e2e].sum()) / FS / 3600.0
    ok, err = 1.0, ""
    try:
        h_pcm_t = torch.empty(tot_e, dtype=torch.int16).pin_memory()
        h_pcm_t.copy_(d_pcm[:tot_e])
        h_out_t = torch.empty(out_e, dtype=torch.float32).pin_memory()
        h_pcm, h_out = h_pcm_t.numpy(), h_out_t.numpy()
        fe.run_packed(h_pcm, off[:n_e2e], lens[:n_e2e], speed_idx=None if speed_idx is None else speed_idx[:n_e2e], out=h_out)      # warm (device staging buffers)
    except Exception as ex:                                             # e.g. not enough pinnable host memory
        ok, err = 0.0, str(ex)[:200]
    if all_min(ok) < 1.0:
        e2e = {"value": None, "unit": UNIT, "error": err or "another rank could not pin its host buffers"}
    else:
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(a.e2e_steps):
            fe.run_packed(h_pcm, off[:n_e2e], lens[:n_e2e], speed_idx=None if speed_idx is None else speed_idx[:n_e2e], out=h_out)  # synchronous for host outputs
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / a.e2e_steps
        td = torch.tensor([dt], dtype=torch.float64, device="cuda")
        th = torch.tensor([hours_e], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(td, op=dist.ReduceOp.MAX)
            dist.all_reduce(th, op=dist.ReduceOp.SUM)
        e2e = {"value": float(th[0]) / float(td[0]), "unit": UNIT, "h2d_bytes_per_step": int(tot_e * 2),
               "d2h_bytes_per_step": out_e * 4, "ms_per_step": float(td[0]) * 1e3,
               "api": "Frontend.run_packed(host ndarray) -> fe_run (C-ABI), pinned host buffers",
               "host_gbs_per_rank": (tot_e * 2 + out_e * 4) / float(td[0]) / 1e9,
               "host_gbs_all_ranks": world * (tot_e * 2 + out_e * 4) / float(td[0]) / 1e9,
               "note": "every rank moves its whole shard between pinned host memory and HBM: PCIe Gen5 x16 at N = 1; at N >= 4 the box's "
                       "aggregate host-memory / root-complex rate (one NUMA node, all GPUs behind it: nvidia-smi topo) is the ceiling",
               "sample": "whole shard" if n_e2e == n_all else
                         "first %d of %d utterances of every shard (%.1f audio-h per GPU): host memory" % (n_e2e, n_all, hours_e)}
        k = min(64, n_e2e)
        chk = float(np.abs(h_out[:int(out_off[k])] - d_out[:int(out_off[k])].cpu().numpy()).max())

    # ---- end to end through the reference's real entry point: process_audios(list of FLAC paths, args) ----
    # (preprocess.py:50-69 takes file paths; files sit in the page cache, as LibriSpeech does on the second pass)
    e2e_files = None
    if a.files_hours > 0 and a.config == "configs4":
        h_pcm_t = h_out_t = h_pcm = h_out = None        # release the pinned 21 GB before the file leg
        try:
            # N > 1: half the set per rank (the ranks share the host's cores for generating and encoding the files)
            e2e_files = files_leg(pkg, a.files_hours if world == 1 else a.files_hours / 2, rank, local, world, dist if world > 1 else None, torch)
        except Exception as ex:
            e2e_files = {"value": None, "unit": UNIT, "error": str(ex)[:200]}

    if rank != 0:
        if world > 1:
            dist.barrier(); dist.destroy_process_group()
        return

    # ---- checker (not timed): 256 utterances spread over the whole shard against the oracle (all host cores) ----
    parity = None
    try:
        import multiprocessing as mp
        pick = np.unique(np.linspace(0, len(lens) - 1, 256).astype(np.int64))
        _CHK["pcm"] = [d_pcm[int(off[i]):int(off[i]) + int(lens[i])].cpu().numpy() for i in pick]
        _CHK["cube"] = [d_out[int(out_off[i]):int(out_off[i]) + int(nfr[i]) * W["planes"]].cpu().numpy().reshape((int(nfr[i]),) + W["shape"]) for i in pick]
        _CHK["kw"] = W["oracle"]
        _CHK["speed"] = None if speeds is None else [speeds[i] for i in pick]
        with mp.get_context("fork").Pool(min(os.cpu_count() or 1, 32)) as pool:
            rows = np.array(pool.map(_chk_worker, range(len(pick))))
        parity = {"utterances": int(len(pick)), "spread": "np.linspace over the %d utterances of the rank-0 shard" % len(lens),
                  "max_abs_err_vs_oracle": float(rows[:, 0].max()), "elements_out_of_tolerance": int(rows[:, 1].sum()),
                  "tolerance": "abs <= 1e-3 OR abs <= 1e-4 |ref| (north_star)", "host_vs_device_max_abs": chk}
        if speeds is not None:
            parity["note"] = ("the resampler re-quantises to int16: FP32 vs FP64 accumulation differs by 1 LSB on ~0.05 % of the samples, "
                              "which the log amplifies in low-energy mel bands (tests/test_gpu_fullsize_parity.py separates the two stages)")
    except Exception as ex:
        parity = {"error": str(ex)[:200]}

    # ---- CPU baseline on this box's host cores, bounded sample of the same shard ----
    cpu = None
    if world == 1 and not a.no_cpu_baseline:
        cores = os.cpu_count() or 1
        # about 15 s of CPU work on all host cores: probe the rate on a small slice, then size the sample
        take = lambda h: [d_pcm[int(off[i]):int(off[i]) + int(lens[i])].cpu().numpy()
                          for i in range(min(int(np.searchsorted(np.cumsum(lens), h * 3600 * FS)) + 1, len(lens)))]
        want_h = a.cpu_sample_hours
        if not want_h:
            pv, _, _, _ = cpu_baseline(take(0.05 * cores))
            want_h = min(pv * 15.0, 0.6 * hours)
        sample = take(want_h)
        v, cores, h_s, secs = cpu_baseline(sample)
        cpu = {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
               "sample": "first %d utterances (%.2f audio-h) of the rank-0 shard, %.1f s on %d processes" % (len(sample), h_s, secs, cores),
               "cpu": cpu_model()}
        # how the reference actually runs (preprocess.py:67: one process, one file at a time): ~3 s on one core
        try:
            one = take(max(v / cores * 3.0, 0.01))
            t1 = time.perf_counter()
            _cpu_worker(one)
            dt1 = time.perf_counter() - t1
            h1 = sum(len(p) for p in one) / FS / 3600.0
            cpu["single_process"] = {"value": h1 / dt1, "unit": UNIT, "sample": "%d utterances (%.3f audio-h), %.1f s, 1 BLAS thread" % (len(one), h1, dt1)}
            # the same process with BLAS threads left at the library's default (what `python3 preprocess.py` gets)
            from oracle import speechpy_ref as _ref
            t1 = time.perf_counter()
            for p_ in one:
                _ref.features_one(p_)
            dt2 = time.perf_counter() - t1
            cpu["single_process_blas_unpinned"] = {"value": h1 / dt2, "unit": UNIT, "sample": "same %d utterances, %.1f s, BLAS threads at default" % (len(one), dt2)}
            cpu["note"] = ("the %d-process figure is %.1fx one process: the numpy port is memory-bound (index gathers, (L, 400) float64 "
                           "temporaries), not core-bound" % (cores, v / (h1 / dt1)))
        except Exception as ex:
            cpu["single_process"] = {"error": str(ex)[:120]}

    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    k1 = float(np.mean(k1_ms)) * 1e-3
    traffic, traffic_source = None, None
    try:
        tj = json.load(open(os.path.join(ROOT, "profiles", "k1_traffic.json")))
        if a.config == "configs4":
            traffic = float(tj["dram_bytes_per_frame"]) * frames
            traffic_source = "replayed: dram__bytes_read.sum + dram__bytes_write.sum per frame of the ncu capture %s x frames of this launch" % tj.get("capture", "")
    except Exception:
        pass
    ach = frames * W["flop_k1"] / k1 / 1e12
    nominal = 148 * 128 * 2 * 1.965e9 / 1e12
    roofline = {
        "kernel": "k_frames_to_statics_u" if os.environ.get("FE_K1T", "2") == "2" else "k_frames_to_statics",
        "bound": "fp32", "achieved": ach, "peak": fp32_peak, "unit": "TFLOP/s",
        "frac": ach / fp32_peak, "traffic": traffic, "traffic_source": traffic_source,
        "peak_source": "measured live, best of four FMA probes (fe_measure_fp32_peaks): what the FP32 pipe sustains depends on the operand "
                       "source -- scalar FFMA with a warp-uniform operand reaches the nominal rate, three register operands do not "
                       "(register-file read bandwidth, tools/ubench_issue2.cu); nominal 148 SM x 128 lanes x 2 x 1.965 GHz = %.1f" % nominal,
        "peak_variants": peak_variants, "peak_nominal": nominal, "frac_of_nominal": ach / nominal,
        "algorithmic_flop_per_frame": W["flop_k1"], "frames_per_launch": frames, "launch_ms": k1 * 1e3,
        "launch_ms_source": "CUDA events on the launching stream around every K1 launch, mean over %d steps run right after the timed region" % a.steps,
        "hbm": {"achieved_gbs": frames * W["bytes_k1"] / k1 / 1e9, "peak_gbs": hbm_peak,
                "peak_source": "MEASURED_PEAKS.json" if peaks else "fallback 6650 (B200_PROFILING.md)",
                "algorithmic_bytes_per_frame": W["bytes_k1"]},
        "whole_pass": {"flop_per_frame": W["flop_all"], "bytes_per_frame": W["bytes_all"],
                       "achieved_tflops": frames_all * W["flop_all"] / (ms_per_step * 1e-3) / 1e12 / world,
                       "achieved_gbs": frames_all * W["bytes_all"] / (ms_per_step * 1e-3) / 1e9 / world},
        "cmvn_delta_pack_ms": float(np.mean(k2_ms)), "resample_ms": k0_ms,
    }
    line = {
        "metric": W["metric"], "value": value, "unit": UNIT, "n_gpus": world, "steps": a.steps, "warmup": max(a.warmup, 3),
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": W["text"] % a.hours_per_gpu,
                   "audio_hours_per_gpu": hours, "utterances_per_gpu": int(len(lens)), "frames_per_gpu": frames,
                   "pcm": "int16 16 kHz, seeded on-device Gaussian noise (broadband), clip(N(12.3,3.8^2),2,35) s",
                   "l2": "inputs (%.1f GB/GPU) far larger than the 126 MB L2; no flush needed" % (total * 2 / 1e9),
                   "sharding": "utterances, no collective (per-utterance CMVN)"},
        "e2e": e2e, "e2e_files": e2e_files, "gpu_launches": int(launches), "roofline": roofline, "cpu_baseline": cpu,
        "clocks": sampler.summary(), "parity_check": parity,
    }
    emit(line)
    if world > 1:
        dist.barrier(); dist.destroy_process_group()


if __name__ == "__main__":
    main()
