"""Fixed device-resident speed-perturbation workload for ncu captures of K0: python tools/profile_k0.py [hours] [reps]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import asr_b200 as A
hours = float(sys.argv[1]) if len(sys.argv) > 1 else 5.0
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
rng = np.random.default_rng(3456)
n = int(hours * 3600 / 12.3) + 1
lens = A.synth.durations(n, 2, 35, rng, "librispeech")
pad = (lens + 7) // 8 * 8
off = np.concatenate(([0], np.cumsum(pad)))[:-1]
d = (torch.randn(int(pad.sum()), device="cuda") * 3000).to(torch.int16)
sp = np.array([(0, 1)[i % 2] for i in range(n)], np.int32)          # every utterance resampled: 0.9 / 1.1
fe = A.Frontend(A.FrontendConfig()); fe.set_profiling(True)
out = None
for it in range(reps):
    out, oo, nfr = fe.run_packed(d, off, lens, speed_idx=sp, out=out); fe.sync()
    ms = fe.kernel_ms()
    outs = float(np.where(sp == 0, np.ceil(lens * 10 / 9), np.ceil(lens * 10 / 11)).sum())
    print(it, ms, "outputs %.3g  -> %.2f TFLOP/s (64 FLOP per output)" % (outs, outs * 64 / (ms["resample"] * 1e-3) / 1e12), flush=True)
