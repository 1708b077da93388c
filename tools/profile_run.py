"""Fixed device-resident workload for ncu captures: python tools/profile_run.py [hours] [feat] [reps]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import asr_b200 as A
hours = float(sys.argv[1]) if len(sys.argv) > 1 else 10.0
feat = sys.argv[2] if len(sys.argv) > 2 else "mfcc"
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
rng = np.random.default_rng(5678)
n = int(hours * 3600 / 12.3) + 1
lens = A.synth.durations(n, 2, 35, rng, "librispeech")
pad = (lens + 7) // 8 * 8
off = np.concatenate(([0], np.cumsum(pad)))[:-1]
g = torch.Generator(device="cuda"); g.manual_seed(1)
d = (torch.randn(int(pad.sum()), device="cuda", generator=g) * 3000).to(torch.int16)
cfg = A.FrontendConfig(feat_type="mfcc", feat_dim=39) if feat == "mfcc39" else A.FrontendConfig(feat_type=feat, feat_dim=13 if feat == "mfcc" else 80)
fe = A.Frontend(cfg); fe.set_profiling(True)
out = None
for it in range(reps):
    out, oo, nfr = fe.run_packed(d, off, lens, out=out); fe.sync()
    ms = fe.kernel_ms(); h = lens.sum() / 16000 / 3600
    print(it, ms, "audio-h %.2f" % h, "audio-h/s %.0f" % (h / (ms["device_pass"] * 1e-3)), flush=True)
