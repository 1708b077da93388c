"""HBM throughput of the bucketed-batch kernel (fe_pad_batches / k_pad_slots) on cubes resident in HBM:
python tools/bench_pad.py [hours] [feat_dim].  Algorithmic bytes = 4 (valid + slot floats) per slot."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import asr_b200 as A
hours = float(sys.argv[1]) if len(sys.argv) > 1 else 30.0
D = int(sys.argv[2]) if len(sys.argv) > 2 else 13
rng = np.random.default_rng(4567)
lens = A.synth.durations(int(hours * 3600 / 8.5), 2, 15, rng)                 # configs[3]: U(2,15) s, all below 1710 frames
nfr = np.maximum((lens - 400) // 160, 0).astype(np.int32)
row = D * 3
off = np.zeros(len(nfr) + 1, np.int64); np.cumsum((nfr.astype(np.int64) * row + 3) // 4 * 4, out=off[1:])
feats = torch.randn(int(off[-1]), device="cuda")
fe = A.Frontend(A.FrontendConfig()); bb = A.bucketing.BucketBatcher(fe)
plan = A.bucketing.plan_batches(nfr)
src, valid, dst, slot, bases, total = bb.layout(plan, nfr, row)
out = torch.empty(total, device="cuda")
fe.set_profiling(True)
ms = []
for it in range(6):
    bb.pad(feats, off[:-1], nfr, row, plan, out=out); fe.sync(); ms.append(bb.pad_ms())
best = min(ms[1:]); byts = 4.0 * (valid.astype(np.int64).sum() + slot.astype(np.int64).sum())
print(json.dumps({"utterances": len(nfr), "batches": len(plan), "feat_dim": D, "in_gb": 4e-9 * float(valid.astype(np.int64).sum()),
                  "out_gb": 4e-9 * float(slot.astype(np.int64).sum()), "ms": best, "gb_per_s": byts / (best * 1e-3) / 1e9,
                  "frac_of_6532": byts / (best * 1e-3) / 1e9 / 6531.9}))
