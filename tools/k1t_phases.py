"""Where do the two warps of a K1T lane group spend their cycles?  clock64 totals per phase (library built with FE_K1_PROF)."""
import os, sys, importlib, json
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
pkg = importlib.import_module("automatic-speech-recognition_b200")
hours = float(sys.argv[1]) if len(sys.argv) > 1 else 30.0
rng = np.random.default_rng(1)
lens = pkg.synth.durations(int(hours * 3600 / 12.3), 2, 35, rng, "librispeech")
pad = (lens + 7) // 8 * 8
off = np.concatenate(([0], np.cumsum(pad)))[:-1].astype(np.int64)
d_pcm = (torch.randn(int(pad.sum()), device="cuda") * 3000).clamp_(-32768, 32767).to(torch.int16)
fe = pkg.Frontend(pkg.FrontendConfig())
out_off, nfr = fe.plan(lens)
d_out = torch.empty(int(out_off[-1]), dtype=torch.float32, device="cuda")
for _ in range(2): fe.run_packed(d_pcm, off, lens, out=d_out)
fe.sync(); fe.debug_counters()
fe.set_profiling(True)
fe.run_packed(d_pcm, off, lens, out=d_out)
fe.sync()
c = fe.debug_counters(); km = fe.kernel_ms()
tiles = float(sum(-(-int(n) // 32) for n in nfr))
names = ["wait_raw", "stage_a", "pair_sync_1", "rows_0_8", "row_pairs", "pair_sync_2", "epilogue", "loop_top"]
print(json.dumps({"k1_ms": km["frames_to_statics"], "tiles": tiles,
                  "cycles_per_tile": {"warp0": {n: c[i] / tiles for i, n in enumerate(names)},
                                      "warp1": {n: c[8 + i] / tiles for i, n in enumerate(names)}}}, indent=1))
