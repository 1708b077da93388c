"""Where do K1's warps spend their cycles?  clock64 totals per phase (FE_K1_DBG=8), per tile / per group."""
import os, sys, importlib, json
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["FE_K1_DBG"] = str(8 | int(os.environ.get("DBG", "0")))
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
tiles = c[4] / 4.0; groups = float(c[13])          # 4 epilogue warps count every tile
epi = {n: c[i] / max(c[4], 1) for i, n in enumerate(["wait_full", "mel", "barrier", "dct"])}
fft = {n: c[8 + i] / max(groups, 1) for i, n in enumerate(["wait_raw", "stage_a", "stage_b", "wait_empty", "post_pass"])}
fft["loop_top"] = c[14] / max(groups, 1)
print(json.dumps({"k1_ms": km["frames_to_statics"], "tiles": tiles, "groups": groups,
                  "epilogue_cycles_per_tile_per_warp": epi, "fft_cycles_per_group_per_warp": fft}, indent=1))
