"""K1 / K2 device time on the bench shard for the library FE_LIB points at (A/B of kernel variants).

    FE_LIB=_scratch/lib_x.so python tools/k1_time.py [hours] [feat_type feat_dim]"""
import importlib, json, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
pkg = importlib.import_module("automatic-speech-recognition_b200")
hours = float(sys.argv[1]) if len(sys.argv) > 1 else 125.0
kw = {}
if len(sys.argv) > 3:
    kw = dict(feat_type=sys.argv[2], feat_dim=int(sys.argv[3]))
rng = np.random.default_rng(5678)
lens = pkg.synth.durations(int(hours * 3600 / 12.3), 2, 35, rng, "librispeech")
pad = (lens + 7) // 8 * 8
off = np.concatenate(([0], np.cumsum(pad)))[:-1].astype(np.int64)
g = torch.Generator(device="cuda"); g.manual_seed(91)
d_pcm = torch.empty(int(pad.sum()), dtype=torch.int16, device="cuda")
CH = 1 << 27
for s in range(0, d_pcm.numel(), CH):
    e = min(d_pcm.numel(), s + CH)
    d_pcm[s:e] = (torch.randn(e - s, device="cuda", generator=g) * 3000.0).clamp_(-32768, 32767).to(torch.int16)
fe = pkg.Frontend(pkg.FrontendConfig(**kw))
out_off, nfr = fe.plan(lens)
d_out = torch.empty(int(out_off[-1]), dtype=torch.float32, device="cuda")
for _ in range(3):
    fe.run_packed(d_pcm, off, lens, out=d_out)
fe.sync()
fe.set_profiling(True)
for _ in range(10):
    fe.run_packed(d_pcm, off, lens, out=d_out)
fe.sync()
km = fe.kernel_ms()
chk = float(d_out[: 1 << 22].double().abs().sum())
frames = int(nfr.sum())
print(json.dumps({"lib": os.environ.get("FE_LIB", "default"), "cfg": kw, "frames": frames, "k1_ms": km["frames_to_statics"],
                  "k2_ms": km["cmvn_delta_pack"], "pass_ms": km.get("device_pass"), "frac_nominal": frames * 14284 / (km["frames_to_statics"] * 1e-3) / 74.45e12,
                  "checksum": chk}))
