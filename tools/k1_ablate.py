"""Phase ablation of K1 (FE_K1_DBG): device-resident timing of the K1 launch with phases removed."""
import os, sys, importlib, json
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
pkg = importlib.import_module("automatic-speech-recognition_b200")
hours = float(sys.argv[1]) if len(sys.argv) > 1 else 30.0
rng = np.random.default_rng(1)
lens = pkg.synth.durations(int(hours * 3600 / 12.3), 2, 35, rng, "librispeech")
pad = (lens + 7) // 8 * 8
off = np.concatenate(([0], np.cumsum(pad)))[:-1].astype(np.int64)
total = int(pad.sum())
d_pcm = (torch.randn(total, device="cuda") * 3000).clamp_(-32768, 32767).to(torch.int16)
res = {}
for dbg in (os.environ.get("ABLATE", "0,1,2").split(",")):
    os.environ["FE_K1_DBG"] = dbg
    fe = pkg.Frontend(pkg.FrontendConfig())
    out_off, nfr = fe.plan(lens)
    d_out = torch.empty(int(out_off[-1]), dtype=torch.float32, device="cuda")
    st = torch.cuda.Stream(); torch.cuda.set_stream(st)
    for _ in range(3): fe.run_packed(d_pcm, off, lens, out=d_out, stream=st.cuda_stream)
    torch.cuda.synchronize()
    fe.set_profiling(True)
    for _ in range(10): fe.run_packed(d_pcm, off, lens, out=d_out, stream=st.cuda_stream)
    torch.cuda.synchronize()
    km = fe.kernel_ms()
    frames = int(nfr.sum())
    res[dbg] = {"k1_ms": km["frames_to_statics"], "k2_ms": km["cmvn_delta_pack"], "frames": frames,
                "cyc_per_frame_smsp": km["frames_to_statics"] * 1e-3 * 1.965e9 * 592 / frames}
    fe.close()
print(json.dumps(res, indent=1))
