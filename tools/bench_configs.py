"""Device-resident throughput of BASELINE.json configs[1] (fbank-80 + CMVN) and configs[2] (MFCC-39 with speed
perturbation 0.9/1.0/1.1) next to configs[4]'s MFCC-39 (what bench.py reports).  python tools/bench_configs.py [hours]"""
import os, sys, importlib, json
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
pkg = importlib.import_module("automatic-speech-recognition_b200")
hours = float(sys.argv[1]) if len(sys.argv) > 1 else 30.0
rng = np.random.default_rng(2345)
lens = pkg.synth.durations(int(hours * 3600 / 12.3), 2, 35, rng, "librispeech")
pad = (lens + 7) // 8 * 8
off = np.concatenate(([0], np.cumsum(pad)))[:-1].astype(np.int64)
d_pcm = (torch.randn(int(pad.sum()), device="cuda") * 3000).clamp_(-32768, 32767).to(torch.int16)
h = float(lens.sum()) / 16000 / 3600
res = {}
def run(name, cfg, speed_idx=None):
    fe = pkg.Frontend(cfg)
    out_off, nfr = fe.plan(lens, speed_idx)
    d_out = torch.empty(int(out_off[-1]), dtype=torch.float32, device="cuda")
    st = torch.cuda.Stream(); torch.cuda.set_stream(st)
    for _ in range(3): fe.run_packed(d_pcm, off, lens, speed_idx=speed_idx, out=d_out, stream=st.cuda_stream)
    torch.cuda.synchronize(); fe.set_profiling(True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): fe.run_packed(d_pcm, off, lens, speed_idx=speed_idx, out=d_out, stream=st.cuda_stream)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10; km = fe.kernel_ms()
    res[name] = {"audio_h_per_s": h / (ms * 1e-3), "ms_per_pass": ms, "frames": int(nfr.sum()), "kernels_ms": km,
                 "out_gb": int(out_off[-1]) * 4 / 1e9}
    fe.close()
run("configs[4] mfcc-39 + cmvn", pkg.FrontendConfig())
run("configs[1] fbank-80 (linear, as shipped) + cmvn", pkg.FrontendConfig(feat_type="fbank", feat_dim=80))
run("configs[1] fbank-80 log + cmvn", pkg.FrontendConfig(feat_type="fbank", feat_dim=80, fbank_log=True))
sp = np.array([(-1, 0, 1)[i % 3] for i in range(len(lens))], np.int32)
run("configs[2] mfcc-39 + speed 0.9/1.0/1.1", pkg.FrontendConfig(), sp)
run("mfcc 40 filters -> 39 cepstra (run.sh default feat_dim), specialised plan", pkg.FrontendConfig(feat_dim=39))
run("fbank-40 + cmvn, specialised plan", pkg.FrontendConfig(feat_type="fbank", feat_dim=40))
run("mfcc-39, Hamming window (window switch), specialised plan", pkg.FrontendConfig(window=np.hamming(400)))
os.environ["FE_K1_GENERIC"] = "1"
run("mfcc-39, generic (run-time plan) epilogue", pkg.FrontendConfig())
run("mfcc 40 -> 39, generic epilogue", pkg.FrontendConfig(feat_dim=39))
run("fbank-40, generic epilogue", pkg.FrontendConfig(feat_type="fbank", feat_dim=40))
print(json.dumps({"audio_hours": h, "results": res}, indent=1))
