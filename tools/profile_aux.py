"""One pass over the kernels around K1 for an ncu capture: K0 (speed 0.9 / 1.1), fbank-80 cube pass, K3 padded
batches, GPU FLAC decode.  python tools/profile_aux.py [hours]"""
import sys, os, tempfile, shutil
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import asr_b200 as A
hours = float(sys.argv[1]) if len(sys.argv) > 1 else 3.0
rng = np.random.default_rng(3456)
n = int(hours * 3600 / 8.5) + 1
lens = A.synth.durations(n, 2, 15, rng)
pad = (lens + 7) // 8 * 8
off = np.concatenate(([0], np.cumsum(pad)))[:-1]
d = (torch.randn(int(pad.sum()), device="cuda") * 3000).to(torch.int16)
sp = np.array([(0, 1)[i % 2] for i in range(n)], np.int32)
fe = A.Frontend(A.FrontendConfig())
out, oo, nfr = fe.run_packed(d, off, lens, speed_idx=sp); fe.sync()                      # K0 fast paths + K1 + K2
out, oo, nfr = fe.run_packed(d, off, lens); fe.sync()
bb = A.bucketing.BucketBatcher(fe)
views, _ = bb.pad(out, oo[:-1], nfr, 39, A.bucketing.plan_batches(nfr)); fe.sync()        # K3
fe80 = A.Frontend(A.FrontendConfig(feat_type="fbank", feat_dim=80))
fe80.run_packed(d, off, lens); fe80.sync()                                                # k_cube_local<80>
root = tempfile.mkdtemp(dir="/dev/shm" if os.path.isdir("/dev/shm") else None)
try:
    pcm = [(p // 16).astype(np.int16) for p in A.synth.corpus(150, 2.0, 15.0, seed=4567)]
    paths = [os.path.join(root, "%05d.flac" % i) for i in range(len(pcm))]
    packed, poff, plens = A.pack_pcm(pcm)
    A.audio_io.write_audio_batch(paths, packed, poff, plens, 16000)
    buf, files, pcm_off, lens2, fs, total = A.audio_io.load_flac_batch(paths * 24)
    pcm_total = int(pcm_off[-1] + (lens2[-1] + 7) // 8 * 8)
    got = fe.decode_flac(buf, files, len(lens2), total, pcm_total); fe.sync()             # k_flac_scan / decode / validate
    print("flac ok", bool(np.array_equal(got[:len(pcm[0])].cpu().numpy(), pcm[0])), "samples", int(lens2.sum()), "bytes", total)
finally:
    shutil.rmtree(root, ignore_errors=True)
print("frames", int(nfr.sum()), "k0 outputs", float(np.where(sp == 0, np.ceil(lens * 10 / 9), np.ceil(lens * 10 / 11)).sum()))
