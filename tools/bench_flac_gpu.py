"""Throughput of the GPU FLAC decoder on a synthetic corpus: python tools/bench_flac_gpu.py [n_utts]"""
import sys, os, json, time, tempfile, shutil
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import asr_b200 as A
n = int(sys.argv[1]) if len(sys.argv) > 1 else 600
rep = int(sys.argv[2]) if len(sys.argv) > 2 else 1                    # every file is listed rep times (bigger batch, same disk set)
pcm = A.synth.corpus(n, 2.0, 15.0, seed=4567)
pcm = [(p // 16).astype(np.int16) for p in pcm]                      # LibriSpeech-like level: FLAC ratio ~0.56
hours = sum(len(p) for p in pcm) / 16000 / 3600
root = tempfile.mkdtemp(dir="/dev/shm" if os.path.isdir("/dev/shm") else None)
try:
    paths = [os.path.join(root, "%05d.flac" % i) for i in range(n)]
    packed, off, lens = A.pack_pcm(pcm)
    A.audio_io.write_audio_batch(paths, packed, off, lens, 16000)
    paths = paths * rep; pcm = pcm * rep; hours *= rep; n *= rep
    t = time.time(); buf, files, pcm_off, lens2, fs, total = A.audio_io.load_flac_batch(paths); t_load = time.time() - t
    fe = A.Frontend(A.FrontendConfig()); fe.set_profiling(True)
    pcm_total = int(pcm_off[-1] + (lens2[-1] + 7) // 8 * 8)
    d_buf = torch.zeros(total + 4096, dtype=torch.uint8, device="cuda"); d_buf[:total] = torch.from_numpy(buf[:total])
    d_pcm = torch.zeros(pcm_total, dtype=torch.int16, device="cuda")
    best = None
    for it in range(4):
        fe.decode_flac(d_buf, files, n, total, pcm_total, pcm=d_pcm); ms = fe.flac_ms()
        tot = ms["scan"] + ms["decode"] + ms["validate"]
        if best is None or tot < best[0]: best = (tot, ms)
    got = d_pcm.cpu().numpy()
    assert all(np.array_equal(got[o:o + m], x) for o, m, x in zip(pcm_off, lens2, pcm))
    t = time.time(); A.audio_io.read_audio_batch(paths); t_cpu = time.time() - t
    print(json.dumps({"utterances": n, "audio_hours": hours, "flac_bytes": total, "ratio": total / (2.0 * float(lens2.sum())),
                      "gpu_ms": best[1], "gpu_audio_h_per_s": hours / (best[0] * 1e-3),
                      "gpu_samples_per_s": float(lens2.sum()) / (best[0] * 1e-3),
                      "host_pool_audio_h_per_s": hours / t_cpu, "host_cores": os.cpu_count(),
                      "file_read_audio_h_per_s": hours / t_load}, indent=1))
finally:
    shutil.rmtree(root, ignore_errors=True)
