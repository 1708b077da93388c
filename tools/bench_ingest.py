"""File boundary: native FLAC decode / encode throughput on the host's cores and process_audios on FLAC files
end to end (decode pool overlapped with the GPU).  python tools/bench_ingest.py [n_utts] [threads]
Rates are audio-hours per second."""
import sys, os, json, time, tempfile, shutil, types
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import asr_b200 as A
n = int(sys.argv[1]) if len(sys.argv) > 1 else 400
threads = int(sys.argv[2]) if len(sys.argv) > 2 else 0
rep = int(sys.argv[3]) if len(sys.argv) > 3 else 1                    # list every file rep times for the process_audios runs
pcm = A.synth.corpus(n, 2.0, 15.0, seed=4567)
hours = sum(len(p) for p in pcm) / 16000 / 3600
root = tempfile.mkdtemp(dir="/dev/shm" if os.path.isdir("/dev/shm") else None)
res = {"utterances": n, "audio_hours": hours, "cores": os.cpu_count()}
try:
    for tag, scale in (("full_scale", 1), ("quiet_1_16", 16)):            # second set: LibriSpeech-like levels compress to ~0.6
        data = [(p // scale).astype(np.int16) for p in pcm]
        paths = [os.path.join(root, "%s-%05d.flac" % (tag, i)) for i in range(n)]
        packed, off, lens = A.pack_pcm(data)
        t = time.time(); A.audio_io.write_audio_batch(paths, packed, off, lens, 16000, n_threads=threads); t_enc = time.time() - t
        size = sum(os.path.getsize(p) for p in paths)
        t = time.time(); A.audio_io.read_audio_batch(paths, n_threads=1); t_dec1 = time.time() - t
        t = time.time(); got, goff, glens, fs = A.audio_io.read_audio_batch(paths, n_threads=threads); t_decN = time.time() - t
        assert all(np.array_equal(got[o:o + m], d) for o, m, d in zip(goff, glens, data))
        r = {"flac_ratio": size / (2.0 * float(lens.sum())), "encode_pool_h_per_s": hours / t_enc,
             "decode_1_thread_h_per_s": hours / t_dec1, "decode_pool_h_per_s": hours / t_decN}
        try:
            import torch
            if torch.cuda.is_available():
                args = types.SimpleNamespace(frame_step=10, frame_length=25, feat_dim=13, feat_type="mfcc", cmvn=True)
                A.process_audios(paths[:8], args, device_decode=False)               # warm-up (handle, tables)
                big = paths * rep
                t = time.time(); feats, featlen = A.process_audios(big, args, n_threads=threads, device_decode=False); t_e2e = time.time() - t
                r["process_audios_files_h_per_s"] = hours * rep / t_e2e
                A.process_audios(paths[:8], args, device_decode=True)
                A.process_audios(big, args, n_threads=threads, device_decode=True)          # grows the device buffers once
                t = time.time(); feats2, featlen2 = A.process_audios(big, args, n_threads=threads, device_decode=True); t_dev = time.time() - t
                assert featlen2 == featlen and all(np.array_equal(a, b) for a, b in zip(feats, feats2))
                r["process_audios_files_device_decode_h_per_s"] = hours * rep / t_dev
                r["process_audios_audio_hours"] = hours * rep
        except ImportError:
            pass
        res[tag] = r
finally:
    shutil.rmtree(root, ignore_errors=True)
print(json.dumps(res, indent=1))
