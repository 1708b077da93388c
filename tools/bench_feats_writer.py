"""SURVEY 8f-1, first half: `process_libri_feats` end to end on a > 30 000-file set (the reference's chunked layout,
preprocess.py:116-130): FLAC files -> features on the GPU -> `{cat}-feats-{i}.pkl` + `{cat}-featlen.npy`.

The cubes reach joblib as VIEWS into the result buffers fe_run filled; joblib writes every element's bytes straight to
the file -- no host repack.  This tool times the stages and compares the pickle rate with a raw write of the same
bytes on the same filesystem.      python tools/bench_feats_writer.py [n_files] [out_json]"""
import importlib, json, multiprocessing as mp, os, shutil, sys, tempfile, time, types
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
pkg = importlib.import_module("automatic-speech-recognition_b200")
pre = importlib.import_module("automatic-speech-recognition_b200.preprocess")
n_files = int(sys.argv[1]) if len(sys.argv) > 1 else 30001
G = {}


def gen(i):
    return pkg.synth.utterance(int(G["lens"][i]), np.random.default_rng([99, i]))


def main():
    rng = np.random.default_rng(30000)
    lens = pkg.synth.durations(n_files, 2, 6, rng)
    G["lens"] = lens
    with mp.get_context("fork").Pool(os.cpu_count()) as pool:
        pcm = pool.map(gen, range(n_files), chunksize=64)
    d = tempfile.mkdtemp(prefix="asr_b200_feats_")
    try:
        paths = [os.path.join(d, "%06d.flac" % i) for i in range(n_files)]
        packed, off, ln = pkg.pack_pcm(pcm)
        t0 = time.perf_counter()
        pkg.audio_io.write_audio_batch(paths, packed, off, ln)
        t_enc = time.perf_counter() - t0
        hours = float(lens.sum()) / 16000 / 3600
        del packed, pcm
        args = types.SimpleNamespace(frame_step=10, frame_length=25, feat_dim=13, feat_type="mfcc", cmvn=True, feat_dir=os.path.join(d, "features"))
        pkg.process_audios(paths[:2000], args)                                   # warm: library, staging buffers, page cache of a part
        stages = {"process_audios_s": 0.0, "joblib_dump_s": 0.0}
        import joblib
        real_pa, real_dump = pre.process_audios, joblib.dump

        def timed_pa(*a, **k):
            t = time.perf_counter(); r = real_pa(*a, **k); stages["process_audios_s"] += time.perf_counter() - t; return r

        def timed_dump(*a, **k):
            t = time.perf_counter(); r = real_dump(*a, **k); stages["joblib_dump_s"] += time.perf_counter() - t; return r
        pre.process_audios, pre.joblib.dump = timed_pa, timed_dump
        t0 = time.perf_counter()
        featlen = pkg.process_libri_feats(paths, "train-100", 4, args)
        total = time.perf_counter() - t0
        pre.process_audios, pre.joblib.dump = real_pa, real_dump
        pk = sorted(os.listdir(args.feat_dir))
        pkl_bytes = sum(os.path.getsize(os.path.join(args.feat_dir, f)) for f in pk if f.endswith(".pkl"))
        back = joblib.load(os.path.join(args.feat_dir, "train-100-feats-0.pkl"))
        assert len(back) == n_files // 4 + 1 and back[0].shape == (featlen[0], 13, 3) and back[0].dtype == np.float32
        flat = np.empty(pkl_bytes // 4, np.float32); flat[:] = 1.0
        t0 = time.perf_counter(); flat.tofile(os.path.join(d, "raw.bin")); t_raw = time.perf_counter() - t0
        res = {"files": n_files, "audio_hours": hours, "flac_encode_s": t_enc, "files_written": pk,
               "process_libri_feats_s": total, "audio_h_per_s": hours / total, **stages,
               "pickle_bytes": pkl_bytes, "joblib_dump_gb_per_s": pkl_bytes / 1e9 / max(stages["joblib_dump_s"], 1e-9),
               "raw_write_same_bytes_s": t_raw, "raw_write_gb_per_s": pkl_bytes / 1e9 / t_raw,
               "note": "joblib.dump receives views into the buffers fe_run filled and writes each element's bytes straight to the "
                       "file; its rate is the filesystem's raw write rate -- there is no host-side repack to remove"}
        print(json.dumps(res, indent=1))
        if len(sys.argv) > 2:
            json.dump(res, open(sys.argv[2], "w"), indent=1)
    finally:
        shutil.rmtree(d, ignore_errors=True)


if __name__ == "__main__":
    main()
