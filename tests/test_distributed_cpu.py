"""world_size-2 run of the sharded front-end host logic over gloo on CPU: every rank extracts
its LPT shard (feature kernel stood in by the oracle -- the GPU path is covered by -m gpu),
rank 0 re-assembles by the inverse permutation and must equal the single-process result."""
import importlib
import os
import socket
import sys

import numpy as np
import pytest

from conftest import PKG, ROOT, make_args


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch.distributed as dist
    dist.init_process_group("gloo", rank=rank, world_size=world)
    pkg = importlib.import_module(PKG)
    from oracle import speechpy_ref as ref
    pcm = pkg.synth.corpus(9, 0.3, 1.2, seed=11)
    args = make_args()

    def extract(pcm_list, args, fs=16000, device=0, **kw):
        return ref.process_audios(pcm_list, args)
    feats, featlen = pkg.sharding.process_pcm_sharded(pcm, args, extract_fn=extract)
    ok = True
    if rank == 0:
        want, want_len = ref.process_audios(pcm, args)
        ok = featlen == want_len and all(np.array_equal(a, b) for a, b in zip(feats, want))
    else:
        ok = feats is None and featlen is None
    # file lists: process_audios_sharded / process_libri_feats_sharded (preprocess.py:112-130 over ranks)
    import joblib
    tmp = os.environ["ASR_TEST_TMP"]
    paths = []
    for i, x in enumerate(pcm):
        p = os.path.join(tmp, "utt%02d.flac" % i)
        if rank == 0:
            pkg.audio_io.write_audio(p, x, 16000)
        paths.append(p)
    dist.barrier()

    def process_files(audio_path, args, device=0, **kw):          # stand-in for the GPU call: native decode + oracle
        return ref.process_audios([pkg.audio_io.read_audio(p)[0] for p in audio_path], args)
    f2, l2 = pkg.sharding.process_audios_sharded(paths, args, process_fn=process_files)
    if rank == 0:
        ok = ok and l2 == want_len and all(np.array_equal(a, b) for a, b in zip(f2, want))
    else:
        ok = ok and f2 is None
    args.feat_dir = os.path.join(tmp, "features")
    for threshold, k in ((100, 1), (4, 4)):                       # one pickle / the reference's chunked layout
        got = pkg.sharding.process_libri_feats_sharded(paths, "train-%d" % k, k, args, process_fn=process_files, threshold=threshold)
        dist.barrier()
        if rank == 0:
            ok = ok and got == want_len and np.load(args.feat_dir + "/train-%d-featlen.npy" % k).tolist() == want_len
            if k == 1:
                back = list(joblib.load(args.feat_dir + "/train-1-feats.pkl"))
            else:
                n = len(paths) // k + 1                           # chunk boundaries of preprocess.py:118
                back = []
                for i in range(k):
                    chunk = joblib.load(args.feat_dir + "/train-%d-feats-%d.pkl" % (k, i))
                    ok = ok and len(chunk) == len(paths[i * n:(i + 1) * n])
                    back += list(chunk)
            ok = ok and all(np.array_equal(a, b) for a, b in zip(back, want))
        else:
            ok = ok and got is None
    q.put(bool(ok))
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_extract_gloo_world2(tmp_path):
    import torch.multiprocessing as mp
    os.environ["ASR_TEST_TMP"] = str(tmp_path)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=180) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert res == [True, True]
