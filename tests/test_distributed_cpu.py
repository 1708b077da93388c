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
    if rank == 0:
        want, want_len = ref.process_audios(pcm, args)
        ok = featlen == want_len and all(np.array_equal(a, b) for a, b in zip(feats, want))
        q.put(bool(ok))
    else:
        q.put(feats is None and featlen is None)
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_extract_gloo_world2():
    import torch.multiprocessing as mp
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
