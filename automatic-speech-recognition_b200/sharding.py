"""Length-balanced partition of utterances across GPUs (one process per GPU).

Per-utterance CMVN needs no global statistic, so the front-end shards by utterance
with NO data-path collective: every rank extracts its own shard and results are
re-assembled by the inverse permutation.  ``torch.distributed`` is only used to
agree on timing / gather small Python objects."""
import heapq

import numpy as np


def frame_counts(lengths, frame_len=400, hop=160):
    lengths = np.asarray(lengths, dtype=np.int64)
    return np.maximum((lengths - frame_len) // hop, 0) * (lengths >= frame_len)


def lpt_partition(costs, world_size):
    """Longest-processing-time-first greedy: returns a list of index arrays, one per
    rank (deterministic: ties broken by index, ranks by lowest load then id)."""
    costs = np.asarray(costs, dtype=np.int64)
    order = np.lexsort((np.arange(costs.size), -costs))
    heap = [(0, r) for r in range(world_size)]
    heapq.heapify(heap)
    bins = [[] for _ in range(world_size)]
    for i in order:
        load, r = heapq.heappop(heap)
        bins[r].append(int(i))
        heapq.heappush(heap, (load + int(costs[i]), r))
    return [np.asarray(sorted(b), dtype=np.int64) for b in bins]


def shard_indices(lengths, rank, world_size, frame_len=400, hop=160):
    """Indices of the utterances rank ``rank`` processes (ascending, so per-shard output
    order is input order)."""
    if world_size == 1:
        return np.arange(len(lengths), dtype=np.int64)
    return lpt_partition(frame_counts(lengths, frame_len, hop) + 1, world_size)[rank]


def imbalance(lengths, world_size, frame_len=400, hop=160):
    """max shard cost / mean shard cost - 1 for the LPT partition."""
    costs = frame_counts(lengths, frame_len, hop) + 1
    parts = lpt_partition(costs, world_size)
    loads = np.asarray([costs[p].sum() for p in parts], dtype=np.float64)
    return float(loads.max() / loads.mean() - 1.0)


def merge_shards(parts, shard_results, n_total):
    """Inverse permutation: ``shard_results[r][j]`` belongs to utterance ``parts[r][j]``."""
    out = [None] * n_total
    for idx, res in zip(parts, shard_results):
        if len(idx) != len(res):
            raise ValueError("shard size mismatch")
        for i, r in zip(idx, res):
            out[int(i)] = r
    if any(o is None for o in out):
        raise ValueError("shards do not cover the batch")
    return out


def _gather_cubes(feats, dist, rank, world, gather_to):
    """Hand every rank's cubes to rank ``gather_to`` WITHOUT pickling them through the process group: a rank writes its
    shard as one flat float32 file in shared memory (/dev/shm; one node, as the whole path is), only the file name and the
    shapes (a few ints per utterance) go through ``gather_object``; the receiver maps the files and returns views --
    one copy on the sending side, none on the receiving side.  Returns the per-rank lists of cubes on ``gather_to``."""
    import os
    import tempfile
    import uuid
    tag = [uuid.uuid4().hex if rank == gather_to else None]
    dist.broadcast_object_list(tag, src=gather_to)
    shm = "/dev/shm" if os.path.isdir("/dev/shm") and os.access("/dev/shm", os.W_OK) else tempfile.gettempdir()
    shapes = [tuple(f.shape) for f in feats]
    sizes = [int(np.prod(sh)) for sh in shapes]
    total = int(sum(sizes))
    path = None
    if rank != gather_to and total > 0:
        path = os.path.join(shm, "asr_b200_%s_%d.f32" % (tag[0], rank))
        mm = np.memmap(path, dtype=np.float32, mode="w+", shape=(total,))
        o = 0
        for f, n in zip(feats, sizes):
            mm[o:o + n] = np.asarray(f, dtype=np.float32).reshape(-1)
            o += n
        mm.flush()
        del mm
    gathered = [None] * world if rank == gather_to else None
    dist.gather_object((path, shapes), gathered, dst=gather_to)
    if rank != gather_to:
        return None
    out = []
    for r, (pth, shp) in enumerate(gathered):
        if r == gather_to:
            out.append(list(feats))
            continue
        if pth is None:
            out.append([np.empty(sh, np.float32) for sh in shp])
            continue
        flat = np.asarray(np.memmap(pth, dtype=np.float32, mode="r+"))
        os.unlink(pth)                                   # the mapping keeps the pages alive for as long as the views live
        cubes, o = [], 0
        for sh in shp:
            n = int(np.prod(sh))
            cubes.append(flat[o:o + n].reshape(sh))
            o += n
        out.append(cubes)
    return out


def process_pcm_sharded(pcm_list, args, fs=16000, gather_to=0, extract_fn=None, **switches):
    """Every rank calls this with the SAME ``pcm_list``; each extracts its LPT shard on
    its own GPU (LOCAL_RANK) and rank ``gather_to`` gets the re-assembled
    (feats, featlen); other ranks get (None, None).  ``extract_fn`` (same signature as
    process_pcm) is the per-shard feature call; tests inject a stand-in to exercise the
    partition / gather logic without a GPU.  Falls back to a single shard when
    torch.distributed is not initialised."""
    import os
    from .preprocess import process_pcm, to_object_array
    try:
        import torch.distributed as dist
        live = dist.is_available() and dist.is_initialized()
    except ImportError:
        dist, live = None, False
    rank = dist.get_rank() if live else 0
    world = dist.get_world_size() if live else 1
    device = int(os.environ.get("LOCAL_RANK", 0))
    lengths = [len(p) for p in pcm_list]
    frame_len = int(round(fs * args.frame_length / 1000.0))
    hop = int(round(fs * args.frame_step / 1000.0))
    parts = lpt_partition(frame_counts(lengths, frame_len, hop) + 1, world) if world > 1 else \
        [np.arange(len(pcm_list), dtype=np.int64)]
    mine = [pcm_list[int(i)] for i in parts[rank]]
    extract = extract_fn or process_pcm
    feats, _ = extract(mine, args, fs=fs, device=device, **switches) if mine else (to_object_array([]), [])
    if world == 1:
        return feats, [len(f) for f in feats]
    gathered = _gather_cubes(feats, dist, rank, world, gather_to)
    if rank != gather_to:
        return None, None
    merged = merge_shards(parts, gathered, len(pcm_list))
    return to_object_array(merged), [len(m) for m in merged]


def _dist():
    try:
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized():
            return dist, dist.get_rank(), dist.get_world_size()
    except ImportError:
        pass
    return None, 0, 1


def process_audios_sharded(audio_path, args, gather_to=0, process_fn=None, lengths=None, **kw):
    """File-list variant of ``process_pcm_sharded``: every rank calls it with the SAME ``audio_path``; the
    files are LPT-partitioned by their sample counts (header probe, no decoding), each rank runs
    ``process_audios`` on its shard on its own GPU, rank ``gather_to`` receives (feats, featlen) in input
    order, the others (None, None).  ``process_fn`` / ``lengths`` let tests stand in for the GPU call and
    the probe."""
    import os
    from . import audio_io
    from .preprocess import process_audios, to_object_array
    dist, rank, world = _dist()
    device = int(os.environ.get("LOCAL_RANK", 0))
    fn = process_fn or process_audios
    if world == 1:
        return fn(list(audio_path), args, device=device, **kw)
    if lengths is None:
        lengths = [max(i["n_samples"], 0) for i in audio_io.probe_batch(list(audio_path))]
    parts = lpt_partition(frame_counts(lengths) + 1, world)
    mine = [audio_path[int(i)] for i in parts[rank]]
    feats, _ = fn(mine, args, device=device, **kw) if mine else (to_object_array([]), [])
    gathered = _gather_cubes(feats, dist, rank, world, gather_to)
    if rank != gather_to:
        return None, None
    merged = merge_shards(parts, gathered, len(audio_path))
    return to_object_array(merged), [len(m) for m in merged]


def process_libri_feats_sharded(audio_path, cat, k, args, process_fn=None, threshold=None, **kw):
    """preprocess.py:112-130 over several GPUs with the SAME files on disk as the single-process run.

    Sets above the reference's 30 000-file threshold are written as k chunk pickles with the reference's
    boundaries (``n = len // k + 1``): rank r extracts and writes chunks r, r + world, ... -- no feature
    leaves its rank; only the per-chunk frame counts (ints) are exchanged so that rank 0 can write
    ``{cat}-featlen.npy`` in file order.  Smaller sets become one ``{cat}-feats.pkl`` written by rank 0
    from the LPT-sharded, gathered result.  Returns featlen on rank 0, None elsewhere."""
    import os
    import joblib
    from . import preprocess as pp
    dist, rank, world = _dist()
    device = int(os.environ.get("LOCAL_RANK", 0))
    fn = process_fn or pp.process_audios
    threshold = pp._SAMPLE_THRESHOLD if threshold is None else threshold
    os.makedirs(args.feat_dir, exist_ok=True)
    if len(audio_path) <= threshold:
        feats, featlen = process_audios_sharded(audio_path, args, 0, process_fn, **kw)
        if rank == 0:
            joblib.dump(feats, args.feat_dir + "/{}-feats.pkl".format(cat))
            np.save(args.feat_dir + "/{}-featlen.npy".format(cat), featlen)
            return featlen
        return None
    n = len(audio_path) // k + 1
    mine = {}
    for i in range(rank, k, world):
        feats, flen = fn(audio_path[i * n:(i + 1) * n], args, device=device, **kw)
        joblib.dump(feats, args.feat_dir + "/{}-feats-{}.pkl".format(cat, i))
        mine[i] = list(flen)
    if world > 1:
        allc = [None] * world if rank == 0 else None
        dist.gather_object(mine, allc, dst=0)
        if rank != 0:
            return None
        mine = {i: v for d in allc for i, v in d.items()}
    featlen = [x for i in range(k) for x in mine[i]]
    np.save(args.feat_dir + "/{}-featlen.npy".format(cat), featlen)
    return featlen
