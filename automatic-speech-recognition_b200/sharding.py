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
    gathered = [None] * world if rank == gather_to else None
    dist.gather_object(list(feats), gathered, dst=gather_to)
    if rank != gather_to:
        return None, None
    merged = merge_shards(parts, gathered, len(pcm_list))
    return to_object_array(merged), [len(m) for m in merged]
