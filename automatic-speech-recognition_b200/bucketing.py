"""Length-bucketed, padded batches built on the GPU: the stage right after the feature path
(/root/reference/tfrecord_data_loader.py:54-106, ``bucket_by_sequence_length`` with
``pad_to_bucket_boundary=True``), without the TFRecord round trip.

The cubes ``fe_run`` left in HBM (or host arrays) are scattered by ONE kernel launch
(``fe_pad_batches`` -> ``k_pad_slots``) into dense ``[B, boundary - 1, D, 3]`` batch tensors with
zero padding; which utterance goes where is planned on the host with the reference's bucket
boundaries and batch sizes.  ``bucketed_batches`` yields what ``iterator.get_next()`` yields in
train.py: ``((feat, featlen), (token, tokenlen))``."""
import ctypes as C

import numpy as np

from .frontend import _device_ptr, _ptr

BUCKETS_TRAIN = [639, 1062, 1275, 1377, 1449, 1506, 1563, 1710]      # tfrecord_data_loader.py:77
BUCKETS_EVAL = [639, 1062, 1275, 1377, 1449, 1506, 1563, 3600]       # :82
BATCH_LIMIT = [96, 48, 48, 48, 48, 48, 48, 48, 48]                   # :85
MAX_TOKENLEN_TRAIN, MAX_TOKENLEN_EVAL = 219, 227                      # :78, :83


def plan_batches(featlen, boundaries=None, batch_sizes=None, drop_long=False):
    """Host plan of the batching: list of (bucket id, int64 index array) in emission order.

    Buckets are [0, b0), [b0, b1), ...; a bucket's window is emitted when it holds its batch size,
    partial windows at the end of the input in ascending bucket order.  An utterance with
    L >= the last boundary is an error, as in TF with pad_to_bucket_boundary (``drop_long=True``
    skips it instead, like create_tfrecord.py:134-136 does before writing)."""
    boundaries = np.asarray(BUCKETS_TRAIN if boundaries is None else boundaries, dtype=np.int64)
    batch_sizes = BATCH_LIMIT if batch_sizes is None else batch_sizes
    if len(batch_sizes) != len(boundaries) + 1:
        raise ValueError("len(bucket_batch_sizes) must equal len(bucket_boundaries) + 1")
    featlen = np.asarray(featlen, dtype=np.int64)
    ids = np.searchsorted(boundaries, featlen, side="right")
    windows = {}
    out = []
    for i, b in enumerate(ids.tolist()):
        if b >= len(boundaries):
            if drop_long:
                continue
            raise ValueError("element %d: length %d >= the last bucket boundary %d" % (i, featlen[i], boundaries[-1]))
        w = windows.setdefault(b, [])
        w.append(i)
        if len(w) == batch_sizes[b]:
            out.append((b, np.asarray(w, dtype=np.int64)))
            del windows[b]
    for b in sorted(windows):
        out.append((b, np.asarray(windows[b], dtype=np.int64)))
    return out


def pad_tokens(tokens, idx, max_tokenlen):
    tok = np.zeros((len(idx), max_tokenlen), dtype=np.int32)
    lens = np.zeros(len(idx), dtype=np.int32)
    for k, i in enumerate(idx):
        t = np.asarray(tokens[int(i)], dtype=np.int32)
        if t.size > max_tokenlen:
            raise ValueError("element %d: %d tokens > max_tokenlen %d" % (i, t.size, max_tokenlen))
        tok[k, :t.size] = t
        lens[k] = t.size
    return tok, lens


class BucketBatcher:
    """Scatter a flat feature buffer into padded batch tensors with one kernel launch."""

    def __init__(self, frontend, boundaries=None, batch_sizes=None):
        self.fe = frontend
        self.boundaries = list(BUCKETS_TRAIN if boundaries is None else boundaries)
        self.batch_sizes = list(BATCH_LIMIT if batch_sizes is None else batch_sizes)

    def layout(self, plan, n_frames, row_floats):
        """Slot tables for a plan: (src index, valid floats, dst offsets, slot floats, batch bases, total)."""
        n_frames = np.asarray(n_frames, dtype=np.int64)
        src, valid, dst, slot, bases = [], [], [], [], []
        off = 0
        for b, idx in plan:
            T = self.boundaries[b] - 1
            sf = T * row_floats
            off = (off + 3) // 4 * 4                       # every batch tensor starts 16-byte aligned
            bases.append(off)
            src.append(idx)
            valid.append(n_frames[idx] * row_floats)
            dst.append(off + np.arange(len(idx), dtype=np.int64) * sf)
            slot.append(np.full(len(idx), sf, dtype=np.int64))
            off += len(idx) * sf
        cat = (lambda xs, dt: np.concatenate(xs).astype(dt) if xs else np.zeros(0, dt))
        return cat(src, np.int64), cat(valid, np.int32), cat(dst, np.int64), cat(slot, np.int32), bases, off

    def pad(self, feats, feat_offsets, n_frames, row_floats, plan, out=None, stream=None):
        """feats: flat float32 buffer (CUDA tensor or numpy), utterance i at feat_offsets[i] with
        n_frames[i] rows of row_floats (= D * 3) floats.  Returns (list of [B, T_pad, row_floats] views
        -- reshape to (B, T_pad, D, 3) -- , flat batch buffer)."""
        src, valid, dst, slot, bases, total = self.layout(plan, n_frames, row_floats)
        if np.any(slot.astype(np.int64) >= 2 ** 31):
            raise ValueError("slot too large")
        feat_offsets = np.asarray(feat_offsets, dtype=np.int64)
        src_off = np.ascontiguousarray(feat_offsets[src], dtype=np.int64)
        dev = _device_ptr(feats)
        if out is None:
            if dev is not None:
                import torch
                out = torch.empty(max(total, 1), dtype=torch.float32, device=feats.device)
            else:
                out = np.empty(max(total, 1), dtype=np.float32)
        out_dev = _device_ptr(out)
        in_ptr = dev if dev is not None else np.ascontiguousarray(feats, dtype=np.float32).ctypes.data
        out_ptr = out_dev if out_dev is not None else out.ctypes.data
        cap = int(out.numel()) if hasattr(out, "numel") else int(out.size)
        self.fe._check(self.fe._lib.fe_pad_batches(
            self.fe._h, C.c_void_p(in_ptr), _ptr(src_off, C.c_int64), _ptr(valid, C.c_int32), _ptr(dst, C.c_int64),
            _ptr(slot, C.c_int32), int(src.size), C.c_void_p(out_ptr), cap,
            C.c_void_p(int(stream)) if stream else None), "fe_pad_batches")
        if stream is None:
            self.fe.sync()                 # device in / device out runs asynchronously on the handle's own stream
        views = []
        for (b, idx), base in zip(plan, bases):
            T = self.boundaries[b] - 1
            views.append(out[base:base + len(idx) * T * row_floats].reshape(len(idx), T, row_floats))
        return views, out

    def pad_ms(self):
        v = C.c_float()
        self.fe._check(self.fe._lib.fe_get_pad_ms(self.fe._h, C.byref(v)), "fe_get_pad_ms")
        return float(v.value)


def bucketed_batches(frontend, feats, feat_offsets, n_frames, tokens, feat_dim, is_training=True, drop_long=False):
    """Generator over ``((feat [B, T_pad, D, 3], featlen [B]), (token [B, max_tokenlen], tokenlen [B]))`` --
    one pass over the set in the reference's bucket geometry (tfrecord_data_loader.py:75-94)."""
    boundaries = BUCKETS_TRAIN if is_training else BUCKETS_EVAL
    max_tok = MAX_TOKENLEN_TRAIN if is_training else MAX_TOKENLEN_EVAL
    bb = BucketBatcher(frontend, boundaries, BATCH_LIMIT)
    plan = plan_batches(n_frames, boundaries, BATCH_LIMIT, drop_long=drop_long)
    views, _ = bb.pad(feats, feat_offsets, n_frames, feat_dim * 3, plan)
    n_frames = np.asarray(n_frames)
    for (b, idx), v in zip(plan, views):
        tok, toklen = pad_tokens(tokens, idx, max_tok)
        yield (v.reshape(v.shape[0], v.shape[1], feat_dim, 3), n_frames[idx].astype(np.int32)), (tok, toklen)
