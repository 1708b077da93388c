"""CPU restatement of the batching stage of /root/reference/tfrecord_data_loader.py:54-106
(``tf.data.experimental.bucket_by_sequence_length`` with ``pad_to_bucket_boundary=True``).

TEST INFRASTRUCTURE ONLY: imported by tests/ (the checker for fe_pad_batches), never by the product.

PARITY UNPINNED against TensorFlow itself: TF 1.13 is not installable here and the reference holds
no fixture for this stage; the function below restates the documented semantics of
bucket_by_sequence_length / group_by_window / padded_batch:

* bucket id of an element of length L = number of boundaries b with b <= L, i.e. buckets are
  [0, b0), [b0, b1), ..., [b_last, inf)                          (tfrecord_data_loader.py:75-83);
* an element joins the open window of its bucket; when the window holds ``bucket_batch_sizes[id]``
  elements it is emitted as one batch, in arrival order; at the end of the input the partial
  windows are emitted (here: in ascending bucket id -- TF iterates a hash map, order unspecified);
* with pad_to_bucket_boundary=True the time axis is padded to ``boundary[id] - 1`` (the longest
  length the bucket admits) and an element of the overflow bucket (L >= last boundary) is an error
  (create_tfrecord.py:28,134-136 drops L >= 1710 before writing for that reason);
* tokens are padded with zeros to ``max_tokenlen`` (padded_shapes, tfrecord_data_loader.py:86)."""
import numpy as np

BUCKETS_TRAIN = [639, 1062, 1275, 1377, 1449, 1506, 1563, 1710]      # tfrecord_data_loader.py:77
BUCKETS_EVAL = [639, 1062, 1275, 1377, 1449, 1506, 1563, 3600]       # :82
BATCH_LIMIT = [96, 48, 48, 48, 48, 48, 48, 48, 48]                   # :85
MAX_TOKENLEN_TRAIN, MAX_TOKENLEN_EVAL = 219, 227                      # :78, :83


def bucket_id(length, boundaries):
    return int(np.searchsorted(np.asarray(boundaries), length, side="right"))


def plan(featlen, boundaries, batch_sizes):
    """-> list of (bucket id, [element indices]) in emission order."""
    open_windows = {}
    out = []
    for i, L in enumerate(featlen):
        b = bucket_id(int(L), boundaries)
        if b >= len(boundaries):
            raise ValueError("element %d: length %d >= the last bucket boundary %d" % (i, L, boundaries[-1]))
        w = open_windows.setdefault(b, [])
        w.append(i)
        if len(w) == batch_sizes[b]:
            out.append((b, w))
            del open_windows[b]
    for b in sorted(open_windows):
        out.append((b, open_windows[b]))
    return out


def batches(feats, tokens, boundaries, batch_sizes, max_tokenlen):
    """-> list of ((feat [B, T_pad, D, 3] float32, featlen [B] int32), (token [B, max_tokenlen] int32, tokenlen [B] int32))."""
    featlen = [len(f) for f in feats]
    res = []
    for b, idx in plan(featlen, boundaries, batch_sizes):
        T = boundaries[b] - 1
        tail = feats[idx[0]].shape[1:]
        x = np.zeros((len(idx), T) + tail, dtype=np.float32)
        tok = np.zeros((len(idx), max_tokenlen), dtype=np.int32)
        for k, i in enumerate(idx):
            x[k, :featlen[i]] = feats[i]
            t = np.asarray(tokens[i], dtype=np.int32)
            if len(t) > max_tokenlen:
                raise ValueError("element %d: %d tokens > max_tokenlen %d" % (i, len(t), max_tokenlen))
            tok[k, :len(t)] = t
        res.append(((x, np.asarray([featlen[i] for i in idx], np.int32)),
                    (tok, np.asarray([len(tokens[i]) for i in idx], np.int32))))
    return res
