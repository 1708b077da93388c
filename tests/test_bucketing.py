"""Length-bucketed padded batches (tfrecord_data_loader.py:54-106): host plan against the CPU restatement
(oracle/bucket_ref.py) here, the padding kernel (fe_pad_batches) against it on the GPU -- bit-exact copies."""
import importlib

import numpy as np
import pytest

from conftest import PKG


@pytest.fixture(scope="module")
def bk(pkg):
    return importlib.import_module(PKG + ".bucketing")


@pytest.fixture(scope="module")
def bref():
    from oracle import bucket_ref
    return bucket_ref


def test_constants_are_the_references(bk, bref):
    assert bk.BUCKETS_TRAIN == bref.BUCKETS_TRAIN == [639, 1062, 1275, 1377, 1449, 1506, 1563, 1710]
    assert bk.BUCKETS_EVAL[-1] == 3600 and bk.BATCH_LIMIT == [96] + [48] * 8
    assert (bk.MAX_TOKENLEN_TRAIN, bk.MAX_TOKENLEN_EVAL) == (219, 227)


def test_plan_matches_restatement(bk, bref):
    rng = np.random.default_rng(0)
    for n in (0, 1, 50, 3000):
        L = rng.integers(1, 1710, n)
        got = bk.plan_batches(L, bk.BUCKETS_TRAIN, bk.BATCH_LIMIT)
        want = bref.plan(L, bref.BUCKETS_TRAIN, bref.BATCH_LIMIT)
        assert [(b, i.tolist()) for b, i in got] == [(b, list(i)) for b, i in want]
        assert sorted(np.concatenate([i for _, i in got]).tolist() if got else []) == list(range(n))
        for b, idx in got:
            lo = 0 if b == 0 else bk.BUCKETS_TRAIN[b - 1]
            assert np.all((L[idx] >= lo) & (L[idx] < bk.BUCKETS_TRAIN[b])) and len(idx) <= bk.BATCH_LIMIT[b]
    edges = np.array([638, 639, 1061, 1062, 1709])
    assert [b for b, _ in bk.plan_batches(edges)] == [0, 1, 2, 7]            # boundaries are exclusive upper bounds
    with pytest.raises(ValueError, match="last bucket boundary"):
        bk.plan_batches([100, 1710])                                          # TF errors; create_tfrecord.py:134 drops these
    assert sum(len(i) for _, i in bk.plan_batches([100, 1710, 5000], drop_long=True)) == 1
    assert len(bk.plan_batches([3262, 3493], bk.BUCKETS_EVAL)) == 1           # dev / test maxima (:81-82) fit the eval buckets


def test_layout_is_dense_and_aligned(bk):
    rng = np.random.default_rng(1)
    L = rng.integers(1, 1710, 500)
    plan = bk.plan_batches(L)
    bb = bk.BucketBatcher(None)
    src, valid, dst, slot, bases, total = bb.layout(plan, L, 39)
    assert np.all(np.asarray(bases) % 4 == 0) and np.all(valid <= slot)
    order = np.argsort(dst)
    assert np.all(dst[order][1:] >= (dst + slot)[order][:-1]) and (dst + slot).max() == total
    assert np.array_equal(valid, L[src] * 39)


def _cubes(pkg, n, D, seed, lo=1, hi=1709):
    rng = np.random.default_rng(seed)
    L = rng.integers(lo, hi, n)
    off = np.zeros(n + 1, np.int64)
    np.cumsum((L * D * 3 + 3) // 4 * 4, out=off[1:])
    flat = rng.standard_normal(int(off[-1])).astype(np.float32)
    feats = [flat[off[i]:off[i] + L[i] * D * 3].reshape(L[i], D, 3) for i in range(n)]
    tokens = [rng.integers(1, 5000, rng.integers(1, 200)).tolist() for _ in range(n)]
    return flat, off[:-1], L.astype(np.int32), feats, tokens


@pytest.mark.gpu
@pytest.mark.parametrize("D", [13, 80])
def test_pad_kernel_matches_restatement(pkg, bk, bref, D):
    import torch
    flat, off, L, feats, tokens = _cubes(pkg, 260, D, seed=D)
    want = bref.batches(feats, tokens, bref.BUCKETS_TRAIN, bref.BATCH_LIMIT, bref.MAX_TOKENLEN_TRAIN)
    fe = pkg.Frontend(pkg.FrontendConfig())
    for resident in (True, False):
        src = torch.from_numpy(flat).cuda() if resident else flat
        got = list(bk.bucketed_batches(fe, src, off, L, tokens, D, is_training=True))
        assert len(got) == len(want)
        for ((x, xl), (t, tl)), ((wx, wxl), (wt, wtl)) in zip(got, want):
            x = x.cpu().numpy() if resident else x
            assert x.shape == wx.shape and x.dtype == np.float32
            assert np.array_equal(x, wx)                                      # bit-exact copy + exact zeros
            assert np.array_equal(xl, wxl) and np.array_equal(t, wt) and np.array_equal(tl, wtl)
    fe.close()


@pytest.mark.gpu
def test_pad_after_fe_run_stays_on_device(pkg, bk, bref, ref):
    """PCM -> fe_run -> fe_pad_batches without leaving HBM; equals oracle features padded by the restatement."""
    import torch
    from conftest import assert_close
    pcm = pkg.synth.corpus(12, 1.0, 9.0, seed=77)
    fe = pkg.Frontend(pkg.FrontendConfig())
    packed, off, lens = pkg.pack_pcm(pcm)
    out, out_off, nfr = fe.run_packed(torch.from_numpy(packed).cuda(), off, lens)
    plan = bk.plan_batches(nfr)
    bb = bk.BucketBatcher(fe)
    fe.set_profiling(True)
    views, _ = bb.pad(out, out_off[:-1], nfr, 39, plan)
    assert bb.pad_ms() > 0
    want = [ref.features_one(p) for p in pcm]
    for (b, idx), v in zip(plan, views):
        v = v.cpu().numpy().reshape(len(idx), -1, 13, 3)
        assert v.shape[1] == bk.BUCKETS_TRAIN[b] - 1
        for k, i in enumerate(idx):
            assert_close(v[k, :nfr[i]], want[i], what="padded batch")
            assert not v[k, nfr[i]:].any()
    with pytest.raises(RuntimeError, match="longer than its slot"):
        bk.BucketBatcher(fe, [10], [4, 4]).pad(out, out_off[:-1], nfr, 39, [(0, np.arange(2))])
    fe.close()
