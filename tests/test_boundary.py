"""Host-side logic of the drop-in boundary: packing, the object-array / pickle format the
reference's consumers read, audio I/O, augmentation naming, length-balanced sharding."""
import importlib
import os

import joblib
import numpy as np
import pytest

from conftest import PKG, make_args


def test_pack_pcm_alignment(pkg):
    rng = np.random.default_rng(0)
    lst = [rng.integers(-100, 100, size=n, dtype=np.int16) for n in (401, 8, 1603, 77)]
    packed, off, lens = pkg.pack_pcm(lst)
    assert lens.tolist() == [401, 8, 1603, 77] and np.all(off % 8 == 0)
    for p, o, n in zip(lst, off, lens):
        assert np.array_equal(packed[o:o + n], p)
    pf, off_f, _ = pkg.pack_pcm([x.astype(np.float32) for x in lst], np.float32)
    assert np.all(off_f % 4 == 0) and pf.dtype == np.float32


def test_object_array_and_pickle_roundtrip(pkg, tmp_path):
    cubes = [np.random.rand(L, 13, 3).astype(np.float32) for L in (5, 5, 9)]   # equal lengths must not collapse
    arr = pkg.to_object_array(cubes)
    assert arr.dtype == object and arr.shape == (3,)
    joblib.dump(arr, tmp_path / "dev-feats.pkl")
    back = joblib.load(tmp_path / "dev-feats.pkl")
    # what create_tfrecord.py:37-38,84-85,131-137 and decode.py:82,123 do with it
    feats = np.append([], back)
    perm = np.random.default_rng(0).permutation(len(feats))
    feats = feats[perm]
    for f, src in zip(feats, [cubes[i] for i in perm]):
        assert len(f.shape) == 3 and f.flatten().shape == (f.shape[0] * 39,) and np.array_equal(f, src)
        assert np.expand_dims(f, 0).shape == (1,) + src.shape
    keep = np.array([len(f) for f in feats]) < 1710
    assert feats[keep].shape == (3,)


def test_audio_io_wav_roundtrip(pkg, tmp_path):
    x = pkg.synth.corpus(1, 0.2, 0.3, seed=1)[0]
    pkg.audio_io.write_audio(str(tmp_path / "a.wav"), x, 16000)
    y, fs = pkg.audio_io.read_audio(str(tmp_path / "a.wav"))
    assert fs == 16000 and y.dtype == np.int16 and np.array_equal(x, y)
    np.save(tmp_path / "b.npy", x)
    z, _ = pkg.audio_io.read_audio(str(tmp_path / "b.npy"))
    assert np.array_equal(z, x)
    with pytest.raises((RuntimeError, Exception)):
        pkg.audio_io.read_audio(str(tmp_path / "missing.xyz"))


def test_augmentation_file_interface(pkg, sox, tmp_path, monkeypatch):
    """Naming / skip-if-exists of utils/augmentation.py:19,24-26,52 (kernel replaced by the oracle here)."""
    aug = importlib.import_module(PKG + ".augmentation")

    class Fake:
        def perturb(self, pcm, speeds=None, gains=None):
            if speeds is not None:
                return [sox.speed_perturb(p, s) for p, s in zip(pcm, speeds)]
            return [sox.volume_perturb(p, g) for p, g in zip(pcm, gains)]

        def perturb_packed(self, packed, off, lens, speeds=None, gains=None):
            ys = self.perturb([packed[o:o + n] for o, n in zip(off, lens)], speeds, gains)
            fr = importlib.import_module(PKG + ".frontend")
            dst, d_off, d_len = fr.pack_pcm(ys)
            return dst, np.concatenate((d_off, [dst.size])), d_len

        def close(self):
            pass
    monkeypatch.setattr(aug, "_frontend", lambda speed=None, device=0: Fake())
    src = []
    for i, x in enumerate(pkg.synth.corpus(2, 0.2, 0.3, seed=2)):
        p = str(tmp_path / ("1-2-%04d.flac" % i))
        pkg.audio_io.write_audio(p, x, 16000)
        src.append(p)
    out = aug.SpeedAugmentation(src, str(tmp_path / "LibriSpeech_speed_aug"), 0.9)
    assert out == [str(tmp_path / "LibriSpeech_speed_aug_0.9" / ("1-2-%04d_0.9.flac" % i)) for i in range(2)]
    y, _ = pkg.audio_io.read_audio(out[0])
    x, _ = pkg.audio_io.read_audio(src[0])
    assert len(y) == -(-len(x) * 10 // 9)
    os.utime(out[0], (1, 1))
    aug.SpeedAugmentation(src, str(tmp_path / "LibriSpeech_speed_aug"), 0.9)          # existing files are skipped
    assert os.path.getmtime(out[0]) == 1
    vout = aug.VolumeAugmentation(src, str(tmp_path / "vol"), [0.8, 1.5], rng=np.random.default_rng(3))
    g = float(os.path.basename(vout[0]).rsplit("_", 1)[1][:-5])
    assert 0.8 <= g <= 1.5 and round(g, 2) == g
    assert np.array_equal(pkg.audio_io.read_audio(vout[0])[0], sox.volume_perturb(x, g))


def test_lpt_partition(pkg):
    sh = pkg.sharding
    rng = np.random.default_rng(0)
    lens = pkg.synth.durations(5000, 2, 15, rng)
    for world in (1, 2, 4, 8):
        parts = sh.lpt_partition(sh.frame_counts(lens) + 1, world)
        allidx = np.concatenate(parts)
        assert sorted(allidx.tolist()) == list(range(5000))
        assert all(np.all(np.diff(p) > 0) for p in parts)
        assert sh.imbalance(lens, world) < 0.01
        assert all(np.array_equal(p, sh.shard_indices(lens, r, world)) for r, p in enumerate(parts))
    merged = sh.merge_shards(parts, [[int(i) * 10 for i in p] for p in parts], 5000)
    assert merged == [i * 10 for i in range(5000)]
    assert sh.frame_counts([399, 400, 559, 560, 16000]).tolist() == [0, 0, 0, 1, 97]


def test_file_batch_planning(pkg):
    """Host batching of process_audios: consecutive ranges, every file exactly once, ~one audio-hour per batch."""
    pp = importlib.import_module(PKG + ".preprocess")
    rng = np.random.default_rng(0)
    lens = rng.integers(32000, 560000, 1000).tolist()
    ranges = pp._plan_file_batches(lens, 57_600_000)
    assert ranges[0][0] == 0 and ranges[-1][1] == 1000 and all(a[1] == b[0] for a, b in zip(ranges, ranges[1:]))
    sums = [sum(lens[lo:hi]) for lo, hi in ranges]
    assert all(s >= 57_600_000 for s in sums[:-1]) and all(s < 57_600_000 + 560000 for s in sums)
    assert pp._plan_file_batches([10], 100) == [(0, 1)] and pp._plan_file_batches([200, 5], 100) == [(0, 1), (1, 2)]
    off, total = pkg.audio_io.plan_batch([5, 8, 9])
    assert off.tolist() == [0, 8, 16] and total == 32


def test_epoch_augmenter_draws_are_reproducible(pkg):
    aug = importlib.import_module(PKG + ".augmentation")
    a = aug.EpochAugmenter(None, speeds=(0.9, 1.0, 1.1), vol_range=(0.8, 1.5), seed=7)
    s0, g0 = a.draw(1000, epoch=0)
    s0b, g0b = a.draw(1000, epoch=0)
    s1, g1 = a.draw(1000, epoch=1)
    assert np.array_equal(s0, s0b) and np.array_equal(g0, g0b) and not np.array_equal(s0, s1)
    assert set(np.unique(s0)) == {0.9, 1.0, 1.1} and abs((s0 == 1.0).mean() - 1 / 3) < 0.06
    assert g0.min() >= 0.8 and g0.max() <= 1.5 and np.array_equal(np.around(g0, 2), g0)     # utils/augmentation.py:48-49
    assert aug.EpochAugmenter(None, seed=7).draw(5, 3)[1] is None
