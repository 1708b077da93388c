"""The reference's offline pipeline end to end on a synthetic LibriSpeech-shaped tree, with only the hot path and
its two neighbours replaced: FLAC files -> (main_libri's calls, /root/reference/preprocess.py:110-179) features +
featlen + tokens on disk -> (create_tfrecord.py:100-140) TFRecords -> (tfrecord_data_loader.py:54-106) bucketed
batches with the shapes train.py feeds the model."""
import glob
import importlib
import os

import numpy as np
import pytest

from conftest import PKG, make_args, assert_close

pytestmark = pytest.mark.gpu


def _make_tree(pkg, root, n_spk=3, n_utt=5):
    """root/{speaker}/{chapter}/{spk}-{chap}-{k:04d}.flac + {spk}-{chap}.trans.txt, as LibriSpeech lays it out."""
    rng = np.random.default_rng(5)
    pcm = pkg.synth.corpus(n_spk * n_utt, 1.0, 5.0, seed=99)
    words = ["THE", "QUICK", "BROWN", "FOX", "JUMPS", "OVER", "A", "LAZY", "DOG", "IT'S"]
    k, truth = 0, {}
    for s in range(n_spk):
        d = os.path.join(root, str(100 + s), str(2000 + s))
        os.makedirs(d)
        lines = []
        for u in range(n_utt):
            uid = "%d-%d-%04d" % (100 + s, 2000 + s, u)
            pkg.audio_io.write_audio(os.path.join(d, uid + ".flac"), pcm[k], 16000)
            text = " ".join(rng.choice(words, rng.integers(2, 7)))
            lines.append(uid + " " + text)
            truth[os.path.join(d, uid + ".flac")] = (pcm[k], text.replace("'", ""))
            k += 1
        open(os.path.join(d, "%d-%d.trans.txt" % (100 + s, 2000 + s)), "w").write("\n".join(lines) + "\n")
    return truth


def _data_preparation(libri_path):
    """What preprocess.py:26-48 returns for this tree (caller-side code, restated for the test)."""
    texts, audio_path = [], []
    for path in sorted(glob.glob(libri_path + "/**/**")):
        for line in open(glob.glob(path + "/*txt")[0]).read().splitlines():
            uid = line.split(" ")[0]
            audio_path.append(path + "/" + uid + ".flac")
            texts.append(line[len(uid) + 1:].replace("'", ""))
    return texts, audio_path


@pytest.mark.parametrize("device_decode", [False, True])
def test_mini_librispeech_offline_pipeline(pkg, ref, tmp_path, device_decode):
    import joblib
    tfr = importlib.import_module(PKG + ".tfrecord")
    bk = importlib.import_module(PKG + ".bucketing")
    corpus = str(tmp_path / "LibriSpeech" / "train-clean-100")
    truth = _make_tree(pkg, corpus)
    texts, audio_path = _data_preparation(corpus)
    assert len(audio_path) == 15 and all(truth[p][1] == t for p, t in zip(audio_path, texts))
    args = make_args(feat_dir=str(tmp_path / "features"), unit="char")
    # preprocess.py:141-155: tokens, then features
    vocab = {c: i + 3 for i, c in enumerate(" ABCDEFGHIJKLMNOPQRSTUVWXYZ")}
    tokens = np.empty(len(texts), dtype=object)
    for i, t in enumerate(texts):
        tokens[i] = [vocab[c] for c in t] + [2]                                          # with_eos
    os.makedirs(args.feat_dir)
    np.save(args.feat_dir + "/train-100-chars.npy", tokens, allow_pickle=True)
    featlen = pkg.process_libri_feats(audio_path, "train-100", 1, args, device_decode=device_decode)
    feats = joblib.load(args.feat_dir + "/train-100-feats.pkl")
    assert np.load(args.feat_dir + "/train-100-featlen.npy").tolist() == featlen == [len(f) for f in feats]
    for f, p in zip(feats, audio_path):
        assert_close(f, ref.features_one(truth[p][0]), what="pipeline features")
    # create_tfrecord.py:43-97 (the caller script create_tfrecord.py:100-140 is a consumer and runs unchanged)
    os.makedirs(str(tmp_path / "tfrecord"))
    perm = np.random.default_rng(0).permutation(len(feats))
    tfr_paths = tfr.create_tfrecords(feats[perm], tokens[perm], str(tmp_path / "tfrecord" / "train-100"), 1)
    records = [r for p in tfr_paths for r in tfr.read_tfrecord(p)]
    assert len(records) == 15
    by_len = {(f.shape[0], tuple(t.tolist())) for f, t in records}
    assert by_len == {(len(f), tuple(tok)) for f, tok in zip(feats, tokens)}                # features stay with their tokens
    # tfrecord_data_loader.py:54-106: what iterator.get_next() hands train.py
    fe = pkg.Frontend(pkg.FrontendConfig())
    cubes = [f for f, _ in records]
    flat = np.concatenate([c.reshape(-1) for c in cubes])
    offs = np.concatenate(([0], np.cumsum([c.size for c in cubes])))[:-1]
    seen = 0
    for (x, xl), (tok, tl) in bk.bucketed_batches(fe, flat, offs, [len(c) for c in cubes], [t for _, t in records], 13):
        assert x.ndim == 4 and x.shape[2:] == (13, 3) and x.shape[1] + 1 in bk.BUCKETS_TRAIN
        assert tok.shape == (x.shape[0], bk.MAX_TOKENLEN_TRAIN) and tok.dtype == np.int32
        for k in range(x.shape[0]):
            assert not x[k, xl[k]:].any() and not tok[k, tl[k]:].any() and tok[k, tl[k] - 1] == 2
        seen += x.shape[0]
    assert seen == 15
    fe.close()
