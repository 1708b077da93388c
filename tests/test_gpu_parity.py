"""GPU parity: the CUDA path, called through the reference-shaped API / the C-ABI, against the CPU
oracle on the same seeded inputs (sizes the oracle finishes in seconds), against the committed golden
vectors, and -- at BASELINE.json's full sizes -- through size-independent properties.

Tolerance (BASELINE.json north_star): frame counts bit-exact; features max-abs <= 1e-3 OR rel <= 1e-4."""
import importlib
import os

import numpy as np
import pytest

from conftest import PKG, assert_close, make_args

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def corpus1(pkg):
    """config 1 in small: U(2,15) s, mfcc-13, cmvn."""
    return pkg.synth.corpus(40, 2.0, 15.0, seed=1234)


def _check_batch(pkg, ref, pcm, args, what, **sw):
    feats, featlen = pkg.process_pcm(pcm, args, **sw)
    want, want_len = ref.process_audios(pcm, args, **sw)
    assert featlen == want_len, what                       # bit-exact frame counts
    assert feats.dtype == object and feats.shape == (len(pcm),)
    worst = 0.0
    for a, b in zip(feats, want):
        assert a.dtype == np.float32 and a.flags.c_contiguous
        worst = max(worst, assert_close(a, b, what=what))
    return worst


def test_config1_mfcc39_cmvn(pkg, ref, corpus1, args):
    _check_batch(pkg, ref, corpus1, args, "config1 mfcc-39")


def test_config2_fbank80_librispeech_lengths(pkg, ref):
    rng = np.random.default_rng(2345)
    lens = pkg.synth.durations(10, 2, 35, rng, "librispeech").tolist() + [35 * 16000, 2 * 16000]
    pcm = [pkg.synth.utterance(n, rng) for n in lens]
    _check_batch(pkg, ref, pcm, make_args(feat_type="fbank", feat_dim=80), "config2 fbank-80 (linear, as mfe)")
    _check_batch(pkg, ref, pcm[:4], make_args(feat_type="fbank", feat_dim=80), "config2 fbank-80 log", fbank_log=True)


def test_config3_speed_perturbation(pkg, ref, sox, corpus1):
    pcm = corpus1[:9]
    speeds = [0.9, 1.0, 1.1] * 3
    fe = pkg.Frontend(pkg.FrontendConfig())
    got_pcm = fe.perturb(pcm, speeds=speeds)
    want_pcm = [sox.speed_perturb(p, s) for p, s in zip(pcm, speeds)]
    for a, b in zip(got_pcm, want_pcm):
        assert len(a) == len(b)                                       # output length bit-exact
        d = np.abs(a.astype(np.int32) - b.astype(np.int32))
        assert d.max() <= 1 and (d > 0).mean() < 0.02                 # fp32 vs fp64 rounding at .5 boundaries
    got = fe.extract(pcm, speeds=speeds)                              # resample fused in front of framing
    for a, p in zip(got, want_pcm):
        assert_close(a, ref.features_one(p), what="speed-perturbed mfcc")
    fe.close()


def test_volume_perturbation(pkg, ref, sox, corpus1):
    gains = [0.8, 1.5, 1.23, 1.0, 3.0]
    fe = pkg.Frontend(pkg.FrontendConfig())
    got_pcm = fe.perturb(corpus1[:5], gains=gains)
    for a, p, g in zip(got_pcm, corpus1, gains):
        assert np.array_equal(a, sox.volume_perturb(p, g))            # integer work: bit-exact
    got = fe.extract(corpus1[:5], gains=gains)
    for a, p, g in zip(got, corpus1, gains):
        assert_close(a, ref.features_one(sox.volume_perturb(p, g)), what="gain")
    fe.close()


@pytest.mark.parametrize("sw", [dict(delta_mode="time_regression"), dict(bin_map="nfft_plus_one"),
                                dict(preemph=0.98), dict(window=np.hamming(400))])
def test_switches(pkg, ref, corpus1, args, sw):
    _check_batch(pkg, ref, corpus1[:8], args, "switch %s" % list(sw), **sw)


def test_cmvn_false_and_feat_dims(pkg, ref, corpus1):
    _check_batch(pkg, ref, corpus1[:6], make_args(cmvn=False), "mfcc no cmvn")
    _check_batch(pkg, ref, corpus1[:6], make_args(feat_dim=39), "run.sh default feat_dim=39")   # run.sh:41-50
    _check_batch(pkg, ref, corpus1[:6], make_args(feat_type="fbank", feat_dim=40), "fbank-40")
    _check_batch(pkg, ref, corpus1[:6], make_args(feat_type="fbank", feat_dim=23, cmvn=False), "fbank-23 (odd)")


def test_float_pcm_input(pkg, ref, corpus1, args):
    pcm = [ref.pcm_to_float(p).astype(np.float32) for p in corpus1[:5]]
    feats, featlen = pkg.process_pcm(pcm, args)
    for a, p in zip(feats, corpus1):
        assert_close(a, ref.features_one(p), what="float32 pcm")


def test_golden_vectors_on_gpu(pkg, golden):
    pcm = [golden["pcm_%d" % i] for i in range(3)]
    for key, a, sw in (("mfcc13", make_args(), {}), ("mfcc13_nocmvn", make_args(cmvn=False), {}),
                       ("mfcc13_timereg", make_args(), dict(delta_mode="time_regression")),
                       ("fbank80", make_args(feat_type="fbank", feat_dim=80), {}),
                       ("fbank40_log", make_args(feat_type="fbank", feat_dim=40), dict(fbank_log=True))):
        feats, _ = pkg.process_pcm(pcm, a, **sw)
        for i, f in enumerate(feats):
            assert_close(f, golden["%s_%d" % (key, i)], what="golden %s[%d]" % (key, i))
    fe = pkg.Frontend(pkg.FrontendConfig())
    y = fe.perturb([pcm[0]], speeds=[0.9])[0]
    assert np.abs(y.astype(int) - golden["speed09_pcm_0"].astype(int)).max() <= 1
    assert np.array_equal(fe.perturb([pcm[0]], gains=[1.23])[0], golden["gain123_pcm_0"])
    fe.close()


def test_edge_cases(pkg, ref, args):
    rng = np.random.default_rng(7)
    mk = lambda n: (rng.normal(size=n) * 3000).astype(np.int16)
    pcm = [mk(470), mk(560), mk(559), mk(720), mk(400), mk(16000), np.zeros(4000, np.int16),
           mk(400 + 160 * 32), mk(400 + 160 * 33), mk(400 + 160 * 31 + 159)]
    feats, featlen = pkg.process_pcm(pcm, args)
    assert featlen == [0, 1, 0, 2, 0, 97, 22, 32, 33, 31]
    assert feats[0].shape == (0, 13, 3) and feats[4].shape == (0, 13, 3)
    assert np.all(feats[1][:, :, 0] == 0)                              # one frame: (x - mean) / (0 + eps) = 0
    for a, p in zip(feats, pcm):
        assert_close(a, ref.features_one(p), what="edge")
        assert np.isfinite(a).all()
    with pytest.raises(ValueError):                                    # N < 400: the reference raises too
        pkg.process_pcm([mk(399)], args)
    f0, l0 = pkg.process_pcm([], args)
    assert len(f0) == 0 and l0 == []


def test_maximum_length_matches_reference_comment(pkg, ref, args):
    # tfrecord_data_loader.py:79: longest test-clean utterance (559 280 samples) has 3493 frames
    rng = np.random.default_rng(8)
    p = pkg.synth.utterance(559280, rng)
    feats, featlen = pkg.process_pcm([p], args)
    assert featlen == [3493] and feats[0].shape == (3493, 13, 3)
    assert_close(feats[0], ref.features_one(p), what="35 s utterance")


@pytest.mark.parametrize("variant", ["1", "2"])
def test_k1t_tensor_memory_kernel(pkg, ref, corpus1, args, monkeypatch, variant):
    """FE_K1T=1 / 2 select the lane = frame kernels whose FFT exchange lives in tensor memory (fe_k1t.cuh; 2 = K1U,
    16 FFT + 4 epilogue warps, the default where it applies), FE_K1T=0 the 8-lanes-per-frame kernel K1: same features to
    FP32 round-off, same parity against the oracle, for the specialised plans."""
    fr = importlib.import_module(PKG + ".frontend")
    cases = [(dict(), args), (dict(feat_type="fbank", feat_dim=80), make_args(feat_type="fbank", feat_dim=80))]
    if variant == "2":
        cases += [(dict(feat_dim=39), make_args(feat_dim=39)), (dict(feat_type="fbank", feat_dim=40), make_args(feat_type="fbank", feat_dim=40))]
    for kw, a in cases:
        monkeypatch.setenv("FE_K1T", "0")
        fe0 = fr.Frontend(fr.FrontendConfig(**kw))
        base = fe0.extract(corpus1[:12])
        fe0.close()
        monkeypatch.setenv("FE_K1T", variant)
        fe1 = fr.Frontend(fr.FrontendConfig(**kw))
        got = fe1.extract(corpus1[:12])
        again = fe1.extract(corpus1[:12])
        fe1.close()
        monkeypatch.delenv("FE_K1T", raising=False)
        want, _ = ref.process_audios(corpus1[:12], a)
        for x, y, z, w in zip(got, again, base, want):
            assert np.array_equal(x, y)                                # deterministic
            assert_close(x, w, what="K1T vs oracle")
            assert_close(x, z, what="K1T vs K1")


def test_bitwise_determinism_and_batch_independence(pkg, corpus1, args):
    a, _ = pkg.process_pcm(corpus1, args)
    b, _ = pkg.process_pcm(corpus1, args)
    assert all(np.array_equal(x, y) for x, y in zip(a, b))
    alone, _ = pkg.process_pcm([corpus1[7]], args)
    assert np.array_equal(alone[0], a[7])                              # no cross-utterance state
    rev, _ = pkg.process_pcm(corpus1[::-1], args)
    assert all(np.array_equal(x, y) for x, y in zip(rev[::-1], a))


def test_sharded_equals_unsharded_bitwise(pkg, corpus1, args):
    full, _ = pkg.process_pcm(corpus1, args)
    lens = [len(p) for p in corpus1]
    for world in (2, 8):
        parts = pkg.sharding.lpt_partition(pkg.sharding.frame_counts(lens) + 1, world)
        res = [list(pkg.process_pcm([corpus1[int(i)] for i in idx], args)[0]) for idx in parts]
        merged = pkg.sharding.merge_shards(parts, res, len(corpus1))
        assert all(np.array_equal(x, y) for x, y in zip(merged, full))


def test_full_size_properties_config1(pkg, ref):
    """BASELINE config 1 at full size (1000 utterances, 2.36 audio-hours): size-independent checks."""
    rng = np.random.default_rng(1234)
    lens = pkg.synth.durations(1000, 2, 15, rng)
    pcm = pkg.synth.noise_corpus_fast(lens, seed=4)
    feats, featlen = pkg.process_pcm(pcm, make_args())
    assert featlen == [int((n - 400) // 160) for n in lens]
    tot = 0
    for f in feats[::7]:
        st = f[:, :, 0].astype(np.float64)
        assert np.abs(st.mean(0)).max() < 2e-4 and np.abs(st.std(0) - 1).max() < 2e-4      # per-utterance CMVN
        d1 = (st[:, np.minimum(np.arange(13) + 1, 12)] + 2 * st[:, np.minimum(np.arange(13) + 2, 12)]) / 10
        assert np.abs(f[:, :, 1] - d1).max() < 1e-5                                          # delta identity
        tot += len(f)
    assert tot > 0
    for i in (0, 499, 999):                                                                 # spot parity
        assert_close(feats[i], ref.features_one(pcm[i]), what="config1 full spot")
    # linearity of the power path: 2x amplitude -> mfcc c1.. unchanged, c0 (log energy) + log 4 (before CMVN)
    x = pcm[3]
    a, _ = pkg.process_pcm([x // 2 * 2], make_args(cmvn=False))
    b, _ = pkg.process_pcm([x // 2], make_args(cmvn=False))
    assert np.abs((a[0][:, 0] - b[0][:, 0]) - np.log(4.0)).max() < 1e-4
    assert np.abs(a[0][:, 1:] - b[0][:, 1:]).max() < 1e-3


def test_speechpy_shims(pkg, ref, corpus1):
    sp = pkg.speechpy_shim
    x = ref.pcm_to_float(corpus1[0])
    m = sp.mfcc(x, 16000, frame_length=0.025, frame_stride=0.010, num_cepstral=13)
    want = ref.mfcc(x, 16000, 0.025, 0.010, 13)
    assert m.dtype == np.float64 and np.abs(m - want).max() < 1e-4
    f, e = sp.mfe(x, 16000, frame_length=0.025, frame_stride=0.010, num_filters=40)
    wf, we = ref.mfe(x, 16000, 0.025, 0.010, 40)
    assert np.max(np.abs(f - wf) / np.abs(wf)) < 1e-4 and np.max(np.abs(e - we) / we) < 1e-4
    c = sp.cmvn(want, True)
    assert_close(c, ref.cmvn(want, True), what="cmvn shim")
    assert_close(sp.cmvn(want, False), ref.cmvn(want, False), what="cmvn shim (mean only)")
    cube = sp.extract_derivative_feature(ref.cmvn(want, True))
    assert_close(cube, ref.extract_derivative_feature(ref.cmvn(want, True)), what="delta shim")
    m20 = sp.mfcc(x, 16000)                                            # speechpy's own defaults: 20 ms frames, 10 ms stride
    assert np.abs(m20 - ref.mfcc(x, 16000)).max() < 1e-4
    with pytest.raises(RuntimeError, match="frame geometry"):
        sp.mfcc(x, 16000, frame_length=0.032)                          # a geometry no kernel is built for fails loudly


def test_process_audios_files_and_pickles(pkg, ref, tmp_path, corpus1, args, monkeypatch):
    paths = []
    for i, p in enumerate(corpus1[:4]):
        path = str(tmp_path / ("utt%d.%s" % (i, "flac" if i % 2 == 0 else "wav")))   # LibriSpeech ships FLAC
        pkg.audio_io.write_audio(path, p, 16000)
        paths.append(path)
    feats, featlen = pkg.process_audios(paths, args)                   # the reference signature
    for a, p in zip(feats, corpus1):
        assert_close(a, ref.features_one(p), what="process_audios")
    monkeypatch.setattr(importlib.import_module(PKG + ".preprocess"), "_BATCH_SAMPLES", 200_000)
    feats2, featlen2 = pkg.process_audios(paths, args)                 # several decode batches, prefetch thread
    assert featlen2 == featlen and all(np.array_equal(a, b) for a, b in zip(feats, feats2))
    short = str(tmp_path / "short.flac")
    pkg.audio_io.write_audio(short, np.zeros(399, np.int16), 16000)
    with pytest.raises(ValueError, match="negative dimensions"):
        pkg.process_audios(paths[:1] + [short], args)
    import joblib
    args.feat_dir = str(tmp_path / "feats")
    pkg.process_libri_feats(paths, "dev", 1, args)
    back = joblib.load(args.feat_dir + "/dev-feats.pkl")
    assert all(np.array_equal(a, b) for a, b in zip(back, feats))
    assert np.load(args.feat_dir + "/dev-featlen.npy").tolist() == featlen


@pytest.mark.gpu
def test_augmentation_files_flac_in_flac_out(pkg, sox, tmp_path):
    """utils/augmentation.py:6-31, 33-56 end to end on FLAC files: native decode -> fe_perturb -> native encode."""
    aug = importlib.import_module(PKG + ".augmentation")
    pcm = pkg.synth.corpus(3, 0.5, 1.5, seed=21)
    src = []
    for i, x in enumerate(pcm):
        p = str(tmp_path / ("103-1240-%04d.flac" % i))
        pkg.audio_io.write_audio(p, x, 16000)
        src.append(p)
    for speed in (0.9, 1.1):
        out = aug.SpeedAugmentation(src, str(tmp_path / "LibriSpeech_speed_aug"), speed)
        assert out == [str(tmp_path / ("LibriSpeech_speed_aug_%s" % speed) / ("103-1240-%04d_%s.flac" % (i, speed)))
                       for i in range(3)]
        for x, p in zip(pcm, out):
            y, fs = pkg.audio_io.read_audio(p)
            want = sox.speed_perturb(x, speed)
            assert fs == 16000 and y.shape == want.shape
            assert np.abs(y.astype(np.int32) - want.astype(np.int32)).max() <= 1       # float32 taps vs float64 oracle
    vout = aug.VolumeAugmentation(src, str(tmp_path / "vol"), [0.8, 1.5], rng=np.random.default_rng(3))
    for x, p in zip(pcm, vout):
        g = float(os.path.basename(p).rsplit("_", 1)[1][:-5])
        assert np.array_equal(pkg.audio_io.read_audio(p)[0], sox.volume_perturb(x, g))


@pytest.mark.gpu
def test_on_the_fly_speed_features_equal_file_level_augmentation(pkg, ref, sox, tmp_path):
    """preprocess.py:158-167: speed_{s} features.  The fused call (resampler in front of the framing) must give
    what the reference's detour gives: SpeedAugmentation writes copies, process_audios reads them back."""
    import joblib
    aug = importlib.import_module(PKG + ".augmentation")
    pcm = pkg.synth.corpus(4, 1.0, 4.0, seed=55)
    src = []
    for i, x in enumerate(pcm):
        p = str(tmp_path / ("1272-128104-%04d.flac" % i))
        pkg.audio_io.write_audio(p, x, 16000)
        src.append(p)
    args = make_args(feat_dir=str(tmp_path / "features"))
    for dev in (False, True):
        fused, flen = pkg.process_audios(src, args, speed=0.9, device_decode=dev)
        copies = aug.SpeedAugmentation(src, str(tmp_path / ("aug%d" % dev)), 0.9)
        via_files, vlen = pkg.process_audios(copies, args)
        assert flen == vlen and all(np.array_equal(a, b) for a, b in zip(fused, via_files))      # same kernels, same bits
    for a, x in zip(fused, pcm):
        assert_close(a, ref.features_one(sox.speed_perturb(x, 0.9)), what="speed 0.9 features")
    lens = importlib.import_module(PKG + ".preprocess").process_speed_augmented(src, args, speed_list=(0.9, 1.1), k=4)
    assert sorted(lens) == [0.9, 1.1]
    back = joblib.load(args.feat_dir + "/speed_1.1-feats.pkl")
    assert np.load(args.feat_dir + "/speed_1.1-featlen.npy").tolist() == lens[1.1] == [len(b) for b in back]
    louder, _ = pkg.process_audios(src, args, gain=1.3)
    for a, x in zip(louder, pcm):
        assert_close(a, ref.features_one(sox.volume_perturb(x, 1.3)), what="gain 1.3 features")


@pytest.mark.gpu
def test_full_size_properties_config4_100k_sharded_bucketed(pkg, tmp_path):
    """BASELINE config 4 at full size: 100 000 utterances of U(2,15) s (236 audio-hours, 27 GB of int16) resident in
    HBM, sharded 2 / 8 ways, bucketed, and handed to the create_tfrecord.py drop-in -- size-independent checks."""
    import torch
    bk = importlib.import_module(PKG + ".bucketing")
    rng = np.random.default_rng(4567)
    lens = pkg.synth.durations(100_000, 2, 15, rng)
    pad = (lens + 7) // 8 * 8
    off = np.concatenate(([0], np.cumsum(pad)))[:-1].astype(np.int64)
    total = int(pad.sum())
    assert 230 < lens.sum() / 16000 / 3600 < 242
    g = torch.Generator(device="cuda"); g.manual_seed(4567)
    d = torch.empty(total, dtype=torch.int16, device="cuda")
    for s in range(0, total, 1 << 27):
        e = min(total, s + (1 << 27))
        d[s:e] = (torch.randn(e - s, device="cuda", generator=g) * 3000.0).clamp_(-32768, 32767).to(torch.int16)
    torch.cuda.synchronize()
    fe = pkg.Frontend(pkg.FrontendConfig())
    out, out_off, nfr = fe.run_packed(d, off, lens)
    fe.sync()
    assert np.array_equal(nfr, (lens - 400) // 160)                                       # frame counts bit-exact
    assert int(out_off[-1]) * 4 > 13e9
    # per-utterance CMVN moments and the delta identity, on the device, for every 97th utterance
    idx13 = torch.arange(13, device="cuda")
    for i in range(0, 100_000, 97):
        c = out[int(out_off[i]):int(out_off[i]) + int(nfr[i]) * 39].view(int(nfr[i]), 13, 3).double()
        st = c[:, :, 0]
        assert float(st.mean(0).abs().max()) < 2e-4 and float((st.std(0, unbiased=False) - 1).abs().max()) < 2e-4
        d1 = (st[:, torch.clamp(idx13 + 1, max=12)] + 2 * st[:, torch.clamp(idx13 + 2, max=12)]) / 10
        assert float((c[:, :, 1] - d1).abs().max()) < 1e-5
    # sharding: a rank's shard processed on its own equals the same utterances of the unsharded run, bit for bit
    sh = pkg.sharding
    for world, rank in ((2, 1), (8, 3)):
        mine = sh.shard_indices(lens, rank, world)
        assert abs(len(mine) - 100_000 / world) < 0.2 * 100_000 / world and sh.imbalance(lens, world) < 0.01
        spad = pad[mine]
        soff = np.concatenate(([0], np.cumsum(spad)))[:-1].astype(np.int64)
        sd = torch.empty(int(spad.sum()), dtype=torch.int16, device="cuda")
        src_idx = torch.cat([torch.arange(int(off[i]), int(off[i]) + int(pad[i]), device="cuda") for i in mine[:200]])
        sd[:src_idx.numel()] = d[src_idx]                                                  # first 200 utterances of the shard
        n_chk = 200
        so, so_off, snfr = fe.run_packed(sd, soff[:n_chk], lens[mine[:n_chk]])
        fe.sync()
        for k in range(0, n_chk, 7):
            i = int(mine[k])
            a = so[int(so_off[k]):int(so_off[k]) + int(snfr[k]) * 39]
            b = out[int(out_off[i]):int(out_off[i]) + int(nfr[i]) * 39]
            assert torch.equal(a, b)
        del sd, so
    # bucketing: every utterance lands in exactly one slot of a batch of its bucket, padding is zero
    plan = bk.plan_batches(nfr)
    assert sum(len(ix) for _, ix in plan) == 100_000 and max(int(nfr.max()), 0) < 1710
    bb = bk.BucketBatcher(fe)
    views, flat = bb.pad(out, out_off[:-1], nfr, 39, plan)
    for bi in range(0, len(plan), 211):
        b, ix = plan[bi]
        v = views[bi]
        assert v.shape == (len(ix), bk.BUCKETS_TRAIN[b] - 1, 39)
        k = len(ix) // 2
        i = int(ix[k])
        assert torch.equal(v[k, :int(nfr[i])].reshape(-1), out[int(out_off[i]):int(out_off[i]) + int(nfr[i]) * 39])
        assert not bool(v[k, int(nfr[i]):].any())
    del views, flat
    # consumer: create_tfrecords (create_tfrecord.py:43-97) on 300 of the cubes, read back
    tfr = importlib.import_module(PKG + ".tfrecord")
    pick = np.arange(0, 100_000, 334)[:300]
    cubes = [out[int(out_off[i]):int(out_off[i]) + int(nfr[i]) * 39].cpu().numpy().reshape(int(nfr[i]), 13, 3) for i in pick]
    X = pkg.to_object_array(cubes)
    y = pkg.to_object_array([[int(i) % 5000, 1, 2] for i in pick])
    paths = tfr.create_tfrecords(X, y, str(tmp_path / "train-100"), num_files=3)
    back = [r for p in paths for r in tfr.read_tfrecord(p)]
    assert len(back) == 300 and all(np.array_equal(f, c) and t.tolist() == list(t0) for (f, t), c, t0 in zip(back, cubes, y))
    fe.close()


def _device_noise(lens, seed):
    import torch
    pad = (lens + 7) // 8 * 8
    off = np.concatenate(([0], np.cumsum(pad)))[:-1].astype(np.int64)
    g = torch.Generator(device="cuda"); g.manual_seed(seed)
    d = (torch.randn(int(pad.sum()), device="cuda", generator=g) * 3000.0).clamp_(-32768, 32767).to(torch.int16)
    torch.cuda.synchronize()
    return d, off


@pytest.mark.gpu
def test_full_size_properties_config2_fbank80(pkg, ref):
    """BASELINE config 2 at full size: 2 000 utterances, clip(N(12.3, 3.8^2), 2, 35) s, 80 mel energies + CMVN."""
    import torch
    rng = np.random.default_rng(2345)
    lens = pkg.synth.durations(2000, 2, 35, rng, "librispeech")
    d, off = _device_noise(lens, 2345)
    idx = torch.arange(80, device="cuda")
    for log in (False, True):
        fe = pkg.Frontend(pkg.FrontendConfig(feat_type="fbank", feat_dim=80, fbank_log=log))
        out, out_off, nfr = fe.run_packed(d, off, lens)
        fe.sync()
        assert np.array_equal(nfr, (lens - 400) // 160)
        for i in range(0, 2000, 41):
            c = out[int(out_off[i]):int(out_off[i]) + int(nfr[i]) * 240].view(int(nfr[i]), 80, 3).double()
            st = c[:, :, 0]
            assert float(st.mean(0).abs().max()) < 5e-4 and float((st.std(0, unbiased=False) - 1).abs().max()) < 5e-4
            d1 = (st[:, torch.clamp(idx + 1, max=79)] + 2 * st[:, torch.clamp(idx + 2, max=79)]) / 10
            assert float((c[:, :, 1] - d1).abs().max()) < 1e-5
        for i in (0, 1999):                                                               # spot parity against the oracle
            p = d[int(off[i]):int(off[i]) + int(lens[i])].cpu().numpy()
            got = out[int(out_off[i]):int(out_off[i]) + int(nfr[i]) * 240].cpu().numpy().reshape(int(nfr[i]), 80, 3)
            assert_close(got, ref.features_one(p, feat_dim=80, feat_type="fbank", fbank_log=log), what="config2 full spot")
        fe.close()


@pytest.mark.gpu
def test_full_size_properties_config3_three_speeds(pkg, ref, sox):
    """BASELINE config 3 at full size: 1 000 utterances x speeds 0.9 / 1.0 / 1.1 in ONE batch call."""
    tables = importlib.import_module(PKG + ".tables")
    rng = np.random.default_rng(3456)
    base = pkg.synth.durations(1000, 2, 15, rng)
    lens = np.repeat(base, 3)
    d1, off1 = _device_noise(base, 3456)
    import torch
    pad = (lens + 7) // 8 * 8
    off = np.concatenate(([0], np.cumsum(pad)))[:-1].astype(np.int64)
    d = torch.zeros(int(pad.sum()), dtype=torch.int16, device="cuda")
    for i in range(1000):                                                                 # three copies of every utterance
        src = d1[int(off1[i]):int(off1[i]) + int(base[i])]
        for k in range(3):
            d[int(off[3 * i + k]):int(off[3 * i + k]) + int(base[i])] = src
    torch.cuda.synchronize()
    speeds = [0.9, 1.0, 1.1] * 1000
    fe = pkg.Frontend(pkg.FrontendConfig())
    out, out_off, nfr = fe.run_packed(d, off, lens, speed_idx=fe.speed_indices(speeds))
    plain, plain_off, plain_nfr = fe.run_packed(d1, off1, base)
    fe.sync()
    want_len = np.array([tables.resampled_length(int(n), s) for n, s in zip(lens, speeds)])
    assert want_len[0] == -(-int(base[0]) * 10 // 9) and want_len[2] == -(-int(base[0]) * 10 // 11)
    assert np.array_equal(nfr, np.maximum((want_len - 400) // 160, 0))                    # lengths and frame counts bit-exact
    for i in range(0, 1000, 13):                                                          # the speed-1.0 copy is untouched
        a = out[int(out_off[3 * i + 1]):int(out_off[3 * i + 1]) + int(nfr[3 * i + 1]) * 39]
        b = plain[int(plain_off[i]):int(plain_off[i]) + int(plain_nfr[i]) * 39]
        assert torch.equal(a, b)
    for j in (0, 2, 2999):                                                                # spot parity: oracle resampler + oracle features
        p = d[int(off[j]):int(off[j]) + int(lens[j])].cpu().numpy()
        got = out[int(out_off[j]):int(out_off[j]) + int(nfr[j]) * 39].cpu().numpy().reshape(int(nfr[j]), 13, 3)
        assert_close(got, ref.features_one(sox.speed_perturb(p, speeds[j])), what="config3 full spot")
    fe.close()


@pytest.mark.gpu
def test_epoch_augmenter_features(pkg, ref, sox):
    aug = importlib.import_module(PKG + ".augmentation")
    pcm = pkg.synth.corpus(9, 1.0, 3.0, seed=91)
    packed, off, lens = pkg.pack_pcm(pcm)
    fe = pkg.Frontend(pkg.FrontendConfig())
    ea = aug.EpochAugmenter(fe, vol_range=(0.8, 1.5), seed=3)
    out, out_off, nfr, sp, gains = ea.extract(packed, off, lens, epoch=2)
    out2, _, _, sp2, gains2 = ea.extract(packed, off, lens, epoch=2)
    assert np.array_equal(out, out2) and np.array_equal(sp, sp2)                             # an epoch is reproducible
    out3, _, nfr3, sp3, _ = ea.extract(packed, off, lens, epoch=3)
    assert not np.array_equal(sp, sp3)
    for i, x in enumerate(pcm):
        y = sox.volume_perturb(sox.speed_perturb(x, float(sp[i])) if sp[i] != 1.0 else x, float(gains[i]))
        got = out[int(out_off[i]):int(out_off[i]) + int(nfr[i]) * 39].reshape(int(nfr[i]), 13, 3)
        want = ref.features_one(y)
        assert got.shape == want.shape
        # gain is applied inside the resampler's rounding (one quantisation) in the kernel, after it in the oracle
        # chain above (two quantisations): +-1 LSB differences are expected, the features agree to tolerance
        assert_close(got, want, abs_tol=3e-3, rel_tol=1e-3, what="epoch augmenter")
    with pytest.raises(ValueError, match="not in FrontendConfig.speeds"):
        aug.EpochAugmenter(fe, speeds=(0.8,))
    fe.close()


@pytest.mark.parametrize("fs,fl,fstep,kw", [(16000, 20, 10, dict()), (16000, 30, 10, dict()), (8000, 25, 10, dict()),
                                           (8000, 25, 10, dict(feat_type="fbank", feat_dim=40)),
                                           (16000, 20, 10, dict(window="hamming"))])
def test_other_frame_geometries_and_sample_rates(pkg, ref, fs, fl, fstep, kw):
    """VERDICT r1 missing 5: the reference takes frame_length / frame_step from its arguments (las/arguments.py:33-40)
    and fs from every file (preprocess.py:69): 20 / 30 ms frames at 16 kHz, 25 / 10 ms at 8 kHz."""
    rng = np.random.default_rng(fs + fl)
    pcm = [pkg.synth.utterance(int(n), rng) for n in (fs * 2 + 37, fs // 2, fs * 3)]
    kw = dict(kw)
    sw = {}
    if kw.pop("window", None):
        sw["window"] = np.hamming(int(round(fs * fl / 1000.0)))
    a = make_args(frame_length=fl, frame_step=fstep, **kw)
    feats, featlen = pkg.process_pcm(pcm, a, fs=fs, **sw)
    flen, hop = int(round(fs * fl / 1000.0)), int(round(fs * fstep / 1000.0))
    assert featlen == [(len(p) - flen) // hop for p in pcm]
    for f, p in zip(feats, pcm):
        want = ref.features_one(p, fs, fl, fstep, a.feat_dim, a.feat_type, a.cmvn, **sw)
        assert_close(f, want, what="geometry %d/%d at %d Hz" % (flen, hop, fs))
    with pytest.raises(RuntimeError, match="frame geometry"):
        pkg.process_pcm(pcm, make_args(frame_length=32, frame_step=10), fs=fs)


def test_mixed_sample_rates_in_one_file_list(pkg, ref, tmp_path, args):
    """preprocess.py:69 reads fs per file; a list mixing 16 kHz and 8 kHz files is processed rate by rate, order kept."""
    rng = np.random.default_rng(77)
    specs = [(16000, 16000 * 2), (8000, 8000 * 3), (16000, 16000 + 123), (8000, 8000 * 2 + 7)]
    paths, pcm = [], []
    for i, (fs, n) in enumerate(specs):
        x = pkg.synth.utterance(n, rng)
        p = str(tmp_path / ("m%d.flac" % i))
        pkg.audio_io.write_audio(p, x, fs)
        paths.append(p); pcm.append((x, fs))
    feats, featlen = pkg.process_audios(paths, args)
    for f, n, (x, fs) in zip(feats, featlen, pcm):
        want = ref.features_one(x, fs)
        assert n == len(want)
        assert_close(f, want, what="mixed rates %d" % fs)
