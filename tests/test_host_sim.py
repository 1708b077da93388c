"""The exact lane-level dataflow of the frames->statics kernel (stage A / exchange /
stage B / real-FFT split / mel / log / DCT), replayed on the CPU by tests/host_sim and
checked against the float64 oracle.  Catches index-math, swizzle and codelet errors
without a GPU; the GPU parity tests (-m gpu) check the real thing."""
import numpy as np
import pytest


@pytest.fixture(scope="module")
def utts(pkg):
    u = pkg.synth.corpus(3, 1.0, 3.0, seed=99)
    u.append(u[0][:400 + 160 * 5 + 17])        # partial 4-frame group
    u.append(u[1][:400 + 160 * 2 + 3])
    return u


def test_mfcc_statics(host_sim, ref, utts):
    for u in utts:
        want = ref.mfcc(ref.pcm_to_float(u), 16000, 0.025, 0.010, 13)
        assert np.abs(host_sim(u) - want).max() < 2e-5


def test_fbank_statics_relative(host_sim, ref, utts):
    for u in utts:
        want, _ = ref.mfe(ref.pcm_to_float(u), 16000, 0.025, 0.010, 80)
        got = host_sim(u, feat_type="fbank", feat_dim=80)
        assert np.max(np.abs(got - want) / np.abs(want)) < 1e-4


@pytest.mark.parametrize("kw", [dict(bin_map="nfft_plus_one"), dict(feat_dim=40), dict(feat_dim=7, num_filters=26),
                                dict(dc_elimination=False), dict(window=np.hamming(400))])
def test_mfcc_variants(host_sim, ref, utts, kw):
    u = utts[0]
    okw = dict(num_cepstral=kw.get("feat_dim", 13), num_filters=kw.get("num_filters", 40),
               bin_map=kw.get("bin_map", "coefficients_plus_one"), window=kw.get("window"),
               dc_elimination=kw.get("dc_elimination", True))
    want = ref.mfcc(ref.pcm_to_float(u), 16000, 0.025, 0.010, **okw)
    assert np.abs(host_sim(u, **kw) - want).max() < 2e-5


def test_float_pcm_and_log_fbank(host_sim, ref, utts):
    u = utts[1]
    f = ref.pcm_to_float(u)
    assert np.abs(host_sim(f.astype(np.float32), pcm_dtype="float32") - ref.mfcc(f, 16000, 0.025, 0.010, 13)).max() < 2e-5
    want = np.log(ref.mfe(f, 16000, 0.025, 0.010, 23)[0])
    assert np.abs(host_sim(u, feat_type="fbank", feat_dim=23, fbank_log=True) - want).max() < 2e-5


def test_digital_silence_hits_eps_exactly(host_sim, ref):
    z = np.zeros(3000, np.int16)
    got = host_sim(z)
    want = ref.mfcc(ref.pcm_to_float(z), 16000, 0.025, 0.010, 13)
    assert np.abs(got - want).max() < 1e-5       # log(2.22e-16) on c0, ~0 elsewhere


def test_k1t_lane_per_frame_dataflow(host_sim, ref, utts):
    """fe_k1t.cuh (lane = frame, exchange in tensor memory) replayed with the exchange as an array: reads of
    unwritten exchange words come back NaN, so a wrong column / row index cannot hide."""
    for u in utts:
        want = ref.mfcc(ref.pcm_to_float(u), 16000, 0.025, 0.010, 13)
        got = host_sim(u, kernel="k1t")
        assert np.isfinite(got).all()
        assert np.abs(got - want).max() < 2e-5
        assert np.abs(got - host_sim(u)).max() < 2e-5                 # and it agrees with the K1 dataflow
    for u in utts[:2]:
        want, _ = ref.mfe(ref.pcm_to_float(u), 16000, 0.025, 0.010, 80)
        got = host_sim(u, kernel="k1t", feat_type="fbank", feat_dim=80)
        assert np.max(np.abs(got - want) / np.abs(want)) < 1e-4
    z = np.zeros(3000, np.int16)
    want = ref.mfcc(ref.pcm_to_float(z), 16000, 0.025, 0.010, 13)
    assert np.abs(host_sim(z, kernel="k1t") - want).max() < 1e-5       # digital silence: exact eps path


# ---- K1U (fe_kernels.cuh: k_frames_to_statics_u): index identities of the chunked raw layout, checked on the CPU ----
K1U_CHUNK_VECS = 191        # kUChunkVecs
K1U_FRAME_VECS = 20         # kTFrameVecs: a frame starts 160 samples = 20 sixteen-byte vectors after the previous one


def k1u_frame_of_lane(lane):
    return 8 * (lane & 3) + 2 * (lane >> 3) + ((lane >> 2) & 1)


def test_k1u_lane_frame_permutation_and_bank_groups():
    frames = [k1u_frame_of_lane(l) for l in range(32)]
    assert sorted(frames) == list(range(32))                        # every frame of the tile has exactly one lane
    for w in range(50):                                             # the 50 vectors of a frame (400 samples)
        for quarter in range(4):                                    # a 16-byte shared load is served per quarter warp
            groups = set()
            for l in range(8 * quarter, 8 * quarter + 8):
                f = frames[l]
                vec = (f >> 3) * K1U_CHUNK_VECS + (f & 7) * K1U_FRAME_VECS + w
                groups.add(vec % 8)                                 # 8 bank groups of 16 bytes
            assert len(groups) == 8, (w, quarter)                   # conflict-free
    # the contiguous layout this replaced (lane = frame, stride 20 vectors) hits two bank groups per quarter warp
    assert len({(20 * l) % 8 for l in range(8)}) == 2


def test_k1u_chunks_cover_their_frames_and_fit():
    assert 7 * K1U_FRAME_VECS + 50 <= K1U_CHUNK_VECS               # a chunk = 8 frames = 190 vectors
    for n_frames in range(1, 33):
        total = 0
        for c in range(4):
            cnt = min(n_frames - 8 * c, 8)
            if cnt > 0:
                samples = (cnt - 1) * 160 + 400                    # what the kernel copies for chunk c
                total += samples
                assert samples * 2 % 16 == 0                        # bulk copies are multiples of 16 bytes
                # last frame of the chunk ends inside the tile's samples
                assert c * 8 * 160 + samples <= (n_frames - 1) * 160 + 400
        # the closed form the kernel passes to mbarrier.expect_tx
        assert total == (n_frames - 1) * 160 + 400 + ((n_frames - 1) >> 3) * 240


def test_k1u_power_rows_cover_the_plans(pkg):
    """The K1U power buffers only keep bins 5..127 (kUPBin0, kUPRows): every bin the four specialised plans read."""
    import importlib, re, os
    src = open(os.path.join(os.path.dirname(pkg.library_path()), "csrc", "fe_plans_gen.h")).read()
    for name in ("PlanMfcc40", "PlanFbank80"):
        body = src[src.index("struct " + name):]
        b0 = [int(x) for x in re.search(r"B0\[\d+\] = \{([^}]*)\}", body).group(1).split(",")]
        n = [int(x) for x in re.search(r" N\[\d+\] = \{([^}]*)\}", body).group(1).split(",")]
        assert min(b0) >= 5 and max(b + k - 1 for b, k in zip(b0, n)) <= 127
