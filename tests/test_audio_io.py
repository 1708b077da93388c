"""Native FLAC / WAV ingest and egress (include/asr_audio_io.h): replaces sf.read at
/root/reference/preprocess.py:69 and SoX's file output at utils/augmentation.py:28,53.

The decoder is pinned by files that FFmpeg's FLAC encoder wrote (tests/golden/flac,
minted by tests/golden/make_flac_golden.py; LPC subframes, partitioned Rice, mid/side
stereo); the encoder was decoded bit-exactly by FFmpeg at mint time (manifest) and
must round-trip through the decoder here.  Everything is bit-exact integer work."""
import hashlib
import importlib
import json
import os
import sys

import numpy as np
import pytest

from conftest import ROOT, PKG, make_args

FLAC_DIR = os.path.join(ROOT, "tests", "golden", "flac")


@pytest.fixture(scope="module")
def manifest():
    return json.load(open(os.path.join(FLAC_DIR, "manifest.json")))


@pytest.fixture(scope="module")
def signals(pkg):
    sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
    return importlib.import_module("make_flac_golden").signals()


def test_decoder_matches_ffmpeg_encoded_fixtures(pkg, manifest, signals):
    assert len(manifest["ffmpeg_encoded"]) >= 8
    for fn, meta in manifest["ffmpeg_encoded"].items():
        path = os.path.join(FLAC_DIR, fn)
        want, ch = signals[meta["signal"]]
        assert hashlib.sha256(want.astype("<i2").tobytes()).hexdigest() == meta["pcm_sha256"], "seeded PCM drifted"
        info = pkg.audio_io.probe(path)
        assert info == {"format": "flac", "sample_rate": 16000, "channels": ch, "bits_per_sample": 16,
                        "n_samples": meta["samples"]}
        got, fs = pkg.audio_io.read_audio(path, check_md5=True)        # MD5 of STREAMINFO verified inside
        assert fs == 16000 and got.dtype == np.int16
        assert got.shape == ((meta["samples"],) if ch == 1 else (meta["samples"], ch))
        assert np.array_equal(got.reshape(-1), want), fn


def test_encoder_was_decoded_by_ffmpeg_and_round_trips(pkg, manifest, signals):
    for name, meta in manifest["our_encoder_decoded_by_ffmpeg"].items():
        assert meta["bit_exact"] is True
        pcm, ch = signals[name]
        data = pkg.audio_io.encode_flac(pcm, 16000, channels=ch)
        assert len(data) == meta["bytes"], "encoder output changed since FFmpeg checked it: re-run make_flac_golden.py"
        back, fs = pkg.audio_io.decode_bytes(data)
        assert fs == 16000 and np.array_equal(back.reshape(-1), pcm)
        # STREAMINFO carries the MD5 of the little-endian PCM
        assert data[26:42] == hashlib.md5(pcm.astype("<i2").tobytes()).digest()


@pytest.mark.parametrize("n", [0, 1, 2, 5, 15, 16, 255, 256, 257, 4095, 4096, 4097, 12289])
def test_round_trip_edge_lengths(pkg, n):
    rng = np.random.default_rng(n)
    cases = [rng.integers(-32768, 32768, n), np.full(n, -32768), np.zeros(n),
             np.cumsum(rng.normal(0, 300, n)).clip(-32768, 32767), (rng.integers(-100, 100, n) * 64)]
    for x in cases:
        x = np.asarray(x).astype(np.int16)
        back, _ = pkg.audio_io.decode_bytes(pkg.audio_io.encode_flac(x))
        assert back.shape == (n,) and np.array_equal(back, x)


def test_corruption_is_detected(pkg):
    x = pkg.synth.corpus(1, 1.0, 1.0, seed=5)[0]
    data = bytearray(pkg.audio_io.encode_flac(x))
    with pytest.raises(pkg.audio_io.AudioFormatError):
        pkg.audio_io.decode_bytes(bytes(data[:len(data) // 2]))          # truncated: sample count short
    bad = bytearray(data); bad[len(bad) // 2] ^= 0x10                     # frame CRC-16
    with pytest.raises(pkg.audio_io.AudioFormatError):
        pkg.audio_io.decode_bytes(bytes(bad))
    bad = bytearray(data); bad[30] ^= 0xff                                # MD5 in STREAMINFO
    with pytest.raises(pkg.audio_io.AudioFormatError):
        pkg.audio_io.decode_bytes(bytes(bad), check_md5=True)
    pkg.audio_io.decode_bytes(bytes(bad), check_md5=False)
    with pytest.raises(pkg.audio_io.AudioFormatError):
        pkg.audio_io.decode_bytes(b"OggS" + bytes(100))
    with pytest.raises((pkg.audio_io.AudioFormatError, OSError)):
        pkg.audio_io.read_audio("/nonexistent/file.flac")


def test_batch_ingest_packs_for_fe_run(pkg, tmp_path):
    pcm = pkg.synth.corpus(9, 0.2, 1.2, seed=8)
    paths = []
    for i, x in enumerate(pcm):
        p = str(tmp_path / ("19-198-%04d.%s" % (i, "flac" if i % 3 else "wav")))
        pkg.audio_io.write_audio(p, x, 16000)
        paths.append(p)
    for threads in (1, 4):
        packed, off, lens, fs = pkg.audio_io.read_audio_batch(paths, n_threads=threads)
        assert fs == 16000 and packed.dtype == np.int16
        assert lens.tolist() == [len(x) for x in pcm] and np.all(off % 8 == 0)        # 16-byte aligned starts
        for x, o, n in zip(pcm, off, lens):
            assert np.array_equal(packed[o:o + n], x)
    want, woff, wlens = importlib.import_module(PKG + ".frontend").pack_pcm(pcm)
    assert np.array_equal(woff, off) and np.array_equal(want[:packed.size], packed[:want.size])
    infos = pkg.audio_io.probe_batch(paths, 3)
    assert [i["n_samples"] for i in infos] == [len(x) for x in pcm]
    # egress: one call, thread pool
    outs = [str(tmp_path / ("o%d.flac" % i)) for i in range(len(pcm))]
    pkg.audio_io.write_audio_batch(outs, packed, off, lens, 16000, n_threads=3)
    for x, p in zip(pcm, outs):
        assert np.array_equal(pkg.audio_io.read_audio(p)[0], x)
    bad = paths[:2] + [str(tmp_path / "missing.flac")]
    with pytest.raises(pkg.audio_io.AudioFormatError, match="missing.flac"):
        pkg.audio_io.read_audio_batch(bad)


def test_stereo_and_rate_errors(pkg, tmp_path):
    st = np.arange(2000, dtype=np.int16).reshape(-1, 2)
    p = str(tmp_path / "st.flac")
    pkg.audio_io.write_flac(p, st, 22050, channels=2)
    got, fs = pkg.audio_io.read_audio(p)
    assert fs == 22050 and got.shape == (1000, 2) and np.array_equal(got, st)
    with pytest.raises(ValueError, match="mono"):
        pkg.audio_io.read_audio_batch([p])
    a, b = str(tmp_path / "a.flac"), str(tmp_path / "b.flac")
    pkg.audio_io.write_audio(a, np.zeros(800, np.int16), 16000)
    pkg.audio_io.write_audio(b, np.zeros(800, np.int16), 8000)
    with pytest.raises(ValueError, match="mixed sample rates"):
        pkg.audio_io.read_audio_batch([a, b])


def test_decoder_survives_mutated_streams(pkg):
    """Hand-written parser hygiene: byte flips, truncations and garbage never crash or over-run the decoder --
    every outcome is either the exact PCM (mutation hit padding / a metadata block) or AudioFormatError."""
    rng = np.random.default_rng(2026)
    x = pkg.synth.corpus(1, 0.6, 0.6, seed=12)[0]
    clean = np.frombuffer(pkg.audio_io.encode_flac(x), dtype=np.uint8)
    ff = np.frombuffer(open(os.path.join(FLAC_DIR, "ffmpeg_tone_plus_noise_l8.flac"), "rb").read(), dtype=np.uint8)
    outcomes = {"ok": 0, "rejected": 0}
    for trial in range(600):
        src = clean if trial % 2 == 0 else ff
        buf = src.copy()
        kind = trial % 5
        if kind == 0:                                      # flip 1..4 random bits
            for _ in range(rng.integers(1, 5)):
                buf[rng.integers(0, buf.size)] ^= 1 << rng.integers(0, 8)
        elif kind == 1:                                    # overwrite a run with random bytes
            a = int(rng.integers(0, buf.size - 64)); buf[a:a + 64] = rng.integers(0, 256, 64, dtype=np.uint8)
        elif kind == 2:                                    # truncate
            buf = buf[:int(rng.integers(1, buf.size))]
        elif kind == 3:                                    # all-ones / all-zeros run (long unary codes)
            a = int(rng.integers(42, buf.size - 256)); buf[a:a + 256] = 0xFF if trial % 2 else 0
        else:                                              # header field damage
            buf[int(rng.integers(4, 42))] = rng.integers(0, 256)
        try:
            got, fs = pkg.audio_io.decode_bytes(buf.tobytes(), check_md5=True)
            outcomes["ok"] += 1
        except pkg.audio_io.AudioFormatError:
            outcomes["rejected"] += 1
    assert outcomes["rejected"] > 400 and outcomes["ok"] + outcomes["rejected"] == 600
    with pytest.raises(pkg.audio_io.AudioFormatError):
        pkg.audio_io.decode_bytes(rng.integers(0, 256, 5000, dtype=np.uint8).tobytes())


def test_flac_round_trip_property(pkg):
    """Property test (hypothesis): any int16 signal -- noise, ramps, constant runs, clipped rails, odd lengths and
    sample rates -- survives encode -> decode bit-exactly, with the MD5 of STREAMINFO verified on the way."""
    from hypothesis import given, settings, strategies as st, HealthCheck

    @st.composite
    def signals(draw):
        parts = []
        for _ in range(draw(st.integers(1, 4))):
            n = draw(st.integers(0, 6000))
            kind = draw(st.sampled_from(["noise", "small", "const", "ramp", "rail", "sine"]))
            seed = draw(st.integers(0, 2 ** 31 - 1))
            rng = np.random.default_rng(seed)
            if kind == "noise":
                x = rng.integers(-32768, 32768, n)
            elif kind == "small":
                x = rng.integers(-3, 4, n)
            elif kind == "const":
                x = np.full(n, draw(st.integers(-32768, 32767)))
            elif kind == "ramp":
                x = np.clip(np.arange(n) * draw(st.integers(-40, 40)) + draw(st.integers(-20000, 20000)), -32768, 32767)
            elif kind == "rail":
                x = rng.choice([-32768, 32767], n)
            else:
                x = np.sin(np.arange(n) * draw(st.floats(0.001, 3.0))) * draw(st.integers(1, 32767))
            parts.append(np.asarray(x).astype(np.int16))
        return np.concatenate(parts) if parts else np.zeros(0, np.int16)

    @settings(max_examples=60, deadline=None, suppress_health_check=[HealthCheck.too_slow, HealthCheck.data_too_large])
    @given(signals(), st.sampled_from([8000, 16000, 22050, 44100, 12345]))
    def check(x, fs):
        data = pkg.audio_io.encode_flac(x, fs)
        back, fs2 = pkg.audio_io.decode_bytes(data, check_md5=True)
        assert fs2 == fs and back.dtype == np.int16 and np.array_equal(back, x)
        assert len(data) <= 2 * x.size + 64 + 24 * (x.size // 4096 + 1)          # never worse than verbatim + framing
    check()


def test_advice_r1_stream_without_sample_count_and_variable_blocking_flag(pkg, tmp_path):
    """a FLAC stream whose STREAMINFO has no total-sample count must still decode on the host path, and a
    stream that sets the variable-blocking bit (0xFFF9) is reported as not fixed-block (the device decoder declines it)."""
    x = pkg.synth.corpus(1, 0.5, 0.6, seed=3)[0]
    data = bytearray(pkg.audio_io.encode_flac(x))
    lay = pkg.audio_io.flac_layout(bytes(data))
    assert lay["min_block"] == lay["max_block"] > 0 and lay["n_samples"] == len(x)
    ff = lay["first_frame"]
    assert data[ff] == 0xFF and data[ff + 1] == 0xF8
    flagged = bytearray(data); flagged[ff + 1] = 0xF9
    assert pkg.audio_io.flac_layout(bytes(flagged))["min_block"] == 0
    # clear the 36-bit total-sample field of STREAMINFO (bytes 4+4+13 .. : low nibble of byte 21, bytes 22..25)
    nosize = bytearray(data)
    nosize[21] &= 0xF0
    nosize[22:26] = b"\x00\x00\x00\x00"
    p = str(tmp_path / "nosize.flac")
    open(p, "wb").write(bytes(nosize))
    assert pkg.audio_io.probe(p)["n_samples"] <= 0
    y, fs = pkg.audio_io.read_audio(p, check_md5=False)
    assert fs == 16000 and np.array_equal(y, x)


def test_advice_r1_process_pcm_rejects_unscaled_integer_dtypes(pkg):
    pre = importlib.import_module(PKG + ".preprocess")
    with pytest.raises(TypeError, match="int16"):
        pre.process_pcm([np.zeros(1000, np.int32)], make_args())
