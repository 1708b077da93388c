"""FLAC decode on the GPU (fe_decode_flac) against the host decoder (pinned by FFmpeg-encoded fixtures) and the
original PCM: bit-exact integer work.  Covers LPC / fixed / constant / verbatim subframes, short last blocks,
corruption, and the second pass after a stray header-like byte run."""
import importlib
import json
import os
import sys

import numpy as np
import pytest

from conftest import ROOT, PKG, make_args

pytestmark = pytest.mark.gpu
FLAC_DIR = os.path.join(ROOT, "tests", "golden", "flac")


def _decode(pkg, fe, paths, host_out=False):
    buf, files, pcm_off, lens, fs, total = pkg.audio_io.load_flac_batch(paths, n_threads=2)
    pcm_total = int(pcm_off[-1] + (lens[-1] + 7) // 8 * 8)
    dst = np.zeros(pcm_total, np.int16) if host_out else None
    pcm = fe.decode_flac(buf, files, len(paths), total, pcm_total, pcm=dst)
    pcm = pcm if host_out else pcm.cpu().numpy()
    return [pcm[o:o + n] for o, n in zip(pcm_off, lens)], fs


def test_ffmpeg_fixtures_decode_on_gpu(pkg):
    man = json.load(open(os.path.join(FLAC_DIR, "manifest.json")))
    sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
    signals = importlib.import_module("make_flac_golden").signals()
    names = [fn for fn, m in man["ffmpeg_encoded"].items() if m["channels"] == 1]
    assert len(names) >= 7
    fe = pkg.Frontend(pkg.FrontendConfig())
    for host_out in (False, True):
        got, fs = _decode(pkg, fe, [os.path.join(FLAC_DIR, fn) for fn in names], host_out)
        assert fs == 16000
        for fn, x in zip(names, got):
            want = signals[man["ffmpeg_encoded"][fn]["signal"]][0]
            assert x.shape == want.shape and np.array_equal(x, want), fn
            assert np.array_equal(x, pkg.audio_io.read_audio(os.path.join(FLAC_DIR, fn))[0])     # == host decoder
    fe.close()


def test_own_encoder_streams_and_edge_blocks(pkg, tmp_path):
    rng = np.random.default_rng(3)
    cases = []
    for n in (1, 7, 8, 4095, 4096, 4097, 8192 + 5, 50001):
        cases.append(rng.integers(-32768, 32768, n).astype(np.int16))                            # verbatim
        cases.append(np.cumsum(rng.normal(0, 200, n)).clip(-32768, 32767).astype(np.int16))      # fixed predictors
    cases.append(np.full(9000, -1234, np.int16))                                                 # constant
    cases.append(np.zeros(5000, np.int16))
    cases += pkg.synth.corpus(6, 1.0, 4.0, seed=9)
    paths = []
    for i, x in enumerate(cases):
        p = str(tmp_path / ("c%02d.flac" % i))
        pkg.audio_io.write_audio(p, x, 16000)
        paths.append(p)
    fe = pkg.Frontend(pkg.FrontendConfig())
    got, _ = _decode(pkg, fe, paths)
    for i, (x, want) in enumerate(zip(got, cases)):
        assert x.shape == want.shape and np.array_equal(x, want), i
    fe.close()


def test_corruption_is_reported_per_file(pkg, tmp_path):
    pcm = pkg.synth.corpus(3, 1.0, 2.0, seed=4)
    paths = []
    for i, x in enumerate(pcm):
        p = str(tmp_path / ("u%d.flac" % i))
        pkg.audio_io.write_audio(p, x, 16000)
        paths.append(p)
    raw = bytearray(open(paths[1], "rb").read())
    raw[len(raw) // 2] ^= 0x20                                                                    # breaks one frame's CRC-16
    open(paths[1], "wb").write(bytes(raw))
    fe = pkg.Frontend(pkg.FrontendConfig())
    with pytest.raises(RuntimeError, match="1 file"):
        _decode(pkg, fe, paths)
    assert fe.flac_status.tolist() == [0, -1, 0]
    open(paths[1], "wb").write(bytes(raw[:len(raw) // 2]))                                        # truncated
    with pytest.raises(RuntimeError):
        _decode(pkg, fe, paths)
    assert fe.flac_status.tolist() == [0, -1, 0]
    with pytest.raises(pkg.audio_io.AudioFormatError, match="u1.flac"):
        pkg.process_audios(paths, make_args(), device_decode=True)
    fe.close()


def test_stray_frame_header_in_the_payload_triggers_second_pass(pkg, tmp_path):
    """White noise is stored verbatim, so sample values can spell a valid-looking frame header (sync, the stream's
    block-size code, frame number 0, correct CRC-8) in the middle of frame 1: the scan reports 4 candidates for 3
    frames, the stray one decodes garbage over frame 0's samples, validation flags the file and the second pass
    restores it."""
    rng = np.random.default_rng(11)
    x = rng.integers(-32768, 32768, 3 * 4096).astype(np.int16)
    hdr = [0xFF, 0xF8, 0xC5, 0x08, 0x00]                      # 4096-sample block, 16 kHz, mono, 16 bit, frame 0
    c = 0
    for b in hdr:
        c ^= b
        for _ in range(8):
            c = ((c << 1) ^ 0x07) & 0xFF if c & 0x80 else (c << 1) & 0xFF
    by = hdr + [c]
    x[5000:5003] = np.array([(by[0] << 8) | by[1], (by[2] << 8) | by[3], (by[4] << 8) | by[5]], np.uint16).astype(np.int16)
    p = str(tmp_path / "stray.flac")
    pkg.audio_io.write_audio(p, x, 16000)
    raw = open(p, "rb").read()
    assert bytes(by) in raw                                   # the block really is verbatim
    fe = pkg.Frontend(pkg.FrontendConfig())
    before = fe.launch_count()
    got, _ = _decode(pkg, fe, [p])
    assert fe.launch_count() - before == 4                    # scan, decode, validate + the second decode pass
    assert np.array_equal(got[0], x)
    fe.close()


def test_process_audios_device_decode_equals_host_decode(pkg, tmp_path, monkeypatch):
    pcm = pkg.synth.corpus(7, 1.0, 6.0, seed=31)
    paths = []
    for i, x in enumerate(pcm):
        p = str(tmp_path / ("84-121123-%04d.flac" % i))
        pkg.audio_io.write_audio(p, x, 16000)
        paths.append(p)
    args = make_args()
    a, alen = pkg.process_audios(paths, args, device_decode=False)
    monkeypatch.setattr(importlib.import_module(PKG + ".preprocess"), "_DEVICE_BATCH_SAMPLES", 150_000)   # several batches
    b, blen = pkg.process_audios(paths, args, device_decode=True)
    assert alen == blen and all(np.array_equal(u, v) for u, v in zip(a, b))
    c, clen = pkg.process_audios(paths, args)                                                     # default: GPU decoder for FLAC lists
    assert clen == alen and all(np.array_equal(u, v) for u, v in zip(a, c))
    with pytest.raises(ValueError, match="flac files only"):
        pkg.process_audios([str(tmp_path / "x.wav")], args, device_decode=True)
    # a stream the device decoder does not take (stereo is rejected earlier; here: 8-bit block size not a multiple
    # of 8 cannot be produced by our encoder, so use a 2-channel file): default mode must not silently mis-handle it
    st = str(tmp_path / "stereo.flac")
    pkg.audio_io.write_flac(st, np.zeros(4000, np.int16), 16000, channels=2)
    with pytest.raises(ValueError, match="mono"):
        pkg.process_audios([st], args)


def test_mutated_streams_never_hang_or_corrupt_neighbours(pkg, tmp_path):
    """64 files, half of them damaged (bit flips, random runs, zero / one runs, truncation): the GPU decoder must
    return, flag exactly the files the host decoder rejects, and decode the intact ones bit-exactly."""
    rng = np.random.default_rng(77)
    pcm = pkg.synth.corpus(64, 0.5, 1.5, seed=13)
    paths, expect_bad = [], []
    for i, x in enumerate(pcm):
        data = bytearray(pkg.audio_io.encode_flac(x))
        lay = pkg.audio_io.flac_layout(bytes(data))
        if i % 2:
            kind = (i // 2) % 4
            a = int(rng.integers(lay["first_frame"] + 8, len(data) - 300))
            if kind == 0:
                data[a] ^= 1 << int(rng.integers(0, 8))
            elif kind == 1:
                data[a:a + 64] = bytes(rng.integers(0, 256, 64, dtype=np.uint8))
            elif kind == 2:
                data[a:a + 256] = bytes([0xFF if i % 4 == 1 else 0]) * 256
            else:
                data = data[:a]
        p = str(tmp_path / ("m%02d.flac" % i))
        open(p, "wb").write(bytes(data))
        paths.append(p)
        try:
            pkg.audio_io.decode_bytes(bytes(data))
            expect_bad.append(0)
        except pkg.audio_io.AudioFormatError:
            expect_bad.append(-1)
    assert sum(expect_bad) <= -28
    fe = pkg.Frontend(pkg.FrontendConfig())
    buf, files, pcm_off, lens, fs, total = pkg.audio_io.load_flac_batch(paths)
    pcm_total = int(pcm_off[-1] + (lens[-1] + 7) // 8 * 8)
    dst = np.zeros(pcm_total, np.int16)
    with pytest.raises(RuntimeError, match="file"):
        fe.decode_flac(buf, files, len(paths), total, pcm_total, pcm=dst)
    assert fe.flac_status.tolist() == expect_bad
    for i, x in enumerate(pcm):
        if expect_bad[i] == 0:
            assert np.array_equal(dst[pcm_off[i]:pcm_off[i] + lens[i]], x), i
    fe.close()
