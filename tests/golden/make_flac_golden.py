"""Mint FLAC fixtures with an INDEPENDENT codec: FFmpeg's libavcodec / libavformat
(the copies bundled inside the opencv wheel of this image), driven through ctypes.

    python tests/golden/make_flac_golden.py            # writes tests/golden/flac/*.flac + manifest.json

Two directions, both against code we did not write:

* ``ffmpeg_*.flac`` -- seeded PCM encoded by FFmpeg's FLAC encoder (LPC subframes,
  partitioned Rice, mid/side stereo; compression levels 0 / 5 / 8 / 12, LPC orders up to 32): the
  committed files pin OUR decoder (aio_decode_*) -- the manifest holds the SHA-256
  of the PCM that went in, tests regenerate it from the seed.
* our encoder's output (aio_encode_flac) is demuxed and decoded by FFmpeg here and
  must give the PCM back bit-exactly (checked at mint time; result in the manifest).

FFmpeg structs are touched only at offsets that have been stable for years
(AVFrame.data[0] @0, AVFrame.nb_samples @112, AVFormatContext.pb @32 / .streams @48,
AVStream.codecpar @16, AVPacket.stream_index @36); everything else goes through
AVOptions and avcodec_parameters_*.  The GPU box never runs this script."""
import ctypes as C
import glob
import hashlib
import importlib
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
OUT = os.path.join(HERE, "flac")
FS = 16000


def _load_ffmpeg():
    cand = glob.glob(os.path.join(os.path.dirname(np.__file__), "..", "opencv_python_headless.libs"))
    if not cand:
        raise SystemExit("no bundled FFmpeg found (opencv_python_headless.libs)")
    d = cand[0]

    def L(pat):
        return C.CDLL(glob.glob(os.path.join(d, pat))[0], mode=C.RTLD_GLOBAL)
    for dep in ("libcrypto*", "libssl*", "libdrm*", "libvpx*", "libaom*", "libpng16*"):
        try:
            L(dep)
        except Exception:
            pass
    avutil = L("libavutil*")
    L("libswresample*")
    avcodec = L("libavcodec*")
    avformat = L("libavformat*")
    P = C.c_void_p
    sig = {
        avcodec: {
            "avcodec_find_decoder_by_name": (P, [C.c_char_p]), "avcodec_find_encoder_by_name": (P, [C.c_char_p]),
            "avcodec_alloc_context3": (P, [P]), "avcodec_open2": (C.c_int, [P, P, P]),
            "avcodec_parameters_to_context": (C.c_int, [P, P]), "avcodec_parameters_from_context": (C.c_int, [P, P]),
            "avcodec_send_packet": (C.c_int, [P, P]), "avcodec_receive_frame": (C.c_int, [P, P]),
            "avcodec_send_frame": (C.c_int, [P, P]), "avcodec_receive_packet": (C.c_int, [P, P]),
            "av_packet_alloc": (P, []), "av_packet_unref": (None, [P]), "avcodec_free_context": (None, [C.POINTER(P)]),
        },
        avutil: {
            "av_frame_alloc": (P, []), "av_frame_unref": (None, [P]), "av_frame_make_writable": (C.c_int, [P]),
            "av_opt_set": (C.c_int, [P, C.c_char_p, C.c_char_p, C.c_int]),
            "av_opt_set_int": (C.c_int, [P, C.c_char_p, C.c_int64, C.c_int]),
            "av_log_set_level": (None, [C.c_int]),
        },
        avformat: {
            "avformat_open_input": (C.c_int, [C.POINTER(P), C.c_char_p, P, P]),
            "avformat_find_stream_info": (C.c_int, [P, P]), "av_read_frame": (C.c_int, [P, P]),
            "avformat_close_input": (None, [C.POINTER(P)]),
            "avformat_alloc_output_context2": (C.c_int, [C.POINTER(P), P, C.c_char_p, C.c_char_p]),
            "avformat_new_stream": (P, [P, P]), "avio_open": (C.c_int, [P, C.c_char_p, C.c_int]),
            "avformat_write_header": (C.c_int, [P, P]), "av_interleaved_write_frame": (C.c_int, [P, P]),
            "av_write_trailer": (C.c_int, [P]), "avio_closep": (C.c_int, [P]), "avformat_free_context": (None, [P]),
        },
    }
    for lib, table in sig.items():
        for name, (res, args) in table.items():
            fn = getattr(lib, name)
            fn.restype, fn.argtypes = res, args
    avutil.av_log_set_level(16)
    return avutil, avcodec, avformat


def _ptr_at(base, off):
    return C.c_void_p.from_address(base + off).value


def ffmpeg_decode(ff, path):
    """Demux + decode a FLAC file with FFmpeg -> (int16 interleaved samples, list of AVFrame* kept alive)."""
    avutil, avcodec, avformat = ff
    fmt = C.c_void_p()
    assert avformat.avformat_open_input(C.byref(fmt), path.encode(), None, None) == 0, "open_input " + path
    assert avformat.avformat_find_stream_info(fmt, None) >= 0
    st = _ptr_at(_ptr_at(fmt.value, 48), 0)
    par = _ptr_at(st, 16)
    dec = avcodec.avcodec_find_decoder_by_name(b"flac")
    ctx = avcodec.avcodec_alloc_context3(dec)
    assert avcodec.avcodec_parameters_to_context(ctx, par) >= 0
    assert avcodec.avcodec_open2(ctx, dec, None) == 0
    pkt = avcodec.av_packet_alloc()
    chunks, frames = [], []

    def drain():
        while True:
            fr = avutil.av_frame_alloc()
            if avcodec.avcodec_receive_frame(ctx, fr) != 0:
                return
            n = C.c_int.from_address(fr + 112).value
            fmt_id = C.c_int.from_address(fr + 116).value
            assert fmt_id == 1, "expected AV_SAMPLE_FMT_S16 (interleaved), got %d" % fmt_id
            linesize = C.c_int.from_address(fr + 64).value
            nch = max(1, linesize // (2 * max(n, 1)))
            data = _ptr_at(fr, 0)
            chunks.append(np.ctypeslib.as_array(C.cast(data, C.POINTER(C.c_int16)), shape=(n * nch,)).copy())
            frames.append(fr)
    while avformat.av_read_frame(fmt, pkt) >= 0:
        assert avcodec.avcodec_send_packet(ctx, pkt) == 0
        avcodec.av_packet_unref(pkt)
        drain()
    avcodec.avcodec_send_packet(ctx, None)
    drain()
    return (np.concatenate(chunks) if chunks else np.zeros(0, np.int16)), frames, (fmt, par)


def ffmpeg_encode(ff, pcm, channels, level, dst, template_dir, opts=()):
    """Encode interleaved int16 with FFmpeg's FLAC encoder + muxer.  Frames are borrowed from an
    FFmpeg decode of a template stream (so no AVFrame field has to be filled in by hand)."""
    avutil, avcodec, avformat = ff
    pkg = importlib.import_module("automatic-speech-recognition_b200")
    n = pcm.size // channels
    # template: 4096-sample frames made by our encoder and decoded by FFmpeg; one per encoder frame
    probe_sizes = [4096, 2304, 2048, 1152, 1024, 576, 512, 256, 192]
    tpl_path = os.path.join(template_dir, "tpl_%d.flac" % channels)
    n_tpl_frames = n // 192 + 16
    pkg.audio_io.write_flac(tpl_path, np.zeros(4096 * n_tpl_frames * channels, np.int16), FS, channels=channels)
    _, frames, (fmt, par) = ffmpeg_decode(ff, tpl_path)
    enc = avcodec.avcodec_find_encoder_by_name(b"flac")
    ctx = avcodec.avcodec_alloc_context3(enc)
    assert avcodec.avcodec_parameters_to_context(ctx, par) >= 0
    avutil.av_opt_set(ctx, b"sample_fmt", b"s16", 0)     # not an AVOption in every build: parameters_to_context set it
    assert avutil.av_opt_set(ctx, b"time_base", b"1/%d" % FS, 0) == 0
    assert avutil.av_opt_set_int(ctx, b"compression_level", level, 0) == 0
    for k, v in opts:
        assert avutil.av_opt_set(ctx, k.encode(), str(v).encode(), 1) == 0, k   # AV_OPT_SEARCH_CHILDREN
    assert avcodec.avcodec_open2(ctx, enc, None) == 0
    oc = C.c_void_p()
    assert avformat.avformat_alloc_output_context2(C.byref(oc), None, b"flac", dst.encode()) >= 0
    ost = avformat.avformat_new_stream(oc, None)
    assert avcodec.avcodec_parameters_from_context(_ptr_at(ost, 16), ctx) >= 0
    assert avformat.avio_open(oc.value + 32, dst.encode(), 2) >= 0
    assert avformat.avformat_write_header(oc, None) >= 0
    pkt = avcodec.av_packet_alloc()

    def drain():
        while avcodec.avcodec_receive_packet(ctx, pkt) == 0:
            assert avformat.av_interleaved_write_frame(oc, pkt) == 0
    # the encoder's frame size: the largest candidate it does not reject as "> frame_size"
    fsize, pos, fi = None, 0, 0
    while pos < n:
        fr = frames[fi]; fi += 1
        assert avutil.av_frame_make_writable(fr) >= 0
        data = C.cast(_ptr_at(fr, 0), C.POINTER(C.c_int16))
        buf = np.ctypeslib.as_array(data, shape=(4096 * channels,))
        if fsize is None:
            for cand in probe_sizes:
                m = min(cand, n - pos)
                buf[:m * channels] = pcm[pos * channels:(pos + m) * channels]
                C.c_int.from_address(fr + 112).value = m
                if avcodec.avcodec_send_frame(ctx, fr) == 0:
                    fsize = cand
                    break
            assert fsize is not None, "no frame size accepted"
            m = min(fsize, n - pos)
        else:
            m = min(fsize, n - pos)
            buf[:m * channels] = pcm[pos * channels:(pos + m) * channels]
            C.c_int.from_address(fr + 112).value = m
            assert avcodec.avcodec_send_frame(ctx, fr) == 0
        pos += m
        drain()
    avcodec.avcodec_send_frame(ctx, None)
    drain()
    assert avformat.av_write_trailer(oc) == 0
    avformat.avio_closep(oc.value + 32)
    avformat.avformat_free_context(oc)
    return fsize


def signals():
    """name -> (int16 interleaved, channels).  Seeded; tests regenerate them from this function."""
    pkg = importlib.import_module("automatic-speech-recognition_b200")
    rng = np.random.default_rng(20261017)
    out = {}
    out["speech_like_3s"] = (pkg.synth.corpus(1, 3.0, 3.0, seed=11)[0], 1)
    n = 23457                                  # not a multiple of any block size
    out["white_full_scale"] = (rng.integers(-32768, 32768, n).astype(np.int16), 1)
    t = np.arange(40000)
    out["tone_plus_noise"] = ((np.sin(t * 0.031) * 12000 + np.sin(t * 0.27) * 3000 + rng.normal(0, 40, t.size)).astype(np.int16), 1)
    x = np.zeros(20000, np.int16)
    x[5000:9000] = 777                         # digital silence + a constant run + low-level noise: constant / wasted-bits paths
    x[12000:] = (rng.integers(-8, 8, 8000) * 16).astype(np.int16)
    out["silence_constant_wasted"] = (x, 1)
    left = (np.sin(t[:30000] * 0.02) * 9000 + rng.normal(0, 100, 30000))
    right = left * 0.8 + rng.normal(0, 60, 30000)
    st = np.stack([left, right], 1).astype(np.int16).reshape(-1)
    out["stereo_correlated"] = (st, 2)
    return out


def main():
    pkg = importlib.import_module("automatic-speech-recognition_b200")
    ff = _load_ffmpeg()
    os.makedirs(OUT, exist_ok=True)
    tmp = os.path.join(ROOT, "_scratch", "flac_tpl")
    os.makedirs(tmp, exist_ok=True)
    manifest = {"ffmpeg_encoded": {}, "our_encoder_decoded_by_ffmpeg": {}}
    levels = {"speech_like_3s": [0, 5, 8, 12], "white_full_scale": [5], "tone_plus_noise": [8, 12],
              "silence_constant_wasted": [5], "stereo_correlated": [5, 8]}
    for name, (pcm, ch) in signals().items():
        sha = hashlib.sha256(pcm.astype("<i2").tobytes()).hexdigest()
        for lv in levels[name]:
            fn = "ffmpeg_%s_l%d.flac" % (name, lv)
            fsize = ffmpeg_encode(ff, pcm, ch, lv, os.path.join(OUT, fn), tmp)
            got, _, _ = ffmpeg_decode(ff, os.path.join(OUT, fn))       # FFmpeg round trip (sanity of this script)
            assert got.size == pcm.size and (got == pcm).all(), fn
            manifest["ffmpeg_encoded"][fn] = {"signal": name, "channels": ch, "samples": int(pcm.size // ch),
                                              "level": lv, "frame_size": fsize, "pcm_sha256": sha,
                                              "bytes": os.path.getsize(os.path.join(OUT, fn))}
        if name == "tone_plus_noise":          # LPC orders 20..32: beyond the FLAC "subset" limit of 12
            fn = "ffmpeg_%s_o32.flac" % name
            opts = (("min_prediction_order", 20), ("max_prediction_order", 32), ("prediction_order_method", "estimation"))
            fsize = ffmpeg_encode(ff, pcm, ch, 8, os.path.join(OUT, fn), tmp, opts=opts)
            got, _, _ = ffmpeg_decode(ff, os.path.join(OUT, fn))
            assert got.size == pcm.size and (got == pcm).all(), fn
            manifest["ffmpeg_encoded"][fn] = {"signal": name, "channels": ch, "samples": int(pcm.size // ch), "level": 8,
                                              "options": dict(opts), "frame_size": fsize, "pcm_sha256": sha,
                                              "bytes": os.path.getsize(os.path.join(OUT, fn))}
        ours = os.path.join(tmp, "ours_%s.flac" % name)
        pkg.audio_io.write_flac(ours, pcm, FS, channels=ch)
        got, _, _ = ffmpeg_decode(ff, ours)
        ok = bool(got.size == pcm.size and (got == pcm).all())
        assert ok, "FFmpeg does not reproduce the PCM from our encoder's stream: " + name
        manifest["our_encoder_decoded_by_ffmpeg"][name] = {"bit_exact": ok, "bytes": os.path.getsize(ours), "pcm_sha256": sha}
    json.dump(manifest, open(os.path.join(OUT, "manifest.json"), "w"), indent=1, sort_keys=True)
    print(json.dumps(manifest, indent=1, sort_keys=True))


if __name__ == "__main__":
    main()
