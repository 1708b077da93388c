"""Drop-in for the feature loop of the reference's preprocess.py.

``process_audios(audio_path, args)`` keeps the reference signature and return value
(/root/reference/preprocess.py:50-91): ``feats`` is a 1-D object ndarray whose element
i is a C-contiguous float32 ``(L_i, feat_dim, 3)`` cube (``(L_i, feat_dim)`` when
``args.cmvn`` is false) and ``featlen`` is a Python list of ints, in input order.
``process_libri_feats`` reproduces the chunked on-disk format of
preprocess.py:112-130 that create_tfrecord.py:32-40 and decode.py:82-83 read back."""
import logging
import os

import joblib
import numpy as np

from . import audio_io
from .frontend import Frontend, FrontendConfig

# When a set holds more utterances than this the reference writes several pickles
# (preprocess.py:17).
_SAMPLE_THRESHOLD = 30000
# host-side batching: PCM samples handed to the GPU per call (~1 audio-hour at 16 kHz)
_BATCH_SAMPLES = 57_600_000
# the device-decode path wants bigger batches: the GPU decoder runs one thread per FLAC frame (14 k frames per
# audio-hour) and the per-batch synchronisations amortise; measured optimum ~4 audio-hours (tools/bench_ingest.py)
_DEVICE_BATCH_SAMPLES = 4 * 57_600_000

_frontends = {}


def _config_key(cfg: FrontendConfig, device):
    return (device, cfg.sample_rate, cfg.frame_length, cfg.frame_step, cfg.feat_dim, cfg.feat_type,
            cfg.cmvn, cfg.num_filters, cfg.delta_mode, cfg.bin_map, cfg.fbank_log, cfg.pcm_dtype,
            cfg.preemph, cfg.dc_elimination, cfg.low_frequency, cfg.high_frequency, tuple(cfg.speeds),
            None if cfg.window is None else cfg.window.tobytes())


def get_frontend(cfg: FrontendConfig, device=0) -> Frontend:
    """One cached handle per (configuration, device) for the calling process."""
    key = _config_key(cfg, device)
    fe = _frontends.get(key)
    if fe is None:
        fe = _frontends[key] = Frontend(cfg, device)
    return fe


def to_object_array(cubes):
    """The reference's ``np.array(feats)`` (preprocess.py:91) relied on numpy < 1.24
    making an object array out of ragged cubes; build it explicitly so equal-length
    batches do not collapse into one dense 4-D array either."""
    out = np.empty(len(cubes), dtype=object)
    for i, c in enumerate(cubes):
        out[i] = c
    return out


def process_pcm(pcm_list, args, fs=16000, device=0, speeds=None, gains=None, **switches):
    """In-memory variant of process_audios: list of int16 (or float) arrays in."""
    dts = {np.asarray(p).dtype for p in pcm_list}
    bad = [d for d in dts if d != np.int16 and not np.issubdtype(d, np.floating)]
    if bad:
        raise TypeError("PCM must be int16 (sample counts) or floating point in [-1, 1): got %s" % sorted(str(d) for d in bad))
    pcm_dtype = "int16" if dts <= {np.dtype(np.int16)} else "float32"
    if pcm_dtype == "float32" and np.dtype(np.int16) in dts:
        # a mixed list: int16 utterances join the float path at the scale soundfile gives them (x / 32768)
        pcm_list = [np.asarray(p, dtype=np.float32) / np.float32(32768.0) if np.asarray(p).dtype == np.int16 else p for p in pcm_list]
    cfg = FrontendConfig.from_args(args, sample_rate=fs, pcm_dtype=pcm_dtype, **switches)
    fe = get_frontend(cfg, device)
    for p in pcm_list:
        if len(p) < cfg.frame_len:
            # speechpy's stack_frames hits np.tile with a negative count here
            raise ValueError("negative dimensions are not allowed")
    cubes, featlen = [], []
    start, acc = 0, 0
    for i, p in enumerate(pcm_list):
        acc += len(p)
        if acc >= _BATCH_SAMPLES or i == len(pcm_list) - 1:
            sl = slice(start, i + 1)
            got = fe.extract(pcm_list[sl], None if speeds is None else speeds[sl],
                             None if gains is None else gains[sl])
            cubes.extend(got)
            featlen.extend(len(g) for g in got)
            start, acc = i + 1, 0
    return to_object_array(cubes), featlen


def _plan_file_batches(lengths, limit=_BATCH_SAMPLES):
    """Consecutive [start, stop) ranges of at most ~limit samples each (at least one file)."""
    ranges, start, acc = [], 0, 0
    for i, n in enumerate(lengths):
        acc += int(n)
        if acc >= limit or i == len(lengths) - 1:
            ranges.append((start, i + 1))
            start, acc = i + 1, 0
    return ranges


def _uniform(value, n):
    return None if value is None else [value] * n


def _process_flac_on_device(audio_path, args, device, n_threads, switches, speed=None, gain=None):
    """FLAC files -> features with the decode on the GPU: raw file bytes are read into page-locked staging
    buffers per ~4-audio-hour batch (host threads, no decoding), uploaded, decoded by ``fe_decode_flac`` into
    HBM and framed there by ``fe_run``; only the cubes come back -- through page-locked bounce buffers into one
    result array per batch, filled (and first-touched) by helper threads while the GPU works on the next batch.
    Nothing is probed up front: a batch's sample counts come out of the bytes the reader thread just loaded."""
    import time
    from concurrent.futures import ThreadPoolExecutor
    sizes = [os.path.getsize(p) for p in audio_path]
    ranges = _plan_file_batches(sizes, int(_DEVICE_BATCH_SAMPLES * 1.1))      # ~0.55 x 2 bytes of FLAC per sample
    fe = None
    # two page-locked staging buffers for the file bytes (batch b uploads while b + 1 is being read)
    fe0 = get_frontend(FrontendConfig.from_args(args, sample_rate=audio_io.DEFAULT_FS, pcm_dtype="int16", **switches), device)
    need = max(sum(sizes[lo:hi]) + 16 * (hi - lo) for lo, hi in ranges) + 64
    stage = getattr(fe0, "_flac_stage", None)
    if stage is None or stage[0].size < need:
        if stage is not None:
            fe0.free_pinned(stage)
        stage = fe0._flac_stage = [fe0.pinned(need + need // 4), fe0.pinned(need + need // 4)]
    trace = {"wait_files": 0.0, "decode": 0.0, "features": 0.0, "views": 0.0} if os.environ.get("FE_TRACE_INGEST") else None
    cubes, featlen = [], []
    # cubes leave the GPU through two page-locked bounce buffers (PCIe rate instead of the pageable-copy rate);
    # helper threads move a finished batch into that batch's result array -- which is also its first touch --
    # while the next batch is decoded and framed
    n_copy = 4
    results = []

    def copy_out(dst, src):
        cuts = np.linspace(0, dst.size, n_copy + 1).astype(np.int64)
        return [copier.submit(np.copyto, dst[cuts[k]:cuts[k + 1]], src[cuts[k]:cuts[k + 1]]) for k in range(n_copy)]

    with ThreadPoolExecutor(max_workers=1) as pool, ThreadPoolExecutor(max_workers=n_copy) as copier:
        nxt = pool.submit(audio_io.load_flac_batch, audio_path[ranges[0][0]:ranges[0][1]], n_threads, stage[0])
        copies = []
        for b, (lo, hi) in enumerate(ranges):
            t0 = time.perf_counter()
            buf, files, pcm_off, lens, fs_b, total = nxt.result()
            t1 = time.perf_counter()
            if b + 1 < len(ranges):
                nxt = pool.submit(audio_io.load_flac_batch, audio_path[ranges[b + 1][0]:ranges[b + 1][1]], n_threads,
                                  stage[(b + 1) & 1])
            if fe is None:
                fs = fs_b
                fe = get_frontend(FrontendConfig.from_args(args, sample_rate=fs, pcm_dtype="int16", **switches), device)
            if fs_b != fs:
                raise MixedSampleRates("mixed sample rates in one call: %d vs %d" % (fs, fs_b))
            if np.any(lens < fe.config.frame_len):
                raise ValueError("negative dimensions are not allowed")      # what speechpy's stack_frames raises
            pcm_total = int(pcm_off[-1] + (lens[-1] + 7) // 8 * 8) if len(lens) else 0
            try:
                d_pcm = fe.decode_flac(buf, files, hi - lo, total, pcm_total)
            except RuntimeError as e:
                bad = [audio_path[lo + int(i)] for i in np.flatnonzero(fe.flac_status)]
                raise audio_io.AudioFormatError("%s: %s" % (", ".join(bad[:4]), e))
            t2 = time.perf_counter()
            sp = fe.speed_indices(_uniform(speed, len(lens)))
            n_out = max(int(fe.plan(lens, sp)[0][-1]), 1)
            bounce = getattr(fe, "_out_stage", None)
            if bounce is None or bounce[0].size < n_out:
                for fs_ in copies:                                            # old buffers may still be read by the copy threads
                    for f in fs_:
                        f.result()
                if bounce is not None:
                    fe.free_pinned(bounce)
                bounce = fe._out_stage = [fe.pinned(n_out + n_out // 4, np.float32), fe.pinned(n_out + n_out // 4, np.float32)]
            if b >= 2:
                for f in copies[b - 2]:                                       # the bounce buffer of batch b - 2 is free again
                    f.result()
            src = bounce[b & 1][:n_out]
            out, out_off, nfr = fe.run_packed(d_pcm, pcm_off, lens, speed_idx=sp, gain=_uniform(gain, len(lens)), out=src)
            res_b = np.empty(n_out, dtype=np.float32)
            results.append(res_b)
            copies.append(copy_out(res_b, src))
            t3 = time.perf_counter()
            cubes.extend(fe.split(res_b, out_off, nfr))                       # views; filled by the copy threads
            featlen.extend(int(L) for L in nfr)
            if trace is not None:
                t4 = time.perf_counter()
                for k, v in zip(("wait_files", "decode", "features", "views"), (t1 - t0, t2 - t1, t3 - t2, t4 - t3)):
                    trace[k] += v
        for fs_ in copies:
            for f in fs_:
                f.result()
    if trace is not None:
        import sys
        sys.stderr.write("ingest trace (s): %s over %d batches\n" % ({k: round(v, 4) for k, v in trace.items()}, len(ranges)))
    return to_object_array(cubes), featlen


def process_audios(audio_path, args, device=0, n_threads=0, device_decode=None, speed=None, gain=None, **switches):
    """Same signature and return value as the reference (preprocess.py:50-91).

    ``device_decode``: True moves the FLAC decode itself to the GPU (FLAC lists only), False keeps it on the
    host thread pool.  The default (None) picks the GPU decoder for an all-FLAC list and says so in the log; if
    the FIRST batch turns out to hold a stream the device decoder does not take (stereo, > 16 bit, variable block
    size), the call continues on the host decoder -- logged, and only for that reason: corrupt files raise.
    ``speed`` / ``gain`` perturb every file of the call on the fly (K0 in front of the framing): the
    features of ``SpeedAugmentation(files, ..., speed)`` without the intermediate audio files.

    FLAC / WAV lists take the batch path: headers are probed once, then each ~1-audio-hour
    batch is decoded by the native thread pool straight into a packed int16 buffer
    (``aio_decode_files``) while the GPU works on the previous batch, and handed to ``fe_run``
    as is -- no per-utterance array exists between the file and the device."""
    audio_path = list(audio_path)
    if not audio_path:
        return to_object_array([]), []
    try:
        return _process_audios_one_rate(audio_path, args, device, n_threads, device_decode, speed, gain, switches)
    except MixedSampleRates:
        pass
    # The reference takes fs from every file (preprocess.py:69-76).  The batch paths want one rate per call (frame
    # geometry and filterbank are per-rate tables): group the files by rate, run each group, put the cubes back in order.
    infos = audio_io.probe_batch(audio_path, n_threads)
    rates = sorted({inf["sample_rate"] for inf in infos})
    logging.info("process_audios: %d sample rates in one list %s: one pass per rate", len(rates), rates)
    cubes, featlen = [None] * len(audio_path), [0] * len(audio_path)
    for fs in rates:
        idx = [i for i, inf in enumerate(infos) if inf["sample_rate"] == fs]
        f, n = _process_audios_one_rate([audio_path[i] for i in idx], args, device, n_threads, device_decode, speed, gain, switches)
        for i, c, m in zip(idx, f, n):
            cubes[i], featlen[i] = c, m
    return to_object_array(cubes), featlen


MixedSampleRates = audio_io.MixedSampleRates


def _process_audios_one_rate(audio_path, args, device, n_threads, device_decode, speed, gain, switches):
    exts = {os.path.splitext(p)[1].lower() for p in audio_path}
    if device_decode:
        if exts != {".flac"}:
            raise ValueError("device_decode=True takes .flac files only")
        return _process_flac_on_device(audio_path, args, device, n_threads, switches, speed, gain)
    if device_decode is None and exts == {".flac"}:
        try:
            logging.info("process_audios: %d FLAC files, decoding on the GPU", len(audio_path))
            return _process_flac_on_device(audio_path, args, device, n_threads, switches, speed, gain)
        except (audio_io.UnsupportedStreamError, ImportError) as e:      # ImportError: no torch for the device PCM buffer
            logging.info("process_audios: host FLAC decoder instead (%s)", e)
    if not exts <= {".flac", ".wav"}:
        pcm_list, fs_seen = [], None
        for p in audio_path:
            audio, fs = audio_io.read_audio(p)
            if fs_seen is None:
                fs_seen = fs
            elif fs != fs_seen:
                raise MixedSampleRates("mixed sample rates in one call: %d vs %d (%s)" % (fs_seen, fs, p))
            pcm_list.append(audio)
        return process_pcm(pcm_list, args, fs=fs_seen, device=device, speeds=_uniform(speed, len(pcm_list)),
                           gains=_uniform(gain, len(pcm_list)), **switches)

    from concurrent.futures import ThreadPoolExecutor
    infos = [audio_io.probe(p) for p in audio_path] if len(audio_path) < 64 else audio_io.probe_batch(audio_path, n_threads)
    fs = infos[0]["sample_rate"]
    cfg = FrontendConfig.from_args(args, sample_rate=fs, pcm_dtype="int16", **switches)
    fe = get_frontend(cfg, device)
    if any(inf["n_samples"] < 0 for inf in infos):
        # a stream whose STREAMINFO carries no total sample count (valid FLAC, what a piping encoder writes) cannot be
        # planned into the packed batch: decode file by file, like the reference's sf.read, then the in-memory path
        pcm_list = [audio_io.read_audio(p)[0] for p in audio_path]
        return process_pcm(pcm_list, args, fs=fs, device=device, speeds=_uniform(speed, len(pcm_list)),
                           gains=_uniform(gain, len(pcm_list)), **switches)
    for p, inf in zip(audio_path, infos):
        if inf["channels"] != 1:
            raise ValueError("%s: mono audio expected" % p)
        if 0 <= inf["n_samples"] < cfg.frame_len:
            raise ValueError("negative dimensions are not allowed")      # what speechpy's stack_frames raises
    ranges = _plan_file_batches([max(inf["n_samples"], 0) for inf in infos])
    cubes, featlen = [], []
    with ThreadPoolExecutor(max_workers=1) as pool:
        nxt = pool.submit(audio_io.read_audio_batch, audio_path[ranges[0][0]:ranges[0][1]], n_threads)
        for b, (lo, hi) in enumerate(ranges):
            packed, off, lens, fs_b = nxt.result()
            if b + 1 < len(ranges):
                nxt = pool.submit(audio_io.read_audio_batch, audio_path[ranges[b + 1][0]:ranges[b + 1][1]], n_threads)
            if fs_b != fs:
                raise MixedSampleRates("mixed sample rates in one call: %d vs %d" % (fs, fs_b))
            out, out_off, nfr = fe.run_packed(packed, off, lens, speed_idx=fe.speed_indices(_uniform(speed, len(lens))),
                                              gain=_uniform(gain, len(lens)))
            got = fe.split(out, out_off, nfr)
            cubes.extend(got)
            featlen.extend(int(L) for L in nfr)
    return to_object_array(cubes), featlen


def process_speed_augmented(audio_path, args, speed_list=(0.9, 1.1), k=4, device=0, **switches):
    """preprocess.py:158-167 without the detour over disk: for each speed the reference writes a resampled
    copy of every training file (SpeedAugmentation) and extracts ``speed_{s}`` features from the copies;
    here the resampler runs in front of the framing inside the same batch call.  Same output files
    (``speed_{s}-feats[-i].pkl``, ``speed_{s}-featlen.npy``)."""
    return {s: process_libri_feats(audio_path, "speed_{}".format(s), k, args, device, speed=s, **switches)
            for s in speed_list}


def process_libri_feats(audio_path, cat, k, args, device=0, **switches):
    """preprocess.py:112-130: chunk sets larger than 30 000 files into k pickles
    ``{cat}-feats-{i}.pkl`` (else one ``{cat}-feats.pkl``) plus ``{cat}-featlen.npy``.

    The cubes reach joblib as views into the result buffers ``fe_run`` filled, and joblib writes every element's bytes
    straight to the file: there is no host-side repack, the pickle goes at the filesystem's write rate (measured
    2.8 of 4.2 GB/s raw on the GPU box, profiles/r02_feats_writer.json) -- but that is still slower than the features
    are produced, so chunk i is pickled by a helper thread while chunk i + 1 is on the GPU."""
    from concurrent.futures import ThreadPoolExecutor
    os.makedirs(args.feat_dir, exist_ok=True)
    if len(audio_path) > _SAMPLE_THRESHOLD:
        featlen = []
        n = len(audio_path) // k + 1
        logging.info("Process %s audios...", cat)
        with ThreadPoolExecutor(max_workers=1) as writer:
            pending = None
            for i in range(k):
                feats, featlen_ = process_audios(audio_path[i * n:(i + 1) * n], args, device, **switches)
                featlen += featlen_
                if pending is not None:
                    pending.result()                                   # at most one chunk waits for the disk
                pending = writer.submit(joblib.dump, feats, args.feat_dir + "/{}-feats-{}.pkl".format(cat, i))
            if pending is not None:
                pending.result()
    else:
        feats, featlen = process_audios(audio_path, args, device, **switches)
        joblib.dump(feats, args.feat_dir + "/{}-feats.pkl".format(cat))
    np.save(args.feat_dir + "/{}-featlen.npy".format(cat), featlen)
    return featlen
