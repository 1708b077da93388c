"""Speed / volume perturbation with the reference's file-level interface
(/root/reference/utils/augmentation.py:6-31, 33-56) on top of the CUDA resampler.

Same arguments, same output naming (``{target_folder}_{speed}/{id}_{speed}.{ext}``,
``{target_folder}/{id}_{volume}.{ext}``), same skip-if-exists rule for speed; the
SoX subprocess per file is replaced by: native thread-pool decode of a batch of files
into one packed int16 buffer (``aio_decode_files``), one batched kernel call
(``fe_perturb``), native thread-pool encode of the results (``aio_write_files``; FLAC in,
FLAC out like SoX).  The resampler is the one defined in DESIGN.md (SoX's own ``rate``
internals are not restated)."""
import os

import numpy as np

from . import audio_io
from .frontend import Frontend, FrontendConfig

_BATCH_FILES = 2048


def _frontend(speed=None, device=0):
    speeds = (float(speed),) if speed is not None and abs(float(speed) - 1.0) > 1e-12 else ()
    return Frontend(FrontendConfig(speeds=speeds), device)


def SpeedAugmentation(filelist, target_folder, speed, device=0):
    """Speed Augmentation (augmentation.py:6-31): returns the list of written paths."""
    audio_path = []
    print("Total audios:", len(filelist))
    target_folder_ = target_folder + "_" + str(speed)
    if not os.path.exists(target_folder_):
        os.makedirs(target_folder_)
    todo = []
    for source_filename in filelist:
        file_id = source_filename.split("/")[-1]
        save_filename = target_folder_ + "/" + file_id.split(".")[0] + "_" + str(speed) + "." + file_id.split(".")[1]
        if os.path.isfile(save_filename):
            print("File exist!")
        else:
            todo.append((source_filename, save_filename))
        audio_path.append(save_filename)
    if todo:
        fe = _frontend(speed, device)
        try:
            for b in range(0, len(todo), _BATCH_FILES):
                chunk = todo[b:b + _BATCH_FILES]
                _perturb_files(fe, [s for s, _ in chunk], [d for _, d in chunk], speeds=[speed] * len(chunk))
        finally:
            fe.close()
    return audio_path


def VolumeAugmentation(filelist, target_folder, vol_range, device=0, rng=None):
    """Volume Augmentation (augmentation.py:33-56).  The reference draws the gain from
    numpy's global, unseeded generator (:48-49); pass ``rng`` for reproducibility."""
    audio_path = []
    print("Total audios:", len(filelist))
    if not os.path.exists(target_folder):
        os.makedirs(target_folder)
    uniform = (rng.uniform if rng is not None else np.random.uniform)
    jobs = []
    for source_filename in filelist:
        volume = np.around(uniform(vol_range[0], vol_range[1]), 2)
        file_id = source_filename.split("/")[-1]
        save_filename = target_folder + "/" + file_id.split(".")[0] + "_" + str(volume) + "." + file_id.split(".")[1]
        jobs.append((source_filename, save_filename, float(volume)))
        audio_path.append(save_filename)
    if jobs:
        fe = _frontend(None, device)
        try:
            for b in range(0, len(jobs), _BATCH_FILES):
                chunk = jobs[b:b + _BATCH_FILES]
                _perturb_files(fe, [s for s, _, _ in chunk], [d for _, d, _ in chunk], gains=[g for _, _, g in chunk])
        finally:
            fe.close()
    return audio_path


def _perturb_files(fe, srcs, dsts, speeds=None, gains=None, n_threads=0):
    """files -> packed int16 -> fe_perturb -> files, without per-utterance arrays when the
    container is FLAC / WAV."""
    native = {os.path.splitext(p)[1].lower() for p in srcs + dsts} <= {".flac", ".wav"}
    same_out = len({os.path.splitext(p)[1].lower() for p in dsts}) == 1
    if native and same_out:
        packed, off, lens, fs = audio_io.read_audio_batch(srcs, n_threads)
        dst, d_off, d_len = fe.perturb_packed(packed, off, lens, speeds=speeds, gains=gains)
        audio_io.write_audio_batch(dsts, dst, d_off[:len(srcs)], d_len, fs, n_threads)
        return
    loaded = [audio_io.read_audio(src) for src in srcs]
    out = fe.perturb([_as_int16(a) for a, _ in loaded], speeds=speeds, gains=gains)
    for dst, y, (_, fs) in zip(dsts, out, loaded):
        audio_io.write_audio(dst, y, fs)


def _as_int16(a):
    a = np.asarray(a)
    if a.dtype == np.int16:
        return a
    return np.clip(np.rint(a.astype(np.float64) * 32768.0), -32768, 32767).astype(np.int16)


class EpochAugmenter:
    """On-the-fly perturbation instead of the reference's three static copies of the training set
    (preprocess.py:158-167 writes speed 0.9 / 1.1 copies once; README.md:31 lists speed perturbation as the
    key improvement): every epoch draws a speed from ``speeds`` and, if ``vol_range`` is given, a gain
    ``np.around(U(lo, hi), 2)`` (utils/augmentation.py:48-49) per utterance, reproducibly from
    ``(seed, epoch)``, and the features come out of the same batch call (K0 in front of the framing).

    ``draw(n, epoch)`` is pure host logic; ``extract(...)`` runs the batch on the GPU."""

    def __init__(self, frontend, speeds=(0.9, 1.0, 1.1), vol_range=None, seed=0):
        self.fe = frontend
        self.speeds = tuple(float(s) for s in speeds)
        self.vol_range = None if vol_range is None else (float(vol_range[0]), float(vol_range[1]))
        self.seed = int(seed)
        if frontend is not None:
            frontend.speed_indices(self.speeds)          # raises if a speed is not configured in the handle

    def draw(self, n, epoch):
        """-> (speeds[n] float64, gains[n] float32 or None), a function of (seed, epoch, n) only."""
        rng = np.random.default_rng([self.seed, int(epoch)])
        sp = np.asarray(self.speeds)[rng.integers(0, len(self.speeds), n)]
        gains = None
        if self.vol_range is not None:
            gains = np.around(rng.uniform(self.vol_range[0], self.vol_range[1], n), 2).astype(np.float32)
        return sp, gains

    def extract(self, packed, offsets, lengths, epoch, out=None, stream=None):
        """One epoch's view of a packed PCM batch (host or device): (out, out_offsets, n_frames, speeds, gains)."""
        sp, gains = self.draw(len(lengths), epoch)
        out, out_off, nfr = self.fe.run_packed(packed, offsets, lengths, speed_idx=self.fe.speed_indices(sp), gain=gains,
                                               out=out, stream=stream)
        return out, out_off, nfr, sp, gains
