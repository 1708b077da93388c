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
