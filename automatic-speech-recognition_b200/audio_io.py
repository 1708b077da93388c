"""Audio ingest for the file-based boundary (replaces ``sf.read`` at
/root/reference/preprocess.py:69 and SoX's file output at utils/augmentation.py:28,53).

16-bit PCM WAV is handled with the standard library; ``.npy`` holds raw int16/float
samples; FLAC needs ``soundfile`` (not installed in this image -- the reference does
not list it in requirements.txt either).  Decoding stays on the host."""
import os
import wave

import numpy as np

DEFAULT_FS = 16000


def read_audio(path):
    """-> (samples, fs).  int16 for 16-bit PCM sources (the kernels scale by 1/32768,
    which equals the float64 ``sf.read`` hands the reference), float32 otherwise."""
    ext = os.path.splitext(path)[1].lower()
    if ext == ".wav":
        with wave.open(path, "rb") as w:
            if w.getnchannels() != 1:
                raise ValueError("%s: mono audio expected" % path)
            if w.getsampwidth() != 2:
                raise ValueError("%s: 16-bit PCM expected" % path)
            fs = w.getframerate()
            data = np.frombuffer(w.readframes(w.getnframes()), dtype="<i2")
        return data.astype(np.int16, copy=False), fs
    if ext == ".npy":
        data = np.load(path)
        if data.ndim != 1:
            raise ValueError("%s: 1-D samples expected" % path)
        if data.dtype != np.int16:
            data = data.astype(np.float32)
        return data, DEFAULT_FS
    try:
        import soundfile as sf
    except ImportError:
        raise RuntimeError("%s: reading %s needs the 'soundfile' package (libsndfile); "
                           "WAV and NPY are supported natively" % (path, ext or "this format"))
    data, fs = sf.read(path, dtype="int16")
    return data, fs


def write_audio(path, pcm16, fs=DEFAULT_FS):
    """16-bit output, container chosen by extension (SoX does the same)."""
    pcm16 = np.ascontiguousarray(pcm16, dtype=np.int16)
    ext = os.path.splitext(path)[1].lower()
    if ext == ".wav":
        with wave.open(path, "wb") as w:
            w.setnchannels(1)
            w.setsampwidth(2)
            w.setframerate(int(fs))
            w.writeframes(pcm16.astype("<i2").tobytes())
        return
    if ext == ".npy":
        with open(path, "wb") as f:
            np.save(f, pcm16)
        return
    try:
        import soundfile as sf
    except ImportError:
        raise RuntimeError("%s: writing %s needs the 'soundfile' package" % (path, ext))
    sf.write(path, pcm16, fs, subtype="PCM_16")
