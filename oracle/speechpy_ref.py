"""CPU oracle for the acoustic front-end  --  TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline /
``--impl reference`` legs may import this module.  The product package
(``automatic-speech-recognition_b200``) never does; it fails loudly when its
CUDA library is missing.

What this restates
------------------
The reference hot path is ``process_audios`` (/root/reference/preprocess.py:50-91):

    audio, fs = sf.read(p)                                   preprocess.py:69
    feat = speechpy.feature.mfcc(audio, fs, 0.025, 0.010, num_cepstral=D)   :72-76
    feat, _ = speechpy.feature.mfe(audio, fs, ..., num_filters=D)           :78-82
    feat = speechpy.processing.cmvn(feat, True)                             :85
    feat = speechpy.feature.extract_derivative_feature(feat)                :86
    feats.append(feat.astype(np.float32))                                   :88

The arithmetic lives in the third-party package ``speechpy`` (requirements.txt:5,
UNPINNED; last PyPI release 2.4), which is not vendored in /root/reference and
is not installable here (no network).  Every function below restates the
published speechpy-2.4 algorithm in float64 numpy; the reference call site that
reaches it is cited per function.

PARITY UNPINNED: the reference ships no tests, fixtures or golden vectors for
this path and genuine speechpy cannot be run in this container, so this oracle
is pinned only by (a) the frame-count evidence in
/root/reference/tfrecord_data_loader.py:78-79 (3262 / 3493 frames for the
longest dev/test utterances <=> L = floor((N-400)/160)), and (b) the output
contract (L, D, 3) float32 consumed at tfrecord_data_loader.py:34,44 and
las/beam_search.py:163-164.  The three behaviours with the least certainty are
exposed as named switches that default to the as-shipped speechpy behaviour:
``bin_map``, ``delta_mode`` and the 300 Hz low-edge floor.
"""
import math

import numpy as np
from scipy.fftpack import dct as _scipy_dct

FLOAT_EPS = np.finfo(float).eps          # speechpy.functions.zero_handling substitute
CMVN_EPS = 2.0 ** -30                    # speechpy.processing.cmvn


# --------------------------------------------------------------------------
# speechpy.functions
# --------------------------------------------------------------------------
def frequency_to_mel(f):
    """speechpy.functions.frequency_to_mel (used by filterbanks)."""
    return 1127 * np.log(1 + f / 700.)


def mel_to_frequency(mel):
    """speechpy.functions.mel_to_frequency (used by filterbanks)."""
    return 700 * (np.exp(mel / 1127.0) - 1)


def triangle(x, left, middle, right):
    """speechpy.functions.triangle: 0 outside (left, right), rising on
    left < x <= middle, falling on middle <= x < right (second assignment wins
    at x == middle)."""
    out = np.zeros(x.shape)
    rising = np.logical_and(left < x, x <= middle)
    out[rising] = (x[rising] - left) / (middle - left)
    falling = np.logical_and(middle <= x, x < right)
    out[falling] = (right - x[falling]) / (right - middle)
    return out


def zero_handling(x):
    """speechpy.functions.zero_handling: exact zeros -> float64 machine eps."""
    return np.where(x == 0, FLOAT_EPS, x)


# --------------------------------------------------------------------------
# speechpy.processing
# --------------------------------------------------------------------------
def preemphasis(signal, shift=1, cof=0.98):
    """speechpy.processing.preemphasis.  NOT called by preprocess.py:69-82 (the
    reference feeds the raw signal); kept because the north-star lists it as a
    kernel option.  Circular (np.roll) as speechpy ships it."""
    return signal - cof * np.roll(signal, shift)


def num_frames(n_samples, frame_len=400, hop=160):
    """Frame-count rule of stack_frames(zero_padding=False) -- one fewer than the
    textbook 1+floor(.); corroborated by tfrecord_data_loader.py:78-79."""
    return int(math.floor((n_samples - frame_len) / float(hop)))


def stack_frames(sig, sampling_frequency, frame_length=0.020, frame_stride=0.020,
                 window=None, zero_padding=False):
    """speechpy.processing.stack_frames as reached from mfe (zero_padding=False,
    rectangular window)."""
    assert sig.ndim == 1
    n = sig.shape[0]
    flen = int(np.round(sampling_frequency * frame_length))
    hop = float(np.round(sampling_frequency * frame_stride))
    if zero_padding:
        nfr = int(math.ceil((n - flen) / hop))
        total = int(nfr * hop + flen)
        signal = np.concatenate((sig, np.zeros((total - n,))))
    else:
        nfr = int(math.floor((n - flen) / hop))
        if nfr < 0:
            raise ValueError("negative dimensions are not allowed")   # np.tile in speechpy
        total = int((nfr - 1) * hop + flen)
        signal = sig[0:max(total, 0)]
    idx = (np.arange(flen)[None, :] + (np.arange(nfr) * hop)[:, None]).astype(np.int32)
    frames = signal[idx] if nfr > 0 else np.zeros((0, flen))
    win = np.ones((flen,)) if window is None else np.asarray(window, dtype=float)
    return frames * win[None, :]


def power_spectrum(frames, fft_points=512):
    """speechpy.processing.power_spectrum: |rfft(frames, n)|^2 / n."""
    spec = np.fft.rfft(frames, n=fft_points, axis=-1)
    return 1.0 / fft_points * np.square(np.absolute(spec))


def cmvn(vec, variance_normalization=False):
    """speechpy.processing.cmvn, called at preprocess.py:85 with True."""
    rows, cols = vec.shape
    mean = np.mean(vec, axis=0)
    centred = vec - np.tile(mean, (rows, 1))
    if variance_normalization:
        std = np.std(centred, axis=0)
        return centred / (np.tile(std, (rows, 1)) + CMVN_EPS)
    return centred


def derivative_extraction(feat, DeltaWindows, delta_mode="speechpy_as_shipped"):
    """speechpy.processing.derivative_extraction.

    ``speechpy_as_shipped`` (default): speechpy pads the COEFFICIENT axis with
    'edge' values and its subtraction sits on a dangling source line, so the
    result is  d[t,k] = sum_{r=1..W} r * x[t, min(k+r, D-1)] / sum_r 2 r^2.
    ``time_regression``: the textbook HTK delta along time with edge
    replication,  d[t,k] = sum_r r (x[t+r,k] - x[t-r,k]) / sum_r 2 r^2.
    """
    rows, cols = feat.shape
    acc = np.zeros(feat.shape, dtype=feat.dtype)
    scale = 0
    if delta_mode == "speechpy_as_shipped":
        padded = np.pad(feat, ((0, 0), (DeltaWindows, DeltaWindows)), "edge")
        for i in range(DeltaWindows):
            r = i + 1
            acc += r * padded[:, DeltaWindows + r:DeltaWindows + r + cols]
            scale += 2 * r ** 2
    elif delta_mode == "time_regression":
        padded = np.pad(feat, ((DeltaWindows, DeltaWindows), (0, 0)), "edge")
        for i in range(DeltaWindows):
            r = i + 1
            acc += r * (padded[DeltaWindows + r:DeltaWindows + r + rows, :]
                        - padded[DeltaWindows - r:DeltaWindows - r + rows, :])
            scale += 2 * r ** 2
    else:
        raise ValueError(delta_mode)
    return acc / scale


# --------------------------------------------------------------------------
# speechpy.feature
# --------------------------------------------------------------------------
def filterbank_edges(num_filter, coefficients, sampling_freq, low_freq=None,
                     high_freq=None, bin_map="coefficients_plus_one", fft_length=512):
    """FFT-bin edges of the triangular filters (num_filter + 2 ints)."""
    high_freq = high_freq or sampling_freq / 2
    low_freq = low_freq or 300            # speechpy: 0/None silently becomes 300 Hz
    assert high_freq <= sampling_freq / 2
    assert low_freq >= 0
    mels = np.linspace(frequency_to_mel(low_freq), frequency_to_mel(high_freq), num_filter + 2)
    hertz = mel_to_frequency(mels)
    if bin_map == "coefficients_plus_one":        # as shipped: (257 + 1) * f / fs
        scale = coefficients + 1
    elif bin_map == "nfft_plus_one":              # textbook: (512 + 1) * f / fs
        scale = fft_length + 1
    else:
        raise ValueError(bin_map)
    return np.floor(scale * hertz / sampling_freq).astype(int)


def filterbanks(num_filter, coefficients, sampling_freq, low_freq=None, high_freq=None,
                bin_map="coefficients_plus_one", fft_length=512):
    """speechpy.feature.filterbanks -> dense (num_filter, coefficients) table."""
    edges = filterbank_edges(num_filter, coefficients, sampling_freq, low_freq, high_freq,
                             bin_map, fft_length)
    fb = np.zeros([num_filter, coefficients])
    for i in range(num_filter):
        left, middle, right = int(edges[i]), int(edges[i + 1]), int(edges[i + 2])
        z = np.linspace(left, right, num=right - left + 1)
        fb[i, left:right + 1] = triangle(z, left=left, middle=middle, right=right)
    return fb


def mfe(signal, sampling_frequency, frame_length=0.020, frame_stride=0.01, num_filters=40,
        fft_length=512, low_frequency=0, high_frequency=None, window=None, preemph=None,
        bin_map="coefficients_plus_one"):
    """speechpy.feature.mfe (preprocess.py:78-82; also inside mfcc).  Returns
    (LINEAR mel energies (L, num_filters), frame energies (L,)).
    ``window`` / ``preemph`` are off in the reference (kernel options only)."""
    signal = np.asarray(signal).astype(float)
    if preemph:
        signal = preemphasis(signal, 1, preemph)
    frames = stack_frames(signal, sampling_frequency, frame_length, frame_stride,
                          window=window, zero_padding=False)
    high_frequency = high_frequency or sampling_frequency / 2
    pspec = power_spectrum(frames, fft_length)
    coefficients = pspec.shape[1]
    frame_energies = zero_handling(np.sum(pspec, 1))
    fb = filterbanks(num_filters, coefficients, sampling_frequency, low_frequency,
                     high_frequency, bin_map, fft_length)
    features = zero_handling(np.dot(pspec, fb.T))
    return features, frame_energies


def lmfe(signal, sampling_frequency, **kw):
    """speechpy.feature.lmfe: log of mfe.  NOT what preprocess.py calls for
    feat_type='fbank' (that is linear mfe); kept for the ``fbank_log`` option."""
    feature, _ = mfe(signal, sampling_frequency, **kw)
    return np.log(feature)


def mfcc(signal, sampling_frequency, frame_length=0.020, frame_stride=0.01, num_cepstral=13,
         num_filters=40, fft_length=512, low_frequency=0, high_frequency=None,
         dc_elimination=True, window=None, preemph=None, bin_map="coefficients_plus_one"):
    """speechpy.feature.mfcc (preprocess.py:72-76)."""
    feature, energy = mfe(signal, sampling_frequency, frame_length, frame_stride, num_filters,
                          fft_length, low_frequency, high_frequency, window, preemph, bin_map)
    if len(feature) == 0:
        return np.empty((0, num_cepstral))
    feature = np.log(feature)
    feature = _scipy_dct(feature, type=2, axis=-1, norm="ortho")[:, :num_cepstral]
    if dc_elimination:
        feature[:, 0] = np.log(energy)
    return feature


def extract_derivative_feature(feature, delta_mode="speechpy_as_shipped"):
    """speechpy.feature.extract_derivative_feature (preprocess.py:86) -> (L, D, 3)."""
    d1 = derivative_extraction(feature, 2, delta_mode)
    d2 = derivative_extraction(d1, 2, delta_mode)
    return np.concatenate((feature[:, :, None], d1[:, :, None], d2[:, :, None]), axis=2)


# --------------------------------------------------------------------------
# preprocess.py:50-91 restated over in-memory PCM (sf.read replaced)
# --------------------------------------------------------------------------
def pcm_to_float(pcm):
    """What soundfile.read (preprocess.py:69) hands over for 16-bit audio:
    float64 = int16 / 32768.  Float input passes through as float64."""
    pcm = np.asarray(pcm)
    if pcm.dtype == np.int16:
        return pcm.astype(np.float64) / 32768.0
    return pcm.astype(np.float64)


def features_one(pcm, fs=16000, frame_length=25, frame_step=10, feat_dim=13, feat_type="mfcc",
                 cmvn_flag=True, delta_mode="speechpy_as_shipped", fbank_log=False,
                 window=None, preemph=None, bin_map="coefficients_plus_one"):
    """One iteration of the loop at preprocess.py:67-89."""
    audio = pcm_to_float(pcm)
    if feat_type == "mfcc":
        feat = mfcc(audio, fs, frame_length=frame_length / 1000, frame_stride=frame_step / 1000,
                    num_cepstral=feat_dim, window=window, preemph=preemph, bin_map=bin_map)
    elif feat_type == "fbank":
        feat, _ = mfe(audio, fs, frame_length=frame_length / 1000,
                      frame_stride=frame_step / 1000, num_filters=feat_dim,
                      window=window, preemph=preemph, bin_map=bin_map)
        if fbank_log:
            feat = np.log(feat)
    else:
        raise ValueError(feat_type)
    if cmvn_flag:
        with np.errstate(invalid="ignore", divide="ignore"):
            feat = cmvn(feat, True) if len(feat) else feat
        feat = extract_derivative_feature(feat, delta_mode)
    return feat.astype(np.float32)


def process_audios(pcm_list, args, fs=16000, **switches):
    """preprocess.py:50-91 with file reading replaced by in-memory PCM.
    ``args`` needs frame_step, frame_length, feat_dim, feat_type, cmvn
    (preprocess.py:59-63).  Returns (object ndarray of cubes, list of lengths);
    the object array is built explicitly because np.array(ragged list)
    (preprocess.py:91) raises on numpy >= 1.24."""
    feats, featlen = [], []
    for pcm in pcm_list:
        f = features_one(pcm, fs, args.frame_length, args.frame_step, args.feat_dim,
                         args.feat_type, args.cmvn, **switches)
        feats.append(f)
        featlen.append(len(f))
    out = np.empty(len(feats), dtype=object)
    for i, f in enumerate(feats):
        out[i] = f
    return out, featlen
