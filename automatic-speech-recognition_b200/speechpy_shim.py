"""speechpy-shaped entry points for the four library calls the reference makes
(/root/reference/preprocess.py:72, 78, 85, 86), same argument names and defaults as
speechpy 2.4, executed by the CUDA front-end.  Results are float64 arrays like speechpy's
(values carry float32 accuracy).  Frame geometry must be 400 / 160 samples (25 ms / 10 ms
at 16 kHz -- what the reference passes); other geometries raise."""
import numpy as np

from .frontend import FrontendConfig
from .preprocess import get_frontend


def _config(sampling_frequency, frame_length, frame_stride, fft_length, low_frequency, high_frequency, **kw):
    return FrontendConfig(sample_rate=int(sampling_frequency), frame_length=frame_length * 1000.0,
                          frame_step=frame_stride * 1000.0, fft_length=fft_length, low_frequency=low_frequency,
                          high_frequency=high_frequency, pcm_dtype="float32", speeds=(), **kw)


def _signal(signal):
    # speechpy: signal.astype(float) -- no rescaling, whatever the dtype
    return np.ascontiguousarray(np.asarray(signal).astype(np.float32))


def mfe(signal, sampling_frequency, frame_length=0.020, frame_stride=0.01, num_filters=40, fft_length=512,
        low_frequency=0, high_frequency=None, device=0):
    """-> (features (L, num_filters), frame_energies (L,))"""
    x = _signal(signal)
    base = dict(sampling_frequency=sampling_frequency, frame_length=frame_length, frame_stride=frame_stride,
                fft_length=fft_length, low_frequency=low_frequency, high_frequency=high_frequency)
    fb = get_frontend(_config(feat_type="fbank", feat_dim=num_filters, cmvn=False, **base), device)
    feats = fb.extract([x])[0]
    en = get_frontend(_config(feat_type="mfcc", feat_dim=1, num_filters=num_filters, cmvn=False, **base), device)
    energies = np.exp(en.extract([x])[0][:, 0].astype(np.float64))      # c0 = log(frame energy)
    return feats.astype(np.float64), energies


def lmfe(signal, sampling_frequency, frame_length=0.020, frame_stride=0.01, num_filters=40, fft_length=512,
         low_frequency=0, high_frequency=None, device=0):
    x = _signal(signal)
    fe = get_frontend(_config(sampling_frequency, frame_length, frame_stride, fft_length, low_frequency,
                              high_frequency, feat_type="fbank", feat_dim=num_filters, cmvn=False,
                              fbank_log=True), device)
    return fe.extract([x])[0].astype(np.float64)


def mfcc(signal, sampling_frequency, frame_length=0.020, frame_stride=0.01, num_cepstral=13, num_filters=40,
         fft_length=512, low_frequency=0, high_frequency=None, dc_elimination=True, device=0):
    x = _signal(signal)
    fe = get_frontend(_config(sampling_frequency, frame_length, frame_stride, fft_length, low_frequency,
                              high_frequency, feat_type="mfcc", feat_dim=num_cepstral, num_filters=num_filters,
                              cmvn=False, dc_elimination=dc_elimination), device)
    out = fe.extract([x])[0]
    if len(out) == 0:
        return np.empty((0, num_cepstral))
    return out.astype(np.float64)


def cmvn(vec, variance_normalization=False, device=0):
    fe = get_frontend(FrontendConfig(speeds=()), device)
    return fe.postprocess([vec], mean=True, var=bool(variance_normalization), deltas=False)[0].astype(np.float64)


def extract_derivative_feature(feature, device=0, delta_mode="speechpy_as_shipped"):
    fe = get_frontend(FrontendConfig(speeds=()), device)
    return fe.postprocess([feature], mean=False, var=False, deltas=True, delta_mode=delta_mode)[0].astype(np.float64)
