"""Regenerates tests/golden/golden_v1.npz:  python tests/golden/make_golden.py

The reference ships no fixtures for this path and cannot run here (speechpy / soundfile /
SoX absent), so these vectors are minted by the CPU oracle (oracle/speechpy_ref.py,
oracle/sox_ref.py) on seeded synthetic PCM.  They pin (a) the oracle against accidental
change and (b) the CUDA path on the GPU box, where /root/reference and the generator's
numpy version are not needed: the PCM itself is stored.  PARITY UNPINNED w.r.t. genuine
speechpy -- see the oracle header."""
import hashlib
import importlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import speechpy_ref as R, sox_ref as S       # noqa: E402
synth = importlib.import_module("automatic-speech-recognition_b200.synth")

out = {}
rng = np.random.default_rng(20261017)
lens = [400 + 160 * 60 + 7, 16000, 21923]
pcm = [synth.utterance(n, rng) for n in lens]
for i, p in enumerate(pcm):
    out["pcm_%d" % i] = p
    out["sha_%d" % i] = np.frombuffer(hashlib.sha256(p.tobytes()).digest(), dtype=np.uint8)
    out["mfcc13_%d" % i] = R.features_one(p)
    out["mfcc13_nocmvn_%d" % i] = R.features_one(p, cmvn_flag=False)
    out["mfcc13_timereg_%d" % i] = R.features_one(p, delta_mode="time_regression")
    out["fbank80_%d" % i] = R.features_one(p, feat_dim=80, feat_type="fbank")
    out["fbank40_log_%d" % i] = R.features_one(p, feat_dim=40, feat_type="fbank", fbank_log=True)
out["speed09_pcm_0"] = S.speed_perturb(pcm[0], 0.9)
out["speed11_pcm_0"] = S.speed_perturb(pcm[0], 1.1)
out["gain123_pcm_0"] = S.volume_perturb(pcm[0], 1.23)
out["speed09_mfcc13_0"] = R.features_one(out["speed09_pcm_0"])
# known answers
out["frame_count_n"] = np.array([559, 560, 720, 16000, 32000, 240000, 522320, 559280, 560000], dtype=np.int64)
out["frame_count_L"] = np.array([0, 1, 2, 97, 197, 1497, 3262, 3493, 3497], dtype=np.int64)
out["edges40"] = R.filterbank_edges(40, 257, 16000).astype(np.int64)
out["edges80"] = R.filterbank_edges(80, 257, 16000).astype(np.int64)
path = os.path.join(ROOT, "tests", "golden", "golden_v1.npz")
np.savez_compressed(path, **out)
print(path, os.path.getsize(path), "bytes")
