"""Quick multi-tile sanity run of the K1 pipeline against the oracle (used under `timeout` while developing)."""
import os, sys, importlib
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
pkg = importlib.import_module("automatic-speech-recognition_b200")
from oracle import speechpy_ref as ref
n = int(sys.argv[1]) if len(sys.argv) > 1 else 400
pcm = pkg.synth.corpus(n, 1.0, 9.0, seed=5)
fe = pkg.Frontend(pkg.FrontendConfig())
got = fe.extract(pcm)
err = max(float(np.abs(g - ref.features_one(p)).max()) for g, p in list(zip(got, pcm))[:: max(1, n // 100)])
print("utts", n, "frames", sum(len(g) for g in got), "max abs err vs oracle (sampled)", err, flush=True)
fe.close()
