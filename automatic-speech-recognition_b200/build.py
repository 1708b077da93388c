"""In-tree build of libasr_frontend.so (nvcc cross-compiles sm_100a without a GPU)."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB_NAME = "libasr_frontend.so"
LIB_PATH = os.path.join(HERE, LIB_NAME)
SOURCES = ["fe_api.cu", "audio_codec.cpp", "record_io.cpp"]
HEADERS = ["fe_core.cuh", "fe_kernels.cuh", "fe_k1t.cuh", "tmem_ops_gen.h", "flac_gpu.cuh", "fe_tables.h", "fe_plans_gen.h", os.path.join("..", "..", "include", "asr_frontend.h"),
           os.path.join("..", "..", "include", "asr_audio_io.h"), os.path.join("..", "..", "include", "asr_record_io.h")]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo", "-O3", "-std=c++17",
    "-shared", "-Xcompiler", "-fPIC",
]


def _nvcc():
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (os.path.isabs(cand) and os.path.exists(cand) or not os.path.isabs(cand)):
            return cand
    return "nvcc"


def is_stale():
    if not os.path.exists(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    deps = [os.path.join(CSRC, s) for s in SOURCES + HEADERS]
    return any(os.path.exists(d) and os.path.getmtime(d) > t for d in deps)


def build_library(force=False, verbose=False, out=None, defines=()):
    """Compile the CUDA sources into automatic-speech-recognition_b200/libasr_frontend.so
    (``out`` / ``defines``: experiment variants, loaded with FE_LIB=<path>)."""
    if out is None and not force and not is_stale():
        return LIB_PATH
    cmd = [_nvcc()] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-D" + d for d in defines]
    if os.environ.get("FE_I2F_HALFSEL"):      # experiment: both int16 halves through I2F.S16 half selectors (no SHF)
        cmd += ["-DFE_I2F_HALFSEL"]
    if os.environ.get("FE_K1_PROF"):          # per-phase clock64 counters in K1 (tools/k1_phases.py); costs registers
        cmd += ["-DFE_K1_PROF"]
    cmd += ["-o", out or LIB_PATH] + [os.path.join(CSRC, s) for s in SOURCES]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed:\n%s\n%s" % (res.stdout, res.stderr))
    if verbose:
        sys.stderr.write(res.stderr)
    return out or LIB_PATH


if __name__ == "__main__":
    print(build_library(force="--force" in sys.argv, verbose="-v" in sys.argv))
