"""K1U (k_frames_to_statics_u) is written around WHEN ptxas keeps warp-uniform values in the uniform datapath
(profiles/r02b_uniform_probe.md).  This compiles the probe for sm_100a (no GPU needed) and checks that the toolchain still
behaves that way: if a variant changes sides, the kernel's performance assumptions need another look."""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
NVCC = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
CUOBJDUMP = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"


@pytest.mark.skipif(not (os.path.exists(NVCC) and os.path.exists(CUOBJDUMP)), reason="needs nvcc and cuobjdump")
def test_uniform_datapath_rules(tmp_path):
    cubin = str(tmp_path / "ubench_uniform.cubin")
    subprocess.run([NVCC, "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-cubin", "-o", cubin,
                    os.path.join(ROOT, "tools", "ubench_uniform.cu")], check=True, capture_output=True)
    uniform = {}
    for v in range(10):
        sass = subprocess.run([CUOBJDUMP, "-sass", "-fun", "_Z1kILi%dEEvPfi" % v, cubin], check=True, capture_output=True,
                              text=True).stdout
        ffma2 = [l for l in sass.splitlines() if "FFMA2" in l]
        assert len(ffma2) == 16, (v, len(ffma2))
        uniform[v] = sum("UR" in l for l in ffma2)
    # uniform-register twiddles: plain warp-derived `if` around straight-line code, vote-guarded loops and spins,
    # tcgen05.st outside non-uniform branches
    assert [v for v in range(10) if uniform[v] == 16] == [0, 5, 7, 9], uniform
    # vector registers: shfl / %warpid, warp-derived loop bounds, plain guards around loops, per-thread spin exits,
    # tcgen05.st under a warp-derived branch
    assert [v for v in range(10) if uniform[v] == 0] == [1, 2, 3, 4, 6, 8], uniform
