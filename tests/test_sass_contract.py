"""Code-generation contract of the default frames -> statics kernel (k_frames_to_statics_u), read from the in-tree library
with cuobjdump (no GPU needed): what profiles/r02b_k1u.md and r02b_sass_excerpts.md claim must be in the binary that ships."""
import os
import re
import shutil
import subprocess

import pytest

CUOBJDUMP = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"


@pytest.fixture(scope="module")
def k1u_sass(pkg):
    if not os.path.exists(CUOBJDUMP):
        pytest.skip("needs cuobjdump")
    out = subprocess.run([CUOBJDUMP, "-sass", pkg.library_path()], check=True, capture_output=True, text=True).stdout
    rows, on = [], False
    for line in out.splitlines():
        if "Function :" in line:
            on = "k_frames_to_statics_uILi1E" in line
        elif on and re.match(r"\s+/\*[0-9a-f]{4,5}\*/", line):
            rows.append(re.sub(r"\s*/\* 0x[0-9a-f]+ \*/\s*$", "", line).strip())
    assert rows, "k_frames_to_statics_u<1> not found in the library"
    return rows


def test_k1u_uses_tensor_memory_tma_and_uniform_operands(k1u_sass):
    def count(pat):
        return sum(1 for r in k1u_sass if re.search(pat, r))
    assert count(r"\bLDTM\b") >= 2 and count(r"\bSTTM\b") >= 32                 # tcgen05.ld / tcgen05.st
    assert all("tmem[UR" in r for r in k1u_sass if re.search(r"\b(LDTM|STTM)\b", r))   # uniform-register addresses
    assert count(r"\bUBLKCP\b") >= 4 and count(r"\bUSETMAXREG\b") == 2          # TMA bulk copies, setmaxnreg for both roles
    ffma2 = [r for r in k1u_sass if "FFMA2" in r]
    assert sum("UR" in r for r in ffma2) >= 70, (len(ffma2), sum("UR" in r for r in ffma2))   # twiddle pairs from uniform registers
    assert count(r"\bLDCU") >= 100                                              # twiddles / weights through uniform loads


def test_k1u_fft_loops_do_not_spill(k1u_sass):
    """The FFT warps' code (between the two setmaxnreg instructions) must be free of local-memory traffic; the 64-register
    epilogue warp may spill a little."""
    idx = [i for i, r in enumerate(k1u_sass) if "USETMAXREG" in r]
    fft = k1u_sass[idx[0]:idx[1]]
    assert len(fft) > 1500
    assert not [r for r in fft if re.search(r"\b(LDL|STL)\b", r)][1:], "spills in the FFT warps' loops"   # one reload of a pointer outside the loops is tolerated
