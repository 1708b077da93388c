"""Host tables of the product package against the oracle's own construction (bit-exact:
the same float64 formulas decide every filter edge)."""
import numpy as np
import pytest


@pytest.mark.parametrize("nf", [13, 23, 40, 80])
@pytest.mark.parametrize("bin_map", ["coefficients_plus_one", "nfft_plus_one"])
def test_filterbank_equals_oracle(pkg, ref, nf, bin_map):
    a = pkg.tables.mel_filterbank_dense(nf, 16000, bin_map=bin_map)
    b = ref.filterbanks(nf, 257, 16000, 0, 8000, bin_map=bin_map)
    assert np.array_equal(a, b)
    rs, b0, w = pkg.tables.filterbank_csr(a)
    dense = np.zeros_like(a)
    for m in range(nf):
        dense[m, b0[m]:b0[m] + rs[m + 1] - rs[m]] = w[rs[m]:rs[m + 1]]
    assert np.array_equal(dense, a)


def test_resampler_taps_equal_oracle(pkg, sox):
    for s in (0.9, 1.1, 0.95):
        assert np.array_equal(pkg.tables.resampler_taps(s), sox.polyphase_taps(s))
        assert pkg.tables.speed_ratio(s) == sox.speed_ratio(s)
        assert pkg.tables.resampled_length(12345, s) == sox.out_length(12345, s)


def test_twiddles(pkg):
    tw = pkg.tables.twiddles_256()
    j, k = 7, 13
    np.testing.assert_allclose(tw[j, k, 0] + 1j * tw[j, k, 1], np.exp(-2j * np.pi * j * k / 256), atol=1e-15)
    t5 = pkg.tables.twiddles_512()
    np.testing.assert_allclose(t5[100], [np.cos(2 * np.pi * 100 / 512), np.sin(2 * np.pi * 100 / 512)], atol=1e-15)


def test_frontend_config_geometry(pkg):
    c = pkg.FrontendConfig()
    assert (c.frame_len, c.hop, c.n_filters, c.out_width) == (400, 160, 40, 39)
    c = pkg.FrontendConfig(feat_type="fbank", feat_dim=80, cmvn=False)
    assert (c.n_filters, c.out_width) == (80, 80)


def test_k0_fast_path_index_math(pkg):
    """The resampler's fast kernel (k_resample_fast) replaces the per-output 64-bit `j*down / up`, `% up` of the generic
    kernel by tile-level constants: tiles start at multiples of 2880 outputs (phase 0, a whole number of input
    samples), group g is the UP consecutive outputs j = UP g + p, phase (p*down) % up, whose windows start at
    g*down + (p*down)//up + (first - a0) of the staged span (a thread owns kK0Groups = 3 consecutive groups).  Check every (tile, thread, phase) against the
    defining formula y[j] = sum_t taps[(j*down) % up][t] * x[floor(j*down/up) - (T/2 - 1) + t]."""
    K0_OUT, UP, T = 2880, 10, pkg.tables.RESAMPLE_TAPS
    assert T == 128                                                     # kK0Taps in fe_kernels.cuh
    HW = T // 2
    threads = K0_OUT // UP
    for down in (9, 11):
        tile_in = K0_OUT * down // UP
        assert tile_in * UP == K0_OUT * down and down % 2 == 1
        nx = ((UP - 1) * down) // UP + T
        span = (threads - 1) * down + nx
        nv = (span + 14) // 8
        assert threads % 3 == 0 and (threads // 3) % 32 == 0            # kK0Groups = 3 groups per thread, whole warps
        assert nv <= 6 * (threads // 3)                                 # at most six staging vectors per thread
        for tile in (0, 1, 7, 12345):
            te_y = tile * K0_OUT
            first = (te_y // K0_OUT) * tile_in - (HW - 1)
            a0 = first & ~7
            assert first == te_y * down // UP - (HW - 1) and 0 <= first - a0 < 8 and a0 % 8 == 0
            hi = 0
            for g in range(threads):
                for p in range(UP):
                    j = te_y + g * UP + p
                    assert (j * down) % UP == (p * down) % UP           # the phase is a function of p only
                    start = g * down + (p * down) // UP + (first - a0)
                    assert j * down // UP - (HW - 1) - a0 == start
                    hi = max(hi, start + T - 1)
            assert hi < nv * 8                                          # the staged vectors cover every window
