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
