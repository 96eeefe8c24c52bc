"""The library's C++ host routines for the pixel-independent multiplier vectors against the
oracle's literal restatement (CPU only, no GPU)."""
import numpy as np
import pytest

from helpers import F32, orc, pkg, time_axis


def _close(a, b):
    # both are f32 evaluations of the same formula; glibc cosf vs rounded double cos may differ by 1 ulp
    np.testing.assert_allclose(a, b, rtol=0, atol=2.5e-7)


@pytest.mark.parametrize("n", [64, 128, 2048, 4096])
def test_default_chain_vectors(n):
    h = pkg().host
    t = time_axis(n)
    m_pre, band, m_post = h.chain_multipliers(t)
    o_pre, o_band, o_post = orc.default_chain_multipliers(t)
    _close(m_pre, o_pre)
    _close(band, o_band)
    _close(m_post, o_post)
    assert np.array_equal(band == 0, o_band == 0)          # exact zeros outside the band
    assert np.array_equal(m_post == 0, o_post == 0)
    assert m_post[-1] == 0 and m_pre[-1] == 0              # default gates zero the last sample
    np.testing.assert_array_equal(h.frequency_axis(t), orc.frequency_axis(t))


@pytest.mark.parametrize("wt", ["AdaptedBlackman", "Blackman", "Hanning", "Hamming", "FlatTop"])
def test_fft_windows(wt):
    """test_window_functions_apply (src/math_tools.rs:757-840)."""
    h = pkg().host
    t = np.linspace(0, 1, 128, dtype=F32)
    w = h.fft_window(t, wt, (0.1, 0.1))
    _close(w, orc.fft_window_multiplier(t, wt, (0.1, 0.1)))
    if wt == "Hamming":
        assert abs(w[0] - 0.08) < 1e-5 and abs(w[-1] - 0.08) < 1e-5
    else:
        assert abs(w[0]) < 1e-5 and abs(w[-1]) < 1e-5
    assert np.max(np.abs(w - w[::-1])) < 1e-5
    if wt == "AdaptedBlackman":
        assert w[64] == 1.0


def test_gate_and_band_indices():
    h = pkg().host
    t = np.linspace(0, 1, 256, dtype=F32)
    m, lower, upper, lo, hi = h.time_gate(t, 0.25, 0.55, 0.0)
    ol, ou, olo, ohi = orc.td_gate_indices(t, 0.25, 0.55)
    assert (lower, upper) == (ol, ou) and lo == pytest.approx(olo) and hi == pytest.approx(ohi)
    _close(m, orc.td_gate_multiplier(t, 0.25, 0.55, 0.0))
    assert np.all(m[:lower] == 0) and np.all(m[upper:] == 0) and np.all(m[lower:upper] == 1)
    # out-of-range gate is clamped to the axis like the filter mutates its fields
    m2, l2, u2, lo2, hi2 = h.time_gate(t, -5.0, 9.0, 0.1)
    assert (l2, u2) == orc.td_gate_indices(t, -5.0, 9.0)[:2] and lo2 == 0.0 and hi2 == 1.0
    f = (np.arange(129, dtype=F32) / F32(50.0)).astype(F32)
    b, bl, bu = h.band_pass(f, 0.1, 1.0, 0.1)
    assert (bl, bu) == orc.fd_band_indices(f, 0.1, 1.0)
    _close(b, orc.fd_band_multiplier(f, 0.1, 1.0, 0.1))


def test_optical_properties_match_oracle():
    """calculate_optical_properties (src/math_tools.rs:663-701), host routine of the library."""
    h = pkg().host
    rng = np.random.default_rng(0)
    f = (np.arange(1, 200, dtype=F32) * F32(0.01)).astype(F32)
    sa, ra = rng.uniform(0.1, 2.0, f.size).astype(F32), rng.uniform(0.5, 3.0, f.size).astype(F32)
    sp, rp = np.cumsum(rng.uniform(-0.5, 0.1, f.size)).astype(F32), np.cumsum(rng.uniform(-0.4, 0.1, f.size)).astype(F32)
    n, al, ka = h.optical_properties(sa, sp, ra, rp, f, 1e-3)
    on, oal, oka = orc.calculate_optical_properties(sa, sp, ra, rp, f, 1e-3)
    np.testing.assert_allclose(n, on, rtol=2e-6, atol=1e-6)
    np.testing.assert_allclose(al, oal, rtol=2e-5, atol=1e-3)
    np.testing.assert_allclose(ka, oka, rtol=2e-5, atol=1e-9)
