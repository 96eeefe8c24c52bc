"""The oracle's C / pthreads twin (the timed CPU baseline) against the numpy oracle."""
import numpy as np
import pytest

from helpers import F32, check_unwrapped_phase, orc, rel_err, slot0, synthetic_cube, time_axis


@pytest.mark.parametrize("n", [64, 1024, 4096])
def test_c_twin_matches_numpy_oracle(n):
    from oracle import c_twin
    w, h = 3, 5
    cube = synthetic_cube(w, h, n, seed=n)
    t = time_axis(n)
    slots = orc.run_default_chain(slot0(cube, t))
    tilt = orc.adapted_blackman_multiplier(t, 0.0, 7.0)
    gb = orc.td_gate_multiplier(t, float(t[0]), float(t[-1]), 2.0)
    win = orc.fft_window_multiplier(t)
    band = orc.fd_band_multiplier(orc.frequency_axis(t))
    ga = orc.td_gate_multiplier(t, float(t[0]), float(t[-1]), 0.1)
    out, img, fft5, amp5, ph4 = c_twin.default_chain(cube, tilt, gb, win, band, ga, want_spectra=True)
    assert rel_err(out, slots[7].data) <= 1e-5
    assert rel_err(img, slots[7].img) <= 1e-5
    assert rel_err(fft5, slots[5].fft) <= 1e-5
    assert rel_err(amp5, slots[5].amplitudes) <= 1e-5
    tol = max(1e-4 * float(np.abs(slots[4].phases).max()), 2e-3)
    check_unwrapped_phase(ph4, slots[4].phases, slots[4].fft, tol)
