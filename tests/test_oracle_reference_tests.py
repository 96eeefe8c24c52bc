"""Pins the CPU oracle against every known-answer test the reference holds for the
hot path: each test below re-creates one `#[test]` of /root/reference
with the same inputs and the same assertions."""
import numpy as np
import pytest

from oracle import thz_oracle as O

F32 = np.float32


def _linspace(a, b, n):
    return O._linspace32(a, b, n)


def test_window_functions_apply():
    """src/math_tools.rs:757-840 `test_window_functions_apply`."""
    size = 128
    time = _linspace(0.0, 1.0, size)
    ones = np.ones(size, F32)
    sig_blackman = ones * O.blackman_multiplier(time)
    sig_hann = ones * O.hanning_multiplier(time)
    sig_hamm = ones * O.hamming_multiplier(time)
    sig_flattop = ones * O.flat_top_multiplier(time)
    sig_adapted = O.apply_adapted_blackman_window(ones.copy(), time, 0.1, 0.1)

    for s in (sig_blackman, sig_hann, sig_flattop, sig_adapted):
        assert s[0] <= 1e-5 and s[-1] <= 1e-5
    expected_hamm_end = 0.54 - 0.46
    assert abs(sig_hamm[0] - expected_hamm_end) <= 1e-5
    assert abs(sig_hamm[-1] - expected_hamm_end) <= 1e-5
    for s in (sig_blackman, sig_hann, sig_hamm, sig_flattop, sig_adapted):
        np.testing.assert_allclose(s, s[::-1], atol=1e-5, rtol=0)
        mid = size // 2
        assert s[mid] >= s[mid - 1] and s[mid] >= s[mid + 1]
    assert abs(sig_adapted[size // 2] - 1.0) <= 1e-5


def _single_pixel(n, signal, time):
    d = O.ScannedImageFilterData()
    d.time = time
    d.data = signal.reshape(1, 1, n).astype(F32)
    nf = n // 2 + 1
    d.frequency = (np.arange(nf, dtype=F32) / F32(50.0)).astype(F32)
    d.phases = np.zeros((1, 1, nf), F32)
    d.amplitudes = np.zeros((1, 1, nf), F32)
    d.fft = np.zeros((1, 1, nf), np.complex64)
    d.width = d.height = 1
    return d


def test_fft_roundtrip():
    """src/math_tools.rs:843-897 `test_fft_roundtrip`."""
    n, k1, k2 = 128, 3, 7
    tt = (np.arange(n, dtype=F32) / F32(n)).astype(F32)
    sig = (np.sin(F32(2.0) * O.PI32 * F32(k1) * tt) + F32(0.5) * np.cos(F32(2.0) * O.PI32 * F32(k2) * tt)).astype(F32)
    inp = _single_pixel(n, sig, _linspace(0.0, 1.0, n))
    cfg = O.ConfigContainer(fft_window=(0.0, 0.0), fft_window_type="AdaptedBlackman", avg_in_fourier_space=False)
    after_fft = O.fft(inp, cfg)
    expected_time = after_fft.data.copy()
    after_ifft = O.ifft(after_fft, cfg)
    np.testing.assert_allclose(after_ifft.data, expected_time, atol=1e-4, rtol=0)
    # window [0,0] only touches sample 0 (bw(0,0)=NaN -> 1) and the last sample (bw(0,0) too)
    np.testing.assert_array_equal(after_fft.data, inp.data)


def test_fd_band_pass_zeros_outside_band_and_keeps_data_inside():
    """src/filters/band_pass_fd.rs:474-567."""
    n, k = 256, 9
    t = np.arange(n, dtype=F32)
    sig = np.sin(F32(2.0) * O.PI32 * F32(k) * t / F32(n)).astype(F32)
    inp = _single_pixel(n, sig, _linspace(0.0, 1.0, n))
    cfg = O.ConfigContainer(fft_window=(0.0, 0.0))
    input_fd = O.fft(inp, cfg)
    peak_idx = int(np.argmax(input_fd.amplitudes[0, 0]))
    assert peak_idx == k
    freq = input_fd.frequency
    i_lo = max(peak_idx - 2, 0)
    i_hi = min(peak_idx + 2, freq.shape[0] - 1)
    flt = O.FrequencyDomainBandPass(low=float(freq[i_lo]), high=float(freq[i_hi]), window_width=0.0)
    out = flt.filter(input_fd)
    assert out.fft.shape == input_fd.fft.shape
    assert out.amplitudes.shape == input_fd.amplitudes.shape
    lower, upper = O.fd_band_indices(out.frequency, flt.low, flt.high)
    assert (lower, upper) == (i_lo, i_hi + 1)
    assert np.all(out.amplitudes[0, 0, :lower] == 0.0)
    assert np.all(out.amplitudes[0, 0, upper:] == 0.0)
    assert np.any(out.amplitudes[0, 0, lower:upper] > 0.0)
    # phases untouched (quirk 4)
    np.testing.assert_array_equal(out.phases, input_fd.phases)


@pytest.mark.parametrize("ww_default", [2.0, 0.1])
def test_td_band_pass_zeros_outside_band_and_keeps_data_inside(ww_default):
    """src/filters/band_pass_td_before_fft.rs:389-443 and band_pass_td_after_fft.rs:388-442."""
    n, k = 256, 9
    t = np.arange(n, dtype=F32)
    sig = np.sin(F32(2.0) * O.PI32 * F32(k) * t / F32(n)).astype(F32)
    inp = _single_pixel(n, sig, _linspace(0.0, 1.0, n))
    flt = O.TimeDomainBandPass(window_width=ww_default)
    flt.window_width = 0.0
    flt.low, flt.high = 0.25, 0.55
    out = flt.filter(inp)
    assert out.time.shape == inp.time.shape and out.data.shape == inp.data.shape
    time = out.time
    safe_low = F32(max(flt.low, 0.0))
    safe_high = F32(min(flt.high, float(time[-1])))
    lower = int(np.nonzero(time >= safe_low)[0][0])
    upper = int(np.nonzero(time <= safe_high)[0][-1]) + 1
    assert np.all(out.data[0, 0, :lower] == 0.0)
    assert np.all(out.data[0, 0, upper:] == 0.0)
    assert np.any(out.data[0, 0, lower:upper] > 0.0)


def _tilt_input(impulse_idx):
    n, dt, width, height = 64, F32(0.05), 2, 2
    data = np.zeros((width, height, n), F32)
    data[1, 1, impulse_idx] = 1.0
    d = O.ScannedImageFilterData()
    d.time = _linspace(0.0, dt * (F32(n) - F32(1.0)), n)
    d.data = data
    d.dx = d.dy = 1.0
    d.width, d.height = width, height
    return d, n


def test_tilt_compensation_extends_time_and_shifts_center_trace():
    """src/filters/tilt_compensation.rs:302-346."""
    impulse_idx = 10
    inp, n = _tilt_input(impulse_idx)
    flt = O.TiltCompensation(tilt_x=10.0, tilt_y=0.0)
    out = flt.filter(inp)
    expected_steps = flt.num_steps(2, 2, 1.0, 1.0)[0]
    assert expected_steps > 0
    assert out.time.shape[0] == n + 2 * expected_steps
    assert int(np.argmax(out.data[1, 1])) == impulse_idx + expected_steps


def test_tilt_compensation_no_tilt_no_extension():
    """src/filters/tilt_compensation.rs:348-389."""
    impulse_idx = 12
    inp, n = _tilt_input(impulse_idx)
    out = O.TiltCompensation(0.0, 0.0).filter(inp)
    assert out.time.shape[0] == n
    assert int(np.argmax(out.data[1, 1])) == impulse_idx


def test_deconvolution_shape_preservation(psf_npz_path):
    """src/filters/deconvolution.rs:1138-1177 (2x2 image -> the <16 early return)."""
    inp, n = _tilt_input(12)
    flt = O.Deconvolution(n_iterations=10, n_filters=20, start_freq=0.25, end_freq=4.0, win_width=0.5)
    psf = O.load_psf(psf_npz_path)
    out = flt.filter(inp, psf)
    assert out.time.shape[0] == n
    assert out.data.shape == inp.data.shape
    np.testing.assert_array_equal(out.data, inp.data)


def test_psf_npz_fixture_matches_survey(psf_npz_path):
    """with the shipped psf.npz, 8 bands 0.1-10 THz and
    dx=dy=0.5 mm the PSF sizes and RL iteration counts are the published ones."""
    psf = O.load_psf(psf_npz_path)
    time = (F32(1000.0) + F32(0.05) * np.arange(2048, dtype=F32)).astype(F32)
    flt = O.Deconvolution(n_iterations=500, n_filters=8, start_freq=0.1, end_freq=10.0, win_width=0.5)
    bands, reason = flt.plan(time, (256, 256, 2048), 0.5, 0.5, psf)
    assert reason is None
    assert [b.psf.shape for b in bands] == [(47, 57), (31, 29), (17, 15), (9, 9), (7, 7), (7, 7), (7, 7), (7, 7)]
    assert [b.n_iter for b in bands] == [423, 251, 127, 46, 13, 4, 3, 1]
    for b in bands:
        assert b.psf.max() == 1.0
        np.testing.assert_array_equal(b.psf, np.outer(b.psf_x, b.psf_y).astype(F32))
