"""BASELINE config 2 (a scan normalised by a reference pulse: amplitude-ratio / phase-difference maps) and the GUI
hand-off from device-resident cubes (selected-pixel trace and spectrum, pixel-mean trace and mean spectra,
src/data_thread.rs:1337-1431) against the oracle's stage-by-stage pipeline."""
import numpy as np
import pytest

from helpers import (F32, TOL_TRACE, check_unwrapped_phase, default_multipliers, orc, pkg, rel_err, slot0,
                     synthetic_cube, time_axis)

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    c = pkg().Context(0)
    yield c
    c.close()


@pytest.mark.parametrize("n", [2048, 1000])
def test_config2_reference_normalised_maps(ctx, n):
    """Config 2 at its stand-in shape (100 x 100 x 2048 + a 1 x 1 reference pulse; here 20 x 15 pixels to
    keep the oracle quick, and N = 1000 for the chirp-z path): per pixel A_s / max(A_r, 1e-12) and phi_s - phi_r,
    the operands of calculate_optical_properties (math_tools.rs:665-701), for every pixel."""
    w, h = 20, 15
    cube = synthetic_cube(w, h, n, seed=n, noise=0.005)
    t = time_axis(n)
    tt = np.arange(n) * 0.05 - 9.0
    ref_pulse = (1.3 * np.exp(-(tt / 0.3) ** 2) * np.cos(2 * np.pi * tt)).astype(F32)
    sig, ramp, rph = ctx.reference_pulse(t, t, ref_pulse)            # OpenRef: window + r2c + |s| + unwrap
    osig, oamp, oph = orc.reference_pulse(t, t, ref_pulse)
    win = pkg().host.fft_window(t)
    ctx.plan_trace(n, win, None, None)                               # the fft stage of the chain: window only
    ctx.plan_reference(ramp, rph)
    ratio, dphi = ctx.trace_forward_normalised(cube)
    s4 = orc.fft(slot0(cube, t), orc.ConfigContainer())
    want_ratio = s4.amplitudes / np.maximum(oamp, F32(1e-12))[None, None, :]
    # where the reference spectrum is ~0 (far stop band) the ratio is huge and meaningless: compare where A_r matters
    ok = oamp > 1e-3 * oamp.max()
    assert rel_err(ratio[:, :, ok], want_ratio[:, :, ok]) <= 5 * TOL_TRACE
    # same arithmetic on the GPU's own operands is exact (IEEE division / subtraction in the epilogue)
    amp_gpu = ctx.trace_forward(cube, want=("amp", "phase"))
    assert np.array_equal(ratio, amp_gpu["amp"] / np.maximum(ramp, F32(1e-12))[None, None, :])
    assert np.array_equal(dphi, amp_gpu["phase"] - rph[None, None, :])
    tol = max(TOL_TRACE * float(np.abs(s4.phases).max()), 2e-3)
    check_unwrapped_phase(dphi + rph[None, None, :], s4.phases, s4.fft, tol)
    # one map slice: the bin nearest 1 THz
    f = orc.frequency_axis(t)
    k = int(np.argmin(np.abs(f - 1.0)))
    P, F = w * h, f.size
    d_ratio = ctx.to_device(ratio)
    m = ctx.spectral_slice(d_ratio.ptr, F, k, P)
    assert np.array_equal(m.reshape(w, h), ratio[:, :, k])
    ctx.plan_reference(None, None)
    with pytest.raises(pkg().ThzError):
        ctx.trace_forward_normalised(cube)


def test_pixel_handoff_and_means_from_device_cubes(ctx):
    """After a fused run only the raw and the filtered cube live on the device.  The plots need: the selected
    pixel's raw trace (slot 0), its spectrum as slot fft + 1 holds it (amplitudes x band, phases untouched), its
    filtered trace (last slot), the pixel-mean filtered trace and the mean spectra of `ifft`."""
    n, w, h = 1024, 9, 8
    cube = synthetic_cube(w, h, n, seed=77)
    t, m_pre, band, m_post = default_multipliers(n)
    slots = orc.run_default_chain(slot0(cube, t))
    ctx.plan_trace(n, m_pre, band, m_post)
    P = w * h
    d_raw = ctx.to_device(cube)
    d_fil = ctx.alloc(cube.nbytes)
    d_img = ctx.alloc(P * 4)
    ctx.trace_fused_dev(d_raw.ptr, d_fil.ptr, d_img.ptr, P)
    px, py = 5, 3
    got = ctx.pixel_handoff(d_raw.ptr, d_fil.ptr, P, px * h + py)
    assert np.array_equal(got["raw"], cube[px, py])
    assert rel_err(got["filtered"], slots[7].data[px, py]) <= TOL_TRACE * float(np.abs(slots[7].data).max() / np.abs(slots[7].data[px, py]).max())
    assert rel_err(got["amp"], slots[5].amplitudes[px, py]) <= TOL_TRACE
    assert rel_err(got["fft"], slots[5].fft[px, py]) <= TOL_TRACE
    tol = max(TOL_TRACE * float(np.abs(slots[5].phases).max()), 2e-3)
    check_unwrapped_phase(got["phase"][None, :], slots[5].phases[px, py][None, :], slots[4].fft[px, py][None, :], tol)
    with pytest.raises(pkg().ThzError):
        ctx.pixel_handoff(d_raw.ptr, d_fil.ptr, P, P)            # out of bounds (data_thread.rs:1344-1356)
    # avg_signal: mean over all pixels of the last slot's data
    avg = ctx.mean_trace(d_fil.ptr, n, P)
    want = slots[7].data.astype(np.float64).mean(axis=(0, 1))
    assert rel_err(avg, want) <= TOL_TRACE
    # mean spectra of the band-passed slot, as `ifft` takes them, without spectral cubes
    afft, aamp, aph = ctx.mean_spectra(d_raw.ptr, P)
    assert rel_err(aamp, slots[6].avg_signal_fft) <= TOL_TRACE
    assert rel_err(afft, slots[6].avg_fft) <= TOL_TRACE
    # mean phase: against the f64 mean of the oracle's own phases, modulo the ambiguous unwrap flips of single
    # pixels (a flip of one pixel moves the mean by 2 pi / P)
    want_ph = slots[5].phases.astype(np.float64).mean(axis=(0, 1))
    k = np.round((aph - want_ph) * P / (2 * np.pi))
    assert np.max(np.abs(aph - want_ph - k * 2 * np.pi / P)) <= tol
