"""The C++ host layer (thzhost::ChainDriver, the mirror of src/main.rs:194-268 chain assembly and
the data_thread.rs:1090-1316 driver loop) against the oracle's stage-by-stage pipeline slots."""
import numpy as np
import pytest

from helpers import (F32, TOL_MAP, TOL_TRACE, check_unwrapped_phase, orc, pkg, rel_err, slot0, synthetic_cube,
                     time_axis)

pytestmark = pytest.mark.gpu

EXPECTED_ORDER = ["scaling", "Tilt Compensation", "Time Domain Band Pass (before FFT)", "fft",
                  "Frequency Domain Band Pass", "ifft", "Time Domain Band Pass (after FFT)", "Deconvolution"]


@pytest.fixture(scope="module")
def ctx():
    c = pkg().Context(0)
    yield c
    c.close()


def test_chain_order_matches_reference(ctx):
    ch = pkg().Chain(ctx)
    assert ch.stages() == EXPECTED_ORDER     # 8 stages -> 9 pipeline slots
    assert ch.get_param("Frequency Domain Band Pass", "low") == pytest.approx(0.2)
    assert ch.get_param("Time Domain Band Pass (before FFT)", "window_width") == 2.0
    assert ch.get_param("Time Domain Band Pass (after FFT)", "window_width") == pytest.approx(0.1)
    assert ch.get_param("Deconvolution", "n_filters") == 25


@pytest.mark.parametrize("n,dxdy", [(1024, True), (2048, False)])
def test_stage_by_stage_slots_match_oracle(ctx, n, dxdy):
    w, h = 6, 9
    cube = synthetic_cube(w, h, n, seed=n)
    t = time_axis(n)
    slots = orc.run_default_chain(slot0(cube, t, dx=0.5 if dxdy else None, dy=0.5 if dxdy else None))
    ch = pkg().Chain(ctx)
    ch.open(t, cube, dx=0.5 if dxdy else None, dy=0.5 if dxdy else None)
    ch.run(start_idx=1)
    assert np.array_equal(ch.slot(0)["img"].shape, (w, h))
    assert rel_err(ch.slot(0)["img"], slots[0].img) <= 1e-6
    assert np.array_equal(ch.slot(1)["data"], slots[1].data)   # scaling (s = 1): clone
    for i in (2, 3):                          # tilt taper, gate before FFT: one f32 multiply each; the taper
        assert rel_err(ch.slot(i)["data"], slots[i].data) <= 1e-6, i   # vectors may differ by 1 ulp (cosf)
    s4 = ch.slot(4)                           # fft
    assert rel_err(s4["data"], slots[4].data) <= 1e-6
    assert rel_err(s4["fft"], slots[4].fft) <= TOL_TRACE
    assert rel_err(s4["amplitudes"], slots[4].amplitudes) <= TOL_TRACE
    tol_rad = max(TOL_TRACE * float(np.abs(slots[4].phases).max()), 2e-3)
    check_unwrapped_phase(s4["phases"], slots[4].phases, slots[4].fft, tol_rad)
    s5 = ch.slot(5)                           # FD band-pass: fft and amplitudes scaled, phases untouched
    assert rel_err(s5["fft"], slots[5].fft) <= TOL_TRACE
    assert rel_err(s5["amplitudes"], slots[5].amplitudes) <= TOL_TRACE
    assert np.array_equal(s5["phases"], s4["phases"])
    s6 = ch.slot(6)                           # ifft (+ pixel means)
    assert rel_err(s6["data"], slots[6].data) <= TOL_TRACE
    assert rel_err(s6["avg_fft"], slots[6].avg_fft) <= TOL_TRACE
    assert rel_err(s6["avg_signal_fft"], slots[6].avg_signal_fft) <= TOL_TRACE
    s7 = ch.slot(7)                           # gate after iFFT
    assert rel_err(s7["data"], slots[7].data) <= TOL_TRACE
    s8 = ch.slot(8)                           # deconvolution inactive: clone of slot 7, driver's intensity image
    assert np.array_equal(s8["data"], s7["data"])
    assert rel_err(s8["img"], slots[7].img) <= TOL_TRACE
    # the fused kernel gives the same last slot
    fused, img = ch.run_fused()
    assert rel_err(fused, s7["data"]) <= 2e-6 and rel_err(img, s8["img"]) <= 1e-5


def test_partial_update_and_inactive_filter(ctx):
    """UpdateType::Filter(idx) recomputes from that stage on; an inactive filter clones its input
    (data_thread.rs:1186-1187)."""
    n, w, h = 512, 4, 5
    cube = synthetic_cube(w, h, n, seed=1)
    t = time_axis(n)
    ch = pkg().Chain(ctx)
    ch.open(t, cube, 0.5, 0.5)
    ch.run(1)
    before = ch.slot(7)["data"]
    ch.set_param("Frequency Domain Band Pass", "high", 2.0)
    ch.run(start_idx=ch.slot_of("Frequency Domain Band Pass"))
    after = ch.slot(7)["data"]
    assert not np.array_equal(before, after)
    p = orc.ChainParams()
    p.band.high = 2.0
    ref = orc.run_default_chain(slot0(cube, t), p)
    assert rel_err(after, ref[7].data) <= TOL_TRACE
    ch.set_active("Time Domain Band Pass (after FFT)", False)
    ch.run(start_idx=ch.slot_of("Time Domain Band Pass (after FFT)"))
    assert np.array_equal(ch.slot(7)["data"], ch.slot(6)["data"])


def test_deconvolution_only_on_apply(ctx, psf_npz_path):
    """The deconvolution runs only when the update starts at it (its Apply button,
    data_thread.rs:1139-1150) and needs the filter to be active and a PSF to be loaded."""
    m = pkg()
    n, w, h = 256, 36, 32
    cube = synthetic_cube(w, h, n, seed=2, noise=0.02)
    t = time_axis(n)
    ch = m.Chain(ctx)
    ch.open(t, cube, 1.0, 1.0)
    ch.set_active("Deconvolution", True)
    ch.set_param("Deconvolution", "n_filters", 4)
    ch.set_param("Deconvolution", "n_iterations", 10)
    ch.run(1, run_deconvolution=True)      # an earlier filter ran first -> deconvolution is skipped
    assert np.array_equal(ch.slot(8)["data"], ch.slot(7)["data"])
    ch.run(ch.slot_of("Deconvolution"), run_deconvolution=True)   # no PSF loaded -> input returned
    assert np.array_equal(ch.slot(8)["data"], ch.slot(7)["data"])
    psf = m.host.PSF.load(psf_npz_path)
    ch.set_psf(psf)
    ch.run(ch.slot_of("Deconvolution"), run_deconvolution=True)
    s7, s8 = ch.slot(7), ch.slot(8)
    assert not np.array_equal(s8["data"], s7["data"])
    assert ch.filter_ms("Deconvolution") > 0
    opsf = orc.load_psf(psf_npz_path)
    f = orc.frequency_axis(t)
    sin = orc.ScannedImageFilterData(time=t, data=s7["data"], frequency=f, img=orc.intensity_image(s7["data"]),
                                     dx=1.0, dy=1.0, width=w, height=h)
    ref = orc.Deconvolution(n_filters=4, n_iterations=10).filter(sin, opsf)
    assert rel_err(s8["data"], ref.data) <= TOL_MAP and rel_err(s8["img"], ref.img) <= TOL_MAP


def test_run_fused_with_deconvolution(ctx, psf_npz_path):
    """ChainDriver::run_fused(run_deconvolution) == the stage-by-stage driver ending in Apply."""
    m = pkg()
    n, w, h = 256, 36, 32
    cube = synthetic_cube(w, h, n, seed=8, noise=0.02)
    t = time_axis(n)
    ch = m.Chain(ctx)
    ch.open(t, cube, 1.0, 1.0)
    ch.set_active("Deconvolution", True)
    ch.set_param("Deconvolution", "n_filters", 4)
    ch.set_param("Deconvolution", "n_iterations", 10)
    ch.set_psf(m.host.PSF.load(psf_npz_path))
    ch.run(1)
    ch.run(ch.slot_of("Deconvolution"), run_deconvolution=True)
    s8 = ch.slot(8)
    out, img = ch.run_fused(run_deconvolution=True)
    assert rel_err(out, s8["data"]) <= 1e-4 and rel_err(img, s8["img"]) <= 1e-4


def test_chain_with_downscaling(ctx):
    """scale_factor = 2: the driver's scaling stage halves the image; later stages run on 3x4 pixels."""
    n, w, h = 512, 6, 8
    cube = synthetic_cube(w, h, n, seed=6)
    t = time_axis(n)
    ch = pkg().Chain(ctx)
    ch.set_config(scale_factor=2)
    ch.open(t, cube, 0.5, 0.5)
    ch.run(1)
    p = orc.ChainParams()
    p.config.scale_factor = 2
    ref = orc.run_default_chain(slot0(cube, t), p)
    ch.shape = (w // 2, h // 2, n)
    assert np.array_equal(ch.slot(1)["data"], ref[1].data)
    assert rel_err(ch.slot(7)["data"], ref[7].data) <= TOL_TRACE


def test_run_fused_honours_scale_factor(ctx, psf_npz_path):
    """run_fused with scale_factor = 2 == the staged run: the first chain slot (`scaling`, math_tools.rs:242-310)
    block-averages the cube and doubles dx / dy before anything else; the fused kernels and the deconvolution
    plan must see the scaled cube (an earlier build fed them the raw one)."""
    m = pkg()
    n, w, h = 256, 72, 64
    cube = synthetic_cube(w, h, n, seed=18, noise=0.02)
    t = time_axis(n)
    ch = m.Chain(ctx)
    ch.set_config(scale_factor=2)
    ch.open(t, cube, 0.5, 0.5)
    ch.set_active("Deconvolution", True)
    ch.set_param("Deconvolution", "n_filters", 4)
    ch.set_param("Deconvolution", "n_iterations", 10)
    ch.set_psf(m.host.PSF.load(psf_npz_path))
    ch.run(1)
    ch.run(ch.slot_of("Deconvolution"), run_deconvolution=True)
    ch.shape = (w // 2, h // 2, n)
    s8 = ch.slot(8)
    assert s8["data"].shape == (w // 2, h // 2, n)
    out, img = ch.run_fused(run_deconvolution=True)
    assert out.shape == s8["data"].shape
    assert rel_err(out, s8["data"]) <= 1e-4 and rel_err(img, s8["img"]) <= 1e-4


def test_chain_driver_non_power_of_two(ctx, psf_npz_path):
    """The whole driver (stage by stage and fused, with deconvolution) on N = 1000 samples."""
    m = pkg()
    n, w, h = 1000, 34, 32
    cube = synthetic_cube(w, h, n, seed=13, noise=0.02)
    t = time_axis(n)
    slots = orc.run_default_chain(slot0(cube, t, 1.0, 1.0))
    ch = m.Chain(ctx)
    ch.open(t, cube, 1.0, 1.0)
    ch.run(1)
    assert rel_err(ch.slot(4)["fft"], slots[4].fft) <= TOL_TRACE
    assert rel_err(ch.slot(7)["data"], slots[7].data) <= TOL_TRACE
    assert rel_err(ch.slot(8)["img"], slots[7].img) <= TOL_TRACE
    ch.set_active("Deconvolution", True)
    ch.set_param("Deconvolution", "n_filters", 4)
    ch.set_param("Deconvolution", "n_iterations", 8)
    ch.set_psf(m.host.PSF.load(psf_npz_path))
    out, img = ch.run_fused(run_deconvolution=True)
    opsf = orc.load_psf(psf_npz_path)
    sin = orc.ScannedImageFilterData(time=t, data=slots[7].data, frequency=orc.frequency_axis(t),
                                     img=slots[7].img, dx=1.0, dy=1.0, width=w, height=h)
    ref = orc.Deconvolution(n_filters=4, n_iterations=8).filter(sin, opsf)
    assert rel_err(out, ref.data) <= TOL_MAP and rel_err(img, ref.img) <= TOL_MAP


def test_non_zero_tilt_extends_the_axis(ctx):
    """TiltCompensation at 10 / 4 degrees: per-pixel shift, axis extended by 2 * num_steps (the reference's
    own tilt tests, tilt_compensation.rs:302-389, check extension length and impulse position); the rest of
    the chain then runs on a non power-of-two length."""
    n, w, h = 256, 6, 5
    cube = synthetic_cube(w, h, n, seed=17)
    cube[2, 3, :] = 0.0
    cube[2, 3, 100] = 1.0                       # impulse
    t = time_axis(n)
    ch = pkg().Chain(ctx)
    ch.set_param("Tilt Compensation", "tilt_x", 10.0)
    ch.set_param("Tilt Compensation", "tilt_y", 4.0)
    ch.open(t, cube, 0.5, 0.5)
    ch.run(1)
    p = orc.ChainParams()
    p.tilt.tilt_x, p.tilt.tilt_y = 10.0, 4.0
    ref = orc.run_default_chain(slot0(cube, t), p)
    n_ext = ref[2].data.shape[2]
    assert n_ext > n
    ch.shape = (w, h, n_ext)
    s2 = ch.slot(2)
    assert s2["data"].shape == ref[2].data.shape
    assert rel_err(s2["data"], ref[2].data) <= 1e-6
    assert int(np.argmax(s2["data"][2, 3])) == int(np.argmax(ref[2].data[2, 3]))
    assert rel_err(ch.slot(7)["data"], ref[7].data) <= TOL_TRACE
    assert rel_err(ch.slot(8)["img"], ref[7].img) <= TOL_TRACE
