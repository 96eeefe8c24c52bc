"""GPU parity tests of the trace pass (libthzgpu through its C ABI) against the CPU oracle.
Tolerances are the north star's: max|a-b|/max|b| <= 1e-4 on spectra and filtered traces."""
import numpy as np
import pytest

from helpers import (F32, TOL_TRACE, check_unwrapped_phase, default_multipliers, orc, pkg, rel_err, slot0,
                     synthetic_cube, time_axis)

pytestmark = pytest.mark.gpu

ALL_N = [64, 128, 256, 512, 1024, 2048, 4096, 8192]


@pytest.fixture(scope="module")
def ctx():
    m = pkg()
    c = m.Context(0)
    yield c
    c.close()


def _chain_oracle(cube, t, dx_dy=True):
    s0 = slot0(cube, t, dx=0.5 if dx_dy else None, dy=0.5 if dx_dy else None)
    return orc.run_default_chain(s0)


@pytest.mark.parametrize("n", ALL_N)
def test_forward_matches_oracle_fft(ctx, n):
    """math_tools::fft: windowed data, spectrum, amplitudes, unwrapped phases; odd trace count
    so the last pair is half empty and the last CTA pass is ragged."""
    w, h = 3, 37 if n <= 2048 else 11
    cube = synthetic_cube(w, h, n, seed=n)
    t = time_axis(n)
    cfg = orc.ConfigContainer()
    ref = orc.fft(slot0(cube, t), cfg)
    win = orc.fft_window_multiplier(t, cfg.fft_window_type, cfg.fft_window)
    ctx.plan_trace(n, m_pre=win)
    got = ctx.trace_forward(cube)
    assert np.array_equal(got["windowed"], ref.data), "windowed trace must be bit-exact (one f32 multiply)"
    assert rel_err(got["fft"], ref.fft) <= TOL_TRACE
    assert rel_err(got["amp"], ref.amplitudes) <= TOL_TRACE
    # phases: 1e-4 of the largest phase, but never tighter than 2e-3 rad
    tol_rad = max(TOL_TRACE * float(np.abs(ref.phases).max()), 2e-3)
    check_unwrapped_phase(got["phase"], ref.phases, ref.fft, tol_rad)
    # float64 ground truth bounds both implementations
    truth = np.fft.rfft(ref.data.astype(np.float64), axis=-1)
    assert rel_err(got["fft"], truth.astype(np.complex128)) <= 2e-6 * np.log2(n)


@pytest.mark.parametrize("n", ALL_N)
def test_fused_chain_matches_oracle(ctx, n):
    """Slots 2..7 of the default chain in one kernel vs the oracle run stage by stage."""
    w, h = 5, 13
    cube = synthetic_cube(w, h, n, seed=100 + n)
    t, m_pre, band, m_post = default_multipliers(n)
    slots = _chain_oracle(cube, t)
    ctx.plan_trace(n, m_pre, band, m_post)
    out, img = ctx.trace_fused(cube)
    assert rel_err(out, slots[7].data) <= TOL_TRACE
    assert rel_err(img, slots[7].img) <= TOL_TRACE
    # gate zeros are exact (the default gates zero the last sample)
    assert np.all(out[..., -1] == 0.0)


@pytest.mark.parametrize("n", [64, 256, 2048, 4096])
def test_inverse_matches_oracle_ifft(ctx, n):
    w, h = 4, 9
    cube = synthetic_cube(w, h, n, seed=7 + n)
    t, m_pre, band, m_post = default_multipliers(n)
    slots = _chain_oracle(cube, t)
    ctx.plan_trace(n, m_pre, band, m_post)
    # plain ifft of the band-passed spectrum (slot 5 -> slot 6)
    out, _ = ctx.trace_inverse(slots[5].fft)
    assert rel_err(out, slots[6].data) <= TOL_TRACE
    # band-pass + ifft + gate fused (slot 4 -> slot 7)
    out, img = ctx.trace_inverse(slots[4].fft, use_band=True, use_post=True, want_img=True)
    assert rel_err(out, slots[7].data) <= TOL_TRACE
    assert rel_err(img, slots[7].img) <= TOL_TRACE


def test_stagewise_device_chain_equals_fused(ctx):
    """forward -> band_apply -> inverse on device buffers == fused kernel (<= 1e-6 of peak), and
    the FD band-pass leaves exact zeros outside [lower, upper) and phases untouched."""
    n, w, h = 2048, 6, 10
    P = w * h
    F = n // 2 + 1
    cube = synthetic_cube(w, h, n, seed=5)
    t, m_pre, band, m_post = default_multipliers(n)
    ctx.plan_trace(n, m_pre, band, m_post)
    d_in = ctx.to_device(cube)
    d_fft = ctx.alloc(P * F * 8)
    d_amp = ctx.alloc(P * F * 4)
    d_ph = ctx.alloc(P * F * 4)
    d_out = ctx.alloc(P * n * 4)
    d_img = ctx.alloc(P * 4)
    ctx.trace_forward_dev(d_in.ptr, None, d_fft.ptr, d_amp.ptr, d_ph.ptr, P)
    ph_before = d_ph.download((P, F))
    ctx.band_apply_dev(d_fft.ptr, d_amp.ptr, P)
    amp = d_amp.download((P, F))
    fftv = d_fft.download((P, F), np.complex64)
    lower, upper = orc.fd_band_indices(orc.frequency_axis(t), 0.2, 5.0)
    assert np.all(amp[:, :lower] == 0) and np.all(amp[:, upper:] == 0)
    assert np.all(fftv[:, :lower] == 0) and np.all(fftv[:, upper:] == 0)
    assert np.all(amp[:, lower + 3:upper - 3] > 0)
    assert np.array_equal(ph_before, d_ph.download((P, F)))
    ctx.trace_inverse_dev(d_fft.ptr, 0, 1, d_out.ptr, d_img.ptr, P)
    staged = d_out.download((P, n))
    fused, img = ctx.trace_fused(cube)
    assert rel_err(staged, fused.reshape(P, n)) <= 2e-6
    assert rel_err(d_img.download((P,)), img.reshape(P)) <= 1e-5
    # in-place fused on the device buffer gives the same bits as the host-pointer path
    ctx.trace_fused_dev(d_in.ptr, d_in.ptr, d_img.ptr, P)
    assert np.array_equal(d_in.download((P, n)), fused.reshape(P, n))


def test_reference_fft_roundtrip(ctx):
    """test_fft_roundtrip (src/math_tools.rs:843-897): 1x1x128 two-tone signal, window
    disabled, ifft(fft(x)) == x to 1e-4 absolute."""
    n = 128
    t = np.linspace(0, 1, n, dtype=F32)
    x = (np.sin(2 * np.pi * 5 * t) + 0.5 * np.sin(2 * np.pi * 12 * t)).astype(F32).reshape(1, 1, n)
    ctx.plan_trace(n)  # fft_window = [0, 0] multiplies by exactly 1
    spec = ctx.trace_forward(x, want=("fft",))["fft"]
    back, _ = ctx.trace_inverse(spec)
    assert np.max(np.abs(back - x)) < 1e-4


def test_reference_fd_bandpass_zeros(ctx):
    """band_pass_fd.rs:474-567: 1x1x256 sine at bin 9, frequency = i/50."""
    n = 256
    i = np.arange(n, dtype=F32)
    x = np.sin(2 * np.pi * 9 * i / n).astype(F32).reshape(1, 1, n)
    freq = (np.arange(n // 2 + 1, dtype=F32) / F32(50.0)).astype(F32)
    band = orc.fd_band_multiplier(freq, 0.1, 1.0, 0.1)
    lower, upper = orc.fd_band_indices(freq, 0.1, 1.0)
    ctx.plan_trace(n, None, band, None)
    d_in = ctx.to_device(x)
    F = n // 2 + 1
    d_fft, d_amp = ctx.alloc(F * 8), ctx.alloc(F * 4)
    ctx.trace_forward_dev(d_in.ptr, None, d_fft.ptr, d_amp.ptr, None, 1)
    ctx.band_apply_dev(d_fft.ptr, d_amp.ptr, 1)
    amp = d_amp.download((F,))
    assert amp.shape == (F,)
    assert np.all(amp[:lower] == 0.0) and np.all(amp[upper:] == 0.0)
    assert np.any(amp[lower:upper] > 0.0)


def test_reference_td_gate_zeros(ctx):
    """band_pass_td_before_fft.rs:389-443: 1x1x256, gate 0.25..0.55, window_width 0 -> exact
    zeros outside, signal inside (the gate is a multiplier vector applied by the kernels)."""
    n = 256
    t = np.linspace(0, 1, n, dtype=F32)
    x = np.ones((1, 1, n), F32)
    gate = orc.td_gate_multiplier(t, 0.25, 0.55, 0.0)
    lower, upper, _, _ = orc.td_gate_indices(t, 0.25, 0.55)
    ctx.plan_trace(n, gate, None, None)
    got = ctx.trace_forward(x, want=("windowed",))["windowed"].reshape(n)
    assert np.all(got[:lower] == 0.0) and np.all(got[upper:] == 0.0)
    assert np.sum(got[lower:upper] ** 2) > 0.0


def test_empty_and_single_trace(ctx):
    n = 1024
    t, m_pre, band, m_post = default_multipliers(n)
    ctx.plan_trace(n, m_pre, band, m_post)
    out, img = ctx.trace_fused(np.zeros((0, 0, n), F32))
    assert out.shape == (0, 0, n) and img.shape == (0, 0)
    cube = synthetic_cube(1, 1, n, seed=3)
    out, img = ctx.trace_fused(cube)
    ref = _chain_oracle(cube, t)[7]
    assert rel_err(out, ref.data) <= TOL_TRACE and rel_err(img, ref.img) <= TOL_TRACE


def test_linearity_and_parseval_large(ctx):
    """Size-independent properties on a device-generated cube (no oracle at this size):
    chain(a x) == a chain(x); Parseval for the un-gated forward transform."""
    n, w, h = 4096, 48, 64
    P = w * h
    F = n // 2 + 1
    t, m_pre, band, m_post = default_multipliers(n)
    ctx.plan_trace(n, m_pre, band, m_post)
    d_x = ctx.alloc(P * n * 4)
    ctx.generate_cube(d_x, w, h, n)
    x = d_x.download((P, n))
    assert np.isfinite(x).all() and float(np.abs(x).max()) > 0.5
    y1, img1 = ctx.trace_fused(x)
    y2, img2 = ctx.trace_fused((F32(2.0) * x).astype(F32))
    assert np.array_equal(y2, F32(2.0) * y1)            # power-of-two scaling is exact in f32
    assert np.array_equal(img2, F32(4.0) * img1)
    ctx.plan_trace(n)                                    # no window
    spec = ctx.trace_forward(x, want=("fft",))["fft"]
    e_t = np.sum(x.astype(np.float64) ** 2, axis=-1)
    wgt = np.full(F, 2.0)
    wgt[0] = wgt[-1] = 1.0
    e_f = np.sum(wgt * np.abs(spec.astype(np.complex128)) ** 2, axis=-1) / n
    assert np.max(np.abs(e_f - e_t) / e_t) < 1e-5


def test_spectral_means(ctx):
    n, w, h = 1024, 7, 9
    P, F = w * h, n // 2 + 1
    cube = synthetic_cube(w, h, n, seed=11)
    t, m_pre, band, m_post = default_multipliers(n)
    slots = _chain_oracle(cube, t)
    ctx.plan_trace(n, m_pre, band, m_post)
    d_in = ctx.to_device(cube)
    d_fft, d_amp, d_ph = ctx.alloc(P * F * 8), ctx.alloc(P * F * 4), ctx.alloc(P * F * 4)
    ctx.trace_forward_dev(d_in.ptr, None, d_fft.ptr, d_amp.ptr, d_ph.ptr, P)
    ctx.band_apply_dev(d_fft.ptr, d_amp.ptr, P)
    a_fft, a_amp, a_ph = ctx.spectral_means(d_fft.ptr, d_amp.ptr, d_ph.ptr, P)
    assert rel_err(a_fft, slots[6].avg_fft) <= TOL_TRACE
    assert rel_err(a_amp, slots[6].avg_signal_fft) <= TOL_TRACE
    # mean phases inherit unwrap ambiguities of single traces: compare with the means of the GPU's own phases
    ph = d_ph.download((P, F)).astype(np.float64).mean(axis=0)
    assert rel_err(a_ph, ph.astype(F32)) <= 1e-5


def test_all_zero_traces_stay_exactly_zero(ctx):
    """A dead pixel (all-zero trace) gives exact zeros in the reference, which transforms every
    trace on its own.  Traces are transformed in pairs here; the zero one must not pick up
    rounding leakage from its neighbour (and its phase must be atan2(0, 0) = 0)."""
    for n in (256, 2048, 4096):
        w, h = 3, 6
        cube = synthetic_cube(w, h, n, seed=21)
        cube[0, 1] = 0.0      # second of a pair
        cube[1, 2] = 0.0      # first of a pair
        cube[2, 4:6] = 0.0    # both of a pair
        t, m_pre, band, m_post = default_multipliers(n)
        ctx.plan_trace(n, m_pre, band, m_post)
        out, img = ctx.trace_fused(cube)
        for (i, j) in [(0, 1), (1, 2), (2, 4), (2, 5)]:
            assert not out[i, j].any() and img[i, j] == 0.0
        assert np.abs(out[0, 0]).max() > 0
        got = ctx.trace_forward(cube, want=("fft", "amp", "phase"))
        for (i, j) in [(0, 1), (1, 2), (2, 4), (2, 5)]:
            assert not got["fft"][i, j].any() and not got["amp"][i, j].any() and not got["phase"][i, j].any()
        spec = got["fft"].copy()
        back, _ = ctx.trace_inverse(spec)
        for (i, j) in [(0, 1), (1, 2), (2, 4), (2, 5)]:
            assert not back[i, j].any()


@pytest.mark.parametrize("s", [2, 3])
def test_scaling_block_mean_is_bit_exact(ctx, s):
    """`scaling` (src/math_tools.rs:242-310): block means of data / amplitudes / phases / fft in the
    reference's accumulation order; width and height not divisible by s drop the remainder."""
    w, h, n = 7, 9, 128
    rng = np.random.default_rng(s)
    cube = rng.standard_normal((w, h, n)).astype(F32)
    spec = (rng.standard_normal((w, h, n // 2 + 1)) + 1j * rng.standard_normal((w, h, n // 2 + 1))).astype(np.complex64)
    t = time_axis(n)
    inp = slot0(cube, t)
    inp.fft = spec
    inp.amplitudes = np.abs(spec).astype(F32)
    inp.phases = np.angle(spec).astype(F32)
    ref = orc.scaling(inp, orc.ConfigContainer(scale_factor=s))
    assert np.array_equal(ctx.scale_blocks(cube, s), ref.data)
    assert np.array_equal(ctx.scale_blocks(inp.amplitudes, s), ref.amplitudes)
    assert np.array_equal(ctx.scale_blocks(spec, s), ref.fft)
    assert ref.data.shape == (w // s, h // s, n)


def test_load_path_bias_subtraction(ctx):
    """src/io.rs:578-596: x <- x - x[0] per trace and the initial intensity image."""
    raw = synthetic_cube(4, 5, 256, seed=2) + F32(0.37)
    t = time_axis(256)
    ref = orc.load_scan(t, raw, 0.5, 0.5)
    data, img = ctx.bias_subtract(raw)
    assert np.array_equal(data, ref.data)
    assert rel_err(img, ref.img) <= 1e-6


def test_roi_polygon_average(ctx):
    """average_polygon_roi (src/math_tools.rs:599-661): same pixels (unsigned ray casting, row flip) and the
    same sequential f32 summation order -> bit-exact."""
    rng = np.random.default_rng(4)
    data = rng.standard_normal((24, 30, 65)).astype(F32)
    for poly, s in [([(3, 2), (20, 4), (25, 18), (8, 21)], 1), ([(10, 4), (50, 10), (40, 40), (6, 30)], 2),
                    ([(0, 0), (29, 0), (29, 23), (0, 23)], 1)]:
        ref = orc.average_polygon_roi(data, poly, s)
        got = ctx.roi_average(data, poly, s)
        assert np.array_equal(got, ref), (poly, s)
    assert np.abs(orc.average_polygon_roi(data, [(3, 2), (20, 4), (25, 18), (8, 21)], 1)).max() > 0


@pytest.mark.parametrize("n", [100, 750, 1000, 2001, 3000, 4095, 4097, 5000, 6001, 8191])
def test_arbitrary_length_traces(ctx, n):
    """Trace lengths that are not a power of two (real scans; realfft accepts any N): chirp-z kernels;
    n > 4096 takes the 16384-point convolution as two 8192-point sub-spectra.
    Forward, fused chain and inverse against the oracle (scipy's pocketfft handles any N)."""
    w, h = 3, 5
    cube = synthetic_cube(w, h, n, seed=n)
    t, m_pre, band, m_post = default_multipliers(n)
    slots = _chain_oracle(cube, t)
    cfg = orc.ConfigContainer()
    win = orc.fft_window_multiplier(t, cfg.fft_window_type, cfg.fft_window)
    ref4 = orc.fft(slot0(cube, t), cfg)
    ctx.plan_trace(n, m_pre=win)
    got = ctx.trace_forward(cube)
    assert np.array_equal(got["windowed"], ref4.data)
    assert rel_err(got["fft"], ref4.fft) <= TOL_TRACE
    assert rel_err(got["amp"], ref4.amplitudes) <= TOL_TRACE
    tol_rad = max(TOL_TRACE * float(np.abs(ref4.phases).max()), 2e-3)
    check_unwrapped_phase(got["phase"], ref4.phases, ref4.fft, tol_rad)
    ctx.plan_trace(n, m_pre, band, m_post)
    out, img = ctx.trace_fused(cube)
    assert rel_err(out, slots[7].data) <= TOL_TRACE
    assert rel_err(img, slots[7].img) <= TOL_TRACE
    back, _ = ctx.trace_inverse(slots[5].fft)
    assert rel_err(back, slots[6].data) <= TOL_TRACE
    back2, img2 = ctx.trace_inverse(slots[4].fft, use_band=True, use_post=True, want_img=True)
    assert rel_err(back2, slots[7].data) <= TOL_TRACE and rel_err(img2, slots[7].img) <= TOL_TRACE
    # dead pixel stays exactly zero here too
    cube[1, 2] = 0.0
    out, img = ctx.trace_fused(cube)
    assert not out[1, 2].any() and img[1, 2] == 0.0


@pytest.mark.parametrize("n,m,shift", [(1024, 1024, 0), (1024, 900, 37), (1024, 1200, -50), (1000, 1024, 12)])
def test_reference_pulse_path(ctx, n, m, shift):
    """ConfigCommand::OpenRef (src/data_thread.rs:372-588): alignment / resize + window + forward transform of
    one reference pulse (BASELINE config 2 normalises sample spectra by it)."""
    dt = F32(0.05)
    st = time_axis(n)
    rt = (F32(1000.0) + dt * F32(shift) + dt * np.arange(m, dtype=F32)).astype(F32)
    tt = np.arange(m) * 0.05 - 12.0
    ref = (np.exp(-(tt / 0.3) ** 2) * np.cos(2 * np.pi * tt) + 0.001 * np.random.default_rng(1).standard_normal(m)).astype(F32)
    sig, amp, ph = ctx.reference_pulse(st, rt, ref)
    osig, oamp, oph = orc.reference_pulse(st, rt, ref)
    assert rel_err(sig, osig) <= 1e-6
    assert rel_err(amp, oamp) <= TOL_TRACE
    spec = np.fft.rfft(osig.astype(np.float64)).astype(np.complex64)
    check_unwrapped_phase(ph[None, :], oph[None, :], spec[None, :], max(TOL_TRACE * float(np.abs(oph).max()), 2e-3))


def test_voxel_opacity_and_threshold(ctx):
    """instance_from_data up to the effective threshold (src/gui/threed_plot.rs:165-219)."""
    cube = synthetic_cube(12, 10, 512, seed=23, noise=0.02)
    cube[3, 4] *= F32(0.01)                       # a faint trace: below the opacity threshold -> zeroed
    ref, rthr = orc.voxel_opacity(cube, 0.1, 2.0, 3.0, 9, max_instances=5000)
    got, thr = ctx.voxel_opacity(cube, 0.1, 2.0, 3.0, 9, max_instances=5000)
    assert rel_err(got, ref) <= 2e-5
    assert not got[3, 4].any()
    # the threshold is an exact order statistic of the GPU's own opacities
    flat = np.sort(got.reshape(-1))[::-1]
    assert thr == flat[4999]
    assert abs(thr - float(rthr)) <= 2e-5
    _, thr0 = ctx.voxel_opacity(cube, 0.1, 2.0, 3.0, 9, max_instances=10 ** 9)
    assert thr0 == 0.0


def test_row_slab_sharding_is_bit_identical(ctx):
    """Row slabs over x (sharding.slab_bounds) processed independently give the same bits as one pass:
    traces are independent and, with an even height, never change partner inside a packed pair."""
    n, w, h = 2048, 10, 6
    cube = synthetic_cube(w, h, n, seed=31)
    t, m_pre, band, m_post = default_multipliers(n)
    ctx.plan_trace(n, m_pre, band, m_post)
    whole, img = ctx.trace_fused(cube)
    sh = pkg().sharding
    for world in (2, 4):
        parts = [ctx.trace_fused(cube[slice(*sh.slab_bounds(w, world, r))]) for r in range(world)]
        assert np.array_equal(np.concatenate([p[0] for p in parts]), whole)
        assert np.array_equal(np.concatenate([p[1] for p in parts]), img)


def test_error_paths(ctx):
    m = pkg()
    L = m.lib
    assert L.thz_plan_trace(ctx.handle, 9000, None, None, None) == -1           # > 8192
    assert b"power of two" in L.thz_last_error(ctx.handle)
    assert L.thz_plan_trace(ctx.handle, 1, None, None, None) == -1
    ctx.plan_trace(1024)
    assert L.thz_trace_fused_dev(ctx.handle, None, None, None, 4) == -1          # null cube
    assert L.thz_trace_fused_dev(ctx.handle, None, None, None, 0) == 0           # nothing to do
    c2 = m.Context(0)
    assert L.thz_trace_fused_dev(c2.handle, None, None, None, 4) == -4           # no plan: THZ_ESTATE
    c2.close()
