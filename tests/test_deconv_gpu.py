"""GPU parity tests of the deconvolution path against the CPU oracle (tolerance from the north
star: max|a-b|/max|b| <= 1e-3 on deconvolved maps after N iterations)."""
import ctypes

import numpy as np
import pytest

from helpers import F32, TOL_MAP, TOL_TRACE, orc, pkg, rel_err, synthetic_cube, time_axis

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    c = pkg().Context(0)
    yield c
    c.close()


@pytest.fixture(scope="module")
def psfs(psf_npz_path):
    return pkg().host.PSF.load(psf_npz_path), orc.load_psf(psf_npz_path)


def _gauss(k, c, w):
    x = np.arange(k, dtype=np.float64) - k // 2
    g = np.exp(-2 * (x - c) ** 2 / w ** 2)
    return (g / g.max()).astype(F32)


@pytest.mark.parametrize("shape,kx,ky", [((100, 77), 7, 7), ((64, 64), 17, 15), ((130, 141), 47, 57), ((33, 200), 3, 9),
                                         ((70, 65), 1, 5)])
def test_conv2d_both_orientations(ctx, shape, kx, ky):
    """Tile kernel vs both branches of `convolve2d`: direct (a correlation) and FFT (a convolution);
    off-centre factors make the two differ."""
    rng = np.random.default_rng(kx * 100 + ky)
    img = rng.uniform(0.1, 1.0, shape).astype(F32)
    px, py = _gauss(kx, 0.6, max(kx / 4, 0.8)), _gauss(ky, -0.9, max(ky / 4, 0.8))
    psf = np.outer(px, py).astype(F32)
    corr = orc.direct_convolve2d(img.astype(np.float64), psf.astype(np.float64))
    conv = orc.direct_convolve2d(img.astype(np.float64), psf[::-1, ::-1].astype(np.float64))
    assert rel_err(ctx.conv2d(img, px, py, direct=True), corr) < 5e-6
    assert rel_err(ctx.conv2d(img, px, py, direct=False), conv) < 5e-6
    # dense (non-separable) kernel on the same PSF plus a perturbation that breaks separability
    dense = (psf + 0.05 * rng.uniform(size=psf.shape)).astype(F32)
    corr_d = orc.direct_convolve2d(img.astype(np.float64), dense.astype(np.float64))
    conv_d = orc.direct_convolve2d(img.astype(np.float64), dense[::-1, ::-1].astype(np.float64))
    assert rel_err(ctx.conv2d(img, dense=dense, direct=True), corr_d) < 2e-5  # f32 accumulation over kx*ky taps
    assert rel_err(ctx.conv2d(img, dense=dense, direct=False), conv_d) < 2e-5
    # and the oracle's FFT branch really is the convolution (pins the orientation convention)
    if kx * ky > 1:
        assert rel_err(orc.fft_convolve2d(img, psf), conv.astype(F32)) < 1e-4


@pytest.mark.parametrize("shape,kx,ky", [((700, 300), 31, 29),     # several segments per strip, 64-column strips
                                         ((90, 1100), 9, 65),      # many strips, one chunk of rows
                                         ((513, 129), 5, 3),       # ragged last strip / last segment
                                         ((300, 260), 95, 65),     # buffers fit once: 128-column strips
                                         ((200, 200), 127, 121)])  # too large for the strip buffers: tile kernel
def test_conv2d_strip_geometry(ctx, shape, kx, ky):
    """The streaming strip kernel (k_rl_stream) over image / PSF shapes that exercise every branch of its
    plan: strip width 64 or 128, warm-up rows, chunks that lie wholly below the image, the fallback."""
    from scipy.signal import correlate2d
    rng = np.random.default_rng(kx * 1000 + ky)
    img = rng.uniform(0.1, 1.0, shape).astype(F32)
    px, py = _gauss(kx, 0.6, max(kx / 4, 0.8)), _gauss(ky, -0.9, max(ky / 4, 0.8))
    # separable reference in f64: rows then columns, zero boundary, 'same' (deconvolution.rs:476-530)
    tmp = correlate2d(img.astype(np.float64), px.astype(np.float64)[:, None], mode="same")
    corr = correlate2d(tmp, py.astype(np.float64)[None, :], mode="same")
    tmp = correlate2d(img.astype(np.float64), px[::-1].astype(np.float64)[:, None], mode="same")
    conv = correlate2d(tmp, py[::-1].astype(np.float64)[None, :], mode="same")
    assert rel_err(ctx.conv2d(img, px, py, direct=True), corr) < 5e-6
    assert rel_err(ctx.conv2d(img, px, py, direct=False), conv) < 5e-6


@pytest.mark.parametrize("band_idx,n_iter", [(0, 30), (1, 60), (2, 127), (4, 13)])
def test_richardson_lucy_matches_oracle(ctx, psfs, band_idx, n_iter):
    """richardson_lucy + clamp + gain against the oracle, which takes the same convolve2d branch the
    reference would (FFT for the two large PSFs, direct below 256 taps)."""
    psf, opsf = psfs
    shape = (96, 80)
    t = time_axis(1024)
    bands, _ = pkg().host.Deconvolution(n_filters=8).plan(t, (2048, 2048), 0.5, 0.5, psf)
    obands, _ = orc.Deconvolution(n_filters=8).plan(t, (2048, 2048, 1024), 0.5, 0.5, opsf)
    b, ob = bands[band_idx], obands[band_idx]
    yy, xx = np.meshgrid(np.arange(shape[1]), np.arange(shape[0]))
    img = (1.0 + 0.5 * ((xx // 8 + yy // 8) % 2) + 0.2 * np.sin(xx / 5.0)).astype(F32)
    ref = np.maximum(orc.richardson_lucy(img, ob.psf, n_iter), 0).astype(F32)
    u, gain = ctx.richardson_lucy(img, n_iter, b.psf_x_np(), b.psf_y_np(), direct=bool(b.direct), want_gain=True)
    assert rel_err(u, ref) <= TOL_MAP
    assert rel_err(gain, np.sqrt(ref / img)) <= TOL_MAP
    # the dense kernel runs the same iteration
    if b.kx * b.ky <= 31 * 29 and n_iter <= 60:
        ud = ctx.richardson_lucy(img, n_iter, dense=ob.psf, direct=bool(b.direct))
        assert rel_err(ud, ref) <= TOL_MAP


@pytest.mark.parametrize("n", [256, 1024, 2048, 8192])
def test_band_energies_match_oracle(ctx, psfs, n):
    psf, opsf = psfs
    w, h = 5, 7
    cube = synthetic_cube(w, h, n, seed=n, noise=0.02)
    t = time_axis(n)
    bands, _ = pkg().host.Deconvolution(n_filters=6).plan(t, (64, 64), 0.5, 0.5, psf)
    obands, _ = orc.Deconvolution(n_filters=6).plan(t, (64, 64, n), 0.5, 0.5, opsf)
    P = w * h
    d_cube = ctx.to_device(cube)
    d_e = ctx.alloc(len(bands) * P * 4)
    ctx.deconv_energies_dev(d_cube.ptr, P, n, bands, d_e.ptr)
    e = d_e.download((len(bands), w, h))
    for i, ob in enumerate(obands):
        filt = orc.filter_scan(cube, ob.fir)
        ref = np.sum(filt * filt, axis=2, dtype=F32)
        assert rel_err(e[i], ref) <= 1e-5, i


def test_full_deconvolution_matches_oracle(ctx, psfs):
    """Deconvolution::filter end to end (FIR bank -> band energies -> RL -> gains -> band sum ->
    intensity) on a 40x36x512 cube, 6 bands, 40 iterations max."""
    psf, opsf = psfs
    w, h, n = 40, 36, 512
    cube = synthetic_cube(w, h, n, seed=9, noise=0.02)
    yy, xx = np.meshgrid(np.arange(h), np.arange(w))
    cube *= (1.0 + 0.5 * ((xx // 6 + yy // 6) % 2)).astype(F32)[:, :, None]
    t = time_axis(n)
    dec = pkg().host.Deconvolution(n_filters=6, n_iterations=40)
    bands, why = dec.plan(t, (w, h), 1.0, 1.0, psf)
    assert why is None
    f = orc.frequency_axis(t)
    s = orc.ScannedImageFilterData(time=t, data=cube, frequency=f, img=orc.intensity_image(cube), dx=1.0, dy=1.0,
                                   width=w, height=h)
    ref = orc.Deconvolution(n_filters=6, n_iterations=40).filter(s, opsf)
    seen = []
    out, img, rc = ctx.deconvolution(cube, bands, progress=lambda frac, _u: seen.append(frac))
    assert rc == 0
    assert rel_err(out, ref.data) <= TOL_MAP
    assert rel_err(img, ref.img) <= TOL_MAP
    assert seen and seen[0] == 0.0 and seen[-1] == 1.0 and all(b >= a for a, b in zip(seen, seen[1:]))


def test_dead_pixel_gives_nan_only_there(ctx, psfs):
    """Quirk 10: gain = sqrt(u/d) is NaN where a band's energy is 0; the reference's output is NaN for
    that pixel only.  Traces are processed in pairs on the GPU: the neighbour must stay finite."""
    psf, _ = psfs
    w, h, n = 32, 32, 256
    cube = synthetic_cube(w, h, n, seed=4, noise=0.02)
    cube[10, 11, :] = 0.0
    bands, _ = pkg().host.Deconvolution(n_filters=4, n_iterations=5).plan(time_axis(n), (w, h), 1.0, 1.0, psf)
    out, img, rc = ctx.deconvolution(cube, bands)
    assert rc == 0
    bad = ~np.isfinite(out).all(axis=2)
    assert bad[10, 11] and bad.sum() == 1
    assert np.isnan(img[10, 11]) and np.isfinite(np.delete(img.ravel(), 10 * h + 11)).all()


def test_abort_flag_stops_the_filter(ctx, psfs):
    psf, _ = psfs
    w, h, n = 32, 32, 256
    cube = synthetic_cube(w, h, n, seed=5)
    bands, _ = pkg().host.Deconvolution(n_filters=4, n_iterations=500).plan(time_axis(n), (w, h), 1.0, 1.0, psf)
    flag = ctypes.c_uint8(1)   # layout of Rust AtomicBool
    out, img, rc = ctx.deconvolution(cube, bands, abort_flag=flag)
    assert rc == 1   # THZ_ABORTED: the shim keeps the previous slot, like the cancellable loops


def test_chain_host_equals_staged_calls(ctx, psfs):
    """thz_chain_host (cube resident on the device between the fused trace pass and the deconvolution,
    chunked copies) == thz_trace_fused_host followed by thz_deconvolution_host."""
    from helpers import default_multipliers
    psf, _ = psfs
    w, h, n = 40, 33, 512
    cube = synthetic_cube(w, h, n, seed=12, noise=0.02)
    t, m_pre, band, m_post = default_multipliers(n)
    ctx.plan_trace(n, m_pre, band, m_post)
    bands, why = pkg().host.Deconvolution(n_filters=5, n_iterations=20).plan(t, (w, h), 1.0, 1.0, psf)
    assert why is None
    fused, fimg = ctx.trace_fused(cube)
    ref_out, ref_img, _ = ctx.deconvolution(fused, bands)
    out, img, rc = ctx.chain(cube, bands)
    assert rc == 0
    assert rel_err(out, ref_out) <= 1e-6 and rel_err(img, ref_img) <= 1e-6
    out0, img0, _ = ctx.chain(cube, None)
    assert np.array_equal(out0, fused) and np.array_equal(img0, fimg)


def test_rl_properties(ctx, psfs):
    """Size-independent properties of the iteration: non-negative input stays non-negative, and the
    iteration is positively homogeneous (RL(a d) = a RL(d)), checked on a 300 x 280 image (no oracle)."""
    psf, _ = psfs
    bands, _ = pkg().host.Deconvolution(n_filters=8).plan(time_axis(1024), (2048, 2048), 0.5, 0.5, psf)
    rng = np.random.default_rng(5)
    img = (rng.uniform(0.2, 1.0, (300, 280)) + ((np.indices((300, 280)).sum(axis=0) // 10) % 2)).astype(F32)
    for b in (bands[0], bands[3]):
        u = ctx.richardson_lucy(img, 25, b.psf_x_np(), b.psf_y_np(), direct=bool(b.direct))
        assert np.isfinite(u).all() and (u >= 0).all()
        u4 = ctx.richardson_lucy((F32(4.0) * img).astype(F32), 25, b.psf_x_np(), b.psf_y_np(), direct=bool(b.direct))
        assert rel_err(u4, F32(4.0) * u) <= 1e-5


def test_chain_host_abort(ctx, psfs):
    from helpers import default_multipliers
    psf, _ = psfs
    w, h, n = 32, 32, 256
    cube = synthetic_cube(w, h, n, seed=3)
    t, m_pre, band, m_post = default_multipliers(n)
    ctx.plan_trace(n, m_pre, band, m_post)
    bands, _ = pkg().host.Deconvolution(n_filters=4, n_iterations=500).plan(t, (w, h), 1.0, 1.0, psf)
    flag = ctypes.c_uint8(1)
    out = np.empty_like(cube)
    img = np.empty((w, h), F32)
    rc = pkg().lib.thz_chain_host(ctx.handle, cube.ctypes.data, w, h, n, bands, len(bands), out.ctypes.data,
                                  img.ctypes.data, ctypes.addressof(flag), None, None)
    assert rc == 1


@pytest.mark.parametrize("n,w,h", [(512, 5, 7), (1024, 3, 5), (4096, 3, 3), (256, 4, 4), (8192, 3, 1)])
def test_gain_application_matches_oracle(ctx, psfs, n, w, h):
    """Pass C alone: out[p] = sum_b g_b[p] (fir_b * x[p]) in 'same' mode (deconvolution.rs:873-905) with random
    per-pixel gains, odd pixel count (ragged last pair), in place.  n >= 512 runs the circular form with the
    wrap-around edge corrections, n = 256 the zero-padded form: both must reproduce the linear convolution.
    n = 8192 is the largest trace length: no monolithic 16384-point plan exists, only the N-point forms."""
    psf, opsf = psfs
    cube = synthetic_cube(w, h, n, seed=100 + n, noise=0.05)
    # energy right up to both ends of the trace, so that the wrapped pieces are far from negligible
    rng = np.random.default_rng(n)
    cube[:, :, :40] += rng.standard_normal((w, h, 40)).astype(F32)
    cube[:, :, -40:] += rng.standard_normal((w, h, 40)).astype(F32)
    t = time_axis(n)
    bands, _ = pkg().host.Deconvolution(n_filters=5).plan(t, (64, 64), 0.5, 0.5, psf)
    obands, _ = orc.Deconvolution(n_filters=5).plan(t, (64, 64, n), 0.5, 0.5, opsf)
    P = w * h
    gains = (0.25 + 2.0 * rng.random((len(bands), P))).astype(F32)
    ref = np.zeros((w, h, n), dtype=np.float64)
    for i, ob in enumerate(obands):
        ref += gains[i].reshape(w, h, 1).astype(np.float64) * orc.filter_scan(cube, ob.fir)
    d_cube = ctx.to_device(cube)
    d_g = ctx.to_device(gains)
    d_img = ctx.alloc(P * 4)
    ctx.deconv_apply_dev(d_cube.ptr, d_g.ptr, P, n, bands, d_cube.ptr, d_img.ptr)
    out = d_cube.download((w, h, n))
    img = d_img.download((w, h))
    assert rel_err(out, ref) <= 2e-6
    # the edges are where the two forms differ: check them on their own scale
    for sl in (slice(0, 249), slice(n - 249, n)):
        assert rel_err(out[:, :, sl], ref[:, :, sl]) <= 5e-6
    assert rel_err(img, np.sum(ref * ref, axis=2)) <= 1e-5


def test_gain_application_forms_agree(psfs, monkeypatch):
    """The circular form of pass C (default), the zero-padded split form (THZ_APPLY_FORM=split), the
    bulk-copy-staged variants (THZ_FIR_STAGING=on) and the transform form of the pass-A edges (THZ_EDGE_MMA=off)
    are routes to the same result."""
    psf, _ = psfs
    w, h, n = 6, 5, 2048
    cube = synthetic_cube(w, h, n, seed=77, noise=0.05)
    rng = np.random.default_rng(5)
    cube[:, :, :60] += rng.standard_normal((w, h, 60)).astype(F32)
    cube[:, :, -60:] += rng.standard_normal((w, h, 60)).astype(F32)
    bands, _ = pkg().host.Deconvolution(n_filters=8).plan(time_axis(n), (64, 64), 0.5, 0.5, psf)
    P = w * h
    gains = (0.25 + 2.0 * rng.random((len(bands), P))).astype(F32)
    outs, energies = [], []
    for env in ({}, {"THZ_APPLY_FORM": "split"}, {"THZ_FIR_STAGING": "on"}, {"THZ_EDGE_MMA": "off"}):
        for k in ("THZ_APPLY_FORM", "THZ_FIR_STAGING", "THZ_EDGE_MMA"):
            monkeypatch.delenv(k, raising=False)
        for k, v in env.items():
            monkeypatch.setenv(k, v)
        c = pkg().Context(0)
        try:
            d_cube, d_g, d_out = c.to_device(cube), c.to_device(gains), c.alloc(cube.nbytes)
            d_img, d_e = c.alloc(P * 4), c.alloc(len(bands) * P * 4)
            c.deconv_energies_dev(d_cube.ptr, P, n, bands, d_e.ptr)
            c.deconv_apply_dev(d_cube.ptr, d_g.ptr, P, n, bands, d_out.ptr, d_img.ptr)
            outs.append((d_out.download((w, h, n)), d_img.download((w, h))))
            energies.append(d_e.download((len(bands), P)))
        finally:
            c.close()
    for out, img in outs[1:]:
        assert rel_err(out, outs[0][0]) <= 2e-6
        assert rel_err(img, outs[0][1]) <= 1e-5
    for e in energies[1:]:
        assert rel_err(e, energies[0]) <= 1e-5


def test_chain_host_many_chunks_on_three_streams(psfs, monkeypatch):
    """thz_chain_host pipelines its chunks over three streams; the wrap-around corrections of pass C live in a
    per-stream workspace (a shared one would let chunk i+1 overwrite what chunk i still reads).  A 16 KiB chunk
    size forces 30+ chunks per pass on a small cube; the result must equal the unchunked staged calls, and the
    first / last 249 samples (where the corrections act) are compared on their own scale."""
    from helpers import default_multipliers
    psf, _ = psfs
    w, h, n = 24, 20, 1024
    cube = synthetic_cube(w, h, n, seed=21, noise=0.05)
    rng = np.random.default_rng(8)
    cube[:, :, :60] += rng.standard_normal((w, h, 60)).astype(F32)
    cube[:, :, -60:] += rng.standard_normal((w, h, 60)).astype(F32)
    t = time_axis(n)
    bands, why = pkg().host.Deconvolution(n_filters=5, n_iterations=8).plan(t, (w, h), 1.0, 1.0, psf)
    assert why is None
    results = []
    for chunk in (None, str(16 * 1024)):
        if chunk is None:
            monkeypatch.delenv("THZ_CHAIN_CHUNK_BYTES", raising=False)
        else:
            monkeypatch.setenv("THZ_CHAIN_CHUNK_BYTES", chunk)
        c = pkg().Context(0)
        try:
            c.plan_trace(n, None, None, None)     # no gates: the trace ends keep their energy
            for _ in range(3):                    # the race, if any, is timing dependent
                out, img, rc = c.chain(cube, bands)
                assert rc == 0
                results.append((out, img))
        finally:
            c.close()
    ref_out, ref_img = results[0]
    for out, img in results[1:]:
        assert rel_err(out, ref_out) <= 1e-6
        for sl in (slice(0, 249), slice(n - 249, n)):
            assert rel_err(out[:, :, sl], ref_out[:, :, sl]) <= 2e-6
        assert rel_err(img, ref_img) <= 1e-5


@pytest.mark.parametrize("band_idx,shape", [(0, (700, 300)), (1, (420, 300))])
def test_richardson_lucy_at_benchmark_iteration_counts(ctx, psfs, band_idx, shape):
    """The iteration counts BASELINE config 5 really runs -- 423 for the 47x57 PSF of band 0, 251 for the 31x29 PSF
    of band 1 -- on images tall enough for several row segments and strips of the streaming kernel, against the
    oracle (FFT branch of `convolve2d`, as the reference takes for these PSFs).  North-star tolerance: 1e-3 on
    deconvolved maps after N iterations; the measured error is printed (-s) and recorded in DESIGN.md."""
    psf, opsf = psfs
    t = time_axis(1024)
    bands, _ = pkg().host.Deconvolution(n_filters=8, n_iterations=500).plan(t, (2048, 2048), 0.5, 0.5, psf)
    obands, _ = orc.Deconvolution(n_filters=8, n_iterations=500).plan(t, (2048, 2048, 1024), 0.5, 0.5, opsf)
    b, ob = bands[band_idx], obands[band_idx]
    assert b.n_iter == (423, 251)[band_idx] and (b.kx, b.ky) == ((47, 57), (31, 29))[band_idx]
    yy, xx = np.meshgrid(np.arange(shape[1]), np.arange(shape[0]))
    rng = np.random.default_rng(band_idx)
    img = (1.0 + 0.5 * ((xx // 16 + yy // 16) % 2) + 0.2 * np.sin(xx / 9.0) + 0.02 * rng.random(shape)).astype(F32)
    ref = np.maximum(orc.richardson_lucy(img, ob.psf, b.n_iter), 0).astype(F32)
    u, gain = ctx.richardson_lucy(img, b.n_iter, b.psf_x_np(), b.psf_y_np(), direct=bool(b.direct), want_gain=True)
    err_u, err_g = rel_err(u, ref), rel_err(gain, np.sqrt(ref / img))
    print(f"RL band {band_idx}: {b.n_iter} iterations on {shape}, rel err u {err_u:.2e}, gain {err_g:.2e}")
    assert err_u <= TOL_MAP and err_g <= TOL_MAP


def test_config3_full_chain_and_deconvolution_matches_oracle(ctx, psfs):
    """BASELINE config 3 at its stand-in shape: 256 x 256 pixels x 2048 samples, dx = dy = 0.5 mm,
    default filter chain, then `Deconvolution{n_filters: 8, n_iterations: 500}` with the shipped psf.npz ->
    423 / 251 / 127 / 46 / 13 / 4 / 3 / 1 iterations.  GPU (fused chain + thz_deconvolution on the device-resident
    cube, through thz_chain_host) against the oracle's stage-by-stage chain and `Deconvolution::filter`
    (deconvolution.rs:766-1041).  Takes a few minutes of host time for the oracle."""
    from helpers import default_multipliers, slot0
    psf, opsf = psfs
    w, h, n = 256, 256, 2048
    cube = synthetic_cube(w, h, n, seed=33, noise=0.01)
    yy, xx = np.meshgrid(np.arange(h), np.arange(w))
    cube *= (1.0 + 0.5 * ((xx // 12 + yy // 12) % 2)).astype(F32)[:, :, None]     # bars: structure for RL
    t, m_pre, band, m_post = default_multipliers(n)
    slots = orc.run_default_chain(slot0(cube, t, dx=0.5, dy=0.5))
    dec = pkg().host.Deconvolution(n_filters=8, n_iterations=500)
    bands, why = dec.plan(t, (w, h), 0.5, 0.5, psf)
    assert why is None and [b.n_iter for b in bands] == [423, 251, 127, 46, 13, 4, 3, 1]
    import os
    ref = orc.Deconvolution(n_filters=8, n_iterations=500).filter(slots[7], opsf, workers=min(16, os.cpu_count() or 1))
    ctx.plan_trace(n, m_pre, band, m_post)
    fused, fimg = ctx.trace_fused(cube)
    assert rel_err(fused, slots[7].data) <= TOL_TRACE
    out, img, rc = ctx.chain(cube, bands)
    assert rc == 0
    e_out, e_img = rel_err(out, ref.data), rel_err(img, ref.img)
    print(f"config 3: deconvolved cube rel err {e_out:.2e}, intensity map {e_img:.2e}")
    assert e_out <= TOL_MAP and e_img <= TOL_MAP


def test_batched_rl_equals_band_after_band(psfs, monkeypatch):
    """Iteration i of every band that still iterates goes into ONE launch per filtering (k_rl_multi, 2 max(n_iter)
    launches instead of 2 sum(n_iter)); THZ_RL_BATCH=off iterates band after band with the single-band kernel.
    Same arithmetic per pixel: identical bits."""
    psf, _ = psfs
    w, h, n = 150, 130, 256
    cube = synthetic_cube(w, h, n, seed=41, noise=0.02)
    yy, xx = np.meshgrid(np.arange(h), np.arange(w))
    cube *= (1.0 + 0.5 * ((xx // 10 + yy // 10) % 2)).astype(F32)[:, :, None]
    bands, why = pkg().host.Deconvolution(n_filters=8, n_iterations=60).plan(time_axis(n), (w, h), 0.5, 0.5, psf)
    assert why is None and len({b.n_iter for b in bands}) > 3
    outs = []
    for mode in ("on", "off"):
        monkeypatch.setenv("THZ_RL_BATCH", mode)
        c = pkg().Context(0)
        try:
            l0 = c.launches
            out, img, rc = c.deconvolution(cube, bands)
            assert rc == 0
            outs.append((out, img, c.launches - l0))
        finally:
            c.close()
    assert np.array_equal(outs[0][0], outs[1][0]) and np.array_equal(outs[0][1], outs[1][1])
    assert outs[0][2] < outs[1][2]          # fewer launches
