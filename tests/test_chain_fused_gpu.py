"""The fused trace + band-energy kernel (k_chain_energy_fused, thz_chain_energies_dev / thz_chain_dev) against the
two-pass route (thz_trace_fused_dev followed by thz_deconv_energies_dev) and against the oracle.

The fused kernel stores the traces the trace pass stores (same arithmetic; the compiler contracts the band multiply
into different FMAs in the two kernels, so they agree to an ulp or two, not bit for bit) and forms the band energies
from what it holds on chip; with the default gate (non-unit only in the first / last four samples) or no gate the even
bins of the 2N-point spectrum come from the filtered spectrum itself, which differs from transforming the stored
trace by f32 rounding only."""
import numpy as np
import pytest

from helpers import F32, TOL_MAP, orc, pkg, rel_err, synthetic_cube, time_axis

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def psfs(psf_npz_path):
    return pkg().host.PSF.load(psf_npz_path), orc.load_psf(psf_npz_path)


def _gates(n, kind):
    """(m_pre, band, m_post) of the default chain with the gate after the inverse transform replaced."""
    t = time_axis(n)
    m_pre, band, m_post = pkg().host.chain_multipliers(t)
    if kind == "default":            # 0.1 ps edges at 0.05 ps steps: non-unit in the first 3 / last 4 samples
        assert m_post is not None
        idx = np.nonzero(m_post != 1.0)[0]
        assert idx.size and np.all((idx < 4) | (idx >= n - 4))
    elif kind == "none":
        m_post = None
    elif kind == "ones":
        m_post = np.ones(n, F32)
    elif kind == "wide":             # a real time gate: zero outside [20 %, 70 %] with a smooth edge
        g = np.zeros(n)
        lo, hi = int(0.2 * n), int(0.7 * n)
        g[lo:hi] = 1.0
        ramp = 0.5 - 0.5 * np.cos(np.pi * np.arange(32) / 32)
        g[lo:lo + 32] = ramp
        g[hi - 32:hi] = ramp[::-1]
        m_post = g.astype(F32)
    elif kind == "ends8":            # non-unit in exactly the eight samples the sparse form covers
        m_post = np.ones(n, F32)
        m_post[:4] = [0.0, 0.3, 0.8, 0.95]
        m_post[-4:] = [0.9, 0.6, 0.2, 0.0]
    return m_pre, band, m_post


def _run(env, monkeypatch, n, cube, bands, kind, fused):
    for k in ("THZ_CHAIN_FUSE", "THZ_CHAIN_EVEN", "THZ_EDGE_MMA"):
        monkeypatch.delenv(k, raising=False)
    for k, v in env.items():
        monkeypatch.setenv(k, v)
    w, h, _ = cube.shape
    P = w * h
    c = pkg().Context(0)
    try:
        c.plan_trace(n, *_gates(n, kind))
        d_in, d_out = c.to_device(cube), c.alloc(cube.nbytes)
        d_img, d_e = c.alloc(P * 4), c.alloc(len(bands) * P * 4)
        if fused:
            c.chain_energies_dev(d_in.ptr, d_out.ptr, d_img.ptr, P, n, bands, d_e.ptr)
        else:
            c.trace_fused_dev(d_in.ptr, d_out.ptr, d_img.ptr, P)
            c.deconv_energies_dev(d_out.ptr, P, n, bands, d_e.ptr)
        c.sync()
        return d_out.download((w, h, n)), d_img.download((w, h)), d_e.download((len(bands), P))
    finally:
        c.close()


@pytest.mark.parametrize("n", [512, 1024, 2048, 4096, 8192])
@pytest.mark.parametrize("kind", ["default", "none", "ones", "ends8", "wide"])
def test_fused_equals_two_passes(psfs, monkeypatch, n, kind):
    """Same stored traces and intensities up to FMA contraction (2e-6 of the peak; bit for bit with THZ_CHAIN_FUSE=off);
    band energies within f32 rounding of the two-pass route (1e-5 of the largest energy of the band plus 1e-6 of the
    trace's total energy)."""
    psf, _ = psfs
    w, h = 7, 9                      # odd trace count: the last pair is half empty
    cube = synthetic_cube(w, h, n, seed=n + len(kind), noise=0.05)
    rng = np.random.default_rng(3)
    cube[:, :, :40] += 0.3 * rng.standard_normal((w, h, 40)).astype(F32)      # energy where the gate acts
    cube[:, :, -40:] += 0.3 * rng.standard_normal((w, h, 40)).astype(F32)
    cube[2, 3] = 0.0                 # dead pixels: exact zeros out, zero energies
    cube[6, 8] = 0.0                 # the unpaired last trace
    bands, _ = pkg().host.Deconvolution(n_filters=8).plan(time_axis(n), (64, 64), 0.5, 0.5, psf)
    ref = _run({"THZ_EDGE_MMA": "off"}, monkeypatch, n, cube, bands, kind, fused=False)
    got = _run({"THZ_EDGE_MMA": "off"}, monkeypatch, n, cube, bands, kind, fused=True)
    alt = _run({"THZ_EDGE_MMA": "off", "THZ_CHAIN_EVEN": "transform"}, monkeypatch, n, cube, bands, kind, fused=True)
    off = _run({"THZ_EDGE_MMA": "off", "THZ_CHAIN_FUSE": "off"}, monkeypatch, n, cube, bands, kind, fused=True)
    for r in (got, alt, off):
        assert rel_err(r[0], ref[0]) <= 2e-6 and rel_err(r[1], ref[1]) <= 2e-6
        assert np.all(r[0][2, 3] == 0.0) and np.all(r[0][6, 8] == 0.0) and r[1][2, 3] == 0.0 and r[1][6, 8] == 0.0
        assert np.all(r[2][:, 2 * h + 3] == 0.0) and np.all(r[2][:, 6 * h + 8] == 0.0)
    assert np.array_equal(off[0], ref[0]) and np.array_equal(off[1], ref[1]) and np.array_equal(off[2], ref[2])
    # f32 rounding scales with the energy of the whole trace, not with that of a weak band
    total = ref[2].sum(axis=0)
    for r in (got, alt):
        for b in range(len(bands)):
            tol = 1e-5 + 1e-6 * float(total.max()) / float(ref[2][b].max())
            assert rel_err(r[2][b], ref[2][b]) <= tol, (kind, n, b)
        assert np.all(np.abs(r[2] - ref[2]) <= 2e-4 * ref[2] + 1e-6 * total[None, :])
    # the general-gate form transforms the stored trace exactly as the two-pass route does
    assert rel_err(alt[2], ref[2]) <= 4e-6


def test_fused_energies_match_oracle(psfs, monkeypatch):
    """Band energies of the fused kernel against the oracle's `filter_scan` on the oracle's filtered traces."""
    psf, opsf = psfs
    n, w, h = 2048, 5, 6
    cube = synthetic_cube(w, h, n, seed=11, noise=0.02)
    t = time_axis(n)
    bands, _ = pkg().host.Deconvolution(n_filters=8).plan(t, (64, 64), 0.5, 0.5, psf)
    obands, _ = orc.Deconvolution(n_filters=8).plan(t, (64, 64, n), 0.5, 0.5, opsf)
    out, img, e = _run({"THZ_EDGE_MMA": "off"}, monkeypatch, n, cube, bands, "default", fused=True)
    _, _, e_mma = _run({}, monkeypatch, n, cube, bands, "default", fused=True)
    m_pre, band, m_post = orc.default_chain_multipliers(t)
    spec = np.fft.rfft(cube.astype(np.float64) * m_pre.astype(np.float64), axis=2) * band.astype(np.float64)
    ref = np.fft.irfft(spec, n, axis=2) * m_post.astype(np.float64)
    assert rel_err(out, ref) <= 1e-4
    for b, ob in enumerate(obands):
        f = orc.filter_scan(ref, ob.fir)
        eb = np.sum(f.astype(np.float64) ** 2, axis=2).reshape(-1)
        assert rel_err(e[b], eb) <= 2e-5, b
        # tensor-core edges (TF32 operands): the weak bands at both ends of the bank -- with the default
        # gate the lowest ones consist of the gate's own edge transient, half of which lies in the cut-off samples,
        # the highest one holds 1e-7 of the trace energy -- show the TF32 rounding of the edge energies at the
        # 1e-4 level; far inside the 1e-3 tolerance of the deconvolved maps
        assert rel_err(e_mma[b], eb) <= (3e-4 if (b < 2 or b == len(obands) - 1) else 3e-5), b


def test_chain_dev_equals_trace_then_deconvolution(psfs, monkeypatch):
    """thz_chain_dev == thz_trace_fused_dev + thz_deconvolution_dev within the deconvolved-map tolerance, and both
    agree with the oracle chain (config 3 shape scaled down)."""
    psf, opsf = psfs
    for k in ("THZ_CHAIN_FUSE", "THZ_CHAIN_EVEN", "THZ_EDGE_MMA"):
        monkeypatch.delenv(k, raising=False)
    n, w, h = 2048, 48, 40
    cube = synthetic_cube(w, h, n, seed=5, noise=0.02)
    t = time_axis(n)
    bands, why = pkg().host.Deconvolution(n_filters=6, n_iterations=40).plan(t, (w, h), 0.5, 0.5, psf)
    assert bands is not None, why
    P = w * h
    c = pkg().Context(0)
    try:
        c.plan_trace(n, *pkg().host.chain_multipliers(t))
        d_in, d_a, d_b = c.to_device(cube), c.alloc(cube.nbytes), c.alloc(cube.nbytes)
        d_ia, d_ib = c.alloc(P * 4), c.alloc(P * 4)
        c.chain_dev(d_in.ptr, w, h, n, bands, d_a.ptr, d_ia.ptr)
        st, km = c.deconv_stage_ms(), c.chain_kernel_ms()
        assert st["rl_iterations"] == sum(max(b.n_iter, 1) for b in bands)
        assert km["energy_spectra_ms"] > 0 and km["trace_ms"] == 0      # the fused kernel ran, no separate trace pass
        c.trace_fused_dev(d_in.ptr, d_b.ptr, d_ib.ptr, P)
        c._check(pkg().lib.thz_deconvolution_dev(c.handle, d_b.ptr, w, h, n, bands, len(bands), d_b.ptr, d_ib.ptr,
                                                 None, None, None))
        c.sync()
        a, ia = d_a.download((w, h, n)), d_ia.download((w, h))
        b, ib = d_b.download((w, h, n)), d_ib.download((w, h))
        # in place as well: d_out aliases d_in
        c.chain_dev(d_in.ptr, w, h, n, bands, d_in.ptr, d_ia.ptr)
        c.sync()
        assert np.array_equal(d_in.download((w, h, n)), a)
    finally:
        c.close()
    assert rel_err(a, b) <= 1e-5 and rel_err(ia, ib) <= 1e-5
    from helpers import slot0
    s0 = slot0(cube, t, dx=0.5, dy=0.5)
    s7 = orc.run_default_chain(s0)[7]
    ref = orc.Deconvolution(n_filters=6, n_iterations=40).filter(s7, opsf)
    assert rel_err(a, ref.data) <= TOL_MAP
    assert rel_err(ia, ref.img) <= TOL_MAP


@pytest.mark.parametrize("n", [512, 1024, 2048, 4096, 8192])
@pytest.mark.parametrize("kind", ["default", "wide"])
def test_spectral_handoff_equals_traces(psfs, monkeypatch, n, kind):
    """thz_chain_begin_dev / thz_chain_end_dev: with the spectral hand-off (default) the first half leaves FFT_N of
    the filtered pairs + their edge samples, and pass C skips its forward transform; THZ_CHAIN_SPECTRAL=off hands
    the filtered traces over.  Same deconvolved traces up to f32 rounding, edges included; same energies bit for bit."""
    psf, _ = psfs
    w, h = 6, 8
    cube = synthetic_cube(w, h, n, seed=n + 7, noise=0.05)
    rng = np.random.default_rng(9)
    cube[:, :, :60] += 0.5 * rng.standard_normal((w, h, 60)).astype(F32)
    cube[:, :, -60:] += 0.5 * rng.standard_normal((w, h, 60)).astype(F32)
    bands, _ = pkg().host.Deconvolution(n_filters=8).plan(time_axis(n), (64, 64), 0.5, 0.5, psf)
    P = w * h
    gains = (0.25 + 2.0 * rng.random((len(bands), P))).astype(F32)
    res = []
    for env in ({}, {"THZ_CHAIN_SPECTRAL": "off"}, {"THZ_EDGE_MMA": "off"}):
        for k in ("THZ_CHAIN_FUSE", "THZ_CHAIN_EVEN", "THZ_EDGE_MMA", "THZ_CHAIN_SPECTRAL"):
            monkeypatch.delenv(k, raising=False)
        for k, v in env.items():
            monkeypatch.setenv(k, v)
        c = pkg().Context(0)
        try:
            c.plan_trace(n, *_gates(n, kind))
            d_in, d_work, d_out = c.to_device(cube), c.alloc(cube.nbytes), c.alloc(cube.nbytes)
            d_img, d_e, d_g = c.alloc(P * 4), c.alloc(len(bands) * P * 4), c.to_device(gains)
            c.chain_begin_dev(d_in.ptr, d_work.ptr, d_img.ptr, P, n, bands, d_e.ptr)
            c.chain_end_dev(d_work.ptr, d_g.ptr, P, n, bands, d_out.ptr, d_img.ptr)
            c.sync()
            first = (d_out.download((w, h, n)), d_img.download((w, h)), d_e.download((len(bands), P)))
            # second half in place (d_out aliases d_work)
            c.chain_begin_dev(d_in.ptr, d_work.ptr, d_img.ptr, P, n, bands, d_e.ptr)
            c.chain_end_dev(d_work.ptr, d_g.ptr, P, n, bands, d_work.ptr, d_img.ptr)
            c.sync()
            assert np.array_equal(d_work.download((w, h, n)), first[0])
            res.append(first)
        finally:
            c.close()
    spec, plain, spec_fft_edges = res
    assert rel_err(spec[0], plain[0]) <= 2e-6
    peak = float(np.max(np.abs(plain[0])))
    for sl in (slice(0, 249), slice(n - 249, n)):
        # on the edges' own scale -- unless a gate has zeroed them (then they hold rounding noise of the whole trace)
        scale = max(float(np.max(np.abs(plain[0][:, :, sl]))), 0.1 * peak)
        assert float(np.max(np.abs(spec[0][:, :, sl].astype(np.float64) - plain[0][:, :, sl]))) <= 5e-6 * scale
    assert rel_err(spec[1], plain[1]) <= 1e-5
    if n >= 2048:      # tensor-core edges (TF32) in both: same bits; the transform edges agree to 1e-5
        assert np.array_equal(spec[2], plain[2])
    assert rel_err(spec_fft_edges[2], plain[2]) <= 3e-5
    assert np.array_equal(spec_fft_edges[0], spec[0])


def test_chain_dev_spectral_and_plain_agree(psfs, monkeypatch):
    """Whole chain on a device-resident cube, Richardson-Lucy included: spectral hand-off vs traces."""
    psf, _ = psfs
    n, w, h = 1024, 40, 36
    cube = synthetic_cube(w, h, n, seed=21, noise=0.02)
    t = time_axis(n)
    bands, why = pkg().host.Deconvolution(n_filters=6, n_iterations=30).plan(t, (w, h), 0.5, 0.5, psf)
    assert bands is not None, why
    outs = []
    for mode in ("on", "off"):
        monkeypatch.setenv("THZ_CHAIN_SPECTRAL", mode)
        c = pkg().Context(0)
        try:
            c.plan_trace(n, *pkg().host.chain_multipliers(t))
            d_in, d_out, d_img = c.to_device(cube), c.alloc(cube.nbytes), c.alloc(w * h * 4)
            c.chain_dev(d_in.ptr, w, h, n, bands, d_out.ptr, d_img.ptr)
            c.sync()
            outs.append((d_out.download((w, h, n)), d_img.download((w, h))))
            # the host-pointer twin goes the same way
            o, i, rc = c.chain(cube, bands)
            assert rc == 0
            assert rel_err(o, outs[-1][0]) <= 1e-6 and rel_err(i, outs[-1][1]) <= 1e-6
        finally:
            c.close()
    assert rel_err(outs[0][0], outs[1][0]) <= 1e-5 and rel_err(outs[0][1], outs[1][1]) <= 1e-5


def test_chain_dev_odd_trace_count_and_long_traces(psfs, monkeypatch):
    """Shapes off the fast path: an odd trace count (no whole pairs: the traces, not their spectra, are handed to pass
    C) and n = 8192 (one 512-thread CTA per pair) -- thz_chain_dev against the stage calls."""
    psf, _ = psfs
    for k in ("THZ_CHAIN_FUSE", "THZ_CHAIN_EVEN", "THZ_EDGE_MMA", "THZ_CHAIN_SPECTRAL"):
        monkeypatch.delenv(k, raising=False)
    for n, w, h in ((1024, 37, 33), (8192, 36, 34)):
        cube = synthetic_cube(w, h, n, seed=n + 1, noise=0.02)
        t = time_axis(n)
        bands, why = pkg().host.Deconvolution(n_filters=5, n_iterations=20).plan(t, (w, h), 1.0, 1.0, psf)
        assert bands is not None, why
        P = w * h
        c = pkg().Context(0)
        try:
            c.plan_trace(n, *pkg().host.chain_multipliers(t))
            d_in, d_a, d_b = c.to_device(cube), c.alloc(cube.nbytes), c.alloc(cube.nbytes)
            d_ia, d_ib = c.alloc(P * 4), c.alloc(P * 4)
            c.chain_dev(d_in.ptr, w, h, n, bands, d_a.ptr, d_ia.ptr)
            c.trace_fused_dev(d_in.ptr, d_b.ptr, d_ib.ptr, P)
            c._check(pkg().lib.thz_deconvolution_dev(c.handle, d_b.ptr, w, h, n, bands, len(bands), d_b.ptr, d_ib.ptr,
                                                     None, None, None))
            c.sync()
            assert rel_err(d_a.download((w, h, n)), d_b.download((w, h, n))) <= 1e-5
            assert rel_err(d_ia.download((w, h)), d_ib.download((w, h))) <= 1e-5
        finally:
            c.close()


def test_chain_dev_abort_and_errors(psfs):
    """abort flag (one byte, Rust AtomicBool layout) -> THZ_ABORTED before Richardson-Lucy; plan / pointer errors."""
    import ctypes
    psf, _ = psfs
    n, w, h = 1024, 40, 36
    cube = synthetic_cube(w, h, n, seed=3)
    t = time_axis(n)
    bands, _ = pkg().host.Deconvolution(n_filters=4, n_iterations=12).plan(t, (w, h), 1.0, 1.0, psf)
    L = pkg().lib
    c = pkg().Context(0)
    try:
        d_in, d_out, d_img = c.to_device(cube), c.alloc(cube.nbytes), c.alloc(w * h * 4)
        # no plan yet
        assert L.thz_chain_dev(c.handle, d_in.ptr, w, h, n, bands, len(bands), d_out.ptr, d_img.ptr, None, None, None) == -4
        c.plan_trace(n, *pkg().host.chain_multipliers(t))
        flag = ctypes.c_uint8(1)
        rc = L.thz_chain_dev(c.handle, d_in.ptr, w, h, n, bands, len(bands), d_out.ptr, d_img.ptr,
                             ctypes.addressof(flag), None, None)
        assert rc == 1                                                  # THZ_ABORTED
        assert L.thz_chain_dev(c.handle, d_in.ptr, w, h, n, bands, 0, d_out.ptr, d_img.ptr, None, None, None) == -1
        assert L.thz_chain_dev(c.handle, d_in.ptr, w, h, n, bands, len(bands), d_out.ptr, None, None, None, None) == -1
        assert L.thz_chain_energies_dev(c.handle, None, d_out.ptr, d_img.ptr, w * h, n, bands, len(bands), d_img.ptr) == -1
        # and the context still works afterwards
        c.chain_dev(d_in.ptr, w, h, n, bands, d_out.ptr, d_img.ptr)
        c.sync()
        assert np.isfinite(d_img.download((w, h))).all()
    finally:
        c.close()
