"""Third-party cross-pins of the CPU oracle.  The reference holds no golden vectors for FIR taps, 2-D filtering or
Richardson-Lucy iterates and cannot be built here (no rustc), so the oracle's restatements of those parts are
pinned against independent implementations of the same published algorithms:

  * `create_filter_bank` / `firwin_kaiser_*` (deconvolution.rs:30-211, a port of scipy's Kaiser-window `firwin`)
    against scipy.signal.firwin / kaiser_atten / kaiser_beta / windows.kaiser and scipy.special.i0;
  * both branches of `convolve2d` (deconvolution.rs:432-545) against scipy.signal.correlate2d / fftconvolve;
  * `richardson_lucy` (deconvolution.rs:620-712) against a float64 Richardson-Lucy written with scipy only
    (np.pad(mode="reflect") + scipy.signal.fftconvolve), at the iteration counts the benchmark runs (423 for the
    47x57 PSF of band 0);
  * `filter_scan` (deconvolution.rs:266-317, 574-609) against scipy.signal.lfilter-free direct convolution
    (np.convolve, "same" alignment).
These do not replace the missing reference vectors (parity stays "unpinned" for absolute values) but a
common-mode misreading of the source by oracle and product would have to be shared by scipy as well."""
import os

import numpy as np
import pytest
import scipy.signal as sig
import scipy.special as sp

from oracle import thz_oracle as O

F32 = np.float32
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_kaiser_design_formulas_match_scipy():
    for ntaps, width_ratio in ((499, 0.05), (499, 0.01), (101, 0.2), (31, 0.02)):
        a = O.kaiser_atten(ntaps, width_ratio)
        assert a == pytest.approx(max(sig.kaiser_atten(ntaps, width_ratio), 0.0), rel=1e-12)
        assert O.kaiser_beta(a) == pytest.approx(sig.kaiser_beta(a), rel=1e-12)
    for x in (0.0, 0.5, 3.0, 19.6, 40.0):
        assert O.i0(x) == pytest.approx(float(sp.i0(x)), rel=1e-10)
    beta = 19.57
    w = np.array([O.kaiser_window_coeff(n, 499, beta) for n in range(499)])
    ws = sig.windows.kaiser(499, beta, sym=True)
    np.testing.assert_allclose(w[1:-1], ws[1:-1], rtol=1e-9, atol=1e-15)
    assert w[0] == 0.0 and w[-1] == 0.0 and ws[0] < 1e-7      # the reference zeroes the (negligible) end points


def test_fir_bank_matches_scipy_firwin():
    """The 8-band bank of BASELINE configs 3 / 5: low-pass, six band-passes, high-pass."""
    n = 4096
    t = (F32(1000.0) + F32(0.05) * np.arange(n, dtype=F32)).astype(F32)
    filters, centers = O.create_filter_bank(8, 0.1, 10.0, 0.5, t)
    fs = 1.0 / float(t[1] - t[0])
    beta = sig.kaiser_beta(sig.kaiser_atten(499, 0.5 / (0.5 * fs)))
    c = [float(v) for v in centers]
    edges = [np.sqrt(c[i] * c[i + 1]) for i in range(7)]
    win = ("kaiser", beta)

    def hp(cut):   # the reference's high-pass: spectral inversion of the unit-DC low-pass (deconvolution.rs:111-131)
        h = -sig.firwin(499, cut, window=win, fs=fs)
        h[249] += 1.0
        return h

    ref = [sig.firwin(499, edges[0], window=win, fs=fs)]                                   # low-pass
    for i in range(1, 7):                                                                   # HP(lo) - HP(hi)
        ref.append(hp(edges[i - 1]) - hp(edges[i]))
    ref.append(hp(edges[6]))                                                                # high-pass
    for i in range(8):
        np.testing.assert_allclose(filters[i], ref[i], atol=3e-8, rtol=0, err_msg=f"band {i}")   # taps are stored as f32 (half an ulp of 0.28 is 1.5e-8)
        assert np.max(np.abs(filters[i] - filters[i][::-1])) <= 1e-9      # linear phase: symmetric taps
    # scipy's own high-pass (scaled for unit gain at Nyquist instead of inverting a unit-DC low-pass) is the same
    # filter wherever the cut-off is resolved by the 0.5 THz transition width: bands 2..7.  Band 1 (cut-off
    # 0.139 THz, inside the transition band) is where the two conventions differ (1.5e-4); the reference's is kept.
    for i in range(2, 7):
        alt = (sig.firwin(499, edges[i - 1], window=win, pass_zero=False, fs=fs)
               - sig.firwin(499, edges[i], window=win, pass_zero=False, fs=fs))
        np.testing.assert_allclose(filters[i], alt, atol=2e-8, rtol=0)
    np.testing.assert_allclose(filters[7], sig.firwin(499, edges[6], window=win, pass_zero=False, fs=fs), atol=2e-8, rtol=0)


def test_filter_scan_matches_direct_convolution():
    rng = np.random.default_rng(3)
    n = 700
    x = rng.standard_normal((2, 3, n)).astype(F32)
    t = (F32(1000.0) + F32(0.05) * np.arange(n, dtype=F32)).astype(F32)
    filters, _ = O.create_filter_bank(5, 0.1, 10.0, 0.5, t)
    for fir in filters[[0, 2, 4]]:
        got = O.filter_scan(x, fir)
        for p in np.ndindex(2, 3):
            full = np.convolve(x[p].astype(np.float64), fir.astype(np.float64), mode="full")
            np.testing.assert_allclose(got[p], full[249:249 + n], atol=2e-6, rtol=0)


@pytest.mark.parametrize("shape,kx,ky", [((40, 33), 7, 5), ((64, 50), 17, 15), ((90, 70), 31, 29)])
def test_convolve2d_branches_match_scipy(shape, kx, ky):
    rng = np.random.default_rng(kx)
    img = rng.uniform(0.1, 1.0, shape).astype(F32)
    gx = np.exp(-2 * ((np.arange(kx) - kx // 2 - 0.4) / (kx / 4)) ** 2)
    gy = np.exp(-2 * ((np.arange(ky) - ky // 2 + 0.7) / (ky / 4)) ** 2)
    psf = np.outer(gx, gy).astype(F32)
    corr = sig.correlate2d(img.astype(np.float64), psf.astype(np.float64), mode="same", boundary="fill")
    conv = sig.fftconvolve(img.astype(np.float64), psf.astype(np.float64), mode="same")
    # direct branch = correlation (deconvolution.rs:441-456), FFT branch = convolution (:517-536); off-centre PSF
    np.testing.assert_allclose(O.direct_convolve2d(img, psf), corr, rtol=0, atol=2e-5 * np.abs(corr).max())
    np.testing.assert_allclose(O.fft_convolve2d(img, psf), conv, rtol=0, atol=2e-5 * np.abs(conv).max())
    assert np.abs(corr - conv).max() > 1e-3 * np.abs(corr).max()      # the two really differ for this PSF
    want = corr if kx * ky <= 256 else conv
    np.testing.assert_allclose(O.convolve2d(img, psf), want, rtol=0, atol=2e-5 * np.abs(want).max())


def _rl_scipy_f64(image, psf, n_iter, direct):
    """Richardson-Lucy with scipy only, float64: reflect pad by the PSF half-extents, zero boundary beyond,
    u0 = d = padded image, crop (the published iteration; `direct` picks correlation-then-convolution as the
    reference's small-PSF branch does, else convolution-then-correlation)."""
    psf = psf.astype(np.float64)
    py, px = psf.shape[0] // 2, psf.shape[1] // 2
    d = np.pad(image.astype(np.float64), ((py, py), (px, px)), mode="reflect")
    u = d.copy()
    k1 = psf[::-1, ::-1] if direct else psf              # fftconvolve convolves: flip to correlate
    k2 = psf if direct else psf[::-1, ::-1]
    for _ in range(n_iter):
        c = sig.fftconvolve(u, k1, mode="same")
        r = d / (c + 1e-12)
        u = u * sig.fftconvolve(r, k2, mode="same")
    return u[py:py + image.shape[0], px:px + image.shape[1]]


@pytest.mark.parametrize("band_idx,shape", [(0, (120, 130)), (1, (90, 80)), (2, (70, 64)), (4, (48, 40))])
def test_richardson_lucy_matches_scipy_f64_at_benchmark_iterations(band_idx, shape):
    """Bands of the C5 plan with their full iteration counts (423, 251, 127, 13): f32 oracle vs f64 scipy.  The
    multiplicative update does not amplify rounding: the f32 iterate stays within 1e-4 of the f64 one."""
    psf = O.load_psf(os.path.join(ROOT, "tests", "golden", "psf.npz"))
    n = 1024
    t = (F32(1000.0) + F32(0.05) * np.arange(n, dtype=F32)).astype(F32)
    bands, why = O.Deconvolution(n_filters=8, n_iterations=500).plan(t, (2048, 2048, n), 0.5, 0.5, psf)
    assert why is None
    b = bands[band_idx]
    assert [bb.n_iter for bb in bands] == [423, 251, 127, 46, 13, 4, 3, 1]
    yy, xx = np.meshgrid(np.arange(shape[1]), np.arange(shape[0]))
    img = (1.0 + 0.5 * ((xx // 8 + yy // 8) % 2) + 0.2 * np.sin(xx / 5.0)).astype(F32)
    got = O.richardson_lucy(img, b.psf, b.n_iter)
    ref = _rl_scipy_f64(img, b.psf, b.n_iter, direct=b.psf.size <= 256)
    err = np.abs(got - ref).max() / np.abs(ref).max()
    assert err <= 1e-4, err
