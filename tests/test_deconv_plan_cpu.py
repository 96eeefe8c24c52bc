"""Host part of the deconvolution filter (C++ in libthzgpu, no GPU needed) against the oracle:
Kaiser FIR bank, PSF model evaluators, per-band PSF / iteration planning, skip conditions."""
import numpy as np
import pytest

from helpers import orc, pkg, time_axis


@pytest.fixture(scope="module")
def psfs(psf_npz_path):
    h = pkg().host
    return h.PSF.load(psf_npz_path), orc.load_psf(psf_npz_path)


def test_psf_model_evaluators(psfs):
    psf, opsf = psfs
    for f in [0.05, 0.1, 0.15, 0.3, 0.77, 1.0, 2.5, 4.99, 5.0, 7.0, 10.0]:
        assert psf.wx(f) == pytest.approx(float(opsf.wx_fit.eval_single(f)), rel=2e-6)
        assert psf.wy(f) == pytest.approx(float(opsf.wy_fit.eval_single(f)), rel=2e-6)
        assert psf.x0(f) == pytest.approx(float(opsf.x0_spline.eval_single_const_extrap(f)), rel=2e-6, abs=1e-7)
        assert psf.y0(f) == pytest.approx(float(opsf.y0_spline.eval_single_const_extrap(f)), rel=2e-6, abs=1e-7)


@pytest.mark.parametrize("nf,start,end,ww", [(25, 0.1, 10.0, 0.5), (8, 0.1, 10.0, 0.5), (5, 0.2, 4.0, 0.3)])
def test_fir_bank_matches_oracle(nf, start, end, ww):
    h = pkg().host
    t = time_axis(2048)
    filt, cen = h.fir_bank(nf, start, end, ww, t)
    ofilt, ocen = orc.create_filter_bank(nf, start, end, ww, t)
    np.testing.assert_allclose(cen, ocen, rtol=1e-7)
    np.testing.assert_allclose(filt, ofilt, rtol=0, atol=1e-9)
    # symmetric (linear phase) taps: the zero-phase spectrum used on the GPU is real
    assert np.max(np.abs(filt - filt[:, ::-1])) < 1e-7
    # the bank sums to (almost) a delta: low-pass + band-passes + high-pass telescope
    total = filt.astype(np.float64).sum(axis=0)
    assert abs(total[249] - 1.0) < 1e-5 and np.max(np.abs(np.delete(total, 249))) < 1e-6


@pytest.mark.parametrize("nf,dx,shape", [(8, 0.5, (2048, 2048)), (25, 1.0, (256, 256)), (25, 0.25, (300, 200)),
                                         (8, 0.5, (96, 80))])
def test_band_plans_match_oracle(psfs, nf, dx, shape):
    h = pkg().host
    psf, opsf = psfs
    t = time_axis(1024)
    bands, why = h.Deconvolution(n_filters=nf).plan(t, shape, dx, dx, psf)
    obands, owhy = orc.Deconvolution(n_filters=nf).plan(t, shape + (1024,), dx, dx, opsf)
    assert why == owhy
    if obands is None:
        assert bands is None
        return
    for b, o in zip(bands, obands):
        assert (b.kx, b.ky, b.n_iter) == (o.psf.shape[0], o.psf.shape[1], o.n_iter)
        assert b.direct == int(o.psf.size <= 256)
        np.testing.assert_allclose(b.psf_x_np(), o.psf_x, atol=2e-7)
        np.testing.assert_allclose(b.psf_y_np(), o.psf_y, atol=2e-7)
        np.testing.assert_allclose(b.fir_np(), o.fir, atol=1e-9)


def test_skip_conditions_mirror_reference(psfs):
    """Deconvolution::filter returns its input unchanged (deconvolution.rs:781-812, 873-885);
    the reference's own test_shape_preservation (2x2x64) hits the <16 early return."""
    h = pkg().host
    psf, opsf = psfs
    t = time_axis(64)
    d = h.Deconvolution()
    assert d.plan(t, (2, 2), 0.5, 0.5, psf) == (None, "image too small")
    assert d.plan(t, (64, 64), None, None, psf) == (None, "no dx/dy")
    assert d.plan(t, (64, 64), 0.5, 0.5, None) == (None, "no psf")
    assert d.plan(t, (20, 20), 0.5, 0.5, psf) == (None, "psf too large")
    assert orc.Deconvolution().plan(t, (20, 20, 64), 0.5, 0.5, opsf)[1] == "psf too large"
