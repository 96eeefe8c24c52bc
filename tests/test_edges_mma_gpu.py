"""Pass-A edge energies on the tensor cores (csrc/thz_edges_mma.cu: tcgen05.mma kind::tf32, accumulators in TMEM)
against the transform kernel (k_fir_edges) and the oracle.  THZ_EDGE_MMA=on selects the GEMM form for n >= 2048."""
import numpy as np
import pytest

from helpers import F32, orc, pkg, rel_err, synthetic_cube, time_axis

pytestmark = pytest.mark.gpu


def _energies(env, monkeypatch, cube, bands, n):
    monkeypatch.setenv("THZ_EDGE_MMA", env)
    c = pkg().Context(0)
    try:
        P = cube.shape[0] * cube.shape[1]
        d_cube = c.to_device(cube)
        d_e = c.alloc(len(bands) * P * 4)
        c.deconv_energies_dev(d_cube.ptr, P, n, bands, d_e.ptr)
        return d_e.download((len(bands), cube.shape[0], cube.shape[1]))
    finally:
        c.close()


@pytest.mark.parametrize("n,w,h,kind", [(2048, 13, 11, "pulse"), (4096, 19, 14, "pulse"), (4096, 16, 17, "noise"),
                                        (2048, 9, 30, "noise")])
def test_edge_energies_tensor_core_form(psf_npz_path, monkeypatch, n, w, h, kind):
    """Odd pixel counts (a ragged last 128-trace tile and several tiles), pulses and white noise right up to both
    ends of the trace (the worst case for the edge correction: it is then ~10 % of a band energy)."""
    psf = pkg().host.PSF.load(psf_npz_path)
    opsf = orc.load_psf(psf_npz_path)
    if kind == "pulse":
        cube = synthetic_cube(w, h, n, seed=n + w, noise=0.02)
    else:
        cube = np.random.default_rng(n + h).standard_normal((w, h, n)).astype(F32)
    cube[2, 3, :] = 0.0                                   # a dead pixel keeps exactly zero energy
    t = time_axis(n)
    bands, _ = pkg().host.Deconvolution(n_filters=8).plan(t, (64, 64), 0.5, 0.5, psf)
    obands, _ = orc.Deconvolution(n_filters=8).plan(t, (64, 64, n), 0.5, 0.5, opsf)
    e_fft = _energies("off", monkeypatch, cube, bands, n)
    e_mma = _energies("on", monkeypatch, cube, bands, n)
    assert not np.array_equal(e_fft, e_mma)               # really another code path
    assert (e_mma[:, 2, 3] == 0).all()
    worst = 0.0
    for i, ob in enumerate(obands):
        filt = orc.filter_scan(cube, ob.fir)
        ref = np.sum(filt.astype(np.float64) ** 2, axis=2)
        err = np.max(np.abs(e_mma[i] - ref) / np.maximum(ref, 1e-30)[...] * (ref > 0))   # per pixel, relative
        worst = max(worst, float(err))
        assert rel_err(e_mma[i], ref) <= 1e-5, i
    print(f"n={n} {kind}: worst per-pixel relative error of a band energy {worst:.2e}")
    assert worst <= 3e-5
