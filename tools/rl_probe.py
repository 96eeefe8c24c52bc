"""Times Richardson-Lucy per band on a W x H image (debug / tuning aid)."""
import sys, time, os
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import importlib
m = importlib.import_module("thz-image-explorer_b200")
W = H = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
ctx = m.Context(0)
t = (np.float32(1000.0) + np.float32(0.05) * np.arange(4096, dtype=np.float32)).astype(np.float32)
psf = m.host.PSF.load(os.path.join(ROOT, "tests", "golden", "psf.npz"))
bands, _ = m.host.Deconvolution(n_filters=8).plan(t, (W, H), 0.5, 0.5, psf)
img = (1.0 + 0.5 * ((np.indices((W, H)).sum(axis=0) // 16) % 2)).astype(np.float32)
d_img = ctx.to_device(img)
d_g = ctx.alloc(img.nbytes)
for b in bands:
    px, py = np.ascontiguousarray(b.psf_x_np()), np.ascontiguousarray(b.psf_y_np())
    iters = 100
    for rep in range(2):
        ctx.sync()
        t0 = time.perf_counter()
        ctx._check(m.lib.thz_rl_separable_dev(ctx.handle, d_img.ptr, W, H, px.ctypes.data, px.size, py.ctypes.data,
                                              py.size, b.direct, iters, None, d_g.ptr, None, None, None, 0.0, 0.0))
        ctx.sync()
        dt = time.perf_counter() - t0
    print(f"psf {b.kx}x{b.ky} n_iter {b.n_iter}: {dt / iters * 1e6 / 2:.1f} us per kernel, {iters / dt:.0f} iters/s")
