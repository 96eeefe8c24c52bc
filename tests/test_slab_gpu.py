"""Row-slab Richardson-Lucy (thz_slab_*): the halo-exchanged iteration over 2 / 3 / 4 slabs equals the unsharded
one.  Every output pixel sees the same taps in the same order in both forms, so the comparison is exact.

On a one-GPU box the slabs are emulated on one device (thz_slab_rl_serial: one stream, dependency order -- the
peer stores, flags and waits are the real ones, only the concurrency is missing); with two or more GPUs the
same run goes over NVLink (test_two_gpus_*)."""
import numpy as np
import pytest

from helpers import F32, pkg, rel_err, time_axis

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    c = pkg().Context(0)
    yield c
    c.close()


@pytest.fixture(scope="module")
def bands8(psf_npz_path):
    psf = pkg().host.PSF.load(psf_npz_path)
    bands, why = pkg().host.Deconvolution(n_filters=8, n_iterations=40).plan(time_axis(1024), (2048, 2048), 0.5, 0.5, psf)
    assert why is None
    return bands


def _bounds(rows, world, uneven=False):
    m = pkg()
    b = [m.sharding.slab_bounds(rows, world, r)[0] for r in range(world)] + [rows]
    if uneven and world > 1:
        b[1] += 5
    return b


def _energy_images(nb, rows, cols, seed):
    rng = np.random.default_rng(seed)
    yy, xx = np.meshgrid(np.arange(cols), np.arange(rows))
    base = 1.0 + 0.5 * ((xx // 8 + yy // 8) % 2) + 0.2 * np.sin(xx / 5.0)
    return np.stack([(base * (1 + 0.1 * b) + 0.05 * rng.random((rows, cols))).astype(F32) for b in range(nb)])


def _unsharded(ctx, e, bands):
    gains, us = [], []
    for b in range(len(bands)):
        u, g = ctx.richardson_lucy(e[b], bands[b].n_iter, bands[b].psf_x_np(), bands[b].psf_y_np(),
                                   direct=bool(bands[b].direct), want_gain=True)
        gains.append(g)
        us.append(u)
    return np.stack(us), np.stack(gains)


def _slab_serial(ctx, e, bands, bounds):
    m = pkg()
    world = len(bounds) - 1
    nb, rows, cols = e.shape
    slabs = [m.Slab(ctx, r, world) for r in range(world)]
    try:
        changed = [s.plan(bounds, cols, bands) for s in slabs]
        assert all(c == 2 for c in changed)
        for r, s in enumerate(slabs):
            s.connect_local(slabs[r - 1] if r > 0 else None, slabs[r + 1] if r + 1 < world else None)
        d_e, d_g, d_u, strides = [], [], [], []
        for r in range(world):
            x0, x1 = bounds[r], bounds[r + 1]
            part = np.ascontiguousarray(e[:, x0:x1, :])
            d_e.append(ctx.to_device(part))
            d_g.append(ctx.alloc(part.nbytes))
            d_u.append(ctx.alloc(part.nbytes))
            strides.append((x1 - x0) * cols)
        outs = []
        for _ in range(2):   # twice: the version counters and the halos carry over from run to run
            m.Slab.rl_serial(slabs, [b.ptr for b in d_e], strides, [b.ptr for b in d_g], [b.ptr for b in d_u])
            for s in slabs:
                s.status()
            g = np.concatenate([d_g[r].download((nb, bounds[r + 1] - bounds[r], cols)) for r in range(world)], axis=1)
            u = np.concatenate([d_u[r].download((nb, bounds[r + 1] - bounds[r], cols)) for r in range(world)], axis=1)
            outs.append((u, g))
        assert np.array_equal(outs[0][0], outs[1][0]) and np.array_equal(outs[0][1], outs[1][1])
        assert slabs[0].plan(bounds, cols, bands) == 0     # same plan again: nothing to do
        return outs[0]
    finally:
        for s in slabs:
            s.close()


@pytest.mark.parametrize("world,rows,cols,uneven", [(2, 300, 200, False), (3, 420, 131, True), (4, 560, 97, False)])
def test_slab_rl_equals_unsharded(ctx, bands8, world, rows, cols, uneven):
    """All 8 bands of the C5 plan (PSF 47x57 ... 7x7, correlation and convolution branches), iteration counts
    capped at 40, uneven slabs, a width that is not a multiple of the strip width."""
    e = _energy_images(len(bands8), rows, cols, seed=world)
    u_ref, g_ref = _unsharded(ctx, e, bands8)
    u, g = _slab_serial(ctx, e, bands8, _bounds(rows, world, uneven))
    assert rel_err(u, u_ref) <= 1e-6 and rel_err(g, g_ref) <= 1e-6
    assert np.array_equal(u, u_ref)      # same arithmetic per pixel in both forms


def test_slab_plan_refuses_thin_slabs(ctx, bands8):
    m = pkg()
    s = m.Slab(ctx, 0, 4)
    try:
        with pytest.raises(m.ThzError):
            s.plan(_bounds(160, 4), 128, bands8)    # 40 rows per slab < 3 x 23
    finally:
        s.close()


def _need_gpus(n):
    if pkg().lib.thz_device_count() < n:
        pytest.skip(f"needs {n} GPUs")


def test_two_gpus_slab_rl_over_nvlink(bands8):
    """The same comparison with the two slabs on two devices of one process: peer stores over NVLink, flags,
    concurrent kernels (one launching thread per device inside thz_group_rl)."""
    _need_gpus(2)
    m = pkg()
    rows, cols = 600, 420
    e = _energy_images(len(bands8), rows, cols, seed=7)
    c0 = m.Context(0)
    try:
        u_ref, g_ref = _unsharded(c0, e, bands8)
    finally:
        c0.close()
    with m.Group([0, 1]) as grp:
        g = grp.rl(e, bands8)
    assert np.array_equal(g, g_ref)


def test_two_gpus_one_process_per_gpu_ipc():
    """One process per GPU (torchrun, as bench.py runs): arenas exchanged as cudaIpc handles, halo rows over NVLink."""
    _need_gpus(2)
    import os
    import socket
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                        "--master-addr", "127.0.0.1", "--master-port", str(port),
                        os.path.join(root, "tools", "mgpu_slab_check.py")], capture_output=True, text=True, timeout=600)
    print(r.stdout[-2000:], r.stderr[-2000:])
    assert r.returncode == 0


def test_two_gpus_group_chain_equals_single_gpu(psf_npz_path):
    """thz_group_chain_host on two devices (one calling thread) == thz_chain_host on one."""
    _need_gpus(2)
    from helpers import default_multipliers, synthetic_cube
    m = pkg()
    w, h, n = 320, 64, 512
    cube = synthetic_cube(w, h, n, seed=2, noise=0.02)
    t, m_pre, band, m_post = default_multipliers(n)
    psf = m.host.PSF.load(psf_npz_path)
    bands, why = m.host.Deconvolution(n_filters=5, n_iterations=30).plan(t, (w, h), 0.5, 0.5, psf)
    assert why is None
    with m.Context(0) as c0:
        c0.plan_trace(n, m_pre, band, m_post)
        ref_out, ref_img, rc = c0.chain(cube, bands)
    with m.Group([0, 1]) as grp:
        out, img, rc = grp.chain(cube, m_pre, band, m_post, bands)
    assert rc == 0
    assert np.array_equal(out, ref_out) and np.array_equal(img, ref_img)


def test_group_on_one_device_emulates_the_slabs(ctx, psf_npz_path):
    """The same group call with three ranks on ONE device (serial emulation): available on a one-GPU box."""
    from helpers import default_multipliers, synthetic_cube
    m = pkg()
    w, h, n = 240, 48, 512
    cube = synthetic_cube(w, h, n, seed=4, noise=0.02)
    t, m_pre, band, m_post = default_multipliers(n)
    psf = m.host.PSF.load(psf_npz_path)
    bands, why = m.host.Deconvolution(n_filters=5, n_iterations=20).plan(t, (w, h), 0.5, 0.5, psf)
    assert why is None
    ctx.plan_trace(n, m_pre, band, m_post)
    ref_out, ref_img, rc = ctx.chain(cube, bands)
    with m.Group([0, 0, 0]) as grp:
        assert grp.row_bounds(w) == [0, 80, 160, 240]
        out, img, rc = grp.chain(cube, m_pre, band, m_post, bands)
        out2, img2, rc2 = grp.chain(cube, m_pre, band, m_post, bands)     # plan cached, counters carry over
    assert rc == 0 and rc2 == 0
    assert np.array_equal(out, ref_out) and np.array_equal(img, ref_img)
    assert np.array_equal(out2, ref_out)
