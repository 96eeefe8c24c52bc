"""world_size-2 / 3 (gloo, CPU) test of the row-slab Richardson-Lucy decomposition that libthzgpu runs over NVLink
(thz_slab_*, csrc/thz_rl.cu): every rank iterates its own rows of the reflect-padded band image and receives the
kx/2 boundary rows of `u` (before the first filtering of an iteration) and of the relative blur (before the
second) from its neighbours.  Here the exchange is torch.distributed send / recv and the arithmetic is the
oracle's direct `convolve2d`, so that sharded == unsharded can be asserted bit for bit without a GPU; the geometry
(own rows of the padded domain, halo placement, reflect padding from the rank's own rows, crop) is the same as in
slab_band_geometry / k_reflect_pad_slab / k_rl_finish_slab.  The collective plan -> export -> connect sequence of
sharding.SlabExchange runs against a fake slab object."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from helpers import F32, orc, pkg  # noqa: E402

ROWS, COLS = 47, 23
PSFS = [((7, 5), 9), ((5, 9), 6), ((1, 3), 3)]     # (kx, ky), iterations; (1, 3): no halo at all


def _psf(kx, ky):
    gx = np.exp(-2 * ((np.arange(kx) - kx // 2 - 0.3) / max(kx / 4, 0.8)) ** 2)
    gy = np.exp(-2 * ((np.arange(ky) - ky // 2 + 0.4) / max(ky / 4, 0.8)) ** 2)
    return np.outer(gx / gx.max(), gy / gy.max()).astype(F32)


def _image():
    yy, xx = np.meshgrid(np.arange(COLS), np.arange(ROWS))
    rng = np.random.default_rng(11)
    return (1.0 + 0.5 * ((xx // 4 + yy // 4) % 2) + 0.1 * rng.random((ROWS, COLS))).astype(F32)


def slab_rl_numpy(img_rows, bounds, rank, world, psf, n_iter, dist):
    """This rank's rows of richardson_lucy(image, psf, n_iter): own rows of the padded domain + halo exchange."""
    kx, ky = psf.shape
    halo, pad_x = kx // 2, ky // 2
    rows_total = bounds[-1]
    x0, x1 = bounds[rank], bounds[rank + 1]
    Hp, Wp = rows_total + 2 * halo, COLS + 2 * pad_x
    own_lo = 0 if rank == 0 else x0 + halo
    own_hi = Hp if rank == world - 1 else x1 + halo
    own = own_hi - own_lo
    assert x1 - x0 > halo and own >= 3 * max(halo, 1)
    # reflect padding from the rank's own image rows (k_reflect_pad_slab)
    d = np.zeros((own + 2 * halo, Wp), F32)
    for o in range(own):
        r = own_lo + o
        sr = halo - r if r < halo else (rows_total - 2 - (r - halo - rows_total) if r >= halo + rows_total else r - halo)
        row = img_rows[sr - x0]
        d[halo + o, pad_x:pad_x + COLS] = row
        for j in range(pad_x):
            d[halo + o, j] = row[pad_x - j]
            d[halo + o, pad_x + COLS + j] = row[COLS - 2 - j]
    u = d.copy()
    mirror = np.ascontiguousarray(psf[::-1, ::-1])
    eps = F32(1e-12)
    own_sl = slice(halo, halo + own)

    def exchange(a):
        """boundary rows -> the neighbours' halos (the k_rl_stream epilogue / k_slab_push stores)"""
        if halo == 0:
            return
        reqs = []
        top, bot = torch.from_numpy(a[halo:2 * halo].copy()), torch.from_numpy(a[own:own + halo].copy())
        rtop, rbot = torch.empty_like(top), torch.empty_like(bot)
        if rank > 0:
            reqs += [dist.isend(top, rank - 1), dist.irecv(rtop, rank - 1)]
        if rank + 1 < world:
            reqs += [dist.isend(bot, rank + 1), dist.irecv(rbot, rank + 1)]
        for q in reqs:
            q.wait()
        if rank > 0:
            a[0:halo] = rtop.numpy()
        if rank + 1 < world:
            a[halo + own:] = rbot.numpy()

    exchange(u)
    for _ in range(n_iter):
        c = orc.direct_convolve2d(u, psf)
        r = np.zeros_like(u)
        r[own_sl] = d[own_sl] / (c[own_sl] + eps)
        exchange(r)
        corr = orc.direct_convolve2d(r, mirror)
        u[own_sl] = u[own_sl] * corr[own_sl]
        exchange(u)
    first = halo + (x0 + halo - own_lo)
    return np.ascontiguousarray(u[first:first + (x1 - x0), pad_x:pad_x + COLS])


class FakeSlab:
    """Stands in for binding.Slab in the SlabExchange bookkeeping test."""

    def __init__(self, rank):
        self.rank, self.calls, self.key = rank, [], None

    def plan(self, bounds, cols, bands):
        key = (tuple(bounds), cols, bands)
        changed = 0 if key == self.key else 2
        self.key = key
        self.calls.append(("plan", changed))
        return changed

    def export(self):
        return bytes([self.rank]) * 64

    def connect_ipc(self, handles):
        self.calls.append(("connect", [h[0] for h in handles]))


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        sh = pkg().sharding
        img = _image()
        bounds = sh.all_slab_bounds(ROWS, world)
        outs = []
        for (kx, ky), it in PSFS:
            outs.append(slab_rl_numpy(img[bounds[rank]:bounds[rank + 1]], bounds, rank, world, _psf(kx, ky), it, dist))
        fake = FakeSlab(rank)
        ex = sh.SlabExchange(fake, dist, rank, world, sync=lambda: None)
        first = ex.plan(ROWS, COLS, "bands-A")
        again = ex.plan(ROWS, COLS, "bands-A")
        q.put((rank, bounds[rank], bounds[rank + 1], outs, first, again, fake.calls))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(300)
@pytest.mark.parametrize("world", [2, 3])
def test_slab_decomposition_equals_unsharded(world):
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctxmp = mp.get_context("spawn")
    q = ctxmp.Queue()
    procs = [ctxmp.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=240) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    img = _image()
    for i, ((kx, ky), it) in enumerate(PSFS):
        ref = orc.richardson_lucy(img, _psf(kx, ky), it, conv=orc.direct_convolve2d)
        got = np.zeros_like(ref)
        for rank, x0, x1, outs, *_ in res:
            got[x0:x1] = outs[i]
        assert np.array_equal(got, ref), (kx, ky)
    for rank, x0, x1, outs, first, again, calls in res:
        assert first == 2 and again == 0
        assert calls == [("plan", 2), ("connect", list(range(world))), ("plan", 0)]
