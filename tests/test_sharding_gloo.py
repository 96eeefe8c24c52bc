"""world_size-2 (gloo, CPU) test of the multi-rank orchestration: row-slab partition, all-gather
of the band-energy images, band-parallel Richardson-Lucy, all-reduce of the gains, slab-local
gain application.  The compute is the oracle, so sharded == unsharded is checked end to end
without a GPU (the same `sharded_deconvolution` drives libthzgpu in bench.py)."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from helpers import F32, orc, pkg, synthetic_cube, time_axis  # noqa: E402

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
W, H, N, NB = 21, 18, 256, 4   # odd width: uneven slabs


def _setup():
    cube = synthetic_cube(W, H, N, seed=3, noise=0.02)
    t = time_axis(N)
    opsf = orc.load_psf(os.path.join(ROOT, "tests", "golden", "psf.npz"))
    bands, why = orc.Deconvolution(n_filters=NB, n_iterations=12).plan(t, (W, H, N), 1.0, 1.0, opsf)
    assert why is None
    return cube, t, opsf, bands


class OracleOps:
    def __init__(self, bands):
        self.bands = bands
        self.filtered = None

    def energies(self, slab):
        self.filtered = [orc.filter_scan(slab, b.fir) for b in self.bands]
        return torch.from_numpy(np.stack([np.sum(f * f, axis=2, dtype=F32).reshape(-1) for f in self.filtered]))

    def rl_gain(self, b, image):
        img = image.numpy()
        u = np.maximum(orc.richardson_lucy(img, self.bands[b].psf, self.bands[b].n_iter), 0)
        with np.errstate(divide="ignore", invalid="ignore"):
            return torch.from_numpy(np.sqrt(u / img).astype(F32).reshape(-1))

    def apply(self, slab, g_slab):
        g = g_slab.numpy().reshape(len(self.bands), slab.shape[0], slab.shape[1])
        out = np.zeros_like(slab)
        for b, f in enumerate(self.filtered):
            out = out + f * g[b][:, :, None]
        return out


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        sh = pkg().sharding
        cube, t, opsf, bands = _setup()
        x0, x1 = sh.slab_bounds(W, world, rank)
        costs = [b.n_iter * (100 + b.psf.shape[0] + b.psf.shape[1]) for b in bands]
        out = sh.sharded_deconvolution(OracleOps(bands), cube[x0:x1], W, H, NB, dist, world, rank, band_costs=costs)
        q.put((rank, x0, x1, out))
    finally:
        dist.destroy_process_group()


def test_band_assignment_balances_iteration_counts():
    sh = pkg().sharding
    costs = [423 * 204, 251 * 160, 127 * 132, 46 * 118, 13 * 114, 4 * 114, 3 * 114, 1 * 114]   # C5 plan
    for world in (1, 2, 3, 4, 8):
        own = sh.assign_bands(costs, world)
        assert sorted(set(own)) == list(range(min(world, len(costs)))) and len(own) == len(costs)
        load = [sum(c for c, o in zip(costs, own) if o == r) for r in range(world)]
        assert max(load) == costs[0] if world > 1 else True     # band 0 alone bounds the makespan
    assert sh.assign_bands(costs, 2) == [0, 1, 1, 1, 1, 1, 1, 1]
    assert sh.assign_bands([], 4) == []


def test_slab_bounds_cover_the_image():
    sh = pkg().sharding
    for width in (1, 7, 8, 2048, 2049):
        for world in (1, 2, 4, 8):
            b = [sh.slab_bounds(width, world, r) for r in range(world)]
            assert b[0][0] == 0 and b[-1][1] == width
            assert all(b[i][1] == b[i + 1][0] for i in range(world - 1))
            sizes = [x1 - x0 for x0, x1 in b]
            assert max(sizes) - min(sizes) <= 1


@pytest.mark.timeout(300)
def test_sharded_equals_unsharded_world2():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctxmp = mp.get_context("spawn")
    q = ctxmp.Queue()
    procs = [ctxmp.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=240) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    cube, t, opsf, bands = _setup()
    ref = np.zeros_like(cube)
    for b in bands:
        f = orc.filter_scan(cube, b.fir)
        e = np.sum(f * f, axis=2, dtype=F32)
        u = np.maximum(orc.richardson_lucy(e, b.psf, b.n_iter), 0)
        ref = ref + f * np.sqrt(u / e).astype(F32)[:, :, None]
    got = np.zeros_like(cube)
    for rank, x0, x1, out in res:
        got[x0:x1] = out
    # identical arithmetic on identical inputs: the decomposition must not change a bit
    assert np.array_equal(got, ref)
