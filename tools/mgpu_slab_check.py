#!/usr/bin/env python
"""One process per GPU (torchrun): the row-slab Richardson-Lucy over cudaIpc-mapped arenas and NVLink peer stores
against the unsharded iteration on rank 0.  Exits non-zero on any mismatch.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
        tools/mgpu_slab_check.py [rows cols]
"""
import importlib
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    m = importlib.import_module("thz-image-explorer_b200")
    rows, cols = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (640, 420)
    t = (np.float32(1000.0) + np.float32(0.05) * np.arange(1024, dtype=np.float32)).astype(np.float32)
    psf = m.host.PSF.load(os.path.join(ROOT, "tests", "golden", "psf.npz"))
    bands, why = m.host.Deconvolution(n_filters=8, n_iterations=60).plan(t, (2048, 2048), 0.5, 0.5, psf)
    assert why is None
    B = len(bands)
    rng = np.random.default_rng(5)
    yy, xx = np.meshgrid(np.arange(cols), np.arange(rows))
    base = 1.0 + 0.5 * ((xx // 8 + yy // 8) % 2) + 0.2 * np.sin(xx / 5.0)
    e = np.stack([(base * (1 + 0.1 * b) + 0.05 * rng.random((rows, cols))).astype(np.float32) for b in range(B)])
    ctx = m.Context(local)
    bounds = m.sharding.all_slab_bounds(rows, world)
    x0, x1 = bounds[rank], bounds[rank + 1]
    slab = m.Slab(ctx, rank, world)
    ex = m.sharding.SlabExchange(slab, dist, rank, world, sync=ctx.sync)
    assert ex.plan(rows, cols, bands) == 2
    part = np.ascontiguousarray(e[:, x0:x1, :])
    d_e, d_g = ctx.to_device(part), ctx.alloc(part.nbytes)
    P = (x1 - x0) * cols
    ok = True
    for rep in range(3):                       # version counters and halos carry over between runs
        slab.rl(d_e.ptr, P, d_g.ptr)
        slab.status()
        g = torch.from_numpy(d_g.download((B, x1 - x0, cols))).cuda()
        parts = [torch.empty((B, bounds[r + 1] - bounds[r], cols), dtype=torch.float32, device="cuda") for r in range(world)]
        if len({p.shape for p in parts}) == 1:
            dist.all_gather(parts, g)
        else:
            for r in range(world):
                src = g if r == rank else parts[r]
                dist.broadcast(src, r)
                parts[r] = src
        if rank == 0:
            full = torch.cat(parts, dim=1).cpu().numpy()
            ref = np.stack([ctx.richardson_lucy(e[b], bands[b].n_iter, bands[b].psf_x_np(), bands[b].psf_y_np(),
                                                direct=bool(bands[b].direct), want_gain=True)[1] for b in range(B)])
            same = np.array_equal(full, ref)
            err = float(np.max(np.abs(full - ref)) / np.max(np.abs(ref)))
            print(f"rep {rep}: {world} ranks, {rows}x{cols}, {B} bands: identical={same} rel err {err:.2e}", flush=True)
            ok = ok and same
    flag = torch.tensor([1 if ok else 0], device="cuda")
    dist.broadcast(flag, 0)
    dist.barrier()
    slab.close()
    ctx.close()
    dist.destroy_process_group()
    sys.exit(0 if int(flag.item()) == 1 else 1)


if __name__ == "__main__":
    main()
