"""Row-slab sharding of the filter chain over the ranks of one node (one process per GPU).

The cube is C-contiguous (x, y, t) with x slowest, the axis the reference parallelises over
(src/math_tools.rs:333-340), so rank r owns the contiguous rows x in [x0, x1).  The trace
passes (fused chain, band energies, gain application) are slab-local and need no
communication.  The Richardson-Lucy images are W x H (tiny next to the cube): the B
band-energy images are all-gathered, the bands are dealt round-robin to the ranks (band b on
rank b % world, each running the full-image iteration), and the gains are summed back with
one all-reduce.  Collectives go through torch.distributed (NCCL on GPUs, gloo in the CPU
tests); the compute is injected through `ops` so that the same orchestration runs with
libthzgpu on device tensors and with the oracle in the world_size-2 CPU test.
"""
from __future__ import annotations

import torch


def slab_bounds(width: int, world: int, rank: int):
    """Contiguous rows [x0, x1) of rank `rank`; the first `width % world` ranks get one more row."""
    base, rem = divmod(width, world)
    x0 = rank * base + min(rank, rem)
    return x0, x0 + base + (1 if rank < rem else 0)


def band_owner(band: int, world: int) -> int:
    return band % world


def assign_bands(costs, world: int):
    """Owner rank of every band: longest-processing-time-first on the given per-band costs (deterministic, so
    every rank computes the same table).  Richardson-Lucy iteration counts fall steeply with frequency
    (423, 251, 127, ... at C5), so round-robin leaves the rank of band 0 with the small bands as well."""
    order = sorted(range(len(costs)), key=lambda b: (-float(costs[b]), b))
    load = [0.0] * world
    owner = [0] * len(costs)
    for b in order:
        r = min(range(world), key=lambda q: (load[q], q))
        owner[b] = r
        load[r] += float(costs[b])
    return owner


def gather_band_images(e_slab: torch.Tensor, width: int, height: int, dist, world: int) -> torch.Tensor:
    """e_slab [B][rows_r * H] on every rank -> [B][W * H] on every rank (uneven slabs allowed)."""
    B = e_slab.shape[0]
    if world == 1:
        return e_slab
    counts = [(slab_bounds(width, world, r)[1] - slab_bounds(width, world, r)[0]) * height for r in range(world)]
    mx = max(counts)
    padded = e_slab.new_zeros((B, mx))
    padded[:, : e_slab.shape[1]] = e_slab
    parts = [e_slab.new_empty((B, mx)) for _ in range(world)]
    dist.all_gather(parts, padded)
    full = e_slab.new_empty((B, width * height))
    off = 0
    for r in range(world):
        full[:, off: off + counts[r]] = parts[r][:, : counts[r]]
        off += counts[r]
    return full


def sharded_deconvolution(ops, slab, width: int, height: int, n_bands: int, dist, world: int, rank: int,
                          timings: dict | None = None, band_costs=None):
    """Deconvolution of this rank's slab.

    ops.energies(slab)            -> tensor [B][rows_r * H]
    ops.rl_gain(band, image_WxH)  -> tensor [W * H]  (gain image of one band)
    ops.apply(slab, gains_slab)   -> whatever the backend returns for the filtered slab

    `timings`, when given, receives the wall-clock seconds of the phases of this rank (the ops synchronise).
    `band_costs` (one number per band, e.g. iterations x taps) balances the bands over the ranks; without it
    band b runs on rank b % world.
    """
    import time

    def lap(name, t0):
        if timings is not None:
            if e_slab.is_cuda:
                torch.cuda.synchronize()
            timings[name] = timings.get(name, 0.0) + time.perf_counter() - t0
        return time.perf_counter()

    t = time.perf_counter()
    e_slab = ops.energies(slab)
    t = lap("energies", t)
    e_full = gather_band_images(e_slab, width, height, dist, world)
    t = lap("gather_energies", t)
    g_full = torch.zeros_like(e_full)
    owners = assign_bands(band_costs, world) if band_costs is not None else [band_owner(b, world) for b in range(n_bands)]
    for b in range(n_bands):
        if owners[b] == rank:
            g_full[b] = ops.rl_gain(b, e_full[b].reshape(width, height)).reshape(-1)
    t = lap("richardson_lucy_own_bands", t)
    if world > 1:
        dist.all_reduce(g_full)
    t = lap("reduce_gains_incl_wait", t)
    x0, x1 = slab_bounds(width, world, rank)
    g_slab = g_full[:, x0 * height: x1 * height].contiguous()
    out = ops.apply(slab, g_slab)
    lap("gain_application", t)
    return out


# --------------------------------------------------------------------------------------
# Row-slab Richardson-Lucy: the band images stay sharded, halo rows travel over NVLink
# --------------------------------------------------------------------------------------
def all_slab_bounds(width: int, world: int):
    """row_bounds[world + 1] as thz_slab_plan takes them."""
    return [slab_bounds(width, world, r)[0] for r in range(world)] + [width]


def exchange_handles(handle: bytes, dist, world: int):
    """Every rank's 64-byte arena handle (thz_slab_export), ordered by rank.  Host-side plumbing only: the handles
    go through torch.distributed's object collective, the data they describe never does."""
    if world == 1:
        return [handle]
    got = [None] * world
    dist.all_gather_object(got, handle)
    return got


class SlabExchange:
    """One rank's thz_slab plus the collective bookkeeping around it (plan -> export -> connect, with the host
    barriers thz_slab_plan asks for).  `slab` needs .plan / .export / .connect_ipc / .rl / .status (binding.Slab)."""

    def __init__(self, slab, dist, rank: int, world: int, sync):
        self.slab, self.dist, self.rank, self.world, self.sync = slab, dist, rank, world, sync

    def plan(self, width: int, height: int, bands):
        self.sync()                       # this rank is idle ...
        if self.world > 1:
            self.dist.barrier()           # ... and so is everybody else: nobody pushes into an arena being zeroed
        changed = self.slab.plan(all_slab_bounds(width, self.world), height, bands)
        if changed == 2:
            self.slab.connect_ipc(exchange_handles(self.slab.export(), self.dist, self.world))
        if changed and self.world > 1:
            self.dist.barrier()
        return changed


def sharded_deconvolution_slab(ops, slab, timings: dict | None = None):
    """Deconvolution of this rank's slab with the halo-exchanged Richardson-Lucy: three slab-local calls, no
    collective on the data path (ops.energies -> ops.slab_rl -> ops.apply).  `timings` receives the device
    milliseconds of the phases when ops.lap() provides them."""
    lap = getattr(ops, "lap", None)
    e = ops.energies(slab)
    if lap:
        lap("energies", timings)
    g = ops.slab_rl(e)
    if lap:
        lap("richardson_lucy_slab", timings)
    out = ops.apply(slab, g)
    if lap:
        lap("gain_application", timings)
    return out
