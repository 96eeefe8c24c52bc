#!/usr/bin/env python
"""bench.py -- throughput of the thz-image-explorer filter-chain hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

One "step" = one pass of the hot path over the synthetic cube BASELINE.json names
(config 5: 2048 x 2048 pixels x 4096 samples): the fused trace pass (window -> rFFT ->
band-pass -> irFFT -> gate -> intensity) and, when built, the PSF deconvolution that ends the
chain.  The cube is row-slab sharded over the ranks (strong scaling: total work fixed).
Prints ONE JSON line (rank 0).  See DESIGN.md "Measurement" for the definitions.

`--impl reference` times the reference's CPU algorithm (the oracle port: the Rust reference
cannot be built here, no rustc) on the host cores, on a bounded sample of the same workload.
"""
from __future__ import annotations

import argparse
import importlib
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
PKG = "thz-image-explorer_b200"

METRIC = "pixel_traces_per_s"
UNIT = "traces/s"


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            d = json.load(open(p))
            return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                 "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons, pw = [], [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
                pw.append(float(f[3]))
            except ValueError:
                continue
            for nm, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons),
                "samples": len(sm), "power_w_max": float(max(pw))}


CONFIGS = {
    # BASELINE.json `configs`, at the stand-in shapes SURVEY.md 8d names (the real scans are LFS pointers)
    "c1": dict(width=200, height=200, samples=2048, deconv=False, spectra=False,
               name="BASELINE config 1 stand-in: 200x200 pixels x 2048 samples, default filter chain"),
    "c2": dict(width=100, height=100, samples=2048, deconv=False, spectra=True,
               name="BASELINE config 2 stand-in: 100x100 pixels x 2048 samples, spectra materialised and normalised "
                    "by a reference pulse (amplitude ratio / phase difference maps)"),
    "c3": dict(width=256, height=256, samples=2048, deconv=True, spectra=False,
               name="BASELINE config 3 stand-in: 256x256 pixels x 2048 samples, default chain + deconvolution "
                    "(psf.npz, 8 bands, n_iterations 500)"),
    "c4": dict(width=1024, height=1024, samples=2048, deconv=False, spectra=False,
               name="BASELINE config 4: synthetic 1024x1024 pixels x 2048 samples, full chain with both time gates"),
    "c5": dict(width=2048, height=2048, samples=4096, deconv=True, spectra=False,
               name="BASELINE config 5: synthetic 2048x2048 pixels x 4096 samples, full chain + deconvolution"),
}


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="c5", choices=sorted(CONFIGS))
    ap.add_argument("--width", type=int, default=0)
    ap.add_argument("--height", type=int, default=0)
    ap.add_argument("--samples", type=int, default=0)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-deconv", action="store_true", help="trace pass only (skip the PSF deconvolution)")
    ap.add_argument("--bands", type=int, default=8)
    ap.add_argument("--rl-iterations", type=int, default=500)
    ap.add_argument("--rl", default="slab", choices=["slab", "bands"],
                    help="several GPUs: halo-exchanged row slabs (default) or the older band-parallel split")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--cpu-rows", type=int, default=0, help="rows of the CPU-baseline slab (0 = auto)")
    a = ap.parse_args()
    cfg = CONFIGS[a.config]
    a.width = a.width or cfg["width"]
    a.height = a.height or cfg["height"]
    a.samples = a.samples or cfg["samples"]
    if not cfg["deconv"]:
        a.no_deconv = True
    a.spectra = cfg["spectra"]
    a.workload = cfg["name"]
    return a


def workload_name(a):
    return f"synthetic {a.width}x{a.height} pixels x {a.samples} samples -- {a.workload}"


# --------------------------------------------------------------------------------------
# reference arm / CPU baseline: the oracle (C twin for the chain, numpy/scipy for the deconvolution)
# on the host cores, on a bounded sample of the same workload
# --------------------------------------------------------------------------------------
def cpu_reference_sample(a, rows_chain=4, rows_scan=1, rl_iters=1, with_deconv=True):
    """Times the restated reference CPU path on a bounded sample of the workload and scales it to the
    full cube (every part is linear in the quantity it is scaled by):
      chain    : slots 1..7 on a rows_chain x H x N slab (oracle/thz_oracle_c.c, reference threading)
      fir scan : `filter_scan` + squares + gain multiply for all bands on a rows_scan x H x N slab
      RL       : rl_iters Richardson-Lucy iterations per band on the FULL W x H image, one thread per
                 band as in the reference (bands run in parallel there, so the wall time is the slowest
                 band: max_b n_iter_b * t_iter_b)
    Returns a dict with the extrapolated full-cube seconds and traces/s."""
    from oracle import thz_oracle as orc     # timed CPU baseline (cpu_baseline / --impl reference only)
    from oracle import c_twin
    cores = min(16, len(os.sched_getaffinity(0)))   # fixed worker count: more scipy workers get slower on the 32-core boxes
    W, H, N = a.width, a.height, a.samples
    P_total = W * H
    rng = np.random.default_rng(1)
    t = (np.float32(1000.0) + np.float32(0.05) * np.arange(N, dtype=np.float32)).astype(np.float32)
    tt = np.arange(N) * 0.05 - 10.0
    pulse = np.exp(-(tt / 0.3) ** 2) * np.cos(2 * np.pi * tt)
    cube = (pulse + 0.01 * rng.standard_normal((rows_chain, H, N))).astype(np.float32)
    tilt = orc.adapted_blackman_multiplier(t, 0.0, 7.0)
    gb = orc.td_gate_multiplier(t, float(t[0]), float(t[-1]), 2.0)
    win = orc.fft_window_multiplier(t)
    band = orc.fd_band_multiplier(orc.frequency_axis(t))
    ga = orc.td_gate_multiplier(t, float(t[0]), float(t[-1]), 0.1)
    t0 = time.perf_counter()
    out, _ = c_twin.default_chain(cube, tilt, gb, win, band, ga, threads=cores)
    t_chain = time.perf_counter() - t0
    parts = {"chain_s_per_trace": t_chain / (rows_chain * H)}
    full_s = parts["chain_s_per_trace"] * P_total
    measured = t_chain
    sample = f"chain: {rows_chain}x{H}x{N} slab (C twin, {cores} threads, reference threading)"
    if with_deconv:
        psf = orc.load_psf(os.path.join(ROOT, "tests", "golden", "psf.npz"))
        bands, why = orc.Deconvolution(n_filters=a.bands, n_iterations=a.rl_iterations).plan(t, (W, H, N), 0.5, 0.5, psf)
        slab = out[:rows_scan]
        t0 = time.perf_counter()
        acc = np.zeros_like(slab)
        for b in bands:
            f = orc.filter_scan(slab, b.fir, workers=cores)
            e = np.sum(f * f, axis=2, dtype=np.float32)
            acc = acc + f * np.sqrt(e)[:, :, None]
        t_scan = time.perf_counter() - t0
        parts["fir_scan_s_per_trace"] = t_scan / (rows_scan * H)
        img = (1.0 + 0.5 * (np.indices((W, H)).sum(axis=0) // 16 % 2)).astype(np.float32)
        t0 = time.perf_counter()
        worst = 0.0
        import scipy.fft as sfft
        with sfft.set_workers(1):
            for b in bands:
                t1 = time.perf_counter()
                orc.richardson_lucy(img, b.psf, rl_iters)
                worst = max(worst, (time.perf_counter() - t1) / rl_iters * b.n_iter)
        t_rl = time.perf_counter() - t0
        parts["rl_s_full_slowest_band"] = worst
        full_s += parts["fir_scan_s_per_trace"] * P_total + worst
        measured += t_scan + t_rl
        sample += (f"; fir scan: {rows_scan}x{H}x{N} slab x {len(bands)} bands (scipy pocketfft c128, {cores} workers); "
                   f"RL: {rl_iters} iteration(s) per band on the full {W}x{H} image, 1 thread per band, "
                   "scaled by n_iter, slowest band counts (bands run in parallel in the reference)")
    return {"value": P_total / full_s, "full_cube_seconds_extrapolated": full_s, "measured_seconds": measured,
            "cores": cores, "parts": parts, "sample": sample}


def run_reference(a):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    res = None
    for _ in range(min(a.warmup, 1)):
        cpu_reference_sample(a, rows_chain=1, rows_scan=1, rl_iters=1, with_deconv=False)
    vals, secs = [], []
    t0 = time.perf_counter()
    for _ in range(a.steps):
        res = cpu_reference_sample(a, with_deconv=not a.no_deconv)
        vals.append(res["value"])
        secs.append(res["measured_seconds"])
    total = time.perf_counter() - t0
    value = float(np.mean(vals))
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": a.gpus, "steps": a.steps,
        "warmup": a.warmup, "ms_per_step": 1e3 * float(np.mean(secs)), "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_name(a), "sample": res["sample"],
                   "note": "restated reference (oracle port; rustc unavailable, realfft/rustfft un-vendored): value = "
                           "traces of the full cube / seconds extrapolated linearly from the bounded sample; "
                           "ms_per_step = measured wall time of the sample"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": res["cores"], "kind": "port", "sample": res["sample"],
                         "parts": res["parts"],
                         "full_cube_seconds_extrapolated": res["full_cube_seconds_extrapolated"]},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "wall_s": total,
    }
    print(json.dumps(line))


# --------------------------------------------------------------------------------------
# our arm
# --------------------------------------------------------------------------------------
def bind_to_gpu_numa(index):
    """Pin this rank to the CPU cores NVML reports as local to its GPU, before any pinned host buffer is
    allocated: first-touch then places the staging memory on the GPU's own NUMA node, which matters once
    several ranks stream 50 GB/s each through the host (the e2e leg at N > 1)."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(index)
        words = (os.cpu_count() + 63) // 64
        mask = pynvml.nvmlDeviceGetCpuAffinity(h, words)
        cpus = {64 * w + b for w, word in enumerate(mask) for b in range(64) if (word >> b) & 1}
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
            return len(cpus)
    except Exception:
        pass
    return 0


def rl_flops(bands, W, H):
    """FLOPs of the separable Richardson-Lucy as timed: per iteration two 2-D filterings of the reflect-padded image,
    each a column pass (ky taps) and a row pass (kx taps), one multiply-add per tap and pixel, plus the division
    and the product of the update (4 FLOP per pixel)."""
    tot = 0.0
    for b in bands:
        hp, wp = W + 2 * (b.kx // 2), H + 2 * (b.ky // 2)
        tot += b.n_iter * (4.0 * (b.kx + b.ky) + 4.0) * hp * wp
    return tot


def latest_traffic():
    """ncu DRAM bytes per launch of the cube kernels, newest committed capture (profiles/r*_traffic.json)."""
    import glob
    files = sorted(glob.glob(os.path.join(ROOT, "profiles", "r*_traffic.json")))
    if not files:
        return None, None
    try:
        return json.load(open(files[-1])), os.path.basename(files[-1])
    except Exception:
        return None, None


def run_ours(a):
    import torch
    m = importlib.import_module(PKG)
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    numa_cpus = bind_to_gpu_numa(local) if world > 1 else 0
    dev = torch.device("cuda", local)
    torch.cuda.set_device(local)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
    else:
        dist = None

    W, H, N = a.width, a.height, a.samples
    row0, row1 = m.sharding.slab_bounds(W, world, rank)   # row slabs over axis 0 (x), uneven last slab allowed
    rows = row1 - row0
    P = rows * H
    P_total = W * H

    ctx = m.Context(local)
    stream = torch.cuda.ExternalStream(ctx.stream, device=dev)
    t_axis = (np.float32(1000.0) + np.float32(0.05) * np.arange(N, dtype=np.float32)).astype(np.float32)
    m_pre, band, m_post = m.host.chain_multipliers(t_axis)
    ctx.plan_trace(N, m_pre, band, m_post)
    F = N // 2 + 1

    cube_bytes = P * N * 4
    d_in = ctx.alloc(max(cube_bytes, 16))
    d_out = ctx.alloc(max(cube_bytes, 16))
    p_max = ((W + world - 1) // world) * H                            # same size on every rank (uneven slabs)
    img_t = torch.zeros(max(p_max, 4), dtype=torch.float32, device=dev)   # intensity map of the slab, gathered in the step
    ctx.generate_cube(d_in, rows, H, N, row0=row0, total_width=W)
    ctx.sync()

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    bands = None
    if not a.no_deconv:
        psf = m.host.PSF.load(os.path.join(ROOT, "tests", "golden", "psf.npz"))
        dec = m.host.Deconvolution(n_filters=a.bands, n_iterations=a.rl_iterations)
        bands, why = dec.plan(t_axis, (W, H), 0.5, 0.5, psf)
        if bands is None:
            raise SystemExit(f"deconvolution plan refused: {why}")
    n_rl_iter = sum(b.n_iter for b in bands) if bands is not None else 0
    B = len(bands) if bands is not None else 0

    # config 2: spectra materialised (fft, amplitudes, phases) + maps normalised by the reference pulse
    spec = None
    if a.spectra:
        spec = {k: torch.empty((P, F) if k != "fft" else (P, 2 * F), dtype=torch.float32, device=dev)
                for k in ("fft", "amp", "phase")}

    # several GPUs: the band images stay sharded; Richardson-Lucy exchanges halo rows over NVLink (thz_slab_*)
    slab = slabx = None
    rl_mode = "single GPU"
    e_t = g_t = None
    if bands is not None and world > 1:
        e_t = torch.empty((B, P), dtype=torch.float32, device=dev)
        g_t = torch.empty((B, P), dtype=torch.float32, device=dev)
        rl_mode = "band-parallel (gather, LPT, all-reduce)"
        if a.rl == "slab":
            slab = m.Slab(ctx, rank, world)
            slabx = m.sharding.SlabExchange(slab, dist, rank, world, sync=ctx.sync)
            try:
                slabx.plan(W, H, bands)
                rl_mode = "row slabs, halo rows pushed over NVLink by the filtering kernels"
            except m.ThzError as ex:      # slabs thinner than three PSF half-heights: every rank takes this branch
                slab.close()
                slab = slabx = None
                rl_mode += f" [slab form refused: {ex}]"

    class GpuOps:
        """band-parallel fall-back: libthzgpu stage calls on torch device tensors"""

        def __init__(self):
            self.taps = [(np.ascontiguousarray(b.psf_x_np()), np.ascontiguousarray(b.psf_y_np())) for b in bands]

        def energies(self, slab_ptr):
            ctx.deconv_energies_dev(slab_ptr, P, N, bands, e_t.data_ptr())
            ctx.sync()
            return e_t

        def rl_gain(self, b, image):
            image = image.contiguous()
            g = torch.empty(W * H, dtype=torch.float32, device=dev)
            torch.cuda.synchronize()
            px, py = self.taps[b]
            ctx._check(m.lib.thz_rl_separable_dev(ctx.handle, image.data_ptr(), W, H, px.ctypes.data, px.size,
                                                  py.ctypes.data, py.size, bands[b].direct, bands[b].n_iter, None,
                                                  g.data_ptr(), None, None, None, 0.0, 0.0))
            ctx.sync()
            return g

        def apply(self, slab_ptr, g_slab):
            torch.cuda.synchronize()
            ctx.deconv_apply_dev(slab_ptr, g_slab.data_ptr(), P, N, bands, slab_ptr, img_t.data_ptr())
            ctx.sync()

    ops = GpuOps() if (bands is not None and world > 1 and slab is None) else None
    band_costs = [b.n_iter * (100 + b.kx + b.ky) for b in bands] if bands is not None else None
    phase_s = {}
    img_full = torch.empty((world, img_t.numel()), dtype=torch.float32, device=dev) if world > 1 else None
    marks = []   # (name, event) of the current step, device-timed phases of this rank

    def mark(name):
        e = torch.cuda.Event(enable_timing=True)
        e.record(stream)
        marks.append((name, e))

    def step():
        marks.clear()
        mark("start")
        if spec is not None:
            ctx.trace_forward_dev(d_in.ptr, d_out.ptr, spec["fft"].data_ptr(), spec["amp"].data_ptr(),
                                  spec["phase"].data_ptr(), P)
            mark("trace_forward_spectra")
            return
        if bands is None or ops is not None:
            ctx.trace_fused_dev(d_in.ptr, d_out.ptr, img_t.data_ptr(), P)
            mark("trace_fused")
        if bands is not None:
            if world == 1:
                # trace pass fused with the band energies -> Richardson-Lucy -> gain application (thz_chain_dev)
                ctx.chain_dev(d_in.ptr, rows, H, N, bands, d_out.ptr, img_t.data_ptr())
                mark("chain_and_deconvolution")
            elif slab is not None:
                ctx.chain_begin_dev(d_in.ptr, d_out.ptr, img_t.data_ptr(), P, N, bands, e_t.data_ptr())
                mark("trace_and_band_energies")
                slab.rl(e_t.data_ptr(), P, g_t.data_ptr())
                mark("richardson_lucy_slab")
                ctx.chain_end_dev(d_out.ptr, g_t.data_ptr(), P, N, bands, d_out.ptr, img_t.data_ptr())
                mark("gain_application")
            else:
                m.sharding.sharded_deconvolution(ops, d_out.ptr, W, H, B, dist, world, rank, timings=phase_s,
                                                 band_costs=band_costs)
        if world > 1:
            # the displayed map: gather of the slabs' intensity rows (NCCL over NVLink, 4 bytes per pixel)
            with torch.cuda.stream(stream):
                dist.all_gather_into_tensor(img_full, img_t)
            mark("gather_map")

    for _ in range(a.warmup):
        step()
    ctx.sync()
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    launches0 = ctx.launches
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(a.steps + 1)]
    barrier()
    ev[0].record(stream)
    phase_s.clear()
    for i in range(a.steps):
        step()
        ev[i + 1].record(stream)
    ctx.sync()
    barrier()
    if slab is not None:
        slab.status()
    launches = ctx.launches - launches0
    total_ms = ev[0].elapsed_time(ev[a.steps])
    clocks = sampler.stop() if rank == 0 else None
    # device-timed phases of the last step on this rank
    phases_ms = {marks[i + 1][0]: marks[i][1].elapsed_time(marks[i + 1][1]) for i in range(len(marks) - 1)}
    if phase_s:
        phases_ms.update({k + "_wall": 1e3 * v / a.steps for k, v in phase_s.items()})
    if dist is not None:
        tt = torch.tensor([total_ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        total_ms = float(tt.item())
    ms_per_step = total_ms / a.steps
    value = P_total / (ms_per_step / 1e3)

    # ---- per-kernel breakdown of one extra step (event pairs inside the library), this rank ----
    peak, peak_src = measured_peaks()
    stages = {}
    fused_chain = (bands is not None and ops is None and N >= 512 and (N & (N - 1)) == 0
                   and os.environ.get("THZ_CHAIN_FUSE", "on") != "off")
    if spec is None and not fused_chain:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 3
        e0.record(stream)
        for _ in range(reps):
            ctx.trace_fused_dev(d_in.ptr, d_out.ptr, img_t.data_ptr(), P)
        e1.record(stream)
        ctx.sync()
        stages["trace_fused"] = {"ms": e0.elapsed_time(e1) / reps, "kernel": f"k_trace_fused<{N}>",
                                 "algorithmic_bytes": (8 * N + 4) * P}
    elif spec is not None:
        stages["trace_forward_spectra"] = {"ms": phases_ms["trace_forward_spectra"], "kernel": f"k_trace_forward<{N}>",
                                           "algorithmic_bytes": (8 * N + 16 * F) * P}
    rl_ms = None
    if bands is not None:
        if world == 1:
            st = ctx.deconv_stage_ms()
            km = ctx.chain_kernel_ms()
            rl_ms = st["rl_ms"]
            stages["stage_totals_ms"] = {"trace_pass_and_band_energies": st["energies_ms"], "richardson_lucy": st["rl_ms"],
                                         "gain_application": st["apply_ms"]}
        else:
            barrier()
            ctx._check(m.lib.thz_kernel_timing_begin(ctx.handle))
            step()
            ms4 = np.zeros(4, np.float32)
            ctx._check(m.lib.thz_kernel_timing_end(ctx.handle, ms4.ctypes.data))
            barrier()
            km = {"energy_spectra_ms": float(ms4[0]), "energy_edges_ms": float(ms4[1]),
                  "apply_edges_ms": float(ms4[2]), "apply_main_ms": float(ms4[3])}
            rl_ms = phases_ms.get("richardson_lucy_slab", phases_ms.get("richardson_lucy_own_bands_wall"))
        # per-kernel algorithmic bytes (SURVEY 8d / DESIGN.md): the spectra pass reads the cube and writes B
        # energies per trace; the edge passes touch 2 x 249 samples per trace; the main gain pass reads and
        # writes the cube, reads B gains and 2 x 249 corrections and writes the intensity
        if fused_chain:
            # one kernel: reads the raw cube, writes the filtered cube, the intensity and B energies per trace
            post_mode = "0" if m_post is None or np.all(m_post == 1) else (
                "1" if np.all(np.delete(m_post, np.r_[0:4, N - 4:N]) == 1) else "2")
            stages["trace_energy_fused"] = {"ms": km["energy_spectra_ms"], "kernel": f"k_chain_energy_fused<{N},{post_mode}>",
                                            "algorithmic_bytes": (8 * N + 4 + 4 * B) * P}
        else:
            stages["deconv_energy_spectra"] = {"ms": km["energy_spectra_ms"], "kernel": f"k_fir_energy_split<{N}>",
                                               "algorithmic_bytes": (4 * N + 4 * B) * P}
        edge_kernel = ("k_fir_edges_mma (tcgen05 kind::tf32)" if (N >= 2048 and os.environ.get("THZ_EDGE_MMA", "on") != "off")
                       else "k_fir_edges")
        stages["deconv_energy_edges"] = {"ms": km["energy_edges_ms"], "kernel": edge_kernel,
                                         "algorithmic_bytes": (4 * 498 + 8 * B) * P}
        stages["deconv_apply_edges"] = {"ms": km["apply_edges_ms"], "kernel": "k_fir_edge_corr",
                                        "algorithmic_bytes": (4 * 498 + 4 * B + 4 * 498) * P}
        spectral = fused_chain and os.environ.get("THZ_CHAIN_SPECTRAL", "on") != "off" and P % 2 == 0
        stages["deconv_apply"] = {"ms": km["apply_main_ms"],
                                  "kernel": f"k_fir_apply_circ<{N}>" + (" (spectral hand-off: no forward transform)" if spectral else ""),
                                  "algorithmic_bytes": (8 * N + 4 * B + 4 * 498 + 4) * P}
    for v in stages.values():
        if "algorithmic_bytes" in v and v["ms"] > 0:
            v["gbs"] = v["algorithmic_bytes"] / (v["ms"] / 1e3) / 1e9
            v["frac_of_hbm_peak"] = v["gbs"] / peak
    # FP32 issue-rate microbenchmark in the same run, and the Richardson-Lucy FLOP fraction against it
    fp32 = {"ffma_tflops": 2 * ctx.fp32_rate(0) / 1e12, "ffma2_packed_tflops": 2 * ctx.fp32_rate(1) / 1e12,
            "fadd_tops": ctx.fp32_rate(2) / 1e12, "fadd2_packed_tops": ctx.fp32_rate(3) / 1e12,
            "fmul_tops": ctx.fp32_rate(4) / 1e12, "fmul2_packed_tops": ctx.fp32_rate(5) / 1e12,
            "how": "thz_fp32_rate: 16 independent chains per thread, 8 x 256 threads per SM, register operands"}
    if bands is not None and rl_ms:
        fl = rl_flops(bands, W, H) / world          # this rank's share
        tf = fl / (rl_ms / 1e3) / 1e12
        stages["richardson_lucy"] = {
            "ms": rl_ms, "kernel": "k_rl_multi<1|2> (batched over the bands; k_rl_stream band after band)", "iterations": n_rl_iter,
            "iters_per_s": n_rl_iter / (rl_ms / 1e3), "algorithm": "separable (row + column pass per filtering), f32",
            "tflops": tf, "flop_frac": tf / fp32["ffma_tflops"] if fp32["ffma_tflops"] > 0 else None,
            "flop_frac_of": "measured scalar FFMA rate of this GPU (fp32_peak.ffma_tflops); tensor cores not applicable",
            "mode": rl_mode}
    dom = max((k for k in stages if "gbs" in stages[k]), key=lambda k: stages[k]["ms"])
    traffic, traffic_src = None, None
    tj, tname = latest_traffic()
    if tj is not None and (W, H, N) == tuple(tj.get("cube", (0, 0, 0))):
        ent = tj["kernels"].get(stages[dom]["kernel"].split("<")[0])
        if ent:
            # captured at one GPU on the full cube; bytes scale with the traces of this rank
            traffic = ent["traffic_bytes_per_launch"] * P / float(P_total)
            traffic_src = tname
    roofline = {"bound": "hbm", "kernel": stages[dom]["kernel"], "stage": dom, "achieved": stages[dom]["gbs"],
                "peak": peak, "unit": "GB/s", "frac": stages[dom]["gbs"] / peak, "traffic": traffic,
                "traffic_source": traffic_src, "peak_source": peak_src,
                "algorithmic_bytes_per_launch": stages[dom]["algorithmic_bytes"], "kernel_ms": stages[dom]["ms"],
                "rank": 0}
    # whole-step roofline: 20N + 8B + 4 bytes per trace for the three cube passes
    per_trace = (8 * N + 16 * F) if spec is not None else ((20 * N + 8 * B + 4) if bands is not None else (8 * N + 4))
    chain_bytes = per_trace * P_total
    chain = {"algorithmic_bytes_per_step": chain_bytes, "gbs": chain_bytes / (ms_per_step / 1e3) / 1e9,
             "frac_of_hbm_peak": chain_bytes / (ms_per_step / 1e3) / 1e9 / (peak * world),
             "note": "bytes of the three cube passes of SURVEY 8d (20N + 8B + 4 per trace)" +
                     ("; the fused trace + energy kernel moves 16N + 8B + 8 per trace" if fused_chain else "")}

    e2e = None
    if not a.no_e2e and spec is None:
        e2e = run_e2e(a, m, ctx, d_in, d_out, P, rows, N, world, dist, barrier, P_total, bands, slab)

    cpu = None
    if rank == 0 and not a.no_cpu:
        res = cpu_reference_sample(a, rows_chain=a.cpu_rows or 4, with_deconv=bands is not None)
        cpu = {"value": res["value"], "unit": UNIT, "cores": res["cores"], "kind": "port", "sample": res["sample"],
               "parts": res["parts"], "full_cube_seconds_extrapolated": res["full_cube_seconds_extrapolated"],
               "measured_seconds": res["measured_seconds"]}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": a.steps, "warmup": a.warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": workload_name(a), "name": a.config,
                       "partition": f"row slabs over x, {world} rank(s)", "richardson_lucy": rl_mode,
                       "host_affinity": (f"rank pinned to the {numa_cpus} cores local to its GPU (NVML)"
                                         if numa_cpus else "default"),
                       "l2": ("inputs larger than L2 (cube >> 126 MB), no flush needed" if cube_bytes > (256 << 20)
                              else "cube smaller than 2 x L2: numbers include L2 hits, parity case rather than a "
                                   "bandwidth measurement"),
                       "stages": (["trace pass fused with the band-energy pass of the deconvolution"] if fused_chain
                                  else ["trace pass (fused)"]) +
                                 ([("deconvolution: " if fused_chain else "deconvolution: band energies, ") +
                                   f"Richardson-Lucy ({n_rl_iter} iterations over {len(bands)} bands), "
                                   "gain application"] if bands is not None else [])},
            "stage_breakdown": stages, "rank0_phases_ms": phases_ms, "chain_roofline": chain,
            "roofline": roofline, "fp32_peak": fp32, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": int(launches),
            "clocks": clocks,
        }
        print(json.dumps(line))
    if slab is not None:
        barrier()
        slab.close()
    d_in.free(); d_out.free()
    ctx.close()
    if dist is not None:
        dist.destroy_process_group()


def run_e2e(a, m, ctx, d_in, d_out, P, rows, N, world, dist, barrier, P_total, bands, slab):
    """The same step through the host-pointer C ABI: pinned host cube in, filtered (and deconvolved)
    cube + intensity map out, copies inside the timed region.  One GPU: thz_chain_host.  Several GPUs: every rank
    runs thz_chain_host_begin (H2D chunks under the trace / energy passes), the halo-exchanged Richardson-Lucy,
    and thz_chain_host_end (gain application under the D2H chunks) on its slab."""
    import ctypes as C
    H = a.height
    try:
        avail = os.sysconf("SC_AVPHYS_PAGES") * os.sysconf("SC_PAGE_SIZE")
    except Exception:
        avail = 64 << 30
    nbytes = P * N * 4
    if nbytes * world > 0.6 * avail:
        return {"value": None, "unit": UNIT, "note": f"host RAM too small for a {nbytes * world >> 30} GiB pinned cube"}
    if world > 1 and bands is not None and slab is None:
        return {"value": None, "unit": UNIT, "note": "band-parallel fall-back has no overlapped host path"}
    hp, hi = C.c_void_p(), C.c_void_p()
    if m.lib.thz_host_alloc(nbytes, C.byref(hp)) != 0 or m.lib.thz_host_alloc(max(P * 4, 16), C.byref(hi)) != 0:
        return None
    B = len(bands) if bands is not None else 0
    try:
        ctx._check(m.lib.thz_copy_d2h(ctx.handle, hp.value, d_in.ptr, nbytes))
        reps = 2
        # the host-pointer calls keep their own device-resident cube: release the bench's output buffer first;
        # d_in stays as the pristine copy from which the (in-place) host buffer is restored, untimed
        d_out.free()

        def once():
            if world == 1:
                ctx._check(m.lib.thz_chain_host(ctx.handle, hp.value, rows, H, N, bands, B, hp.value, hi.value,
                                                None, None, None))
            else:
                de, dg = C.c_void_p(), C.c_void_p()
                ctx._check(m.lib.thz_chain_host_begin(ctx.handle, hp.value, rows, H, N, bands, B, hp.value,
                                                      C.byref(de), C.byref(dg)))
                if B:
                    slab.rl(de.value, P, dg.value)
                    slab.status()
                ctx._check(m.lib.thz_chain_host_end(ctx.handle, rows, H, N, bands, B, hp.value, hi.value))

        once()   # warm-up (allocates the device-resident cube)
        dt = 0.0
        for _ in range(reps):
            ctx._check(m.lib.thz_copy_d2h(ctx.handle, hp.value, d_in.ptr, nbytes))   # restore the input, untimed
            barrier()
            t0 = time.perf_counter()
            once()
            ctx.sync()
            barrier()
            dt += (time.perf_counter() - t0) / reps
        if dist is not None:
            import torch
            tt = torch.tensor([dt], device="cuda", dtype=torch.float64)
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            dt = float(tt.item())
        return {"value": P_total / dt, "unit": UNIT, "h2d_bytes_per_step": int(nbytes * world),
                "d2h_bytes_per_step": int((nbytes + P * 4) * world), "seconds_per_step": dt,
                "note": ("pinned host cube -> H2D chunks overlapped with the fused trace pass and the band-energy pass "
                         "-> Richardson-Lucy -> gain application overlapped with D2H chunks"
                         + ("" if world == 1 else "; per rank on its row slab (thz_chain_host_begin / thz_slab_rl / "
                            "thz_chain_host_end), wall clock, max over ranks"))}
    finally:
        m.lib.thz_host_free(hp.value)
        m.lib.thz_host_free(hi.value)


def main():
    a = parse_args()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_ours(a)


if __name__ == "__main__":
    main()
