#!/usr/bin/env python
"""bench.py -- RLZ timesteps/sec (BASELINE.json metric) on the north-star configuration.

    python bench.py --gpus N --steps K --warmup W            # our arm (one process per GPU under torchrun)
    python bench.py --impl reference --gpus N --steps K ...  # CPU restatement of the reference on host cores

Workload: RLZ, 334 radial cells (rings 8..4012 points), 64 levels (43 modes),
N = 128,897,280 points, V = 3 (LinearAdvectionRLZ), synthetic vortex.  With N GPUs the radial
domain is scaled (num_cells x sqrt(N)) so every GPU owns one C4-sized tile (weak scaling, C5) and
`value` counts C4-equivalent tile-timesteps per second over all ranks.
A "step" = one iteration of model_loop: K3 tileTransform! -> equation set + AB3 -> K1
spectralTransform! -> shared-spectral sum (NCCL all-reduce when N>1) -> K2 splineTransform!.
"""
from __future__ import annotations

import argparse
import json
import math
import os
import subprocess
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

METRIC = "RLZ timesteps/sec"
UNIT = "timesteps/s"
C4_CELLS, ZDIM, NVARS = 334, 64, 3
XMAX, ZMAX = 1.0e6, 2.0e4
TS, KDIFF = 2.0, 500.0


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--cells", type=int, default=0, help="override radial cells per GPU tile (debug)")
    ap.add_argument("--equation-set", default="LinearAdvectionRLZ",
                    choices=["LinearAdvectionRLZ", "Oneway_ShallowWater_HeightResolvedBL"],
                    help="C4 step to time: the transform-dominated linear set (headline) or the 6-variable TC boundary-layer set")
    ap.add_argument("--no-tcbl", action="store_true", help="skip the secondary C4 run of the TC boundary-layer equation set")
    ap.add_argument("--k3-slots", default="fused", choices=["fused", "needed", "all"],
                    help="slots the in-step tileTransform! produces: what the equation-set kernel reads, with the equation set + "
                         "AB3 fused into the last transform stage where such a kernel is built (product default); the same slots "
                         "with separate kernels; or all D slots of every variable (the reference's materialised dataflow).  The "
                         "default run times the first and the last")
    ap.add_argument("--no-materialised", action="store_true", help="skip the secondary all-slots timing (ncu captures)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--tile-of", default="", help="debug, one GPU: 'N:i' = own only tile i of the N-GPU weak-scaling patch (per-kernel "
                                                  "times of an outer tile without N GPUs; the state is not meaningful)")
    ap.add_argument("--no-preflight", action="store_true", help="N > 1: skip the multi-rank parity preflight against the oracle")
    ap.add_argument("--no-strong", action="store_true", help="N > 1: skip the strong-scaling run (C4 itself cut into N tiles)")
    return ap.parse_args()


# ------------------------------------------------------------------ synthetic initial state
def tile_rings(xmin, DX, num_cells, patch_offset):
    g = math.sqrt(3.0 / 5.0)
    c = xmin + (np.arange(num_cells) + 0.5) * DX
    r = (c[:, None] + 0.5 * DX * np.array([-g, 0.0, g])[None, :]).reshape(-1)
    ri = np.arange(1, 3 * num_cells + 1) + patch_offset
    return r, ri


def synthetic_state(xmin, DX, num_cells, patch_offset, zDim, zmax, out=None):
    """Rankine-like vortex + wavenumber-2 asymmetry, exp(-z/H) decay; h Gaussian.  Deterministic.
    Returns [N, 3] Fortran-ordered (h, u, v), N = sum_r (4+4 ri) * zDim with z fastest."""
    r, ri = tile_rings(xmin, DX, num_cells, patch_offset)
    n = 4 + 4 * ri
    rr = np.repeat(r, n)
    lam = np.concatenate([0.5 * (2 * np.pi / k) * (i - 1) + (2 * np.pi / k) * np.arange(k) for k, i in zip(n, ri)])
    z = 0.5 * zmax * (1.0 - np.cos(np.pi * np.arange(zDim) / (zDim - 1)))
    Rmax, Vmax = 5.0e4, 50.0
    vbar = np.where(rr < Rmax, Vmax * rr / Rmax, Vmax * Rmax / rr)
    hh = 100.0 * np.exp(-(rr / 2.0e5) ** 2) * (1.0 + 0.1 * np.cos(2.0 * lam))
    uu = -0.05 * vbar * (1.0 + 0.2 * np.sin(lam))
    vv = vbar * (1.0 + 0.1 * np.cos(2.0 * lam))
    zh = 1.0 + 0.3 * np.cos(np.pi * z / zmax)
    zw = np.exp(-z / 8.0e3)
    N = rr.size * zDim
    if out is None:
        out = np.empty((N, 3), order="F")
    out[:, 0] = (hh[:, None] * zh[None, :]).reshape(-1)
    out[:, 1] = (uu[:, None] * zw[None, :]).reshape(-1)
    out[:, 2] = (vv[:, None] * zw[None, :]).reshape(-1)
    return out


def synthetic_state_tcbl(xmin, DX, num_cells, patch_offset, zDim, zmax):
    """h, ug, vg as synthetic_state; boundary-layer winds ub, vb relax to the gradient wind above ~1 km; wb = 0."""
    base = synthetic_state(xmin, DX, num_cells, patch_offset, zDim, zmax)
    z = 0.5 * zmax * (1.0 - np.cos(np.pi * np.arange(zDim) / (zDim - 1)))
    prof = np.tile(1.0 - np.exp(-(z + 50.0) / 300.0), base.shape[0] // zDim)
    out = np.empty((base.shape[0], 6), order="F")
    out[:, :3] = base
    out[:, 3] = base[:, 1] * prof - 0.2 * base[:, 2] * (1.0 - prof)
    out[:, 4] = base[:, 2] * prof
    out[:, 5] = 0.0
    return out


def dims(num_cells, patch_offset=0, zDim=ZDIM):
    ri = np.arange(1, 3 * num_cells + 1) + patch_offset
    hp = int((4 + 4 * ri).sum())
    W = int((1 + 2 * ri).sum())
    bz = min(zDim, (2 * zDim - 1) // 3 + 1)
    kDim = 3 * num_cells + patch_offset
    S = bz * (num_cells + 3) * (1 + 2 * kDim)
    return dict(N=hp * zDim, hp=hp, W=W, bz=bz, S=S, D=7)


# ------------------------------------------------------------------ clocks
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.path = index, None, f"/tmp/sb_clocks_{os.getpid()}.csv"

    def start(self):
        try:
            self.f = open(self.path, "w")
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=self.f,
                                         stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        self.proc.wait()
        self.f.close()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in open(self.path):
            p = [x.strip() for x in line.split(",")]
            if len(p) < 8:
                continue
            try:
                sm.append(float(p[1])); mx.append(float(p[2]))
            except ValueError:
                continue
            for nm, val in zip(names, p[4:8]):
                if val.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------ CPU arm (oracle = the reference restated)
def sample_case(cells):
    """The bounded sample of the C4 workload the CPU arm runs: the inner `cells` radial cells of the same grid."""
    from oracle import grids as G
    from oracle import model as M
    gp = G.GridParameters(geometry="RLZ", xmin=0.0, xmax=XMAX * cells / C4_CELLS, num_cells=cells, zmin=0.0, zmax=ZMAX,
                          zDim=ZDIM, vars={"h": 1, "u": 2, "v": 3})
    return gp, M


def cpu_baseline(steps=2, warmup=1, cells=64, workers=None):
    """Oracle ModelRun on a bounded sample of the same workload, all host threads, scaled by points.
    Returns (record, seconds per step, final var_np1 of the sample [N, V], steps taken)."""
    workers = workers or os.cpu_count() or 1
    gp, M = sample_case(cells)
    mp = M.ModelParameters(ts=TS, integration_time=TS * (steps + warmup), equation_set="LinearAdvectionRLZ",
                           grid_params=gp, physical_params={"K": KDIFF})
    ic = synthetic_state(0.0, XMAX / C4_CELLS, cells, 0, ZDIM, ZMAX)
    run = M.ModelRun(mp, 1, ic, workers=workers)
    run.run(warmup)
    t0 = time.perf_counter()
    run.run(steps)
    dt = (time.perf_counter() - t0) / steps
    n_s, n_full = dims(cells)["N"], dims(C4_CELLS)["N"]
    value = (1.0 / dt) * n_s / n_full
    sample = (f"oracle (NumPy/SciPy pocketfft restatement of the reference's algorithm, NOT Scythe.jl itself: Julia is not in "
              f"the image) ModelRun, RLZ {cells} cells x {ZDIM} levels x {NVARS} vars = {n_s} points, {steps} timed steps "
              f"({dt * 1e3:.0f} ms/step), EXTRAPOLATED by points to the {n_full}-point C4 grid")
    rec = {"value": value, "unit": UNIT, "cores": workers, "kind": "port", "extrapolated": True,
           "label": "NumPy restatement, extrapolated from a 1/27 sample by points", "sample": sample,
           "sample_points": n_s, "sample_ms_per_step": dt * 1e3}
    return rec, dt, np.array(run.mtiles[0].var_np1), steps + warmup


def gpu_sample_parity(ostate, nsteps, cells=64, device=0):
    """The GPU path on the CPU arm's sample (same grid, same synthetic state, same number of steps, default fused
    dataflow) against the oracle's final state: the parity check on the benchmarked level count and kernels."""
    import scythe_jl_b200 as S
    DX = XMAX / C4_CELLS
    gp = S.GridParameters(geometry="RLZ", xmin=0.0, xmax=DX * cells, num_cells=cells, zmin=0.0, zmax=ZMAX, zDim=ZDIM,
                          vars={"h": 1, "u": 2, "v": 3})
    mp = S.ModelParameters(ts=TS, integration_time=TS * nsteps, equation_set="LinearAdvectionRLZ", grid_params=gp,
                           physical_params={"K": KDIFF})
    m = S.Model(mp, num_tiles=1, device=device)
    m.initialize_tiles([synthetic_state(0.0, DX, cells, 0, ZDIM, ZMAX)])
    m.run(nsteps)
    got = m.state(0, "var_np1")
    m.close()
    scale = np.abs(ostate).max(axis=0)
    err = float((np.abs(got - ostate).max(axis=0) / np.where(scale > 0, scale, 1.0)).max())
    return {"rel_err": err, "tol": 1e-9, "ok": bool(err <= 1e-9), "steps": int(nsteps), "points": int(ostate.shape[0]),
            "what": f"GPU (default fused K3+K4 step) vs oracle on the cpu_baseline sample: RLZ {cells} cells x {ZDIM} levels, "
                    "max over variables of max|d var_np1| / max|var_np1|"}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cb, dt, _, _ = cpu_baseline(steps=max(min(args.steps, 3), 1), warmup=1)
    line = {"impl": "reference", "metric": METRIC, "value": cb["value"] * 1.0, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 / cb["value"] if cb["value"] else None,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workload_config(args.gpus, C4_CELLS), "cpu_baseline": cb, "extrapolated": True,
            "e2e": {"value": cb["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))


def kernel_source_hash():
    """sha256 over the kernel sources: profiles/step_kernels_*.json carry the hash they were captured at
    (profiles/make_step_table.py), and `roofline.traffic` is only printed when it matches."""
    import hashlib
    h = hashlib.sha256()
    for f in sorted((ROOT / "scythe_jl_b200" / "csrc").glob("sb_*")):
        h.update(f.name.encode())
        h.update(f.read_bytes())
    return h.hexdigest()[:16]


def workload_config(ngpus, cells_per_tile, eq="LinearAdvectionRLZ", nvars=3):
    total_cells = int(round(cells_per_tile * math.sqrt(ngpus)))
    return {"workload": f"C4 RLZ {eq}" if ngpus == 1 else "C5 RLZ radius-scaled, one C4-sized tile per GPU",
            "geometry": "RLZ", "num_cells": total_cells, "zDim": ZDIM, "b_zDim": 43, "vars": nvars,
            "equation_set": eq, "tiles": ngpus, "ts": TS,
            "exchange": ("none (one tile)" if ngpus == 1 else
                         "z-mode planes of the spline solve dealt over ranks; kernels store into the owners' buffers over "
                         "CUDA-IPC peer mappings (NVLink), two one-element rendezvous per step, no data collective"),
            "l2": "working set >> 126 MB L2 (physical 21.7 GB/GPU); no flush needed",
            "units": "C4-equivalent (128.9 M-point) tile-timesteps, summed over ranks"}


# ------------------------------------------------------------------ multi-rank parity preflight (N > 1)
def multi_rank_preflight(S, dist, rank, world, local_rank, cells=36, nsteps=3):
    """N ranks, one radial tile each, the DEFAULT exchange (CUDA-IPC peer stores, `columns-p2p`), default fused step,
    64 levels: every rank's tile state after `nsteps` against the oracle integrated with the same N tiles
    (reference: /root/reference/src/semiimplicit.jl:320-329, 279-285).  Returns the record rank 0 prints."""
    import torch
    from oracle import grids as G
    from oracle import model as M
    cells = max(cells, 4 * world)
    DX = XMAX / C4_CELLS
    ogp = G.GridParameters(geometry="RLZ", xmin=0.0, xmax=DX * cells, num_cells=cells, zmin=0.0, zmax=ZMAX, zDim=ZDIM,
                           vars={"h": 1, "u": 2, "v": 3})
    ic = synthetic_state(0.0, DX, cells, 0, ZDIM, ZMAX)
    gp = S.GridParameters(geometry="RLZ", xmin=0.0, xmax=DX * cells, num_cells=cells, zmin=0.0, zmax=ZMAX, zDim=ZDIM,
                          vars={"h": 1, "u": 2, "v": 3})
    mp = S.ModelParameters(ts=TS, integration_time=TS * nsteps, equation_set="LinearAdvectionRLZ", grid_params=gp,
                           physical_params={"K": KDIFF})
    m = S.Model(mp, num_tiles=world, device=local_rank, distributed=True)
    exchange = m.exchange
    tp = m.tile_params
    pts = np.concatenate([[0], np.cumsum(tp[4]).astype(np.int64)])
    m.initialize_tiles([ic[pts[rank]:pts[rank + 1]]])
    m.run(nsteps)
    got = m.state(0, "var_np1")
    m.close()
    omp = M.ModelParameters(ts=TS, integration_time=TS * nsteps, equation_set="LinearAdvectionRLZ", grid_params=ogp,
                            physical_params={"K": KDIFF})
    orun = M.ModelRun(omp, world, ic, workers=max(1, (os.cpu_count() or 8) // world))
    orun.run(nsteps)
    want = np.array(orun.mtiles[rank].var_np1)
    scale = np.abs(want).max(axis=0)
    err = float((np.abs(got - want).max(axis=0) / np.where(scale > 0, scale, 1.0)).max())
    t = torch.tensor([err], device="cuda", dtype=torch.float64)
    allerr = [torch.zeros_like(t) for _ in range(world)]
    dist.all_gather(allerr, t)
    errs = [float(x.item()) for x in allerr]
    return {"ranks": world, "rel_err_max": max(errs), "rel_err_per_rank": errs, "tol": 1e-9, "ok": bool(max(errs) <= 1e-9),
            "exchange": exchange, "steps": nsteps, "tile_cells": [int(c) for c in tp[2]],
            "what": f"RLZ {cells} cells x {ZDIM} levels cut into {world} tiles, one per rank, default exchange and fused "
                    "step: each rank's var_np1 vs the oracle's tile of the same N-tile integration"}


def pcie_ceiling(torch, nbytes=1 << 30):
    """This rank's concurrent H2D + D2H rate with plain pinned copies on two streams (the denominator of e2e)."""
    h_in = torch.empty(nbytes, dtype=torch.uint8, pin_memory=True)
    h_out = torch.empty(nbytes, dtype=torch.uint8, pin_memory=True)
    d_in = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
    d_out = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    res = {}
    for name, both in (("h2d_alone", False), ("both", True)):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        for rep in range(2):          # second repetition is the measurement
            with torch.cuda.stream(s1):
                e0.record(); d_in.copy_(h_in, non_blocking=True); e1.record()
            if both:
                with torch.cuda.stream(s2):
                    f0.record(); h_out.copy_(d_out, non_blocking=True); f1.record()
            torch.cuda.synchronize()
        res[name] = {"h2d_GBps": nbytes / 1e6 / e0.elapsed_time(e1)}
        if both:
            res[name]["d2h_GBps"] = nbytes / 1e6 / f0.elapsed_time(f1)
    del h_in, h_out, d_in, d_out
    return res


# ------------------------------------------------------------------ our arm
def run_ours(args):
    import torch

    import scythe_jl_b200 as S

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    distributed = world > 1
    assert torch.cuda.is_available(), "bench.py needs a CUDA device (no CPU fallback)"
    torch.cuda.set_device(local_rank)
    if distributed:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    preflight = None
    if distributed and not args.no_preflight:
        preflight = multi_rank_preflight(S, dist, rank, world, local_rank)
    ntiles = world
    tile_of = None
    if args.tile_of:
        assert not distributed
        ntiles, ti = (int(x) for x in args.tile_of.split(":"))
        tile_of = ti
    cells_tile = args.cells or C4_CELLS
    total_cells = int(round(cells_tile * math.sqrt(ntiles)))
    DX = XMAX / C4_CELLS
    tcbl = args.equation_set != "LinearAdvectionRLZ"
    global NVARS
    if tcbl:
        NVARS = 6
        names = ["h", "u", "v", "ub", "vb", "wb"]
        gp = S.GridParameters(geometry="RLZ", xmin=0.0, xmax=DX * total_cells, num_cells=total_cells, zmin=0.0, zmax=ZMAX,
                              zDim=ZDIM, vars={n: i + 1 for i, n in enumerate(names)},
                              BCL={"h": S.CubicBSpline.R1T1, "u": S.CubicBSpline.R1T0, "v": S.CubicBSpline.R1T0,
                                   "ub": S.CubicBSpline.R1T0, "vb": S.CubicBSpline.R1T0, "wb": S.CubicBSpline.R1T1})
        mp = S.ModelParameters(ts=0.5, integration_time=500.0, equation_set=args.equation_set, grid_params=gp,
                               physical_params=dict(g=9.81, Kh=1500.0, Cd=2.4e-3, Hfree=2000.0, f=5e-5, Um=3.0, Vm=-2.0))
    else:
        gp = S.GridParameters(geometry="RLZ", xmin=0.0, xmax=DX * total_cells, num_cells=total_cells, zmin=0.0, zmax=ZMAX,
                              zDim=ZDIM, vars={"h": 1, "u": 2, "v": 3})
        mp = S.ModelParameters(ts=TS, integration_time=TS * 1000, equation_set="LinearAdvectionRLZ", grid_params=gp,
                               physical_params={"K": KDIFF})
    if tile_of is not None:
        m = S.Model(mp, num_tiles=ntiles, tile_first=tile_of, tile_count=1, device=local_rank)
    else:
        m = S.Model(mp, num_tiles=ntiles, device=local_rank, distributed=distributed)
    cfg_exchange = m.exchange
    # spline coefficients this rank SOLVES per variable: the whole patch when the solve is replicated (one tile, or the
    # reference's all-reduce scheme), its share of the z-mode planes when the columns are dealt over the ranks
    Sp = m.patch.S / world if (m.columns and world > 1) else m.patch.S
    tp = m.tile_params
    tcells, tsil = int(tp[2, m.tile_first]), int(tp[3, m.tile_first])
    ic = (synthetic_state_tcbl if tcbl else synthetic_state)(tp[0, m.tile_first], DX, tcells, (tsil - 1) * 3, ZDIM, ZMAX)
    m.initialize_tiles([ic])
    m.sync()
    tile = m.tiles[0]
    d = dict(N=tile.N, S=tile.S, D=tile.D)
    lib = m.lib
    import ctypes as C

    def barrier():
        torch.cuda.synchronize()
        if distributed:
            dist.barrier()
            torch.cuda.synchronize()

    def timed(fn, n, tail=None, model=None):
        model = model or m
        barrier()
        lib.check(lib.sb_timer_start(model.patch.handle))
        for _ in range(n):
            fn()
        if tail:
            tail()           # stream-ordered only: makes the timed stream wait for work issued on the copy streams
        ms = C.c_float()
        lib.check(lib.sb_timer_stop(model.patch.handle, C.byref(ms)))
        barrier()
        t = torch.tensor([ms.value], device="cuda")
        if distributed:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # warm-up, then K timed steps with per-kernel event profiling on
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()              # nvidia-smi needs ~0.3 s to produce its first sample: start it before the warm-up
    m.set_k3_slots(args.k3_slots)
    for _ in range(args.warmup):
        m.step()
    l0 = m.launch_count()
    m.profile(True)
    total_ms = timed(m.step, args.steps)
    m.profile(False)
    clocks = sampler.stop() if rank == 0 else None
    prof = m.profile_report()
    launches = m.launch_count() - l0
    ms_step = total_ms / args.steps
    value = ntiles * 1e3 / ms_step
    # the same step with tileTransform! producing all D slots of every variable (the reference's dataflow)
    mat = None
    if args.k3_slots != "all" and not args.no_materialised:
        m.set_k3_slots("all")
        nmat = max(3, min(args.steps, 10))
        for _ in range(2):
            m.step()
        m.profile(True)
        mat_ms = timed(m.step, nmat) / nmat
        m.profile(False)
        mat = {"ms_per_step": mat_ms, "value": ntiles * 1e3 / mat_ms, "steps": nmat, "prof": m.profile_report()}
        m.set_k3_slots(args.k3_slots)

    # transforms/sec (the second half of the BASELINE metric): K1 alone and K2+K3 alone on the tile
    def k1():
        lib.check(lib.sb_spectral_transform(tile.handle))

    def k23():
        lib.check(lib.sb_spline_transform(m.patch.handle, m.patch.handle))
        lib.check(lib.sb_tile_transform(m.patch.handle, tile.handle))
    k1(); k23()
    k1_ms = timed(k1, 3) / 3
    lib.check(lib.sb_model_spline_transform(m.handle))
    k23_ms = timed(k23, 3) / 3

    # end to end through the public API with HOST buffers: pinned state in -> one model_loop cycle -> state out
    e2e = None
    if not args.no_e2e:
        hin = torch.empty((NVARS, tile.N), dtype=torch.float64, pin_memory=True)
        hout = torch.empty((NVARS, tile.N), dtype=torch.float64, pin_memory=True)
        a_in, a_out = hin.numpy().T, hout.numpy().T      # [N, V] Fortran views
        m.get_state_into(0, a_in)

        def e2e_step():
            m.set_state(0, a_in)
            m.cycle()
            m.get_state_into(0, a_out)
        e2e_step()
        ns = max(2, min(args.steps, 3))
        s_ms = timed(e2e_step, ns) / ns
        # the same three operations per step through the copy streams (Model.cycle_host: sb_model_stage_in / cycle /
        # sb_model_stage_out): step i+1's H2D, step i's kernels and step i-1's D2H overlap; pipeline fill and drain
        # are inside the timed region
        def e2e_pipe():
            m.cycle_host([a_in], [a_out])
        npipe = max(4, min(args.steps, 20))
        try:
            e2e_pipe()
            m.drain()
            p_ms = timed(e2e_pipe, npipe, tail=lambda: m.drain(block=False)) / npipe
            m.drain()
            pipe_note = None
        except S.ScytheError as exc:       # (e.g. no room for the 4 staging buffers) report the blocking form, say so
            p_ms, npipe, pipe_note = s_ms, ns, "pipelined form unavailable: " + str(exc)[:160]
        e2e = {"value": ntiles * 1e3 / p_ms, "unit": UNIT, "h2d_bytes_per_step": int(a_in.nbytes) * ntiles,
               "d2h_bytes_per_step": int(a_out.nbytes) * ntiles, "steps": npipe, "ms_per_step": p_ms,
               "what": "per rank and per step: Model.cycle_host = stage_in(host pinned [N_tile,V]) -> cycle() -> "
                       "stage_out(host pinned [N_tile,V]); every step copies its whole state in and out, copies of "
                       "neighbouring steps overlap the kernels on two copy streams (fill + drain inside the timed "
                       "region); bytes summed over ranks",
               "serial": {"value": ntiles * 1e3 / s_ms, "ms_per_step": s_ms, "steps": ns,
                          "what": "the same step with blocking copies: set_state -> cycle -> get_state, nothing overlapped"}}
        if pipe_note:
            e2e["note"] = pipe_note
        del hin, hout, a_in, a_out
        barrier()
        ceil = pcie_ceiling(torch)      # every rank at the same time: what the box gives this rank while all ranks copy
        ct = torch.tensor([ceil["h2d_alone"]["h2d_GBps"], ceil["both"]["h2d_GBps"], ceil["both"]["d2h_GBps"]], device="cuda")
        if distributed:
            dist.all_reduce(ct, op=dist.ReduceOp.SUM)
        agg = [float(x) for x in ct.tolist()]
        floor_ms = max(e2e["h2d_bytes_per_step"] / (agg[1] * 1e6), e2e["d2h_bytes_per_step"] / (agg[2] * 1e6))
        e2e["host_link_ceiling"] = {
            "h2d_alone_GBps_sum_over_ranks": agg[0], "h2d_concurrent_GBps_sum_over_ranks": agg[1],
            "d2h_concurrent_GBps_sum_over_ranks": agg[2], "copy_floor_ms_per_step": floor_ms,
            "frac_of_copy_floor": floor_ms / p_ms,
            "what": "1 GiB pinned cudaMemcpyAsync H2D alone, then H2D and D2H concurrently on two streams, all ranks at once; "
                    "copy_floor = this step's bytes at the concurrent rates (full duplex): e2e cannot beat it on this box"}

    per_rank = None
    strong = None
    if distributed:   # every rank's per-kernel time: the step is as slow as the slowest tile
        mine = {k: round(v["ms"] / args.steps, 3) for k, v in prof.items()}
        per_rank = [None] * world
        dist.all_gather_object(per_rank, mine)
        if not args.no_strong and not tcbl and not args.cells:
            # strong scaling: C4 ITSELF cut into N equal-gridpoint tiles, one per rank
            # (/root/reference/src/semiimplicit.jl:141-169): true timesteps/s of the 128.9 M-point problem
            exch = m.exchange
            m.close()
            gps = S.GridParameters(geometry="RLZ", xmin=0.0, xmax=XMAX, num_cells=C4_CELLS, zmin=0.0, zmax=ZMAX,
                                   zDim=ZDIM, vars={"h": 1, "u": 2, "v": 3})
            mps = S.ModelParameters(ts=TS, integration_time=TS * 1000, equation_set="LinearAdvectionRLZ", grid_params=gps,
                                    physical_params={"K": KDIFF})
            ms_ = S.Model(mps, num_tiles=world, device=local_rank, distributed=True)
            tps = ms_.tile_params
            sc, ssil = int(tps[2, ms_.tile_first]), int(tps[3, ms_.tile_first])
            ms_.initialize_tiles([synthetic_state(tps[0, ms_.tile_first], XMAX / C4_CELLS, sc, (ssil - 1) * 3, ZDIM, ZMAX)])
            ms_.set_k3_slots(args.k3_slots)
            for _ in range(max(args.warmup, 3)):
                ms_.step()
            nst = max(3, min(args.steps, 10))
            ms_.profile(True)
            st_ms = timed(ms_.step, nst, model=ms_) / nst
            ms_.profile(False)
            sp = {k: round(v["ms"] / nst, 3) for k, v in ms_.profile_report().items()}
            sp_all = [None] * world
            dist.all_gather_object(sp_all, sp)
            strong = {"what": "strong scaling: the C4 problem itself (334 cells, 128,897,280 points) cut into N equal-gridpoint "
                              "radial tiles, one per rank; value = true timesteps/s of that problem (compare with the N=1 line)",
                      "value": 1e3 / st_ms, "unit": UNIT, "ms_per_step": st_ms, "steps": nst, "tile_cells": [int(c) for c in tps[2]],
                      "tile_points": [int(c) * ZDIM for c in tps[4]], "per_rank_kernel_ms": sp_all, "exchange": ms_.exchange}
            ms_.close()
            m = None
    if rank != 0:
        if distributed:
            dist.destroy_process_group()
        return

    # ---- roofline: K3 (tileTransform!) dominates; algorithmic bytes per SURVEY 8(d)
    peaks_path = ROOT / "MEASURED_PEAKS.json"
    if peaks_path.exists():
        peak, peak_src = float(json.loads(peaks_path.read_text())["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs (of measured)"
    else:
        peak, peak_src = 6650.0, "B200_PROFILING.md fallback (of fallback)"
    N, Sg, D, V = d["N"], d["S"], d["D"], NVARS
    # (variable, slot) pairs the equation-set kernel reads / variables it reads at all (sb_model.cu: equation_set_needs)
    ns_needed, v_read = (21, 5) if tcbl else (7, 3)

    def tables(prof_, steps_, ns, vr, fused=False):
        """per-K-group time and HBM-roofline fraction on ALGORITHMIC bytes; ns = slots K3 writes and K4
        reads (all: V*D), vr = variables K3 transforms; fused: K3's last stage and K4 are one kernel (no slot traffic)"""
        k4_hist = 4 * V + (1 if tcbl else 0)              # K4 kernel: read 2 history arrays, write var_np1 + expdot_n (+ wb)
        groups = {
            "K3 tileTransform! (inv_r+inv_l+inv_z)": (["inv_r", "inv_l", "inv_z"], 8.0 * (vr * Sg + N * ns)),
            "K1 spectralTransform! (fwd_z+fwd_l+fwd_r)": (["fwd_z", "fwd_l", "fwd_r"], 8.0 * V * (N + Sg)),
            "K2 splineTransform! (spline_solve)": (["spline_solve"], 8.0 * V * 2 * Sp),
            f"K4 equation set + AB3 ({args.equation_set})": (["equation_set"], 8.0 * N * (ns_needed + k4_hist)),
        }
        if fused:     # read A and 2 history arrays, write var_np1 + expdot_n: nothing else is algorithmically required
            del groups["K3 tileTransform! (inv_r+inv_l+inv_z)"], groups[f"K4 equation set + AB3 ({args.equation_set})"]
            groups[f"K3+K4 tileTransform! + {args.equation_set} + AB3 (inv_r+inv_l+inv_z_k4, fused)"] = (
                ["inv_r", "inv_l", "inv_z_k4"], 8.0 * (vr * Sg + N * k4_hist))      # every variable with its history
        det = {}
        for name, (ks, nbytes) in groups.items():
            ms = sum(prof_.get(k, {"ms": 0.0})["ms"] for k in ks) / steps_
            det[name] = {"ms_per_step": ms, "algorithmic_GB": nbytes / 1e9,
                         "achieved_GBps": (nbytes / 1e9) / (ms / 1e3) if ms > 0 else None,
                         "frac": ((nbytes / 1e9) / (ms / 1e3) / peak) if ms > 0 else None}
        # timestep, SURVEY 8(d): K3 (S + N slots) + K4 (N slots + 5N) + K1 (N + S) + K2 (2S); all slots: 8V(2ND + 6N + 4S);
        # fused minimum: 8V(6N + 4S)
        step_b = 8.0 * ((vr * Sg + N * ns) + N * (ns + 5 * V) + V * (N + Sg) + 2 * V * Sg)
        if fused:
            step_b = 8.0 * V * (6 * N + 4 * Sg)
        return det, {k: v["ms"] / steps_ for k, v in prof_.items()}, step_b

    is_fused = args.k3_slots == "fused" and "inv_z_k4" in prof
    ns_run, vr_run = (V * D, V) if args.k3_slots == "all" else (ns_needed, v_read)
    detail, kern_ms, step_bytes = tables(prof, args.steps, ns_run, vr_run, is_fused)
    top = max(detail, key=lambda k: detail[k]["ms_per_step"])
    # measured DRAM traffic / FP64-pipe activity of the same step from the committed ncu pass (profiles/), if present
    traffic, ncu_note = None, None
    prof_name = {"fused": "step_kernels_fused.json", "needed": "step_kernels_needed.json"}.get(args.k3_slots, "step_kernels_all.json")
    prof_file = ROOT / "profiles" / prof_name
    src_hash = kernel_source_hash()
    if prof_file.exists() and json.loads(prof_file.read_text()).get("kernel_src_sha16") != src_hash:
        ncu_note = {"file": "profiles/" + prof_name, "stale": True, "kernel_src_sha16_now": src_hash,
                    "note": "the committed ncu per-step table was taken at a different state of csrc/: traffic not reported"}
    elif prof_file.exists() and world == 1 and cells_tile == C4_CELLS and not tcbl:
        pk = json.loads(prof_file.read_text())["one_step"]
        sel = {"K3": ("k_inv_r", "k_inv_l", "k_inv_z"), "K1": ("k_fwd_z", "k_fwd_l", "k_fwd_r"), "K2": ("k_spline",),
               "K4": ("k_pointwise",)}[top[:2]]      # ("K3+K4 ..." starts with K3: k_inv_z matches k_inv_z_advection too)
        rows = [v for k, v in pk.items() if any(s_ in k for s_ in sel)]
        traffic = 1e9 * sum(r["dram_read_GB"] + r["dram_write_GB"] for r in rows)
        fft = [v for k, v in pk.items() if "k_inv_l" in k]
        ncu_note = {"file": "profiles/" + prof_name, "kernel_src_sha16": src_hash,
                    "k_inv_l_fp64_pipe_pct": sum(r["fp64_pipe_pct"] * r["ms"] for r in fft) / max(sum(r["ms"] for r in fft), 1e-9),
                    "k_inv_l_share_of_step_under_ncu": sum(r["share"] for r in fft),
                    "note": "the ring FFT inside K3 is FP64-pipe bound (Bluestein), not HBM bound; measured FP64 peak "
                            "36.7 TFLOP/s DFMA = DMMA (profiles/fp64_peak_b200.json)"}
    # FP64-pipe roofline of the ring FFTs (they are bound by FP64 issue, not by HBM): FP64 instructions the Bluestein
    # convolutions execute (16 complex values per thread, ~1256 instructions per thread and sequence for the classes with two
    # strided radix-16 passes; counted from SASS, ncu inst_executed_pipe_fp64 agrees) against the measured peak
    # of the pipe (profiles/fp64_peak_b200.json: 36.7 TFLOP/s DFMA = 1.835e13 FP64 instructions/s)
    def fft_fp64(ms, rows_per_ring_zb):
        if not ms or tcbl:
            return None
        ri = np.arange(1, 3 * tcells + 1) + (tsil - 1) * 3
        m = ri + 1
        L = np.array([1 << int(np.ceil(np.log2(max(2 * mm - 1, 4)))) for mm in m])
        big = m > 64                                          # register-resident kernels (L >= 256); the rest is < 1 % of the work
        thread_seqs = float(((L / 16.0) * big).sum()) * 43 * rows_per_ring_zb * 2
        instr = thread_seqs * 1256.0
        peak_i = 36.7e12 / 2.0
        return {"bound": "fp64", "fp64_instructions": instr, "achieved_instr_per_s": instr / (ms / 1e3), "peak_instr_per_s": peak_i,
                "frac": instr / (ms / 1e3) / peak_i, "peak_source": "profiles/fp64_peak_b200.json (measured DFMA peak of this pool)"}
    fp64_roof = {"inv_l": fft_fp64(kern_ms.get("inv_l"), 7 if not (args.k3_slots == "all") else 15),
                 "fwd_l": fft_fp64(kern_ms.get("fwd_l"), 3),
                 "note": "the ring FFTs (Bluestein, exact for every ring length) are FP64-issue bound and move ~1.5 TB/s: their HBM "
                         "fraction is low by construction; this is their own roofline"}
    roof = {"bound": "hbm", "kernel": top, "achieved": detail[top]["achieved_GBps"], "peak": peak, "unit": "GB/s",
            "frac": detail[top]["frac"], "traffic": traffic, "peak_source": peak_src,
            "algorithmic_bytes_per_launch": detail[top]["algorithmic_GB"] * 1e9,
            "algorithmic_bytes_note": "SURVEY 8(d) figure: read A and two history arrays, write var_np1 and expdot_n for EVERY variable "
                                      "(8 (V S + 4 V N)).  Since round 2 the fused kernel leaves the history arrays of u and v alone "
                                      "while they hold their initial zeros (LinearAdvectionRLZ gives them no tendency, "
                                      "src/testModels.jl:93): what it must touch is 8 (V S + 6 N) = "
                                      f"{8.0 * (V * Sg + 6 * N) / 1e9:.2f} GB; `achieved` / `frac` keep the SURVEY figure" if is_fused and not tcbl else None,
            "share_of_step": detail[top]["ms_per_step"] / ms_step, "ncu": ncu_note, "ring_fft_fp64": fp64_roof}
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic", "config": workload_config(world, cells_tile, args.equation_set, NVARS), "clocks": clocks, "e2e": e2e,
            "gpu_launches": int(launches), "per_rank_kernel_ms": per_rank, "parity_nranks": preflight, "strong": strong,
            "roofline": roof, "roofline_detail": detail, "kernel_ms_per_step": kern_ms,
            "timestep_algorithmic_GB": step_bytes / 1e9,
            "timestep_frac_of_hbm_roofline": (step_bytes / 1e9) / (ms_step / 1e3) / peak,
            "transforms_per_s": {"spectralTransform_K1": 1e3 / k1_ms, "gridTransform_K2K3": 1e3 / k23_ms,
                                 "K1_ms": k1_ms, "K2K3_ms": k23_ms, "vars": V}}
    what = {"fused": f"fused: the in-step tileTransform! produces the {ns_needed} (variable, slot) pairs the {args.equation_set} kernel "
                     f"reads, not all {V * D}" + (", and the equation set + AB3 run in the epilogue of its last stage (no slot is "
                     "written to HBM)" if is_fused else " (no fused kernel for this equation set: separate K4 kernel)"),
            "needed": f"needed: the in-step tileTransform! produces the {ns_needed} (variable, slot) pairs the {args.equation_set} "
                      f"kernel reads, not all {V * D}; separate K4 kernel",
            "all": f"all: every one of the {V * D} (variable, slot) pairs is produced"}[args.k3_slots]
    if args.k3_slots != "all":
        what += "; state bit-identical to the all-slots step (tests/test_gpu_parity.py::test_needed_slots_*), which is timed in 'materialised'"
    line["config"]["zero_history"] = ("variables the equation set gives no tendency (LinearAdvectionRLZ: u, v; boundary-layer set: the diagnostic "
                                      "wb) keep the all-zero expdot history they were allocated with; the step kernels neither read it nor "
                                      "write zeros back (same AB3 arithmetic on zeros, state and history bit-identical to the general path: "
                                      "tests test_passive_history_*; SB_PASSIVE=0 restores the reads and writes)")
    line["config"]["k3_slots"] = what
    if mat:
        mdet, mkern, mbytes = tables(mat["prof"], mat["steps"], V * D, V)
        mtop = max(mdet, key=lambda k: mdet[k]["ms_per_step"])
        line["materialised"] = {
            "what": "same step, tileTransform! producing all D slots of all variables (SURVEY 8(d) 'materialised dataflow')",
            "value": mat["value"], "unit": UNIT, "ms_per_step": mat["ms_per_step"], "steps": mat["steps"],
            "kernel_ms_per_step": mkern, "roofline_detail": mdet, "timestep_algorithmic_GB": mbytes / 1e9,
            "timestep_frac_of_hbm_roofline": (mbytes / 1e9) / (mat["ms_per_step"] / 1e3) / peak,
            "roofline": {"bound": "hbm", "kernel": mtop, "achieved": mdet[mtop]["achieved_GBps"], "peak": peak, "unit": "GB/s",
                         "frac": mdet[mtop]["frac"], "algorithmic_bytes_per_launch": mdet[mtop]["algorithmic_GB"] * 1e9}}
    if m is not None:
        m.close()               # release this model's ~60 GB before the secondary runs
    if world == 1 and not tcbl and not args.no_tcbl and not args.cells:
        # secondary number (north_star item 4): the same C4 grid stepped with the 6-variable height-resolved TC
        # boundary-layer set, in a fresh process
        try:
            r = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--equation-set", "Oneway_ShallowWater_HeightResolvedBL",
                                "--steps", "5", "--warmup", "3", "--no-cpu-baseline", "--no-e2e"], capture_output=True,
                               text=True, timeout=600)
            t2 = json.loads(r.stdout.strip().splitlines()[-1])
            line["tcbl"] = {"equation_set": "Oneway_ShallowWater_HeightResolvedBL", "vars": 6, "value": t2["value"], "unit": UNIT,
                            "ms_per_step": t2["ms_per_step"], "kernel_ms_per_step": t2["kernel_ms_per_step"],
                            "timestep_frac_of_hbm_roofline": t2["timestep_frac_of_hbm_roofline"],
                            "k3_slots": t2["config"].get("k3_slots"),
                            "materialised": {k: t2["materialised"][k] for k in ("value", "ms_per_step", "kernel_ms_per_step",
                                                                                "timestep_frac_of_hbm_roofline")}
                            if "materialised" in t2 else None}
        except Exception as e:  # the headline line must still print
            line["tcbl"] = {"error": repr(e)[:200]}
    if not args.no_cpu_baseline and world == 1:
        # the CPU arm on its bounded sample; its final state then checks the GPU path on the same sample (64 levels, the
        # benchmarked kernels): `parity`
        line["cpu_baseline"], _, ostate, onsteps = cpu_baseline()
        if not tcbl:
            try:
                line["parity"] = gpu_sample_parity(ostate, onsteps, device=local_rank)
            except Exception as e:
                line["parity"] = {"error": repr(e)[:300], "ok": False}
    else:
        line["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": 0, "kind": "port", "sample": "timed at N=1 only"}
    line["config"]["exchange_mode"] = cfg_exchange
    print(json.dumps(line))
    if distributed:
        dist.destroy_process_group()


if __name__ == "__main__":
    a = parse()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_ours(a)
