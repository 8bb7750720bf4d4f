#!/usr/bin/env python3
"""bench.py -- grid cells/sec for one 20-minute fullchem chemistry step on B200 (BASELINE.json metric).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--grid 4x5] [--hstart warm|cold]
  python bench.py --impl reference ...      # the CPU restatement of the reference on the host cores

A "step" = one pass of the hot path (Update_RCONST + Integrate over every cell of the grid) on
synthetic inputs (geos_chem_b200/grid.py).  N=1 is BASELINE config 2: 4x5 global, 72 levels,
238,464 cells.  N>1 (launched with torchrun, one rank per GPU): every rank integrates its own
4x5-sized share of the columns of an N-times larger grid, dealt round-robin (weak scaling, no data-path collective;
NCCL only reduces the step-count diagnostics).
value = whole-job cells/s with inputs resident in HBM (device entry point);
e2e   = the same through the host-buffer C-ABI call (pinned host arrays, H2D + D2H inside the timed region).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

FLOP_PER_ATTEMPT = 126572.0   # one Rodas3 attempt (LU + 4 solves + 2 Fun + vector ops)
FLOP_PER_ACCEPT = 23235.0     # + 1 Fun + 1 Jac per accepted step
# algorithmic HBM bytes of the integrator kernel per cell: C in 356*8 + RCONST 1058*8 (written by the
# Update_RCONST kernel, read once) + C out 356*8 + ISTATUS 32 + RSTATUS 32 + IERR 4 + hstart 8 = 14236
ALG_BYTES_PER_CELL = 14236.0


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d.get("hbm_gbs", 6650.0), "measured", d.get("sm_max_mhz")
    return 6650.0, "fallback", None


class ClockSampler:
    """nvidia-smi clocks + throttle reasons sampled during the timed region (B200_PROFILING.md recipe)"""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.p = index, [], None

    def start(self):
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                       "--format=csv,noheader,nounits", "-lms", "200"],
                                      stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except OSError:
            self.p = None

    def _read(self):
        for ln in self.p.stdout:
            self.rows.append([x.strip() for x in ln.split(",")])

    def stop(self):
        if not self.p:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.p.terminate()
        try:
            self.p.wait(timeout=2)
        except Exception:
            self.p.kill()
        sm = [float(r[1]) for r in self.rows if len(r) >= 8 and r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in self.rows if len(r) >= 8 and r[2].replace(".", "").isdigit()]
        reasons = set()
        for r in self.rows:
            if len(r) >= 8:
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[4:8]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def make_inputs(args, rank, world):
    from geos_chem_b200 import grid
    NX, NY, NZ = grid.GRIDS[args.grid]
    # weak scaling: the global grid is `world` times wider; its (I,J) columns are dealt round-robin over the
    # ranks (every level of a column stays on one GPU), so each rank sees the global day/night mix
    shape = (NX * world, NY, NZ)
    cells = grid.column_shard(shape, rank, world)
    if args.cells:
        cells = cells[:: max(1, cells.shape[0] // args.cells)][: args.cells]
    g = grid.make_cells(cells, shape, hstart=args.hstart)
    return g, shape


def run_reference(args, rank, world):
    """--impl reference: the CPU restatement of the reference's path (oracle/, C + OpenMP with
    schedule(dynamic,24) like fullchem_mod.F90:541-542) on all host threads.  The reference itself
    is Fortran 90 and no Fortran compiler exists in this image, so oracle/_ref cannot be built."""
    if rank != 0:
        return
    from oracle.pyoracle import Oracle
    o = Oracle()
    g, shape = make_inputs(args, 0, 1)
    n_all = g["conc"].shape[1]
    stride = max(1, n_all // args.ref_cells)
    idx = np.arange(0, n_all, stride)
    sub = lambda a: np.ascontiguousarray(a[..., idx])
    conc, hs = sub(g["conc"]), sub(g["hstart"])
    temp, numden, h2o, photol, khet = map(sub, (g["temp"], g["numden"], g["h2o"], g["photol"], g["khet"]))
    cores = host_cores()      # explicit: torchrun exports OMP_NUM_THREADS=1, which must not throttle the CPU arm
    times = []
    for it in range(args.warmup + args.steps):
        t0 = time.perf_counter()
        rc = o.update_rconst("fullchem", temp, numden, h2o, photol, khet, nthreads=cores)
        o.integrate("fullchem", 0.0, 1200.0, conc, rc, g["atol"], g["rtol"], g["icntrl"], g["rcntrl"], hstart=hs,
                    nthreads=cores)
        if it >= args.warmup:
            times.append(time.perf_counter() - t0)
    T = sum(times)
    val = len(idx) * args.steps / T
    sample = "every %d-th cell of the %s grid (%d of %d cells) per step" % (stride, args.grid, len(idx), n_all)
    line = {"impl": "reference", "metric": "grid cells/sec per 20-min fullchem chemistry step", "value": val,
            "unit": "cells/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1e3 * T / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload_name(args, 1), "hstart": args.hstart, "sample": sample},
            "cpu_baseline": {"value": val, "unit": "cells/s", "cores": cores, "kind": "port", "sample": sample,
                             "note": "C/OpenMP restatement of the reference algorithm, not gfortran/ifort output"},
            "e2e": {"value": val, "unit": "cells/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count()


def workload_name(args, world):
    from geos_chem_b200 import grid
    NX, NY, NZ = grid.GRIDS[args.grid]
    n = NX * NY * NZ if not args.cells else args.cells
    return "fullchem Rodas3, one 1200 s chemistry step, %s x %d levels = %d cells per GPU, hstart %s" % (
        args.grid, NZ, n, args.hstart)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--grid", default="4x5")
    ap.add_argument("--hstart", default="warm", choices=["warm", "cold"])
    ap.add_argument("--cells", type=int, default=0, help="debug: subsample the per-GPU grid to this many cells")
    ap.add_argument("--ref-cells", type=int, default=60000, help="cells per step of the CPU arms (bounded sample)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--option", action="append", default=[], help="solver option key=value")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 0)

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        return run_reference(args, rank, world)

    import torch
    import torch.distributed as dist
    from geos_chem_b200 import kpp
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the B200 path has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    g, shape = make_inputs(args, rank, world)
    ncell = g["conc"].shape[1]
    solver = kpp.KppSolver("fullchem", device=local, max_cells=ncell)
    for kv in args.option:
        k, v = kv.split("=")
        solver.set_option(k, int(v))
    stream = torch.cuda.current_stream()
    solver.set_stream(stream.cuda_stream)
    fp64_peak = kpp.fp64_peak(local) if rank == 0 else None

    host = {k: torch.from_numpy(np.ascontiguousarray(g[k])).pin_memory()
            for k in ("conc", "temp", "numden", "h2o", "photol", "khet", "hstart")}
    d = {k: v.to(dev, non_blocking=True) for k, v in host.items()}
    out = {"conc": torch.empty_like(d["conc"]), "ist": torch.empty((8, ncell), dtype=torch.int32, device=dev),
           "rst": torch.empty((4, ncell), dtype=torch.float64, device=dev),
           "ierr": torch.empty((ncell,), dtype=torch.int32, device=dev)}
    torch.cuda.synchronize()

    def step_device():
        solver.Integrate(0.0, 1200.0, d["conc"], None, g["atol"], g["rtol"], g["icntrl"], g["rcntrl"],
                         hstart=d["hstart"], TEMP=d["temp"], NUMDEN=d["numden"], H2O=d["h2o"], PHOTOL=d["photol"],
                         khet=d["khet"], C_out=out["conc"], ISTATUS=out["ist"], RSTATUS=out["rst"], IERR=out["ierr"])
        return solver.last_stats()

    hn = {k: v.numpy() for k, v in host.items()}
    h_out = {"conc": torch.empty((g["conc"].shape[0], ncell), dtype=torch.float64).pin_memory().numpy(),
             "ist": torch.empty((8, ncell), dtype=torch.int32).pin_memory().numpy(),
             "rst": torch.empty((4, ncell), dtype=torch.float64).pin_memory().numpy(),
             "ierr": torch.empty((ncell,), dtype=torch.int32).pin_memory().numpy()}
    import ctypes as C
    P = lambda a: a.ctypes.data_as(C.c_void_p)
    ic = np.ascontiguousarray(g["icntrl"], np.int32); rcn = np.ascontiguousarray(g["rcntrl"], np.float64)

    def step_host():
        rc = solver.L.gckpp_gpu_integrate(solver.h, ncell, 0.0, 1200.0, P(hn["conc"]), None, P(hn["temp"]),
                                          P(hn["numden"]), P(hn["h2o"]), P(hn["photol"]), P(hn["khet"]),
                                          P(g["atol"]), P(g["rtol"]), P(ic), P(rcn), P(hn["hstart"]), None,
                                          P(h_out["conc"]), P(h_out["ist"]), P(h_out["rst"]), P(h_out["ierr"]))
        if rc < 0:
            raise SystemExit("gckpp_gpu_integrate failed: %d %s" % (rc, solver.L.gckpp_gpu_last_error().decode()))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        """K steps bracketed by barrier+sync, timed with CUDA events on the launching stream; max over ranks"""
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        w0 = time.perf_counter()
        e0.record(stream)
        stats = [fn() for _ in range(steps)]
        e1.record(stream)
        barrier()
        wall = time.perf_counter() - w0
        ms = e0.elapsed_time(e1)
        t = torch.tensor([ms, wall * 1e3], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t[0]), float(t[1]), stats

    for _ in range(args.warmup):
        step_device()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    dev_ms, dev_wall_ms, stats = timed(step_device, args.steps)
    clocks = sampler.stop() if rank == 0 else None
    # the device-entry call synchronises its stream before returning, so events and wall clock agree
    for _ in range(min(args.warmup, 1)):
        step_host()
    e2e_ms, e2e_wall_ms, _ = timed(step_host, args.steps)

    # diagnostics reduce (the only collective on this path): step counts, failures
    ist = out["ist"].to(torch.float64)
    diag = torch.stack([ist[2].sum(), ist[3].sum(), ist[4].sum(), (out["ierr"] != 1).sum().to(torch.float64),
                        torch.tensor(float(ncell), device=dev, dtype=torch.float64)])
    nmax = ist[2].max()
    if world > 1:
        dist.all_reduce(diag)
        dist.all_reduce(nmax, op=dist.ReduceOp.MAX)
    assert torch.equal(out["conc"].cpu(), torch.from_numpy(h_out["conc"])), "host and device entry points disagree"

    if rank == 0:
        total_cells = float(diag[4])
        sum_nstp, sum_nacc = float(diag[0]), float(diag[1])
        value = total_cells * args.steps / (dev_ms * 1e-3)
        e2e = total_cells * args.steps / (e2e_ms * 1e-3)
        kern_ms = float(np.mean([s["integrate_ms"] for s in stats]))      # dominant kernel, per launch (rank 0)
        rconst_ms = float(np.mean([s["rconst_ms"] for s in stats]))
        hbm_peak, peak_src, _ = peaks()
        my_nstp, my_nacc = float(ist[2].sum()), float(ist[3].sum())
        flops = FLOP_PER_ATTEMPT * my_nstp + FLOP_PER_ACCEPT * my_nacc
        alg_bytes = ALG_BYTES_PER_CELL * ncell
        traffic = None
        tp = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tp):
            traffic = json.load(open(tp)).get("dram_bytes_per_cell", 0.0) * ncell or None   # ncu capture, scaled per cell
        h2d = sum(int(v.numel() * v.element_size()) for v in host.values())
        d2h = sum(int(v.nbytes) for v in h_out.values())
        line = {
            "metric": "grid cells/sec per 20-min fullchem chemistry step", "value": value, "unit": "cells/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dev_ms / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload_name(args, world), "grid_shape_global": list(shape), "seed": 20190701,
                       "timing": "inputs (%.0f MB per GPU) are larger than L2, no flush needed" % (h2d / 1e6),
                       "solver_options": args.option},
            "e2e": {"value": e2e, "unit": "cells/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": e2e_ms / args.steps},
            "gpu_launches": int(sum(s["launches"] for s in stats)),
            "clocks": clocks,
            "roofline": {"bound": "hbm", "achieved": alg_bytes / (kern_ms * 1e-3) / 1e9, "peak": hbm_peak,
                         "unit": "GB/s", "frac": alg_bytes / (kern_ms * 1e-3) / 1e9 / hbm_peak, "traffic": traffic,
                         "peak_source": "MEASURED_PEAKS.json hbm_gbs (%s)" % peak_src,
                         "kernel": "ros_smem_kernel<fullchem> (shared-memory Rodas3 integrator, one launch per step)",
                         "kernel_ms": kern_ms, "algorithmic_bytes_per_cell": ALG_BYTES_PER_CELL,
                         "note": "neither HBM nor tensor bound: per-cell state is resident in shared memory, the kernel "
                                 "is bound by dependent-instruction latency on the sparse-LU/solve chains (see "
                                 "DESIGN.md); the algorithmic-byte HBM fraction is reported as the contract asks and "
                                 "the FP64-pipe fraction next to it",
                         "fp64": {"achieved": flops / (kern_ms * 1e-3) / 1e12, "peak": fp64_peak, "unit": "TFLOP/s",
                                  "frac": flops / (kern_ms * 1e-3) / 1e12 / fp64_peak if fp64_peak else None,
                                  "peak_source": "DFMA chain micro-benchmark measured in this run",
                                  "flop_model": "126572*Nstp + 23235*Nacc per cell (SURVEY.md 8d)"},
                         "hbm_scratch": {"achieved": (traffic / (kern_ms * 1e-3) / 1e9) if traffic else None,
                                         "peak": hbm_peak, "unit": "GB/s",
                                         "frac": (traffic / (kern_ms * 1e-3) / 1e9 / hbm_peak) if traffic else None}},
            "diagnostics": {"mean_nstp": sum_nstp / total_cells, "mean_nacc": sum_nacc / total_cells,
                            "max_nstp": float(nmax), "failed_cells": float(diag[3]), "update_rconst_ms": rconst_ms,
                            "wall_ms_per_step": dev_wall_ms / args.steps},
        }
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline(args, g)
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def cpu_baseline(args, g):
    """the oracle (C/OpenMP restatement of the reference) timed on this box's host cores on a bounded
    sample of the same workload -- a reported baseline, used only as the checker/baseline, never shipped"""
    from oracle.pyoracle import Oracle
    o = Oracle()
    n_all = g["conc"].shape[1]
    stride = max(1, n_all // args.ref_cells)
    idx = np.arange(0, n_all, stride)
    sub = lambda a: np.ascontiguousarray(a[..., idx])
    conc, hs = sub(g["conc"]), sub(g["hstart"])
    temp, numden, h2o, photol, khet = map(sub, (g["temp"], g["numden"], g["h2o"], g["photol"], g["khet"]))
    best = None
    cores = host_cores()
    for _ in range(2):
        t0 = time.perf_counter()
        rc = o.update_rconst("fullchem", temp, numden, h2o, photol, khet, nthreads=cores)
        o.integrate("fullchem", 0.0, 1200.0, conc, rc, g["atol"], g["rtol"], g["icntrl"], g["rcntrl"], hstart=hs,
                    nthreads=cores)
        dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
    return {"value": len(idx) / best, "unit": "cells/s", "cores": cores, "kind": "port",
            "sample": "every %d-th cell of the workload (%d cells), best of 2" % (stride, len(idx)),
            "note": "C/OpenMP (schedule(dynamic,24)) restatement of the reference algorithm, not gfortran/ifort output"}


if __name__ == "__main__":
    main()
