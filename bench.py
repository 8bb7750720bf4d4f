#!/usr/bin/env python3
"""bench.py -- grid cells/sec for one 20-minute fullchem chemistry step on B200 (BASELINE.json metric).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--config 1|2|3|4|5-hg|5-carbon|5-ar] [--hstart warm|cold]
  python bench.py --impl reference ...      # the CPU restatement of the reference on the host cores

A "step" = one pass of the hot path (Update_RCONST + Integrate over every cell of the grid) on
synthetic inputs (geos_chem_b200/grid.py).  N=1 defaults to BASELINE config 2: 4x5 global, 72 levels,
238,464 cells.  N>1 (launched with torchrun, one rank per GPU) defaults to config 3, STRONG scaling: the 943,488
cells of the 2x2.5 grid, (I,J) columns dealt round-robin over the ranks, no data-path collective (NCCL only
reduces the step-count diagnostics and the per-rank times); --scaling weak gives every rank a 4x5-sized share.
value = whole-job cells/s with inputs resident in HBM (device entry point);
e2e   = the same through the host-buffer C-ABI call (pinned host arrays, H2D + D2H inside the timed region, waves).
At N=1 the line also carries cpu_baseline (the oracle on the host cores) and parity (every checked cell of the GPU
result against the oracle: violations of the 1e-4 bar, step-count histograms side by side).
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


CONFIGS = {
    # BASELINE.json configs: name -> (mechanism, grid, description)
    "1": ("fullchem", "4x5", "KPP/standalone Beijing sample cell replicated over a 4x5x72 grid (zero divergence)"),
    "2": ("fullchem", "4x5", "4x5 global, 72 levels"),
    "3": ("fullchem", "2x2.5", "2x2.5 global, 72 levels, sharded by (I,J) columns over the GPUs"),
    "4": ("fullchem", "c180", "C180 cubed-sphere equivalent, 72 levels, sharded by columns, processed in waves"),
    "5-hg": ("Hg", "4x5", "Hg mechanism (small-mechanism path) on the 4x5x72 cells"),
    "5-carbon": ("carbon", "4x5", "carbon mechanism (forward Euler) on the 4x5x72 cells"),
    "5-ar": ("fullchem", "4x5", "4x5 global, 72 levels, auto-reduce solver (ICNTRL(12)=1, threshold 100)"),
}
KERNEL_NAMES = {0: "ros_generic_kernel (table-driven, one cell per lane, workspace in HBM)",
                1: "ros_smem_kernel<mech> (shared-memory Rodas3 integrator, one launch per step)",
                2: "ros_warp_kernel<mech> (warp-group shared-memory Rodas3 integrator)",
                3: "ros_lane_kernel (one cell per lane, streamed workspace)",
                4: "ros_unrolled_kernel<mech> (one cell per thread, generated straight-line code)"}


def resolve(args, world):
    """fill in config / scaling defaults: N=1 -> config 2 (the configuration the metric is quoted on);
    N>1 -> config 3 strong-scaled (943,488 cells sharded by column), --scaling weak keeps a 4x5-sized share per GPU"""
    if args.config is None:
        args.config = "2" if ((world == 1 and not args.shard_of) or args.scaling == "weak") else "3"
    if args.scaling is None:
        args.scaling = "strong" if (world > 1 or args.shard_of) else "weak"
    args.mech, cfg_grid, args.desc = CONFIGS[args.config]
    if args.grid is None:
        args.grid = cfg_grid
    args.dt = 3600.0 if args.mech in ("Hg", "carbon") else 1200.0


def make_inputs(args, rank, world):
    """the cells of this rank and their synthetic inputs (geos_chem_b200/grid.py)"""
    from geos_chem_b200 import grid
    NX, NY, NZ = grid.GRIDS[args.grid]
    vworld, vrank = (args.shard_of, args.shard_rank) if args.shard_of else (world, rank)
    if args.scaling == "weak":
        # the global grid is `world` times wider; every rank integrates a share of the size of the named grid
        shape = (NX * vworld, NY, NZ)
    else:
        shape = (NX, NY, NZ)
    # (I,J) columns are dealt round-robin over the ranks (every level of a column stays on one GPU), so each rank
    # sees the global day/night mix
    cells = grid.column_shard(shape, vrank, vworld)
    if args.cells:
        cells = cells[:: max(1, cells.shape[0] // args.cells)][: args.cells]
    if args.config == "1":
        g = grid.replicate_fixture(cells.shape[0])
        g["cells"] = cells
    elif args.mech in ("Hg", "carbon"):
        g = grid.make_small_mech(args.mech, cells)
        g["hstart"] = None
    else:
        g = grid.make_cells(cells, shape, hstart=args.hstart)
        if args.config == "5-ar":
            g["icntrl"] = g["icntrl"].copy(); g["rcntrl"] = g["rcntrl"].copy()
            g["icntrl"][11] = 1; g["rcntrl"][11] = 100.0
    return g, shape


def oracle_run(o, args, g, idx, cores):
    """one pass of the CPU restatement over the cells idx of g; returns (seconds, conc_out, istatus, ierr)"""
    sub = lambda a: None if a is None else np.ascontiguousarray(a[..., idx])
    if args.config == "5-ar" and not getattr(o, "_keep_set", False):
        from geos_chem_b200 import grid as _grid
        from geos_chem_b200.kppgen import ir as _ir
        o.set_keep_active("fullchem", _grid.keep_active_indices(_ir.load("fullchem").spc_names[:353]))
        o._keep_set = True
    conc, hs = sub(g["conc"]), sub(g.get("hstart"))
    t0 = time.perf_counter()
    if "rconst" in g:
        rc = sub(g["rconst"])
    else:
        rc = o.update_rconst(args.mech, sub(g["temp"]), sub(g["numden"]), sub(g["h2o"]), sub(g["photol"]), sub(g["khet"]),
                             nthreads=cores)
    co, ist, rst, ierr = o.integrate(args.mech, 0.0, args.dt, conc, rc, g["atol"], g["rtol"], g["icntrl"], g["rcntrl"],
                                     hstart=hs, nthreads=cores)
    return time.perf_counter() - t0, co, ist, ierr


def pick_oracle(args, g, cores):
    """the faster of the two CPU builds on this box (4000-cell probe): -O3 -march=native built here, or the shipped
    -O2 -march=x86-64-v3 -- the CPU arm gets whichever serves it best"""
    from oracle.pyoracle import Oracle
    n_all = g["conc"].shape[1]
    probe = np.arange(0, n_all, max(1, n_all // 4000))
    best = None
    for variant in ("native", "strict"):
        try:
            o = Oracle(variant)
        except Exception:
            continue
        oracle_run(o, args, g, probe[:200], cores)
        dt, _, _, _ = oracle_run(o, args, g, probe, cores)
        if best is None or dt < best[0]:
            best = (dt, o)
    desc = {"native": "gcc -O3 -march=native -fopenmp, built on this box", "strict": "gcc -O2 -march=x86-64-v3 -fopenmp"}
    best[1].build_desc = "oracle/ %s build (%s, no FMA contraction; the faster of the two builds on a 4000-cell probe)" % (
        best[1].variant, desc[best[1].variant])
    return best[1]


def cpu_sample(o, args, g, cores, passes, budget_s):
    """cells of the CPU arms: the whole workload when `passes` passes fit the time budget, else a uniform sample of it
    (rate estimated on a 4000-cell probe)"""
    n_all = g["conc"].shape[1]
    if args.ref_cells:
        n = min(args.ref_cells, n_all)
    else:
        probe = np.arange(0, n_all, max(1, n_all // 4000))
        dt, _, _, _ = oracle_run(o, args, g, probe, cores)
        rate = len(probe) / dt
        n = n_all if n_all * passes / rate <= budget_s else int(rate * budget_s / passes)
    idx = np.arange(n_all) if n >= n_all else np.arange(0, n_all, n_all / n).astype(np.int64)[:n]
    what = ("the whole workload (%d cells)" % n_all) if len(idx) == n_all else \
        "a uniform sample of the workload (%d of %d cells)" % (len(idx), n_all)
    return idx, what


def run_reference(args, rank, world):
    """--impl reference: the CPU restatement of the reference's path (oracle/, C + OpenMP with
    schedule(dynamic,24) like fullchem_mod.F90:541-542) on all host threads, built -O3 -march=native on this box.
    The reference itself is Fortran 90 and no Fortran compiler exists in this image, so oracle/_ref cannot be built."""
    if rank != 0:
        return
    resolve(args, args.gpus)
    # the same cells as the B200 arm: the whole grid of the config (all ranks' shards together)
    save = (args.shard_of, args.shard_rank)
    args.shard_of, args.shard_rank = (1, 0) if not args.shard_of else save
    g, shape = make_inputs(args, 0, 1)
    from geos_chem_b200 import grid as _grid
    # the B200 arm's own description of the workload (cells per GPU = rank 0's shard), so that the two lines name the same config
    per_gpu = g["conc"].shape[1] if args.scaling == "weak" or args.cells else _grid.column_shard(shape, 0, max(args.gpus, 1)).shape[0]
    cores = host_cores()      # explicit: torchrun exports OMP_NUM_THREADS=1, which must not throttle the CPU arm
    o = pick_oracle(args, g, cores)
    warm = min(args.warmup, 1)
    idx, what = cpu_sample(o, args, g, cores, warm + args.steps, 420.0)
    times = []
    for it in range(warm + args.steps):
        dt, _, _, _ = oracle_run(o, args, g, idx, cores)
        if it >= warm:
            times.append(dt)
    T = sum(times)
    val = len(idx) * args.steps / T
    sample = "%s per step, %d untimed warm-up pass(es)" % (what, warm)
    line = {"impl": "reference", "metric": metric_name(args), "value": val,
            "unit": "cells/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1e3 * T / args.steps, "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload_name(args, per_gpu, args.gpus), "name": args.config, "sample": sample},
            "cpu_baseline": {"value": val, "unit": "cells/s", "cores": cores, "kind": "port", "sample": sample,
                             "build": o.build_desc,
                             "note": "C/OpenMP restatement of the reference algorithm, not gfortran/ifort output"},
            "e2e": {"value": val, "unit": "cells/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count()


def metric_name(args):
    if args.mech == "fullchem":
        return "grid cells/sec per 20-min fullchem chemistry step"
    return "grid cells/sec per 60-min %s chemistry step" % args.mech


def workload_name(args, ncell, world):
    meth = "forward Euler" if args.mech == "carbon" else "Rodas3"
    return "config %s: %s; %s %s, one %g s chemistry step, %d cells per GPU on %d GPU(s), hstart %s" % (
        args.config, args.desc, args.mech, meth, args.dt, ncell, world, args.hstart)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", default=None, choices=sorted(CONFIGS), help="BASELINE.json config (default: 2 on one GPU, 3 on several)")
    ap.add_argument("--scaling", default=None, choices=["weak", "strong"], help="N>1: strong (default) shards the config's grid, weak gives every GPU a grid-sized share")
    ap.add_argument("--grid", default=None, help="override the config's grid (4x5, 2x2.5, c180)")
    ap.add_argument("--hstart", default="warm", choices=["warm", "cold"])
    ap.add_argument("--cells", type=int, default=0, help="debug: subsample the per-GPU cells to this many")
    ap.add_argument("--shard-of", type=int, default=0, help="debug: take shard --shard-rank of this many (e.g. one eighth of C180 on one GPU)")
    ap.add_argument("--shard-rank", type=int, default=0)
    ap.add_argument("--ref-cells", type=int, default=0, help="cells per step of the CPU arms (default: the whole workload if it fits the time budget)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--option", action="append", default=[], help="solver option key=value")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 0)

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        return run_reference(args, rank, world)
    resolve(args, world)

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
    if ncell > 4_000_000:
        raise SystemExit("bench.py: %d cells on one GPU -- config 4 is meant for 8 GPUs (or --shard-of 8)" % ncell)
    solver = kpp.KppSolver(args.mech, device=local, max_cells=ncell)
    for kv in args.option:
        k, v = kv.split("=")
        solver.set_option(k, int(v))
    if args.config == "5-ar":
        from geos_chem_b200 import grid as _grid
        solver.set_keep_active(_grid.keep_active_indices(kpp.spc_names("fullchem")[:353]))      # keepSpcActive: the halogens
    # the rate constants of a wave are computed on the device into a bounded scratch (config 4: 1.75 M cells per GPU)
    solver.set_option("device_wave_cells", 262144)
    stream = torch.cuda.current_stream()
    solver.set_stream(stream.cuda_stream)
    fp64_peak = kpp.fp64_peak(local) if rank == 0 else None

    in_keys = [k for k in ("conc", "rconst", "temp", "numden", "h2o", "photol", "khet", "hstart") if g.get(k) is not None]
    host = {k: torch.from_numpy(np.ascontiguousarray(g[k])).pin_memory() for k in in_keys}
    d = {k: v.to(dev, non_blocking=True) for k, v in host.items()}
    nspec = g["conc"].shape[0]
    out = {"conc": torch.empty_like(d["conc"]), "ist": torch.empty((8, ncell), dtype=torch.int32, device=dev),
           "rst": torch.empty((4, ncell), dtype=torch.float64, device=dev),
           "ierr": torch.empty((ncell,), dtype=torch.int32, device=dev)}
    torch.cuda.synchronize()

    def step_device():
        solver.Integrate(0.0, args.dt, d["conc"], d.get("rconst"), g["atol"], g["rtol"], g["icntrl"], g["rcntrl"],
                         hstart=d.get("hstart"), TEMP=d.get("temp"), NUMDEN=d.get("numden"), H2O=d.get("h2o"),
                         PHOTOL=d.get("photol"), khet=d.get("khet"), C_out=out["conc"], ISTATUS=out["ist"],
                         RSTATUS=out["rst"], IERR=out["ierr"])
        return solver.last_stats()

    hn = {k: v.numpy() for k, v in host.items()}
    h_out = {"conc": torch.empty((nspec, ncell), dtype=torch.float64).pin_memory().numpy(),
             "ist": torch.empty((8, ncell), dtype=torch.int32).pin_memory().numpy(),
             "rst": torch.empty((4, ncell), dtype=torch.float64).pin_memory().numpy(),
             "ierr": torch.empty((ncell,), dtype=torch.int32).pin_memory().numpy()}
    import ctypes as C
    P = lambda a: None if a is None else a.ctypes.data_as(C.c_void_p)
    ic = np.ascontiguousarray(g["icntrl"], np.int32); rcn = np.ascontiguousarray(g["rcntrl"], np.float64)
    atol = np.ascontiguousarray(g["atol"], np.float64); rtol = np.ascontiguousarray(g["rtol"], np.float64)

    # the call INTEGRATION.md prescribes for Do_FullChem: InChemGrid mask (all cells are in the chemistry grid here) and
    # the retry policy on
    active = torch.ones(ncell, dtype=torch.uint8).pin_memory().numpy()

    def step_host(src=None):
        a = src or hn
        rc = solver.L.gckpp_gpu_integrate(solver.h, ncell, 0.0, args.dt, P(a["conc"]), P(a.get("rconst")), P(a.get("temp")),
                                          P(a.get("numden")), P(a.get("h2o")), P(a.get("photol")), P(a.get("khet")),
                                          P(atol), P(rtol), P(ic), P(rcn), P(a.get("hstart")), P(active),
                                          P(h_out["conc"]), P(h_out["ist"]), P(h_out["rst"]), P(h_out["ierr"]))
        if rc < 0:
            raise SystemExit("gckpp_gpu_integrate failed: %d %s" % (rc, solver.L.gckpp_gpu_last_error().decode()))
        return solver.last_stats()

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
        mine = ms
        t = torch.tensor([ms, wall * 1e3], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t[0]), float(t[1]), stats, mine

    for _ in range(args.warmup):
        step_device()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    dev_ms, dev_wall_ms, stats, my_dev_ms = timed(step_device, args.steps)
    clocks = sampler.stop() if rank == 0 else None
    # the device-entry call synchronises its stream before returning, so events and wall clock agree
    solver.set_option("retry", 1)
    for _ in range(min(args.warmup, 1)):
        step_host()
    e2e_ms, e2e_wall_ms, hstats, _ = timed(step_host, args.steps)
    # the same call from PAGEABLE host arrays (what Fortran's State_Chm arrays are), page-locked by the library for
    # the duration of the call ("pin" option); one step, reported next to the pinned number
    pageable = {k: np.array(v, copy=True) for k, v in hn.items()}
    solver.set_option("pin", 1)
    step_host(pageable)
    pg_ms, _, _, _ = timed(lambda: step_host(pageable), 1)
    solver.set_option("pin", 0)
    solver.set_option("retry", 0)

    # diagnostics reduce (the only collective on this path): step counts, failures, per-rank times
    ist = out["ist"].to(torch.float64)
    diag = torch.stack([ist[2].sum(), ist[3].sum(), ist[4].sum(), (out["ierr"] != 1).sum().to(torch.float64),
                        torch.tensor(float(ncell), device=dev, dtype=torch.float64)])
    nmax = ist[2].max()
    rank_ms = torch.zeros(world, dtype=torch.float64, device=dev)
    rank_ms[rank] = my_dev_ms / args.steps
    if world > 1:
        dist.all_reduce(diag)
        dist.all_reduce(nmax, op=dist.ReduceOp.MAX)
        dist.all_reduce(rank_ms)
    assert torch.equal(out["conc"].cpu(), torch.from_numpy(h_out["conc"])), "host and device entry points disagree"

    if rank == 0:
        total_cells = float(diag[4])
        sum_nstp, sum_nacc = float(diag[0]), float(diag[1])
        value = total_cells * args.steps / (dev_ms * 1e-3)
        e2e = total_cells * args.steps / (e2e_ms * 1e-3)
        kern_ms = float(np.mean([s["integrate_ms"] for s in stats]))      # integrator launches of a step (rank 0)
        rconst_ms = float(np.mean([s["rconst_ms"] for s in stats]))
        kern = int(stats[-1].get("kernel", -1))
        hbm_peak, peak_src, _ = peaks()
        my_nstp, my_nacc = float(ist[2].sum()), float(ist[3].sum())
        alg_bpc = 8.0 * (2 * nspec + solver.dims["nreact"]) + 32 + 32 + 4 + (8 if g.get("hstart") is not None else 0)
        alg_bytes = alg_bpc * ncell
        traffic = None
        tp = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tp) and args.mech == "fullchem" and kern == 1:
            traffic = json.load(open(tp)).get("dram_bytes_per_cell", 0.0) * ncell or None   # ncu capture, scaled per cell
        h2d = sum(int(v.numel() * v.element_size()) for v in host.values())
        d2h = sum(int(v.nbytes) for v in h_out.values())
        hbm = {"bound": "hbm", "achieved": alg_bytes / (kern_ms * 1e-3) / 1e9, "peak": hbm_peak, "unit": "GB/s",
               "frac": alg_bytes / (kern_ms * 1e-3) / 1e9 / hbm_peak, "traffic": traffic,
               "peak_source": "MEASURED_PEAKS.json hbm_gbs (%s)" % peak_src, "algorithmic_bytes_per_cell": alg_bpc}
        if args.mech == "fullchem":
            # the larger fraction binds -- for the 353-species mechanism that is the FP64 pipe
            flops = FLOP_PER_ATTEMPT * my_nstp + FLOP_PER_ACCEPT * my_nacc
            roof = {"bound": "fp64", "achieved": flops / (kern_ms * 1e-3) / 1e12, "peak": fp64_peak, "unit": "TFLOP/s",
                    "frac": flops / (kern_ms * 1e-3) / 1e12 / fp64_peak if fp64_peak else None, "traffic": traffic,
                    "peak_source": "DFMA micro-benchmark measured in this run (MEASURED_PEAKS.json has no FP64 entry)",
                    "flop_model": "126572*Nstp + 23235*Nacc per cell (SURVEY.md 8d; an upper bound with auto-reduce)",
                    "hbm": hbm,
                    "note": "the per-cell state is on chip, so neither DRAM traffic nor the FP64 pipe is saturated: the "
                            "kernel is bound by dependent-instruction latency on the sparse LU / solve chains (DESIGN.md)"}
        else:
            roof = dict(hbm, note="small mechanism: a cell is 1-2 KB of input and a few kflop per attempt; the "
                                  "algorithmic-byte HBM fraction is what a fully fused kernel would be held to")
        roof["kernel"] = KERNEL_NAMES.get(kern, "feuler_kernel (forward Euler)" if args.mech == "carbon" else "?")
        roof["kernel_ms"] = kern_ms
        line = {
            "metric": metric_name(args), "value": value, "unit": "cells/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dev_ms / args.steps,
            "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload_name(args, ncell, world), "name": args.config, "grid_shape_global": list(shape),
                       "total_cells": int(total_cells), "seed": 20190701,
                       "timing": "inputs (%.0f MB per GPU) are larger than L2, no flush needed" % (h2d / 1e6),
                       "solver_options": args.option},
            "e2e": {"value": e2e, "unit": "cells/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": e2e_ms / args.steps, "waves": int(hstats[-1].get("waves", 0)),
                    "call": "gckpp_gpu_integrate with the InChemGrid mask and retry=1, as INTEGRATION.md prescribes; pinned host arrays",
                    "pageable_host_value": total_cells / (pg_ms * 1e-3),
                    "pageable_host_note": "same call from pageable arrays, page-locked per call by the library (option pin=1), one step"},
            "gpu_launches": int(sum(s["launches"] for s in stats)),
            "clocks": clocks,
            "roofline": roof,
            "diagnostics": {"mean_nstp": sum_nstp / total_cells, "mean_nacc": sum_nacc / total_cells,
                            "max_nstp": float(nmax), "failed_cells": float(diag[3]), "update_rconst_ms": rconst_ms,
                            "wall_ms_per_step": dev_wall_ms / args.steps,
                            "rank_ms_per_step": [round(float(x), 2) for x in rank_ms.cpu()]},
        }
        if world > 1:
            slow, mean = float(rank_ms.max()), float(rank_ms.mean())
            line["diagnostics"]["limiter"] = ("the slowest rank sets the step time: %.1f ms vs %.1f ms mean over ranks "
                                              "(load imbalance %.1f %%); no data-path collective" % (slow, mean, 100.0 * (slow / mean - 1.0)))
        if world == 1 and not args.no_cpu_baseline:
            cb, par = cpu_baseline(args, g, out)
            line["cpu_baseline"] = cb
            line["parity"] = par
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def cpu_baseline(args, g, out):
    """the oracle (C/OpenMP restatement of the reference, -O3 -march=native build of this box) timed on the host cores
    on the whole workload when one pass fits ~30 s, else on a uniform sample -- a reported baseline and the parity
    checker of the line (every compared cell against the GPU result), never shipped"""
    cores = host_cores()
    o = pick_oracle(args, g, cores)
    idx, what = cpu_sample(o, args, g, cores, 1, 30.0)
    dt, co, isto, ierro = oracle_run(o, args, g, idx, cores)
    cb = {"value": len(idx) / dt, "unit": "cells/s", "cores": cores, "kind": "port", "sample": what + ", one pass",
          "build": o.build_desc,
          "note": "C/OpenMP (schedule(dynamic,24)) restatement of the reference algorithm, not gfortran/ifort output"}
    c = out["conc"][:, idx].cpu().numpy() if len(idx) < out["conc"].shape[1] else out["conc"].cpu().numpy()
    ist = out["ist"].cpu().numpy()[:, idx]
    ierr = out["ierr"].cpu().numpy()[idx]
    big = np.abs(co) > 1e3
    rel = np.abs(c - co)[big] / np.abs(co[big])
    hist = lambda a: np.bincount(a, minlength=48)[:48].tolist()
    par = {"checked_cells": int(len(idx)), "what": what, "tolerance": "every species above 1e3 molec/cm3 within 1e-4 relative",
           "violations": int((rel > 1e-4).sum()), "max_rel_err": float(rel.max()) if rel.size else 0.0,
           "ierr_equal": bool(np.array_equal(ierr, ierro)),
           "cells_with_different_step_counts": int((~np.all(ist == isto, axis=0)).sum()),
           "nstp_histogram_gpu": hist(ist[2]), "nstp_histogram_cpu": hist(isto[2])}
    return cb, par


if __name__ == "__main__":
    main()
