"""GPU bring-up check of the shared-memory kernel against the table-driven kernel and the oracle.
Usage: python tools/smem_debug.py [stage]   (stage: small | grid | full)"""
import sys, time
import numpy as np
sys.path.insert(0, ".")
from geos_chem_b200 import grid, kpp
from oracle.pyoracle import Oracle

stage = sys.argv[1] if len(sys.argv) > 1 else "small"
fx = grid.load_fixture()
s = kpp.KppSolver("fullchem", 0, max_cells=1 << 18)
o = Oracle()


def run(kernel, *a, warps=12, **k):
    s.set_option("kernel", kernel)
    t0 = time.time()
    out = s.Integrate(*a, **k)
    return out, time.time() - t0, s.last_stats()


if stage == "small":
    for n in (1, 3, 7, 64):
        r = grid.replicate_fixture(n, fx)
        args = (0.0, r["dt"], r["conc"], r["rconst"], r["atol"], r["rtol"], r["icntrl"], r["rcntrl"])
        (c1, i1, r1, e1, _), t1, st1 = run(1, *args)
        (c0, i0, r0, e0, _), t0, st0 = run(0, *args)
        big = np.abs(c0) > 1e3
        rel = np.abs(c1 - c0)[big] / np.abs(c0[big])
        print("n=%d smem ierr %s ist %s Hexit %.6f | generic ist %s Hexit %.6f | max rel %.3e | ms %.2f vs %.2f"
              % (n, e1[:3], i1[:, 0], r1[1, 0], i0[:, 0], r0[1, 0], rel.max() if rel.size else -1, st1["integrate_ms"], st0["integrate_ms"]), flush=True)
if stage in ("small", "grid"):
    g = grid.make_grid("4x5", limit=6000)
    for hs in ("warm",):
        args = (0.0, 1200.0, g["conc"], None, g["atol"], g["rtol"], g["icntrl"], g["rcntrl"])
        kw = dict(hstart=g["hstart"], TEMP=g["temp"], NUMDEN=g["numden"], H2O=g["h2o"], PHOTOL=g["photol"], khet=g["khet"])
        (c1, i1, r1, e1, _), t1, st1 = run(1, *args, **kw)
        (c0, i0, r0, e0, _), t0, st0 = run(0, *args, **kw)
        big = np.abs(c0) > 1e3
        rel = np.zeros_like(c0); rel[big] = np.abs(c1 - c0)[big] / np.abs(c0[big])
        same = np.all(i1 == i0, axis=0)
        print("grid 6000: ierr equal %s, cells with different steps %d, max rel %.3e (cell %d), ms smem %.1f generic %.1f"
              % (np.array_equal(e1, e0), int((~same).sum()), rel.max(), int(np.argmax(rel.max(axis=0))), st1["integrate_ms"], st0["integrate_ms"]), flush=True)
        print("  sum nstp smem %d generic %d" % (i1[2].sum(), i0[2].sum()))
if stage == "full":
    import torch
    g = grid.make_grid("4x5", hstart="warm")
    n = g["conc"].shape[1]
    dev = torch.device("cuda:0")
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    conc, temp, numden, h2o, photol, khet, hs = map(t, (g["conc"], g["temp"], g["numden"], g["h2o"], g["photol"], g["khet"], g["hstart"]))
    for warps in (12,):
        s.set_option("kernel", 1)
        for it in range(2):
            torch.cuda.synchronize(); t0 = time.time()
            out = s.Integrate(0.0, 1200.0, conc, None, g["atol"], g["rtol"], g["icntrl"], g["rcntrl"], hstart=hs, TEMP=temp, NUMDEN=numden, H2O=h2o, PHOTOL=photol, khet=khet)
            torch.cuda.synchronize(); dt = time.time() - t0
            st = s.last_stats()
            print("warps", warps, "iter", it, "wall %.3fs integrate %.1f ms" % (dt, st["integrate_ms"]), "cells/s %.0f" % (n / (st["integrate_ms"] / 1e3)),
                  "sum_nstp", st["sum_nstp"], "ierr ok", bool((out[3] == 1).all()), flush=True)
