#!/usr/bin/env python3
"""Quick GPU-vs-oracle check with statistics (run on the GPU box): parity percentiles, step-count
histograms side by side, kernel time.  Usage: python tools/gpu_check.py [ncells] [warm|cold]"""
import os, sys, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
import numpy as np
from geos_chem_b200 import grid, kpp
from oracle.pyoracle import Oracle

n = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
hstart = sys.argv[2] if len(sys.argv) > 2 else "warm"
g = grid.make_grid("4x5", hstart=hstart)
ncell = g["conc"].shape[1]
rng = np.random.default_rng(123)
idx = np.sort(rng.choice(ncell, min(n, ncell), replace=False))
sub = lambda a: np.ascontiguousarray(a[..., idx])
conc, hs = sub(g["conc"]), sub(g["hstart"])
o = Oracle()
s = kpp.KppSolver("fullchem", 0, max_cells=len(idx))
for kv in sys.argv[3:]:
    k, v = kv.split("="); s.set_option(k, int(v))
t0 = time.time()
rc_gpu = s.Update_RCONST(sub(g["temp"]), sub(g["numden"]), sub(g["h2o"]), sub(g["photol"]), sub(g["khet"]))
rc = o.update_rconst("fullchem", sub(g["temp"]), sub(g["numden"]), sub(g["h2o"]), sub(g["photol"]), sub(g["khet"]))
t0 = time.time()
co, isto, rsto, ierro = o.integrate("fullchem", 0.0, 1200.0, conc, rc, g["atol"], g["rtol"], g["icntrl"], g["rcntrl"], hstart=hs)
tcpu = time.time() - t0
print("oracle: %.2fs, %.0f cells/s on %d threads" % (tcpu, len(idx) / tcpu, os.cpu_count()))
for rcname, rcu in (("oracle-rconst", rc), ("gpu-rconst", rc_gpu)):
    t0 = time.time()
    c, ist, rst, ierr, nf = s.Integrate(0.0, 1200.0, conc, rcu, g["atol"], g["rtol"], g["icntrl"], g["rcntrl"], hstart=hs)
    tg = time.time() - t0
    st = s.last_stats()
    big = np.abs(co) > 1e3
    rel = np.abs(c - co)[big] / np.abs(co[big])
    same = np.all(ist == isto, axis=0)
    print("[%s] gpu e2e %.3fs (%.0f cells/s) kernel %.1f ms (%.0f cells/s) copies %.1f ms" % (
        rcname, tg, len(idx) / tg, st["integrate_ms"], len(idx) / st["integrate_ms"] * 1e3, st["copy_ms"]))
    print("   ierr equal:", np.array_equal(ierr, ierro), " cells with different step counts:", int((~same).sum()),
          " rel err max %.3e p99.9 %.3e median %.3e  violations(>1e-4): %d" % (
              rel.max(), np.percentile(rel, 99.9), np.median(rel), int((rel > 1e-4).sum())))
print("Nstp histogram (oracle | gpu):")
ho, hg = np.bincount(isto[2], minlength=64), np.bincount(ist[2], minlength=64)
print("  ", ho[:64].tolist()); print("  ", hg[:64].tolist())
print("mean Nstp %.2f Nacc %.2f Nrej %.2f" % (ist[2].mean(), ist[3].mean(), ist[4].mean()))
