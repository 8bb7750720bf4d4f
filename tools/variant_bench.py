"""time the integrator on the full 4x5 grid (device-resident inputs) under a few launch conditions"""
import os, sys, time
import numpy as np
sys.path.insert(0, ".")
import torch
from geos_chem_b200 import grid, kpp
mode = sys.argv[1] if len(sys.argv) > 1 else "own"
g = grid.make_grid("4x5", hstart="warm", limit=int(os.environ.get("VB_CELLS", "0")) or None)
n = g["conc"].shape[1]
s = kpp.KppSolver("fullchem", 0, max_cells=n)
for kv in sys.argv[2:]:
    k, v = kv.split("="); s.set_option(k, int(v))
dev = torch.device("cuda:0")
if "torchstream" in mode:
    s.set_stream(torch.cuda.current_stream().cuda_stream)
if "peak" in mode:
    print("fp64 peak", kpp.fp64_peak(0))
t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
conc, temp, numden, h2o, photol, khet, hs = map(t, (g["conc"], g["temp"], g["numden"], g["h2o"], g["photol"], g["khet"], g["hstart"]))
kw = {}
if "prealloc" in mode:
    kw = dict(C_out=torch.empty_like(conc), ISTATUS=torch.empty((8, n), dtype=torch.int32, device=dev),
              RSTATUS=torch.empty((4, n), dtype=torch.float64, device=dev), IERR=torch.empty((n,), dtype=torch.int32, device=dev))
for it in range(int(os.environ.get("VB_ITERS", "4"))):
    out = s.Integrate(0.0, 1200.0, conc, None, g["atol"], g["rtol"], g["icntrl"], g["rcntrl"], hstart=hs, TEMP=temp, NUMDEN=numden, H2O=h2o, PHOTOL=photol, khet=khet, **kw)
    st = s.last_stats()
    print(mode, it, "integrate %.1f ms" % st["integrate_ms"], "cells/s %.0f" % (n / st["integrate_ms"] * 1e3), flush=True)
