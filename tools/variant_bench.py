"""time the integrator of one library variant on the full 4x5 grid (device-resident inputs)"""
import sys, time
import numpy as np
sys.path.insert(0, ".")
import torch
from geos_chem_b200 import grid, kpp
g = grid.make_grid("4x5", hstart="warm")
n = g["conc"].shape[1]
s = kpp.KppSolver("fullchem", 0, max_cells=n)
dev = torch.device("cuda:0")
t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
conc, temp, numden, h2o, photol, khet, hs = map(t, (g["conc"], g["temp"], g["numden"], g["h2o"], g["photol"], g["khet"], g["hstart"]))
for it in range(2):
    out = s.Integrate(0.0, 1200.0, conc, None, g["atol"], g["rtol"], g["icntrl"], g["rcntrl"], hstart=hs, TEMP=temp, NUMDEN=numden, H2O=h2o, PHOTOL=photol, khet=khet)
    st = s.last_stats()
print(kpp.LIB_PATH.split("/")[-1], "integrate %.1f ms" % st["integrate_ms"], "cells/s %.0f" % (n / st["integrate_ms"] * 1e3), "sum_nstp", st["sum_nstp"], "ok", bool((out[3] == 1).all()))
