for b in 2 4 5; do
echo "=== blocks_per_sm=$b"
python - <<PY
import sys, time, numpy as np
sys.path.insert(0,'.')
from geos_chem_b200 import grid, kpp
import torch
g = grid.make_grid("4x5", hstart="warm")
n = g["conc"].shape[1]
s = kpp.KppSolver("fullchem", 0, max_cells=n)
s.set_option("blocks_per_sm", $b)
dev = torch.device("cuda:0")
t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
conc, temp, numden, h2o, photol, khet, hs = map(t, (g["conc"], g["temp"], g["numden"], g["h2o"], g["photol"], g["khet"], g["hstart"]))
for it in range(2):
    torch.cuda.synchronize(); t0 = time.time()
    out = s.Integrate(0.0, 1200.0, conc, None, g["atol"], g["rtol"], g["icntrl"], g["rcntrl"], hstart=hs, TEMP=temp, NUMDEN=numden, H2O=h2o, PHOTOL=photol, khet=khet)
    torch.cuda.synchronize(); dt = time.time() - t0
    st = s.last_stats()
    print("iter", it, "wall %.3fs" % dt, "integrate %.1f ms rconst %.1f ms" % (st["integrate_ms"], st["rconst_ms"]), "cells/s %.0f" % (n / dt), "sum_nstp", st["sum_nstp"], "ierr ok", bool((out[3] == 1).all()))
PY
done
