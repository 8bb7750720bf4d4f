"""Throughput of the small mechanisms (config 5) per integrator kernel, device time of the integration only.
usage: python tools/mech_bench.py Hg|carbon [ncells] [kernel ...]"""
import sys
import numpy as np
sys.path.insert(0, ".")
from geos_chem_b200 import grid, kpp
mech = sys.argv[1] if len(sys.argv) > 1 else "Hg"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 238464
kernels = [int(k) for k in sys.argv[3:]] or [3, 1, 0]
g = grid.make_small_mech(mech, np.arange(n))
s = kpp.KppSolver(mech, 0, max_cells=n)
ref = None
for k in kernels:
    s.set_option("kernel", k)
    best = 1e30
    for it in range(3):
        c, ist, rst, ierr, _ = s.Integrate(0.0, g["dt"], g["conc"], g["rconst"], g["atol"], g["rtol"], g["icntrl"], g["rcntrl"])
        best = min(best, s.last_stats()["integrate_ms"])
    if ref is None:
        ref = c
    rel = np.abs(c - ref)[np.abs(ref) > 1e3] / np.abs(ref[np.abs(ref) > 1e3])
    print("%s kernel %d: %.2f ms (incl. copies in waves) -> %.0f cells/s; ierr==1: %d of %d; mean Nstp %.2f; max rel diff to first kernel %.2e" % (
        mech, k, best, n / best * 1e3, int((ierr == 1).sum()), n, ist[2].mean(), rel.max() if rel.size else 0.0))
