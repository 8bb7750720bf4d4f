"""one launch of the shared-memory kernel on the replicated fixture (for ncu captures)"""
import sys
import numpy as np
sys.path.insert(0, ".")
from geos_chem_b200 import grid, kpp
n = int(sys.argv[1]) if len(sys.argv) > 1 else 444
fx = grid.load_fixture()
s = kpp.KppSolver("fullchem", 0, max_cells=1 << 16)
s.set_option("kernel", int(sys.argv[2]) if len(sys.argv) > 2 else 2)
r = grid.replicate_fixture(n, fx)
out = s.Integrate(0.0, r["dt"], r["conc"], r["rconst"], r["atol"], r["rtol"], r["icntrl"], r["rcntrl"])
print("ierr ok", bool((out[3] == 1).all()), "nstp", out[1][2, 0], "ms", s.last_stats()["integrate_ms"])
