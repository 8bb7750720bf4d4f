import sys
import numpy as np
sys.path.insert(0, ".")
from geos_chem_b200 import kpp
from oracle.pyoracle import Oracle
o = Oracle()
s = kpp.KppSolver("Hg", device=0, max_cells=4096)
d = s.dims
rng = np.random.default_rng(20190701)
n = 5
conc = 10.0 ** rng.uniform(2, 8, size=(d["nspec"], n))
rconst = 10.0 ** rng.uniform(-16, -11, size=(d["nreact"], n))
atol, rtol = np.full(d["nvar"], 1e-2), np.full(d["nvar"], 1e-2)
icntrl = np.zeros(20, np.int32); icntrl[[0, 2, 6, 14]] = [1, 4, 1, -1]
rcntrl = np.zeros(20)
co, isto, rsto, ierro = o.integrate("Hg", 0.0, 3600.0, conc, rconst, atol, rtol, icntrl, rcntrl)
print("oracle ierr", ierro, "ist", isto[:, 0], "rst", rsto[:, 0])
for k in (1, 0):
    s.set_option("kernel", k)
    c, ist, rst, ierr, _ = s.Integrate(0.0, 3600.0, conc, rconst, atol, rtol, icntrl, rcntrl)
    print("kernel", k, "ierr", ierr, "ist", ist[:, 0], "rst", rst[:, 0], "finite", np.isfinite(c).all())
