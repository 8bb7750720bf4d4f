"""Ad-hoc checker: two identical calls of one kernel must give bit-identical results."""
import sys
import numpy as np
sys.path.insert(0, ".")
from geos_chem_b200 import grid, kpp
n = int(sys.argv[1]) if len(sys.argv) > 1 else 8000
kern = int(sys.argv[2]) if len(sys.argv) > 2 else 2
g = grid.make_grid("4x5", limit=n)
s = kpp.KppSolver("fullchem", 0, max_cells=n)
s.set_option("kernel", kern)
for kv in sys.argv[3:]:
    k, v = kv.split("="); s.set_option(k, int(v))
args = (0.0, 1200.0, g["conc"], None, g["atol"], g["rtol"], g["icntrl"], g["rcntrl"])
kw = dict(hstart=g["hstart"], TEMP=g["temp"], NUMDEN=g["numden"], H2O=g["h2o"], PHOTOL=g["photol"], khet=g["khet"])
c1 = s.Integrate(*args, **kw)[0]
c2 = s.Integrate(*args, **kw)[0]
d = (c1 != c2).any(axis=0)
print("kernel %d lib %s: cells differing between two identical calls: %d of %d, max rel %.3e" % (
    kern, kpp.os.environ.get("GCKPP_B200_LIB", "default"), int(d.sum()), n, (np.abs(c1 - c2) / np.maximum(np.abs(c1), 1e-300)).max()))
