"""Ad-hoc checker: the wave-tiled host entry against the one-pass call, per kernel (how many cells differ, where)."""
import sys
import numpy as np
sys.path.insert(0, ".")
from geos_chem_b200 import grid, kpp

n = int(sys.argv[1]) if len(sys.argv) > 1 else 24001
g = grid.make_grid("4x5", limit=n)
s = kpp.KppSolver("fullchem", 0, max_cells=n)
args = (0.0, 1200.0, g["conc"], None, g["atol"], g["rtol"], g["icntrl"], g["rcntrl"])
kw = dict(hstart=g["hstart"], TEMP=g["temp"], NUMDEN=g["numden"], H2O=g["h2o"], PHOTOL=g["photol"], khet=g["khet"])
for kern in (2, 1, 0):
    s.set_option("kernel", kern)
    s.set_option("chunks", 1)
    c1, i1, r1, e1, _ = s.Integrate(*args, **kw)
    c1b, i1b, r1b, e1b, _ = s.Integrate(*args, **kw)
    print("kernel %d: repeat of the one-wave call: cells differing %d" % (kern, int((c1 != c1b).any(axis=0).sum())))
    for k in (4, 7):
        s.set_option("chunks", k)
        c, i, r, e, _ = s.Integrate(*args, **kw)
        bad = np.nonzero((c != c1).any(axis=0) | (i != i1).any(axis=0) | (e != e1))[0]
        print("kernel %d chunks %d: cells differing %d of %d; first %s; ierr there %s; max rel %.3e; waves %s" % (
            kern, k, bad.size, n, bad[:8], e[bad[:8]], (np.abs(c - c1) / np.maximum(np.abs(c1), 1e-300)).max() if bad.size else 0.0,
            s.last_stats().get("waves")))
        if bad.size:
            print("   nstp one-wave %s waves %s" % (i1[2, bad[:8]], i[2, bad[:8]]))
