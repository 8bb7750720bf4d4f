"""GPU bring-up check of the warp-per-cell kernel (kernel=2) against the table-driven kernel (kernel=0).
Usage: python tests/gpu_tools/warp_debug.py [small|grid]"""
import sys, time
import numpy as np
sys.path.insert(0, ".")
from geos_chem_b200 import grid, kpp

stage = sys.argv[1] if len(sys.argv) > 1 else "small"
fx = grid.load_fixture()
s = kpp.KppSolver("fullchem", 0, max_cells=1 << 18)


def run(kernel, *a, **k):
    s.set_option("kernel", kernel)
    out = s.Integrate(*a, **k)
    return out, s.last_stats()


if stage == "small":
    for n in (1, 3, 7, 64, 500):
        r = grid.replicate_fixture(n, fx)
        args = (0.0, r["dt"], r["conc"], r["rconst"], r["atol"], r["rtol"], r["icntrl"], r["rcntrl"])
        (c1, i1, r1, e1, _), st1 = run(2, *args)
        (c0, i0, r0, e0, _), st0 = run(0, *args)
        big = np.abs(c0) > 1e3
        rel = np.abs(c1 - c0)[big] / np.abs(c0[big])
        print("n=%d warp ierr %s ist %s Hexit %.6f | generic ist %s Hexit %.6f | max rel %.3e | finite %s | ms %.2f vs %.2f"
              % (n, e1[:3], i1[:, 0], r1[1, 0], i0[:, 0], r0[1, 0], rel.max() if rel.size else -1, np.isfinite(c1).all(),
                 st1["integrate_ms"], st0["integrate_ms"]), flush=True)
g = grid.make_grid("4x5", limit=6000)
args = (0.0, 1200.0, g["conc"], None, g["atol"], g["rtol"], g["icntrl"], g["rcntrl"])
kw = dict(hstart=g["hstart"], TEMP=g["temp"], NUMDEN=g["numden"], H2O=g["h2o"], PHOTOL=g["photol"], khet=g["khet"])
(c1, i1, r1, e1, _), st1 = run(2, *args, **kw)
(c0, i0, r0, e0, _), st0 = run(0, *args, **kw)
big = np.abs(c0) > 1e3
rel = np.zeros_like(c0); rel[big] = np.abs(c1 - c0)[big] / np.abs(c0[big])
same = np.all(i1 == i0, axis=0)
print("grid 6000: ierr equal %s, cells with different steps %d, max rel %.3e (cell %d), ms warp %.1f generic %.1f"
      % (np.array_equal(e1, e0), int((~same).sum()), rel.max(), int(np.argmax(rel.max(axis=0))), st1["integrate_ms"], st0["integrate_ms"]), flush=True)
print("  sum nstp warp %d generic %d" % (i1[2].sum(), i0[2].sum()))
# Hg through the same kernel
h = kpp.KppSolver("Hg", 0, max_cells=4096)
d = h.dims
rng = np.random.default_rng(20190701)
n = 1001
conc = 10.0 ** rng.uniform(2, 8, size=(d["nspec"], n))
rconst = 10.0 ** rng.uniform(-16, -11, size=(d["nreact"], n))
atol, rtol = np.full(d["nvar"], 1e-2), np.full(d["nvar"], 1e-2)
icntrl = np.zeros(20, np.int32); icntrl[[0, 2, 6, 14]] = [1, 4, 1, -1]
res = {}
for k in (2, 0):
    h.set_option("kernel", k)
    res[k] = h.Integrate(0.0, 3600.0, conc, rconst, atol, rtol, icntrl, np.zeros(20))
    print("Hg kernel", k, "ms", h.last_stats()["integrate_ms"], "ierr ok", bool((res[k][3] == 1).all()))
big = np.abs(res[0][0]) > 1e3
print("Hg max rel", (np.abs(res[2][0] - res[0][0])[big] / np.abs(res[0][0][big])).max(), "steps differ", int((~np.all(res[2][1] == res[0][1], axis=0)).sum()))
