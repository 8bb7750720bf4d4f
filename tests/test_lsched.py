"""The stream tables of the lane kernel (geos_chem_b200/kppgen/lsched.py) executed by a numpy emulation of the
kernel's consumers (same operation order, same read-before-write pipelining, same 16-batch prefetch distance)
must reproduce the oracle's Fun, Jac_SP, KppDecomp + KppSolve (KPP/fullchem/gckpp_Function.F90,
gckpp_Jacobian.F90, gckpp_LinearAlgebra.F90:46-83, 644-2309)."""
import numpy as np
import pytest

from geos_chem_b200 import grid
from geos_chem_b200.kppgen import ir, lsched
from oracle.pyoracle import Oracle


@pytest.mark.parametrize("mech", ["fullchem", "Hg"])
def test_lane_schedule_matches_oracle(mech):
    m = ir.load(mech)
    s = lsched.LaneSchedule(m)
    o = Oracle()
    rng = np.random.default_rng(5)
    if mech == "fullchem":
        fx = grid.load_fixture()
        C = fx["C"] * 10 ** rng.uniform(-0.3, 0.3, m.nspec)
        R = fx["R"].copy()
    else:
        C = 10 ** rng.uniform(3, 9, m.nspec)
        R = 10 ** rng.uniform(-14, -10, m.nreact)
    vdot_o, A_o = o.fun(mech, C, R)
    A = s.emulate_rates(s.rates_a, s.vec(C), s.rcx(R), m.nreact)
    np.testing.assert_array_equal(A[:m.nreact], A_o)
    vdot = s.emulate_fun(C, R)
    assert vdot.shape == (m.nvar,)
    scale = np.zeros(m.nvar)
    for i, e in enumerate(m.Vdot):
        for t in e:
            c = [float(v) for k, v in t.factors if k == "N"]
            idx = [v for k, v in t.factors if k == "A"][0]
            scale[i] = max(scale[i], abs((c[0] if c else 1.0) * A_o[idx]))
    assert np.all(np.abs(vdot - vdot_o) <= 4e-10 * (scale + 1e-300))
    jvs_o = o.jac(mech, C, R)
    ghinv = 1.0 / (300.0 * 0.5)
    G = s.emulate_jac(C, R, ghinv)
    Gk = -jvs_o.copy()
    Gk[np.array(m.lu_diag)] += ghinv
    Gref = s.ga_from_kpp(Gk)
    np.testing.assert_allclose(G, Gref, rtol=1e-11, atol=1e-13 * np.abs(Gref).max())
    lu_o, ier = o.decomp(mech, Gk)
    assert ier == 0
    b = rng.standard_normal(m.nvar) * np.abs(vdot_o).max()
    x_o = o.solve(mech, lu_o, b)
    ga = Gref.copy()
    assert not s.emulate_lu(ga)
    d = np.array(m.lu_diag)
    np.testing.assert_allclose(ga[np.array(s.dslot)], 1.0 / lu_o[d], rtol=1e-10)
    x = s.emulate_solve(ga, b.copy())
    np.testing.assert_allclose(x, x_o, rtol=1e-9, atol=1e-12 * np.abs(x_o).max())


def test_lane_schedule_structure():
    s = lsched.build("fullchem")
    # every GA position is loaded once per row sweep and finalised once
    tab = s.lu
    kinds = tab[:, 4] & 7
    fin = np.isin(kinds, (lsched.K_FINL, lsched.K_FIND, lsched.K_FINU))
    q = (tab[fin][:, :4] & 0x1fff)[(tab[fin][:, :4] & lsched.LU_VALID) != 0]
    assert sorted(q.tolist()) == list(range(s.ng))
    assert tab.shape[0] % lsched.CHUNK == 0 and s.fwd.shape[0] % lsched.CHUNK == 0
    print("lu", s.lu_stats, "fwd", s.fwd_stats, "bwd", s.bwd_stats)
