"""The bundle streams of the warp-per-cell kernel (geos_chem_b200/kppgen/wsched.py) executed by a numpy
emulation of the kernel's bundle engine must reproduce the oracle's Fun, Jac_SP, KppDecomp + KppSolve
(KPP/fullchem/gckpp_Function.F90, gckpp_Jacobian.F90, gckpp_LinearAlgebra.F90:46-83, 644-2309)."""
import numpy as np
import pytest

from geos_chem_b200 import grid
from geos_chem_b200.kppgen import ir, wsched
from oracle.pyoracle import Oracle
from test_sched import eval_terms


@pytest.mark.parametrize("mech", ["fullchem", "Hg"])
def test_wschedule_matches_oracle(mech):
    m = ir.load(mech)
    s = wsched.WSchedule(m)
    o = Oracle()
    rng = np.random.default_rng(11)
    if mech == "fullchem":
        fx = grid.load_fixture()
        C = fx["C"] * 10 ** rng.uniform(-0.3, 0.3, m.nspec)
        R = fx["R"].copy()
    else:
        C = 10 ** rng.uniform(3, 9, m.nspec)
        R = 10 ** rng.uniform(-14, -10, m.nreact)
    V, F = C[:m.nvar], C[m.nvar:]
    vdot_o, A_o = o.fun(mech, C, R)
    A = eval_terms(m.A, V, F, R)
    vdot = s.emulate_fun(A)
    big = np.zeros(m.nvar)
    for i, e in enumerate(m.Vdot):
        for t in e:
            c, idx = 1.0, None
            for k, v in t.factors:
                if k == "N":
                    c = float(v)
                else:
                    idx = v
            big[i] = max(big[i], abs(c * A[idx]))
    assert np.all(np.abs(vdot - vdot_o) <= 4e-10 * (big + 1e-300))
    B = eval_terms(m.B, V, F, R)
    jvs_o = o.jac(mech, C, R)
    ghinv = 1.0 / (300.0 * 0.5)
    G = s.emulate_jac(B, ghinv)
    assert G[-1] == 0.0                       # the zero slot the padding terms point at
    Gref = -jvs_o.copy()
    Gref[np.array(m.lu_diag)] += ghinv
    np.testing.assert_allclose(G[:-1], Gref, rtol=1e-11, atol=1e-13 * np.abs(Gref).max())
    lu_o, ier = o.decomp(mech, Gref)
    assert ier == 0
    b = rng.standard_normal(m.nvar) * np.abs(vdot_o).max()
    x_o = o.solve(mech, lu_o, b)
    Glu, sing = s.emulate_lu(np.append(Gref, 0.0))
    assert not sing and Glu[-1] == 0.0
    x = s.emulate_solve(Glu, b.copy())
    np.testing.assert_allclose(x, x_o, rtol=1e-9, atol=1e-12 * np.abs(x_o).max())
    d = np.array(m.lu_diag)
    np.testing.assert_allclose(Glu[:-1][d], 1.0 / lu_o[d], rtol=1e-10)


def test_wschedule_structure():
    """every target of a phase is written once; a bundle never reads what its own level writes"""
    m = ir.load("fullchem")
    s = wsched.WSchedule(m)
    for name in ("lu", "fwd", "bwd"):
        seen = set()
        level_writes = set()
        for b in s.phase[name]:
            if b.sync:
                level_writes = set()
            reads = set()
            for l in range(32):
                if b.flags[l] & wsched.F_WRITE:
                    t = b.hdr[l] & 0xffff
                    assert t not in seen
                    seen.add(t)
                for hi, lo in b.terms[l]:
                    reads.add(("G", hi))
                    reads.add(("G" if name == "lu" else "X", lo))
                if b.flags[l] & wsched.F_MUL:
                    reads.add(("G", b.hdr[l] >> 16))
            tk = "G" if name == "lu" else "X"
            assert not (reads & {(tk, t) for t in level_writes})
            for l in range(32):
                if b.flags[l] & wsched.F_WRITE:
                    level_writes.add(b.hdr[l] & 0xffff)
    # singular matrix is flagged
    G = np.zeros(s.nnz + 1)
    G[np.array(m.lu_diag)] = 1.0
    G[m.lu_diag[5]] = 0.0
    _, sing = s.emulate_lu(G)
    assert sing
