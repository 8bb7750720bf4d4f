"""The static round/bundle schedules (geos_chem_b200/kppgen/sched.py) executed by a numpy emulation of
the kernel's bundle engine must reproduce the oracle's Fun, Jac_SP, KppDecomp+KppSolve
(KPP/fullchem/gckpp_Function.F90, gckpp_Jacobian.F90, gckpp_LinearAlgebra.F90:46-83, 644-2309)."""
import numpy as np
import pytest

from geos_chem_b200 import grid
from geos_chem_b200.kppgen import ir, sched
from oracle.pyoracle import Oracle


def eval_terms(terms_list, V, F, R):
    out = np.zeros(len(terms_list))
    for i, e in enumerate(terms_list):
        if e is None:
            continue
        assert len(e) == 1
        x = 1.0
        for k, v in e[0].factors:
            x *= {"R": lambda: R[v], "V": lambda: V[v], "F": lambda: F[v], "N": lambda: float(v)}[k]()
        out[i] = x
    return out


@pytest.mark.parametrize("mech", ["fullchem", "Hg"])
def test_schedule_matches_oracle(mech):
    m = ir.load(mech)
    s = sched.Schedule(m)
    o = Oracle()
    rng = np.random.default_rng(7)
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
    np.testing.assert_allclose(A, A_o, rtol=1e-14)
    vdot = s.emulate_fun(A)
    scale = np.abs(vdot_o) + 1e-30
    # aggregate, re-associated sums: compare against the magnitude of the largest contribution
    big = np.zeros(m.nvar)
    for i, e in enumerate(m.Vdot):
        for t in e:
            c = 1.0
            idx = None
            for k, v in t.factors:
                if k == "N":
                    c = float(v)
                else:
                    idx = v
            big[i] = max(big[i], abs(c * A[idx]))
    assert np.all(np.abs(vdot - vdot_o) <= 1e-12 * (big + 1e-300) * 400), np.max(np.abs(vdot - vdot_o) / (big + 1e-300))
    # Jacobian
    B = eval_terms(m.B, V, F, R)
    jvs_o = o.jac(mech, C, R)
    H, gam = 300.0, 0.5
    ghinv = 1.0 / (H * gam)
    G = s.emulate_jac(B, ghinv)
    assert G[-1] == 0.0          # the zero slot the padding terms point at
    G = G[:-1]
    Gref = -jvs_o.copy()
    Gref[np.array(m.lu_diag)] += ghinv
    # RECIP_DIAG tables: the diagonals of head pivots that no LU update touches leave the Jacobian round as reciprocals
    Gcmp = G.copy()
    Gcmp[s.inv_jvs] = 1.0 / Gcmp[s.inv_jvs]
    np.testing.assert_allclose(Gcmp, Gref, rtol=1e-11, atol=1e-13 * np.abs(Gref).max())
    # LU + solve against the oracle's KppDecomp / KppSolve
    lu_o, ier = o.decomp(mech, Gref)
    assert ier == 0
    b = rng.standard_normal(m.nvar) * np.abs(vdot_o).max()
    x_o = o.solve(mech, lu_o, b)
    Gin = np.append(Gref, 0.0)
    Gin[s.inv_jvs] = 1.0 / Gin[s.inv_jvs]
    Glu = s.emulate_lu(Gin)
    x = s.emulate_solve(Glu, b.copy())
    assert Glu[-1] == 0.0
    Glu = Glu[:-1]
    np.testing.assert_allclose(x, x_o, rtol=1e-9, atol=1e-12 * np.abs(x_o).max())
    # L multipliers are the reference's; U rows are the reference's divided by the pivot
    d = np.array(m.lu_diag)
    np.testing.assert_allclose(Glu[d], 1.0 / lu_o[d], rtol=1e-10)


def test_bank_aware_placement_reduces_modelled_conflicts():
    """kppgen/sched.py: the shared-memory bank model (64-bit gathers, half-warp by half-warp, 16 bank pairs) and the term
    placement against it.  The placement must (1) leave every bundle's terms intact as a multiset, (2) not lengthen any
    bundle, (3) cut the modelled wavefronts of the LU and sweep rounds by at least 15 % (measured on B200: bank conflicts
    839 M -> 550 M, profiles/r02aj_ncu_*.csv)."""
    m = ir.load("fullchem")
    saved = sched.BANK_OPT
    try:
        sched.BANK_OPT = 0
        s0 = sched.Schedule(m)
        sched.BANK_OPT = 1
        s1 = sched.Schedule(m)
    finally:
        sched.BANK_OPT = saved
    assert len(s0.bundles) == len(s1.bundles) and s0.rounds == s1.rounds

    def terms(s, r):
        b0, b1, kind = s.rounds[r]
        out = []
        for b in range(b0, b1):
            B = s.bundles[b]
            for lane in range(32):
                row = B.lw[lane] & 0x1fff
                out += [(row, w) for w in B.pieces[lane] if w != B.pad]
        return sorted(out)

    for name in ("lu", "fwd", "bwd"):
        w0 = w1 = 0
        for r in range(*s0.phase[name]):
            b0, b1, kind = s0.rounds[r]
            if kind & sched.K_DIV:
                continue
            # the same (target row, term) pairs, however the lanes and steps were rearranged
            assert terms(s0, r) == terms(s1, r), (name, r)
            for b in range(b0, b1):
                assert s1.bundles[b].maxlen == s0.bundles[b].maxlen
            w0 += sum(sched.bundle_wavefronts(s0.bundles[b], kind)[0] for b in range(b0, b1))
            w1 += sum(sched.bundle_wavefronts(s1.bundles[b], kind)[0] for b in range(b0, b1))
        assert w1 <= 0.85 * w0, (name, w0, w1)


def test_reciprocal_diagonal_flags():
    """RECIP_DIAG tables: every head pivot's diagonal is flagged (bit 30 of the lane word) exactly once -- in the Jacobian
    round when no LU update touches it, else in the LU round of its last update -- and no tail diagonal is."""
    m = ir.load("fullchem")
    s = sched.Schedule(m)
    assert s.recip_diag
    diag = list(m.lu_diag)
    flagged = {}
    for name in ("jvs", "lu"):
        for r in range(*s.phase[name]):
            b0, b1, kind = s.rounds[r]
            if kind & sched.K_DIV:
                continue
            for b in range(b0, b1):
                B = s.bundles[b]
                for lane in range(32):
                    lw = B.lw[lane]
                    if (lw >> 28) & 1 and (lw >> 30) & 1:
                        flagged.setdefault(lw & 0x1fff, []).append((name, r))
    head = set(diag[:s.h])
    assert set(flagged) == head and all(len(v) == 1 for v in flagged.values())
    assert sorted(p for p, v in flagged.items() if v[0][0] == "jvs") == sorted(int(x) for x in s.inv_jvs)
    # a pivot flagged in an LU round must not be updated in any later round
    last = {}
    for r in range(*s.phase["lu"]):
        b0, b1, kind = s.rounds[r]
        if kind & sched.K_DIV:
            continue
        for b in range(b0, b1):
            B = s.bundles[b]
            for lane in range(32):
                if (B.lw[lane] >> 28) & 1:
                    last[B.lw[lane] & 0x1fff] = r
    for p, v in flagged.items():
        if v[0][0] == "lu":
            assert last[p] == v[0][1]
        else:
            assert p not in last
