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
    assert G[-2] == 0.0 and G[-1] == 1.0      # the slots padding terms / unscaled targets point at
    Gref = -jvs_o.copy()
    Gref[np.array(m.lu_diag)] += ghinv
    np.testing.assert_allclose(G[:-2], Gref, rtol=1e-11, atol=1e-13 * np.abs(Gref).max())
    lu_o, ier = o.decomp(mech, Gref)
    assert ier == 0
    b = rng.standard_normal(m.nvar) * np.abs(vdot_o).max()
    x_o = o.solve(mech, lu_o, b)
    Glu, sing = s.emulate_lu(s.gbuf(Gref))
    assert not sing and Glu[-2] == 0.0 and Glu[-1] == 1.0
    x = s.emulate_solve(Glu, b.copy())
    np.testing.assert_allclose(x, x_o, rtol=1e-9, atol=1e-12 * np.abs(x_o).max())
    d = np.array(m.lu_diag)
    np.testing.assert_allclose(Glu[:-2][d], 1.0 / lu_o[d], rtol=1e-10)


def test_wschedule_structure():
    """every target of a phase is written once; a bundle never reads what its own level writes"""
    m = ir.load("fullchem")
    s = wsched.WSchedule(m)
    for name in ("lu", "fwd", "bwd"):
        seen = set()
        levels = []
        for b in s.phase[name]:
            if b.sync:
                levels.append((set(), set()))
            reads, writes = levels[-1]
            for l in range(32):
                if b.flags[l] & wsched.F_WRITE:
                    t = b.hdr[l] & 0xffff
                    assert t not in seen
                    seen.add(t)
                    writes.add(t)
                for hi, lo in b.terms[l]:
                    reads.add(("G", hi))
                    reads.add(("G" if name == "lu" else "X", lo))
                if b.flags[l] & wsched.F_MUL:
                    reads.add(("G", b.hdr[l] >> 16))
        tk = "G" if name == "lu" else "X"
        for reads, writes in levels:          # the bundles of a level run concurrently on the warps of a group
            assert not (reads & {(tk, t) for t in writes})
    # singular matrix is flagged
    G = s.gbuf()
    G[np.array(m.lu_diag)] = 1.0
    G[m.lu_diag[5]] = 0.0
    _, sing = s.emulate_lu(G)
    assert sing


@pytest.mark.parametrize("mech", ["fullchem", "Hg"])
def test_host_plan_replay(mech, lib):
    """The per-warp streams the C++ host plan builds (bundles of a dependency level dealt over the warps of a group,
    group barrier after a warp's last bundle of a level) replayed level by level must give what the single
    stream gives: same LU factors and solution, and every warp-stream must hold the same number of barriers."""
    from geos_chem_b200 import kpp
    m = ir.load(mech)
    s = wsched.WSchedule(m)
    plan = kpp.warp_plan(mech)
    wg = plan["wg"]
    assert wg == (4 if mech == "fullchem" else 1)
    SEG = ["vdot", "jvs", "jvs2", "lu", "fwd", "bwd", "fwd", "bwd", "vdot", "fwd", "bwd", "vdot", "fwd", "bwd"]
    PH = wsched.PHASES

    def seg_levels(w, seg):
        """row chunks of warp-stream w for segment seg, one chunk (possibly empty) per dependency level"""
        W = plan["warps"][w]
        rows = plan["stream"][W["off"]:W["off"] + W["rows"]]
        ph = PH.index(SEG[seg])
        nb, nlev = W["nb"][ph], plan["nlev"][ph]
        r = W["seg_off"][seg] if nb else None
        out, cur = [], []
        for _ in range(nb):
            meta = int(rows[r, 0, 1]); T = meta & 63
            n = 1 + (max(T - 2, 0) + 3) // 4
            pre = (meta >> 13) & 31
            assert not (pre and cur), "barriers before a bundle only at the start of a level"
            out.extend([None] * pre)             # levels without a bundle of this warp: barrier only
            cur.append(rows[r:r + n])
            r += n
            if meta & wsched.F_SYNC:
                out.append(np.concatenate(cur)); cur = []
        assert not cur, "a warp-stream must end every level with a barrier"
        assert len(out) <= nlev
        out.extend([None] * (nlev - len(out)))    # trailing barriers are made up by the kernel
        if nb:                                    # segments follow each other in the cyclic stream
            k = (seg + 1) % len(SEG)
            while W["nb"][PH.index(SEG[k])] == 0:
                k = (k + 1) % len(SEG)
            assert r % W["rows"] == W["seg_off"][k]
        return out

    def replay(seg, hi, lo, tgt, mode, ghinv=0.0):
        lv = [seg_levels(w, seg) for w in range(wg)]
        assert len({len(x) for x in lv}) == 1, "all warps of a group execute the same number of barriers"
        sing = False
        for i in range(len(lv[0])):
            for w in reversed(range(wg)):          # any order within a level
                if lv[w][i] is not None:
                    sg, _ = wsched.run_rows(lv[w][i], hi, lo, tgt, mode, ghinv)
                    sing |= sg
        return sing

    if wg > 1:      # table rows are spread evenly over the warps of a group
        rows = [W["rows"] for W in plan["warps"]]
        assert max(rows) < 1.25 * min(rows), rows

    rng = np.random.default_rng(5)
    A = rng.standard_normal(m.nreact)
    X = s.xbuf(np.zeros(m.nvar))
    replay(0, s.coefs, A, X, "vdot")
    np.testing.assert_array_equal(X[:m.nvar], s.emulate_fun(A))
    X2 = s.xbuf(np.zeros(m.nvar))
    replay(8, s.coefs, A, X2, "vdot")
    np.testing.assert_array_equal(X2, X)
    B = rng.standard_normal(len(m.B))
    Bp = np.concatenate([B, np.zeros(2 * s.nscr - len(B))])
    G = s.gbuf()
    replay(1, s.coefs, Bp[:s.nscr], G, "jvs", 0.25)
    replay(2, s.coefs, Bp[s.nscr:], G, "jvs", 0.25)
    np.testing.assert_array_equal(G, s.emulate_jac(B, 0.25))
    # a diagonally dominant matrix on the pattern
    G = s.gbuf(rng.standard_normal(s.nnz) * 0.1)
    G[np.array(m.lu_diag)] = 3.0 + rng.uniform(size=m.nvar)
    Gref, _ = s.emulate_lu(G.copy())
    Gw = G.copy()
    assert not replay(3, Gw, Gw, Gw, "lu")
    # the head phase only: compare entries outside the tail block (the tail LU is the kernel's register code)
    tail = set(int(p) for p in s.tposT.reshape(-1) if p != wsched.NONE)
    head = np.array([k for k in range(s.nnz) if k not in tail], dtype=np.int64)
    if head.size:
        np.testing.assert_array_equal(Gw[head], Gref[head])
    for seg in (4, 6, 9, 12):
        b = rng.standard_normal(m.nvar)
        xr = s.xbuf(b); wsched.run_rows(s.phase_rows("fwd"), Gref, xr, xr, "solve")
        xw = s.xbuf(b); replay(seg, Gref, xw, xw, "solve")
        np.testing.assert_array_equal(xw, xr)
        xr = s.xbuf(b); wsched.run_rows(s.phase_rows("bwd"), Gref, xr, xr, "solve")
        xw = s.xbuf(b); replay(seg + 1, Gref, xw, xw, "solve")
        np.testing.assert_array_equal(xw, xr)
