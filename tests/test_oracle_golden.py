"""The oracle against the reference's own known-answer vector for this path:
KPP/standalone/Beijing_L1_20190701_0040.txt (committed as tests/golden/beijing_l1_20190701_0040.json
by tools/make_golden.py).  This is what pins the oracle."""
import numpy as np
import pytest

from geos_chem_b200 import grid


def test_reaction_rates_bit_equal(oracle, fx):
    """A(1:1058) written by Fun(...,Aout) in the 3-D model (kppsa_interface_mod.F90:668-672)"""
    vdot, a = oracle.fun("fullchem", fx["C"], fx["R"])
    assert np.array_equal(a, fx["A"])
    # Q1: reactions 126 and 735 run on inlined literals although RCONST holds 0 for them
    assert fx["R"][125] == 0 and fx["R"][734] == 0 and a[125] > 0 and a[734] > 0


def test_standalone_acceptance(oracle, fx):
    """kpp_standalone.F90:147-167: step count equal to the 3-D run's, Hexit within 0.1 %"""
    r = grid.replicate_fixture(1, fx)
    c, ist, rst, ierr = oracle.integrate_cell("fullchem", 0.0, r["dt"], fx["C"], fx["R"], r["atol"], r["rtol"],
                                              r["icntrl"], r["rcntrl"])
    assert ierr == 1
    assert ist[2] == fx["fileTotSteps"] == 12
    assert abs(rst[1] - fx["Hexit"]) / fx["Hexit"] <= 1e-3
    assert round(rst[1], 4) == 497.8023
    assert list(ist[:8]) == [33, 9, 12, 9, 0, 12, 48, 0]      # Q4: the 3 early rejections are not counted
    assert rst[0] == 1200.0
    # fixed species untouched; two species end slightly negative before the caller's MAX(C,0) (Q5)
    assert np.array_equal(c[353:], fx["C"][353:])
    assert (c < 0).sum() == 2
    names = fx["names"]
    for nm, v in (("O3", 1.3772380995e12), ("OH", 6.6532318280e6), ("NO2", 5.0276639423e10)):
        assert abs(c[names.index(nm)] / v - 1) < 1e-9


@pytest.mark.parametrize("hstart,expect", [(0.0, 24)])
def test_cold_start_variant(oracle, fx, hstart, expect):
    """Hstart=0 -> H starts at 1e-5 s (gckpp_Integrator.F90:420-421) and 24 steps"""
    r = grid.replicate_fixture(1, fx)
    rc = r["rcntrl"].copy(); rc[2] = hstart
    c, ist, rst, ierr = oracle.integrate_cell("fullchem", 0.0, r["dt"], fx["C"], fx["R"], r["atol"], r["rtol"], r["icntrl"], rc)
    assert ierr == 1 and ist[2] == expect


def test_option_errors(oracle, fx):
    """error conventions of Rosenbrock() (gckpp_Integrator.F90:358-477)"""
    r = grid.replicate_fixture(1, fx)
    ic = r["icntrl"].copy(); ic[2] = 9
    assert oracle.integrate_cell("fullchem", 0.0, 1200.0, fx["C"], fx["R"], r["atol"], r["rtol"], ic, r["rcntrl"])[3] == -2
    ic = r["icntrl"].copy(); ic[3] = -5
    assert oracle.integrate_cell("fullchem", 0.0, 1200.0, fx["C"], fx["R"], r["atol"], r["rtol"], ic, r["rcntrl"])[3] == -1
    rt = r["rtol"].copy(); rt[7] = 2.0
    assert oracle.integrate_cell("fullchem", 0.0, 1200.0, fx["C"], fx["R"], r["atol"], rt, r["icntrl"], r["rcntrl"])[3] == -5
    ic = r["icntrl"].copy(); ic[3] = 3      # Max_no_steps = 3 -> IERR -6 once Nstp > 3
    out = oracle.integrate_cell("fullchem", 0.0, 1200.0, fx["C"], fx["R"], r["atol"], r["rtol"], ic, r["rcntrl"])
    assert out[3] == -6 and out[1][2] == 4


def test_all_methods_agree(oracle, fx):
    """every ICNTRL(3) method integrates the fixture to the same answer within the tolerance class"""
    r = grid.replicate_fixture(1, fx)
    ref = None
    for m in (4, 1, 2, 3, 5, 6):
        ic = r["icntrl"].copy(); ic[2] = m
        c, ist, rst, ierr = oracle.integrate_cell("fullchem", 0.0, 1200.0, fx["C"], fx["R"], r["atol"], r["rtol"], ic, r["rcntrl"])
        assert ierr == 1 and rst[0] == 1200.0
        if ref is None:
            ref = c
        big = np.abs(ref) > 1e6
        assert np.max(np.abs(c - ref)[big] / np.abs(ref[big])) < 0.1


def test_update_rconst_against_fixture(oracle, fx):
    """gas-phase laws reproduce the fixture's R to the precision its header allows (T printed to 2
    decimals -> ~1e-4 through exp(c/T)); photolysis and externally supplied entries pass through exactly"""
    photol, khet = grid.fixture_inputs(fx)
    T, numden = fx["temperature_K"], fx["numden"]
    rc = oracle.update_rconst("fullchem", np.array([T]), np.array([numden]), np.array([fx["h2o_vmr"] * numden]),
                              photol[:, None], khet[:, None])[:, 0]
    gas, phot, ext, null, nphot = grid.rate_layout("fullchem")
    assert len(gas) == 766 and len(phot) == 177 and len(ext) == 113 and null == [125, 734]
    for r, k in phot:
        assert rc[r] == fx["R"][r]
    for r in ext:
        assert rc[r] == fx["R"][r]
    assert rc[125] == 0 and rc[734] == 0
    g = np.array(gas)
    rel = np.abs(rc[g] - fx["R"][g]) / np.abs(fx["R"][g])
    assert np.median(rel) < 1e-5 and rel.max() < 5e-3, (np.median(rel), rel.max())


def test_fma_build_agrees(fx):
    """the same oracle built with FMA contraction allowed follows the same step sequence"""
    from oracle.pyoracle import Oracle
    o = Oracle("fma")
    r = grid.replicate_fixture(1, fx)
    c, ist, rst, ierr = o.integrate_cell("fullchem", 0.0, r["dt"], fx["C"], fx["R"], r["atol"], r["rtol"], r["icntrl"], r["rcntrl"])
    assert ierr == 1 and ist[2] == 12 and round(rst[1], 4) == 497.8023


def test_autoreduce_oracle_runs_and_keeps_halogens(oracle, fx):
    """auto-reduce restatement (ros_yIntegrator): with the threshold at 0 nothing is removed and the result is the
    standard integrator's bit for bit; with the reference's OH-target setting most species change, the kept
    halogens are integrated implicitly, and RSTATUS(NARthr) reports the threshold.  (Unpinned by the reference.)"""
    from geos_chem_b200 import grid
    names = fx["names"]
    r = grid.replicate_fixture(1, fx)
    c0, ist0, rst0, ierr0 = oracle.integrate_cell("fullchem", 0.0, r["dt"], r["conc"][:, 0], r["rconst"][:, 0], r["atol"],
                                                  r["rtol"], r["icntrl"], r["rcntrl"])
    ic, rc = r["icntrl"].copy(), r["rcntrl"].copy()
    ic[11] = 1
    ic[13] = 0                           # the sample's own ICNTRL(14) = ind_OH selects the target-species threshold
    rc[11] = 1e-300                      # nothing is below this threshold
    c1, ist1, rst1, ierr1 = oracle.integrate_cell("fullchem", 0.0, r["dt"], r["conc"][:, 0], r["rconst"][:, 0], r["atol"],
                                                  r["rtol"], ic, rc)
    assert ierr1 == 1 and np.array_equal(ist1[:8], ist0[:8]) and np.array_equal(c1, c0)
    ic[13] = [n.upper() for n in names].index("OH") + 1
    rc[11], rc[13] = 0.0, 5e-5
    oracle.set_keep_active("fullchem", grid.keep_active_indices(names[:353]))
    try:
        c2, ist2, rst2, ierr2 = oracle.integrate_cell("fullchem", 0.0, r["dt"], r["conc"][:, 0], r["rconst"][:, 0],
                                                      r["atol"], r["rtol"], ic, rc)
    finally:
        oracle.set_keep_active("fullchem", [])
    assert ierr2 == 1 and rst2[3] > 0.0 and ist2[2] >= 1
    assert np.mean(c2[:353] != c0[:353]) > 0.5
    big = np.abs(c0) > 1e8               # the abundant species barely notice the reduction
    assert np.max(np.abs(c2 - c0)[big] / np.abs(c0[big])) < 5e-2
