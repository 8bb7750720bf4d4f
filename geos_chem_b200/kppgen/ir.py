"""Mechanism IR: loader + expression parser.

The IR (geos_chem_b200/mech/<mech>.json, written by tools/extract_mech.py) keeps every
generated statement of the reference mechanism as a normalised sum of products, with the
term order and factor order of the generated Fortran (KPP/<mech>/gckpp_Function.F90,
gckpp_Jacobian.F90).  Evaluation is left to right; both the CPU oracle and the CUDA code
are emitted from the parsed form below so the operation order is the reference's.
"""
import json
import os
import re

MECH_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "mech")

_TOK = re.compile(r"([ARVFB])(\d+)|(\d+\.?\d*(?:[eE][+-]?\d+)?|\.\d+(?:[eE][+-]?\d+)?)|([+\-*])")


class Term:
    """sign * f0 * f1 * ...  ; factor = (kind, value) with kind in A R V F B or 'N' (numeric literal text)"""
    __slots__ = ("neg", "factors")

    def __init__(self, neg, factors):
        self.neg = neg
        self.factors = factors

    def __repr__(self):
        return ("-" if self.neg else "+") + "*".join("%s%s" % f if f[0] != "N" else f[1] for f in self.factors)


def parse_expr(s):
    """'A3+0.44*A7-R5*2*V9' -> [Term, ...] ; '0' -> []"""
    if s is None:
        return None
    terms = []
    neg = False
    cur = []
    expect_factor = True
    for m in _TOK.finditer(s):
        if m.group(1):
            cur.append((m.group(1), int(m.group(2))))
            expect_factor = False
        elif m.group(3):
            cur.append(("N", m.group(3)))
            expect_factor = False
        else:
            op = m.group(4)
            if op == "*":
                expect_factor = True
                continue
            if cur:
                terms.append(Term(neg, cur))
                cur = []
            elif not expect_factor:
                raise ValueError("bad expression %r" % s)
            neg = op == "-"
            expect_factor = True
    if cur:
        terms.append(Term(neg, cur))
    # a bare literal zero means "structural zero"
    if len(terms) == 1 and terms[0].factors == [("N", "0")]:
        return []
    return terms


class Mechanism:
    def __init__(self, name):
        path = os.path.join(MECH_DIR, name + ".json")
        with open(path) as f:
            d = json.load(f)
        self.raw = d
        self.name = d["name"]
        self.nspec, self.nvar, self.nfix = d["nspec"], d["nvar"], d["nfix"]
        self.nreact, self.lu_nonzero = d["nreact"], d["lu_nonzero"]
        self.spc_names = d["spc_names"]
        self.ind = d.get("ind", {})
        self.has_jac = "lu_icol" in d
        if self.has_jac:
            self.lu_icol, self.lu_crow, self.lu_diag = d["lu_icol"], d["lu_crow"], d["lu_diag"]
        self.A = [parse_expr(e) for e in d["A"]]
        self.P_VAR = [parse_expr(e) for e in d["P_VAR"]]
        self.D_VAR = [parse_expr(e) for e in d["D_VAR"]]
        self.Vdot = [parse_expr(e) for e in d["Vdot"]]
        if self.has_jac:
            self.B = [parse_expr(e) for e in d["B"]]
            self.JVS = [parse_expr(e) for e in d["JVS"]]
        self.rconst = d["rconst"]
        self.ohreact = d.get("ohreact", [])   # Get_OHreactivity terms [coef, reaction, species or -1] (gckpp_Util.F90)
        # which ODE-function form the mechanism's FunTemplate calls:
        #   fullchem: Fun_SPLIT (KPP/fullchem/gckpp_Integrator.F90:2503)
        #   Hg, carbon: aggregate Fun (KPP/Hg/gckpp_Integrator.F90:1342-1370, KPP/carbon/gckpp_Integrator.F90)
        self.fun_form = "split" if self.name == "fullchem" else "agg"

    # ---- derived structure -------------------------------------------------------
    def lu_schedule(self):
        """row-wise LU of KppDecomp (gckpp_LinearAlgebra.F90:46-83) as static lists:
        for each row k: [(pos_of_L_entry, pivot_row_j, [(pos_in_row_j_U, pos_in_row_k)...])]"""
        crow, diag, icol = self.lu_crow, self.lu_diag, self.lu_icol
        sched = []
        for k in range(self.nvar):
            colpos = {icol[p]: p for p in range(crow[k], crow[k + 1])}
            steps = []
            for p in range(crow[k], diag[k]):
                j = icol[p]
                upd = []
                for q in range(diag[j] + 1, crow[j + 1]):
                    c = icol[q]
                    # KPP's symbolic LU guarantees fill-in slots exist
                    upd.append((q, colpos[c]))
                steps.append((p, j, upd))
            sched.append(steps)
        return sched


def load(name):
    return Mechanism(name)
