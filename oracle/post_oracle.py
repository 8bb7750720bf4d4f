"""numpy restatement of the pieces of Do_FullChem around the integration (TEST INFRASTRUCTURE ONLY, like the rest
of oracle/).  Plain IEEE operations in the reference's order, cell-fastest arrays [row, cell].  Parity unpinned by
the reference (it ships no vectors for these lines); the OH-reactivity term list is machine-extracted from the
reference's generated source (tools/extract_mech.py)."""
import numpy as np


def zero_species(conc, ids0):
    """GeosCore/fullchem_mod.F90:941-946: DO F = 1, NFAM; IF (KppID > 0) C(KppID) = 0"""
    c = conc.copy()
    for k in ids0:
        c[k] = 0.0
    return c


def post_integrate(conc, scale_ids0=(), scale_div=(), spc_mask=None, negatives=None):
    """fullchem_ConvertEquivToAlk (KPP/fullchem/fullchem_SulfurChemFuncs.F90:94-105: C = C / (MW * 7.0e-5)), then
    GeosCore/fullchem_mod.F90:1326-1348: count negatives (REAL*4 diagnostic) and C = MAX(C, 0) for mapped species"""
    c = conc.copy()
    for k, d in zip(scale_ids0, scale_div):
        c[k] = c[k] / d
    neg = None if negatives is None else negatives.astype(np.float32).copy()
    cnt = np.zeros(c.shape[1], np.float32)
    for s in range(c.shape[0]):
        if spc_mask is not None and not spc_mask[s]:
            continue
        isneg = c[s] < 0.0
        cnt += isneg.astype(np.float32)
        c[s] = np.maximum(c[s], 0.0)
    if neg is not None:
        neg += cnt
    return c, neg


def prod_loss(conc, dt, ids0):
    """GeosCore/fullchem_mod.F90:1463-1492: Loss(I,J,L,S) = C(KppID) / DT"""
    return np.stack([conc[k] / dt for k in ids0]) if len(ids0) else np.zeros((0, conc.shape[1]))


def oh_reactivity(terms, conc, rconst):
    """Get_OHreactivity (KPP/fullchem/gckpp_Util.F90:983-1040): the generated sum, left to right"""
    acc = None
    for coef, r, s in terms:
        v = rconst[r]
        if coef != 1.0:
            v = coef * v
        if s >= 0:
            v = v * conc[s]
        acc = v.copy() if acc is None else acc + v
    return acc
