"""Classification of the Update_RCONST statements (KPP/<mech>/gckpp_Rates.F90:408-1503) of the IR."""
import re

# gas-phase rate laws the generated Update_RCONST may call (csrc/ratelaws.cuh)
GAS_LAWS = [
    "GCARR_ab", "GCARR_ac", "GCARR_abc", "ARRPLUS_ade", "ARRPLUS_abde", "TUNPLUS_abcde", "GC_ISO1", "GC_ISO2",
    "GC_EPO_a", "GC_PAN_abab", "GC_PAN_acac", "GC_NIT", "GC_ALK", "GC_HO2HO2_acac", "GC_TBRANCH_1_acac",
    "GC_RO2HO2_aca", "GC_DMSOH_acac", "GC_GLYXNO3_ac", "GC_GLYCOH_A_a", "GC_GLYCOH_B_a", "GC_HACOH_A_ac",
    "GC_HACOH_B_ac", "GC_RO2NO_A1_ac", "GC_RO2NO_B1_ac", "GC_RO2NO_A2_aca", "GC_RO2NO_B2_aca", "GCJPLEQ_acabab",
    "GCJPLPR_aa", "GCJPLPR_aba", "GCJPLPR_abab", "GCJPLPR_abcabc", "GCJPLAC_ababac",
]


def classify_rate(expr):
    """'gas' (function of TEMP/NUMDEN/H2O only), 'photol', 'ext' (needs K_MT/K_CLD/State_Het/C:
    supplied by the caller through khet_in), 'null' (Q1: never assigned)"""
    if expr is None:
        return "null"
    if re.fullmatch(r"PHOTOL\[\d+\]", expr):
        return "photol"
    if "State_Het" in expr or "K_MT" in expr or "K_CLD" in expr or "SR_MW" in expr or re.search(r"\bC\[", expr) \
            or "k_Trop" in expr or "k_Strat" in expr or "TROP" in expr:
        return "ext"
    return "gas"


def rate_layout(mech):
    """index maps for Update_RCONST: which entries are gas / photol / external; external ones are
    numbered in ascending reaction order (fullchem: 113 = K_MT(6) + K_CLD(6) + 101 heterogeneous)"""
    gas, phot, ext, null = [], [], [], []
    for r, e in enumerate(mech.rconst):
        c = classify_rate(e)
        if c == "gas":
            gas.append(r)
        elif c == "photol":
            phot.append((r, int(re.findall(r"\d+", e)[0])))
        elif c == "ext":
            ext.append(r)
        else:
            null.append(r)
    nphot = max([k for _, k in phot], default=-1) + 1
    return gas, phot, ext, null, nphot


