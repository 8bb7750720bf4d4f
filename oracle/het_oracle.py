"""Scalar Python restatement of the heterogeneous rate laws the GPU evaluates itself (TEST INFRASTRUCTURE ONLY).
Transcribed from the reference's Fortran, one cell at a time (small cases only):
  KPP/fullchem/rateLawUtilFuncs.F90     Ars_L1k :77-101, kIIR1Ltd :103-140, SafeDiv :459-478, Is_SafeDiv :480-495
  KPP/fullchem/fullchem_RateLawFuncs.F90  HBrUptkBySALA/SALC :1423-1455, HO2uptk1stOrd :1461-1484, Iuptk* / Ibrkdn* :2371-2542,
                                          OHuptkBySALACl/SALCCl :3244-3280, GLYX / Epox / IEPOX / MGLY / VOC uptake :3286-3458
  GeosCore/fullchem_mod.F90:2139-2160   Set_Kpp_GridBox_Values (RELHUM, FOUR_RGASLATM_T)
Parity unpinned by the reference: it ships no vectors for these functions."""
import math
import re

DU1, SUL, BKC, ORC, SSA, SSC, SLA, IIC = 1, 8, 9, 10, 11, 12, 13, 14
CRITRH = 35.0
HET_MIN_LIFE = 1.0e-3
HET_MIN_RATE = 1.0 / HET_MIN_LIFE
RGASLATM = 8.2057e-2
BOLTZ = 1.38064852e-23
CONSVAP = 6.1078e+03 / (BOLTZ * 1e+7)


def exponent(x):
    return 0 if x == 0.0 else math.frexp(x)[1]


def safe_div(num, denom, alt):
    ediff = exponent(num) - exponent(denom)
    if ediff > 1023 or denom == 0.0:
        return alt
    if ediff < -1020:
        return 0.0
    return num / denom


def is_safe_div(num, denom):
    ediff = exponent(num) - exponent(denom)
    return not (ediff < -1020 or ediff > 1023 or denom == 0.0)


class Cell:
    """the module variables of one grid box: met scalars, the HetState fields, SR_MW and C"""

    def __init__(self, temp, numden, h2o, het, sr_mw, conc, fields):
        self.TEMP, self.NUMDEN, self.H2O = temp, numden, h2o
        self.SR_TEMP = math.sqrt(temp)
        self.FOUR_RGASLATM_T = 4.0 * RGASLATM * temp
        consexp = 17.2693882 * (temp - 273.16) / (temp - 35.86)
        vpresh2o = CONSVAP * math.exp(consexp) / temp
        self.RELHUM = (h2o / vpresh2o) * 100.0
        for k, name in enumerate(fields):
            setattr(self, name, het[k])
        self.xArea = [None] + [getattr(self, "xArea%d" % k) for k in range(1, 15)]      # 1-based like the Fortran
        self.xRadi = [None] + [getattr(self, "xRadi%d" % k) for k in range(1, 15)]
        self.SR_MW, self.C = sr_mw, conc

    def ars_l1k(self, area, radius, gamma, srmw):
        if gamma < 1.0e-30 or radius < 1.0e-30:
            return 0.0
        dfkg = (9.45e+17 / self.NUMDEN) * self.SR_TEMP * math.sqrt(3.472e-2 + 1.0 / (srmw * srmw))
        return area / ((radius / dfkg) + 2.749064e-4 * srmw / (gamma * self.SR_TEMP))


def kIIR1Ltd(conc_gas, conc_educt, k_source):
    if conc_educt < 1.0:
        return 0.0
    if not is_safe_div(conc_gas * k_source, conc_educt):
        return 0.0
    k_gas = k_source
    k_educt = k_gas * conc_gas / conc_educt
    kii = k_gas / conc_educt
    if k_gas > 0.0:
        life_a = safe_div(1.0, k_gas, 0.0)
        life_b = safe_div(1.0, k_educt, 0.0)
        if life_a < life_b and life_a < HET_MIN_LIFE:
            kii = safe_div(HET_MIN_RATE, conc_educt, 0.0)
        elif life_b < HET_MIN_LIFE:
            kii = safe_div(HET_MIN_RATE, conc_gas, 0.0)
    return kii


def VOCuptk1stOrd(c, ind, srmw, gamma):
    k = 0.0
    if c.RELHUM >= CRITRH:
        for a in (SUL, BKC, ORC, SSA, SSC, SLA, IIC):
            k = k + c.ars_l1k(c.xArea[a], c.xRadi[a], gamma, srmw)
    return k


def EpoxUptkGamma(c, srmw):
    aervol = (c.xArea[SUL] * c.xRadi[SUL]) / 3.0
    xmms = math.sqrt((2.117e+8 * c.TEMP) / (srmw * srmw))
    kpart = (3.6e-2 * c.H_PLUS) + (2.0e-4 * c.H_PLUS * (c.NO3_molal + c.SO4_molal)) + (7.3e-4 * c.HSO4_molal) + 0.0
    val1 = (c.xRadi[SUL] * xmms) / (4.0 * 1.0e-1)
    val2 = 1.0 / 1.0e-1
    valtmp = 0.0
    if c.xArea[SUL] > 0.0 and xmms > 0.0:
        valtmp = (c.FOUR_RGASLATM_T * aervol * 1.7e+7 * kpart) / (c.xArea[SUL] * xmms)
    val3 = 1.0 / valtmp if valtmp > 0.0 else 0.0
    gamma = 1.0 / (val1 + val2 + val3) if kpart >= 1.e-8 else 0.0
    return max(gamma, 0.0)


def IEPOXuptk1stOrd(c, ind, srmw, do_scale):
    k = 0.0
    if c.RELHUM >= CRITRH:
        gamma = EpoxUptkGamma(c, srmw)
        if do_scale and c.H_PLUS > 8.0e-5:
            gamma = gamma / 30.0
        k = k + c.ars_l1k(c.xArea[SUL], c.xRadi[SUL], gamma, srmw)
    return k


def MGLYuptk1stOrd(c, ind, srmw):
    return (0.0 + c.ars_l1k(c.xArea[SUL], c.xRadi[SUL], 3.6e-7, srmw)) if c.RELHUM >= CRITRH else 0.0


def GLYXuptk1stOrd(c, ind, srmw):
    if c.RELHUM >= CRITRH:
        gamma = 4.4e-3 if c.SUNCOS > 0.0 else 8.0e-6
        return 0.0 + c.ars_l1k(c.xArea[SUL], c.xRadi[SUL], gamma, srmw)
    return 0.0


def IuptkBySulf1stOrd(c, ind, srmw, gamma):
    return c.ars_l1k(c.xArea[SUL], c.xRadi[SUL], gamma, srmw) + c.ars_l1k(c.xArea[SLA], c.xRadi[SLA], gamma, srmw)


def IuptkBySALA1stOrd(c, ind, srmw, gamma):
    return 0.0 if c.stratBox else c.ars_l1k(c.xArea[SSA], c.xRadi[SSA], gamma, srmw)


def IuptkByAlkSALA1stOrd(c, ind, srmw, gamma):
    if c.stratBox or not c.SSA_is_Alk:
        return 0.0
    return c.ars_l1k(c.f_Alk_SSA * c.xArea[SSA], c.xRadi[SSA], gamma, srmw)


def IuptkBySALC1stOrd(c, ind, srmw, gamma):
    return 0.0 if c.stratBox else c.ars_l1k(c.xArea[SSC], c.xRadi[SSC], gamma, srmw)


def IuptkByAlkSALC1stOrd(c, ind, srmw, gamma):
    if c.stratBox or not c.SSC_is_Alk:
        return 0.0
    return c.ars_l1k(c.f_Alk_SSC * c.xArea[SSC], c.xRadi[SSC], gamma, srmw)


def _ibrkdn(c, srmw, conc, gamma, acid, frac, a, yld, educt):
    if c.stratBox or not acid:
        return 0.0
    k = yld * c.ars_l1k(frac * c.xArea[a], c.xRadi[a], gamma, srmw)
    return kIIR1Ltd(conc, educt, k)


def IbrkdnByAcidBrSALA(c, ind, srmw, conc, gamma):
    return _ibrkdn(c, srmw, conc, gamma, c.SSA_is_Acid, c.f_Acid_SSA, SSA, 0.15, c.C[ind["BrSALA"]])


def IbrkdnByAcidBrSALC(c, ind, srmw, conc, gamma):
    return _ibrkdn(c, srmw, conc, gamma, c.SSC_is_Acid, c.f_Acid_SSC, SSC, 0.15, c.C[ind["BrSALC"]])


def IbrkdnByAcidSALACl(c, ind, srmw, conc, gamma):
    return _ibrkdn(c, srmw, conc, gamma, c.SSA_is_Acid, c.f_Acid_SSA, SSA, 0.85, c.C[ind["SALACL"]])


def IbrkdnByAcidSALCCl(c, ind, srmw, conc, gamma):
    return _ibrkdn(c, srmw, conc, gamma, c.SSC_is_Acid, c.f_Acid_SSC, SSC, 0.85, c.C[ind["SALCCL"]])


def HO2uptk1stOrd(c, ind):
    srmw = c.SR_MW[ind["HO2"]]
    k = 0.0
    for a in (1, 2, 3, 4, 5, 6, 7, SUL, BKC, ORC, SSA, SSC):
        k = k + c.ars_l1k(c.xArea[a], c.xRadi[a], c.gamma_HO2, srmw)
    return k


def HBrUptkBySALA(c, ind):
    if c.stratBox:
        return 0.0
    gamma = 1.3e-8 * math.exp(4290.0 / c.TEMP)
    return c.ars_l1k(c.ClearFr * c.aClArea, c.aClRadi, gamma, c.SR_MW[ind["HBr"]])


def HBrUptkBySALC(c, ind):
    if c.stratBox:
        return 0.0
    gamma = 1.3e-8 * math.exp(4290.0 / c.TEMP)
    return c.ars_l1k(c.ClearFr * c.xArea[SSC], c.xRadi[SSC], gamma, c.SR_MW[ind["HBr"]])


def OHuptkBySALACl(c, ind):
    if c.stratBox:
        return 0.0
    k = c.ars_l1k(c.aClArea, c.aClRadi, 0.04 * c.Cl_conc_SSA, c.SR_MW[ind["OH"]])
    return kIIR1Ltd(c.C[ind["OH"]], c.C[ind["SALACL"]], k)


def OHuptkBySALCCl(c, ind):
    if c.stratBox:
        return 0.0
    k = c.ars_l1k(c.xArea[SSC], c.xRadi[SSC], 0.04 * c.Cl_conc_SSC, c.SR_MW[ind["OH"]])
    return kIIR1Ltd(c.C[ind["OH"]], c.C[ind["SALCCL"]], k)


LAWS = {f.__name__: f for f in (VOCuptk1stOrd, IEPOXuptk1stOrd, MGLYuptk1stOrd, GLYXuptk1stOrd, IuptkBySulf1stOrd,
                                IuptkBySALA1stOrd, IuptkByAlkSALA1stOrd, IuptkBySALC1stOrd, IuptkByAlkSALC1stOrd,
                                IbrkdnByAcidBrSALA, IbrkdnByAcidBrSALC, IbrkdnByAcidSALACl, IbrkdnByAcidSALCCl,
                                HO2uptk1stOrd, HBrUptkBySALA, HBrUptkBySALC, OHuptkBySALACl, OHuptkBySALCCl)}


def evaluate(rconst_exprs, ind, cell):
    """{reaction index: value} for every Update_RCONST entry whose law is restated here"""
    out = {}
    for r, e in enumerate(rconst_exprs):
        if e is None:
            continue
        m = re.fullmatch(r"(\w+)\((.*)\)", e.strip())
        if not m or m.group(1) not in LAWS:
            continue
        args = []
        for a in [x.strip() for x in m.group(2).split(",")]:
            if a == "State_Het":
                continue
            m1, m2 = re.fullmatch(r"SR_MW\[(\d+)\]", a), re.fullmatch(r"C\[(\d+)\]", a)
            args.append(cell.SR_MW[int(m1.group(1))] if m1 else cell.C[int(m2.group(1))] if m2 else float(a))
        out[r] = LAWS[m.group(1)](cell, ind, *args)
    return out
