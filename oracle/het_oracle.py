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


# ---------------------------------------------------------------------------------------------------------------------
# Second part (cloud / halogen uptake): fullchem_RateLawFuncs.F90:803-3238 and rateLawUtilFuncs.F90 CloudHet :142-250,
# coth :423-432, ReactoDiff_Corr :434-453.  Transcribed function by function from the Fortran; the cell object `c`
# additionally carries MW, HENRY_K0, HENRY_CR (gckpp_Global) and the second block of HetState fields.
PI = 3.14159265358979323
CON_ATM_BAR = 1.0 / 1.01325
INV_T298 = 1.0 / 298.15
CON_R = 0.083144598
RSTARG = 8.3144598
(N2O5_plus_H2O, N2O5_plus_HCl, ClNO3_plus_H2O, ClNO3_plus_HCl, ClNO3_plus_HBr, BrNO3_plus_H2O, BrNO3_plus_HCl,
 HOCl_plus_HCl, HOCl_plus_HBr, HOBr_plus_HCl, HOBr_plus_HBr) = range(1, 12)


def _prep2(c):
    """the module variables of Set_Kpp_GridBox_Values the second part reads, and 1-based views of the array fields"""
    if getattr(c, "_prep2_done", False):
        return
    c.INV_TEMP = 1.0 / c.TEMP
    c.FOUR_R_T = 4.0 * CON_R * c.TEMP
    c.EIGHT_RSTARG_T = 8.0 * RSTARG * c.TEMP
    c.KHETI_SLA = [None] + [getattr(c, "KHETI_SLA%d" % k) for k in range(1, 12)]
    c.pHSSA = [None, c.pHSSA1, c.pHSSA2]
    c.aWater = [None, c.aWater1, c.aWater2]
    c._prep2_done = True


def coth(x):
    y = math.exp(-2.0 * x)
    return (1.0 + y) / (1.0 - y)


def ReactoDiff_Corr(radius, l):
    x = radius / l
    if x > 1000.0:
        return 1.0
    if x < 0.1:
        return x / 3.0
    return coth(x) - (1.0 / x)


def Br2_Yield(br_over_cl):
    y = 0.0
    if br_over_cl > 0.0:
        y = 0.41 * math.log10(br_over_cl) + 2.25
        y = max(min(y, 0.9), 0.0)
    return y


def CloudHet(c, srMw, gamLiq, gamIce, brLiq, brIce):
    tauc = 3600.0
    if c.CldFr < 0.0001 or (c.aLiq + c.aIce <= 0.0):
        return 0.0
    kI = kIb = 0.0
    if brLiq > 0.0:
        area = safe_div(c.aLiq, c.CldFr, 0.0)
        if area > 0.0:
            ktmp = c.ars_l1k(area, c.rLiq, gamLiq, srMw)
            kI = kI + ktmp
            kIb = kIb + (ktmp * brLiq)
    if brIce > 0.0:
        area = safe_div(c.aIce, c.CldFr, 0.0)
        if area > 0.0:
            ktmp = c.ars_l1k(area, c.rIce, gamIce, srMw)
            kI = kI + ktmp
            kIb = kIb + (ktmp * brIce)
    branch = safe_div(kIb, kI, 0.0)
    if not branch > 0.0:
        return 0.0
    kk = kI * tauc
    ff = safe_div(c.CldFr, c.ClearFr, 1.0e+30)
    ff = min(ff, 1.0e+30)
    xx = (ff - kk - 1.0) / 2.0 + math.sqrt(1.0 + ff * ff + kk * kk + 2.0 * ff + 2.0 * kk - 2.0 * ff * kk) / 2.0
    xx = max(xx, 0.0)
    kHet = kI / (1.0 + safe_div(1.0, xx, 1.0e+30))
    return kHet * branch


def _cavg(c, ind, name):
    M_X = c.MW[ind[name]] * 1.0e-3
    return math.sqrt(c.EIGHT_RSTARG_T / (PI * M_X)) * 100.0


def BrNO3uptkByH2O(c, ind):
    k = 0.0
    gamLiq = 0.0021 * c.TEMP - 0.561
    gamIce = 5.3e-4 * math.exp(1100.0 / c.TEMP)
    srMw = c.SR_MW[ind["BrNO3"]]
    gamma = gamLiq
    k = k + c.ars_l1k(c.ClearFr * c.xArea[SUL], c.xRadi[SUL], gamma, srMw)
    k = k + c.ars_l1k(c.ClearFr * c.xArea[SSA], c.xRadi[SSA], gamma, srMw)
    k = k + c.ars_l1k(c.ClearFr * c.xArea[SSC], c.xRadi[SSC], gamma, srMw)
    k = k + c.xArea[SLA] * c.KHETI_SLA[BrNO3_plus_H2O]
    gamma = 0.3
    if c.natSurface:
        gamma = 0.001
    k = k + c.ars_l1k(c.ClearFr * c.xArea[IIC], c.xRadi[IIC], gamma, srMw)
    k = k + CloudHet(c, srMw, gamLiq, gamIce, 1.0, 1.0)
    return kIIR1Ltd(c.C[ind["BrNO3"]], c.C[ind["H2O"]], k)


def BrNO3uptkByHCl(c, ind):
    k = 0.0
    srMw = c.SR_MW[ind["BrNO3"]]
    if c.stratBox:
        k = k + c.ars_l1k(c.xArea[SUL], c.xRadi[SUL], 0.9, srMw)
        k = k + c.xArea[SLA] * c.KHETI_SLA[BrNO3_plus_HCl]
        k = k + c.ars_l1k(c.xArea[IIC], c.xRadi[IIC], 0.3, srMw)
    k = kIIR1Ltd(c.C[ind["BrNO3"]], c.C[ind["HCl"]], k)
    if c.TurnOffHetRates:
        k = 0.0
    return k


def Gam_ClNO2(c, ind, radius, pH, C_Cl, C_Br):
    INV_AB = 1.0 / 0.01
    D_l = 1.0e-5
    cavg = _cavg(c, ind, "ClNO2")
    H_X = 4.5e-2 * CON_ATM_BAR
    k_Cl = 1.0e+7 * C_Cl
    if pH >= 2.0:
        k_Cl = 0.0
    k_Br = (1.01e-1 / (H_X * H_X * D_l)) * C_Br
    k_tot = k_Cl + k_Br
    gamma = branchCl = branchBr = 0.0
    if k_tot > 0.0:
        l_r = math.sqrt(D_l / k_tot)
        gb_tot = c.FOUR_R_T * H_X * l_r * k_tot / cavg
        gb_tot = gb_tot * ReactoDiff_Corr(radius, l_r)
        gamma = 1.0 / (INV_AB + 1.0 / gb_tot)
        branchCl = k_Cl / k_tot
        branchBr = k_Br / k_tot
    return gamma, branchCl, branchBr


def ClNO2uptkByBrSALA(c, ind):
    k = 0.0
    srMw = c.SR_MW[ind["ClNO2"]]
    if not c.stratBox:
        gamma, _, branchBr = Gam_ClNO2(c, ind, c.rLiq, c.pHCloud, c.Cl_conc_Cld, c.Br_conc_Cld)
        branch = branchBr * c.frac_Br_CldA
        k = k + CloudHet(c, srMw, gamma, 0.0, branch, 0.0)
    gamma, _, branchBr = Gam_ClNO2(c, ind, c.aClRadi, c.pHSSA[1], c.Cl_conc_SSA, c.Br_conc_SSA)
    area = c.ClearFr * c.aClArea
    k = k + c.ars_l1k(area, c.aClRadi, gamma, srMw) * branchBr
    return kIIR1Ltd(c.C[ind["ClNO2"]], c.C[ind["BrSALA"]], k)


def ClNO2uptkByBrSALC(c, ind):
    k = 0.0
    srMw = c.SR_MW[ind["ClNO2"]]
    if not c.stratBox:
        gamma, _, branchBr = Gam_ClNO2(c, ind, c.rLiq, c.pHCloud, c.Cl_conc_Cld, c.Br_conc_Cld)
        branch = branchBr * c.frac_Br_CldC
        k = k + CloudHet(c, srMw, gamma, 0.0, branch, 0.0)
    gamma, _, branchBr = Gam_ClNO2(c, ind, c.xRadi[SSC], c.pHSSA[2], c.Cl_conc_SSC, c.Br_conc_SSC)
    area = c.ClearFr * c.xArea[SSC]
    k = k + c.ars_l1k(area, c.xRadi[SSC], gamma, srMw) * branchBr
    return kIIR1Ltd(c.C[ind["ClNO2"]], c.C[ind["BrSALC"]], k)


def ClNO2uptkByHBr(c, ind):
    k = 0.0
    srMw = c.SR_MW[ind["ClNO2"]]
    if not c.stratBox:
        gamma, _, branchBr = Gam_ClNO2(c, ind, c.rLiq, c.pHCloud, c.Cl_conc_Cld, c.Br_conc_Cld)
        branch = branchBr * c.frac_Br_CldG
        k = k + CloudHet(c, srMw, gamma, 0.0, branch, 0.0)
    return kIIR1Ltd(c.C[ind["ClNO2"]], c.C[ind["HBr"]], k)


def ClNO2uptkBySALACL(c, ind):
    k = 0.0
    srMw = c.SR_MW[ind["ClNO2"]]
    if not c.stratBox:
        gamma, branchCl, _ = Gam_ClNO2(c, ind, c.rLiq, c.pHCloud, c.Cl_conc_Cld, c.Br_conc_Cld)
        branch = branchCl * c.frac_Cl_CldA
        k = k + CloudHet(c, srMw, gamma, 0.0, branch, 0.0)
    gamma, branchCl, _ = Gam_ClNO2(c, ind, c.aClRadi, c.pHSSA[1], c.Cl_conc_SSA, c.Br_conc_SSA)
    area = c.ClearFr * c.aClArea
    k = k + c.ars_l1k(area, c.aClRadi, gamma, srMw) * branchCl
    return kIIR1Ltd(c.C[ind["ClNO2"]], c.C[ind["SALACL"]], k)


def ClNO2uptkBySALCCL(c, ind):
    k = 0.0
    srMw = c.SR_MW[ind["ClNO2"]]
    if not c.stratBox:
        gamma, branchCl, _ = Gam_ClNO2(c, ind, c.rLiq, c.pHCloud, c.Cl_conc_Cld, c.Br_conc_Cld)
        branch = branchCl * c.frac_Cl_CldC
        k = k + CloudHet(c, srMw, gamma, 0.0, branch, 0.0)
    return kIIR1Ltd(c.C[ind["ClNO2"]], c.C[ind["SALCCL"]], k)


def ClNO2uptkByHCl(c, ind):
    k = 0.0
    srMw = c.SR_MW[ind["ClNO2"]]
    if not c.stratBox:
        gamma, branchCl, _ = Gam_ClNO2(c, ind, c.rLiq, c.pHCloud, c.Cl_conc_Cld, c.Br_conc_Cld)
        branch = branchCl * c.frac_Cl_CldG
        k = k + CloudHet(c, srMw, gamma, 0.0, branch, 0.0)
    return kIIR1Ltd(c.C[ind["ClNO2"]], c.C[ind["HCl"]], k)


def Gam_ClNO3_Aer(c, ind, C_Br):
    INV_AB = 1.0 / 0.108
    K_0 = 1.2e+5 ** 2.0
    D_l = 5.0e-6
    cavg = _cavg(c, ind, "ClNO3")
    k_Br = 1.0e+12 * C_Br
    k_tot = K_0 + k_Br
    gb_tot = c.FOUR_R_T * math.sqrt(k_tot * D_l) / cavg
    gamma = 1.0 / (INV_AB + 1.0 / gb_tot)
    return gamma, k_Br / k_tot


def Gam_ClNO3_Ice(c, ind):
    twenty = 1.0 / 0.5
    g1 = 0.24 * c.HCl_theta
    g2 = 0.56 * c.HBr_theta
    cavg = _cavg(c, ind, "ClNO3")
    H2Os = 1e+15 - (3.0 * 2.7e+14 * c.HNO3_theta)
    kks = 4.0 * 5.2e-17 * math.exp(2032.0 / c.TEMP)
    g3 = 1.0 / (twenty + cavg / (kks * H2Os))
    gamma = g1 + g2 + g3
    return gamma, g1 / gamma, g2 / gamma, g3 / gamma


def ClNO3uptkByH2O(c, ind):
    k = 0.0
    srMw = c.SR_MW[ind["ClNO3"]]
    gamma, branchBr = Gam_ClNO3_Aer(c, ind, c.Br_conc_SSA)
    area = c.ClearFr * c.aClArea
    branchLiq = (1.0 - branchBr) * (1.0 - c.frac_SALACL)
    k = k + c.ars_l1k(area, c.aClRadi, gamma, srMw) * branchLiq
    k = k + c.xArea[SLA] * c.KHETI_SLA[ClNO3_plus_H2O]
    gamma = 0.3
    if c.natSurface:
        gamma = 0.004
    k = k + c.ars_l1k(c.xArea[IIC], c.xRadi[IIC], gamma, srMw)
    if not c.stratBox:
        gamma, branchBr = Gam_ClNO3_Aer(c, ind, c.Br_conc_Cld)
        branchLiq = 1.0 - branchBr
        gammaIce, _, _, branchIce = Gam_ClNO3_Ice(c, ind)
        k = k + CloudHet(c, srMw, gamma, gammaIce, branchLiq, branchIce)
    return kIIR1Ltd(c.C[ind["ClNO3"]], c.C[ind["H2O"]], k)


def ClNO3uptkByHCl(c, ind):
    k = 0.0
    srMw = c.SR_MW[ind["ClNO3"]]
    if c.stratBox:
        gamma = 0.1e-4
        k = k + c.ars_l1k(c.xArea[SUL], c.xRadi[SUL], gamma, srMw)
        k = k + c.xArea[SLA] * c.KHETI_SLA[ClNO3_plus_HCl]
        gamma = 0.3
        if c.natSurface:
            gamma = 0.2
        k = k + c.ars_l1k(c.xArea[IIC], c.xRadi[IIC], gamma, srMw)
    else:
        gammaIce, branchIce, _, _ = Gam_ClNO3_Ice(c, ind)
        k = k + CloudHet(c, srMw, 0.0, gammaIce, 0.0, branchIce)
    return kIIR1Ltd(c.C[ind["ClNO3"]], c.C[ind["HCl"]], k)


def ClNO3uptkByHBr(c, ind):
    k = 0.0
    srMw = c.SR_MW[ind["ClNO3"]]
    if c.stratBox:
        k = k + c.xArea[SLA] * c.KHETI_SLA[ClNO3_plus_HBr]
        gamma = 0.3
        k = k + c.ars_l1k(c.xArea[IIC], c.xRadi[IIC], gamma, srMw)
    else:
        gamma, branchBr = Gam_ClNO3_Aer(c, ind, c.Br_conc_Cld)
        branchLiq = branchBr * c.frac_Br_CldG
        gammaIce, _, branchIce, _ = Gam_ClNO3_Ice(c, ind)
        k = CloudHet(c, srMw, gamma, gammaIce, branchLiq, branchIce)
    k = kIIR1Ltd(c.C[ind["ClNO3"]], c.C[ind["HBr"]], k)
    if c.TurnOffHetRates:
        k = 0.0
    return k


def ClNO3uptkByBrSALA(c, ind):
    k = 0.0
    srMw = c.SR_MW[ind["ClNO3"]]
    if not c.stratBox:
        gamma, branchBr = Gam_ClNO3_Aer(c, ind, c.Br_conc_Cld)
        branch = branchBr * c.frac_Br_CldA
        k = k + CloudHet(c, srMw, gamma, 0.0, branch, 0.0)
    gamma, branchBr = Gam_ClNO3_Aer(c, ind, c.Br_conc_SSA)
    area = c.ClearFr * c.aClArea
    k = k + c.ars_l1k(area, c.aClRadi, gamma, srMw) * branchBr
    k = kIIR1Ltd(c.C[ind["ClNO3"]], c.C[ind["BrSALA"]], k)
    if c.TurnOffHetRates:
        k = 0.0
    return k


def ClNO3uptkByBrSALC(c, ind):
    k = 0.0
    srMw = c.SR_MW[ind["ClNO3"]]
    if not c.stratBox:
        gamma, branchBr = Gam_ClNO3_Aer(c, ind, c.Br_conc_Cld)
        branch = branchBr * c.frac_Br_CldC
        k = k + CloudHet(c, srMw, gamma, 0.0, branch, 0.0)
    gamma, branchBr = Gam_ClNO3_Aer(c, ind, c.Br_conc_SSC)
    area = c.ClearFr * c.xArea[SSC]
    k = k + c.ars_l1k(area, c.xRadi[SSC], gamma, srMw) * branchBr
    k = kIIR1Ltd(c.C[ind["ClNO3"]], c.C[ind["BrSALC"]], k)
    if c.TurnOffHetRates:
        k = 0.0
    return k


def ClNO3uptkBySALACL(c, ind):
    k = 0.0
    if c.stratBox:
        return k
    gamma, branchBr = Gam_ClNO3_Aer(c, ind, c.Br_conc_SSA)
    srMw = c.SR_MW[ind["ClNO3"]]
    area = c.ClearFr * c.aClArea
    branch = (1.0 - branchBr) * c.frac_SALACL
    k = k + c.ars_l1k(area, c.aClRadi, gamma, srMw) * branch
    return kIIR1Ltd(c.C[ind["ClNO3"]], c.C[ind["SALACL"]], k)


def ClNO3uptkBySALCCL(c, ind):
    k = 0.0
    if c.stratBox:
        return k
    gamma, branchBr = Gam_ClNO3_Aer(c, ind, c.Br_conc_SSC)
    srMw = c.SR_MW[ind["ClNO3"]]
    area = c.ClearFr * c.xArea[SSC]
    branch = 1.0 - branchBr
    k = k + c.ars_l1k(area, c.xRadi[SSC], gamma, srMw) * branch
    return kIIR1Ltd(c.C[ind["ClNO3"]], c.C[ind["SALCCL"]], k)


def _henry(c, ind, name, inv_temp):
    return (c.HENRY_K0[ind[name]] * CON_ATM_BAR) * math.exp(c.HENRY_CR[ind[name]] * (inv_temp - INV_T298))


def Gam_HOBr_Aer(c, ind, radius, C_Hp, C_Clm, C_Brm):
    INV_AB = 1.0 / 0.6
    D_l = 1.4e-5
    H_X = _henry(c, ind, "HOBr", 1.0 / c.TEMP)
    cavg = _cavg(c, ind, "HOBr")
    C_Hp1 = max(min(C_Hp, 1.0e-6), 1.0e-9)
    C_Hp2 = max(min(C_Hp, 1.0e-2), 1.0e-6)
    k_HOBr_Cl = 2.3e+10 * C_Clm * C_Hp1
    k_HOBr_Br = 1.6e+10 * C_Brm * C_Hp2
    k_tot = k_HOBr_Cl + k_HOBr_Br
    gamma = 0.0
    if k_tot > 0.0:
        l_r = math.sqrt(D_l / k_tot)
        gb_tot = c.FOUR_R_T * H_X * l_r * k_tot / cavg
        gb_tot = gb_tot * ReactoDiff_Corr(radius, l_r)
        gamma = 1.0 / (INV_AB + 1.0 / gb_tot)
    return gamma


def Gam_HOBr_Cld(c, ind):
    """-> gamma, k_tot, k_HOBr_Cl, k_HOBr_Br, k_HOBr_HSO3, k_HOBr_HSO3_2"""
    INV_AB = 1.0 / 0.6
    D_l = 1.4e-5
    H_X = _henry(c, ind, "HOBr", 1.0 / c.TEMP)
    cavg = _cavg(c, ind, "HOBr")
    C_Hp1 = min(c.H_conc_LCl, 1.0e-6)
    C_Hp2 = min(c.H_conc_LCl, 1.0e-2)
    C_Hp1 = max(C_Hp1, 1.0e-9)
    C_Hp2 = max(C_Hp2, 1.0e-6)
    k_Cl = 2.3e+10 * c.Cl_conc_Cld * C_Hp1
    k_Br = 1.6e+10 * c.Br_conc_Cld * C_Hp2
    k_HSO3 = 2.6e+7 * c.HSO3_aq
    k_HSO3_2 = 5.0e+9 * c.SO3_aq
    k_tot = k_Cl + k_Br + k_HSO3 + k_HSO3_2
    gamma = 0.0
    if k_tot > 0.0:
        l_r = math.sqrt(D_l / k_tot)
        gb_tot = c.FOUR_R_T * H_X * l_r * k_tot / cavg
        gb_tot = gb_tot * ReactoDiff_Corr(c.rLiq, l_r)
        gamma = 1.0 / (INV_AB + 1.0 / gb_tot)
    return gamma, k_tot, k_Cl, k_Br, k_HSO3, k_HSO3_2


def Gam_HOBr_Ice(c):
    gamma_HCl = c.HCl_theta * 0.25
    gamma_HBr = c.HBr_theta * 4.8e-4 * math.exp(1240.0 / c.TEMP)
    gamma = gamma_HCl + gamma_HBr
    branch_HCl = branch_HBr = 0.0
    if gamma > 0.0:
        branch_HCl = gamma_HCl / gamma
        branch_HBr = gamma_HBr / gamma
    return gamma, branch_HCl, branch_HBr


def HOBrUptkByHBr(c, ind):
    k = 0.0
    srMw = c.SR_MW[ind["HOBr"]]
    if c.stratBox:
        gammaLiq = 0.25
        k = k + c.ars_l1k(c.xArea[SUL], c.xRadi[SUL], gammaLiq, srMw)
        k = k + c.xArea[SLA] * c.KHETI_SLA[HOBr_plus_HBr]
        gammaIce = 0.3
        if c.natSurface:
            gammaIce = 0.001
        k = k + c.ars_l1k(c.xArea[IIC], c.xRadi[IIC], gammaIce, srMw)
    else:
        gammaLiq, k_tot, k_Cl, k_Br, _, _ = Gam_HOBr_Cld(c, ind)
        branch_0 = (k_Cl + k_Br) / k_tot
        branch = branch_0 * 0.9
        if c.Br_over_Cl_Cld <= 5.0e-4:
            branch = branch_0 * Br2_Yield(c.Br_over_Cl_Cld)
        brLiq = branch * c.frac_Br_CldG
        gammaIce, _, brIce = Gam_HOBr_Ice(c)
        k = k + CloudHet(c, srMw, gammaLiq, gammaIce, brLiq, brIce)
    return kIIR1Ltd(c.C[ind["HOBr"]], c.C[ind["HBr"]], k)


def HOBrUptkByHCl(c, ind):
    k = 0.0
    srMw = c.SR_MW[ind["HOBr"]]
    if c.stratBox:
        gammaLiq = 0.2
        k = k + c.ars_l1k(c.xArea[SUL], c.xRadi[SUL], gammaLiq, srMw)
        k = k + c.xArea[SLA] * c.KHETI_SLA[HOBr_plus_HCl]
        gammaIce = 0.3
        if c.natSurface:
            gammaIce = 0.1
        k = k + c.ars_l1k(c.xArea[IIC], c.xRadi[IIC], gammaIce, srMw)
    else:
        gammaLiq, k_tot, k_Cl, k_Br, _, _ = Gam_HOBr_Cld(c, ind)
        branch_0 = (k_Cl + k_Br) / k_tot
        branch = branch_0 * 0.1
        if c.Br_over_Cl_Cld <= 5.0e-4:
            branch = branch_0 * (1.0 - Br2_Yield(c.Br_over_Cl_Cld))
        brLiq = branch * c.frac_Cl_CldG
        gammaIce, brIce, _ = Gam_HOBr_Ice(c)
        k = k + CloudHet(c, srMw, gammaLiq, gammaIce, brLiq, brIce)
    return kIIR1Ltd(c.C[ind["HOBr"]], c.C[ind["HCl"]], k)


def _HOBr_seasalt(c, ind, to_br, coarse):
    """HOBrUptkByBrSALA / BrSALC / SALACL / SALCCL share this shape in the Fortran (:1755-1965); the four copies
    differ in the yield (0.9 and Br2_Yield for the bromide channel, 0.1 and 1 - Br2_Yield for chloride), the cloud
    fraction field and the sea-salt mode."""
    k = 0.0
    srMw = c.SR_MW[ind["HOBr"]]
    if not c.stratBox:
        gammaLiq, k_tot, k_Cl, k_Br, _, _ = Gam_HOBr_Cld(c, ind)
        branch_0 = (k_Cl + k_Br) / k_tot
        if to_br:
            branch = branch_0 * 0.9
            if c.Br_over_Cl_Cld <= 5.0e-4:
                branch = branch_0 * Br2_Yield(c.Br_over_Cl_Cld)
            frac = c.frac_Br_CldC if coarse else c.frac_Br_CldA
        else:
            branch = branch_0 * 0.1
            if c.Br_over_Cl_Cld <= 5.0e-4:
                branch = branch_0 * (1.0 - Br2_Yield(c.Br_over_Cl_Cld))
            frac = c.frac_Cl_CldC if coarse else c.frac_Cl_CldA
        brLiq = branch * frac
        k = k + CloudHet(c, srMw, gammaLiq, 0.0, brLiq, 0.0)
    if (c.SSC_is_Acid if coarse else c.SSA_is_Acid):
        if coarse:
            gammaAer = Gam_HOBr_Aer(c, ind, c.xRadi[SSC], c.H_conc_SSC, c.Cl_conc_SSC, c.Br_conc_SSC)
            ratio = c.Br_over_Cl_SSC
        else:
            gammaAer = Gam_HOBr_Aer(c, ind, c.aClRadi, c.H_conc_SSA, c.Cl_conc_SSA, c.Br_conc_SSA)
            ratio = c.Br_over_Cl_SSA
        if to_br:
            branch = 0.9
            if ratio <= 5.0e-4:
                branch = Br2_Yield(ratio)
        else:
            branch = 0.1
            if ratio <= 5.0e-4:
                branch = 1.0 - Br2_Yield(ratio)
        if coarse:
            area = c.ClearFr * c.xArea[SSC] * c.f_Acid_SSC
            k = k + c.ars_l1k(area, c.xRadi[SSC], gammaAer, srMw) * branch
        else:
            area = c.ClearFr * c.aClArea * c.f_Acid_SSA
            k = k + c.ars_l1k(area, c.aClRadi, gammaAer, srMw) * branch
    educt = ("BrSALC" if coarse else "BrSALA") if to_br else ("SALCCL" if coarse else "SALACL")
    return kIIR1Ltd(c.C[ind["HOBr"]], c.C[ind[educt]], k)


def HOBrUptkByBrSALA(c, ind):
    return _HOBr_seasalt(c, ind, True, False)


def HOBrUptkByBrSALC(c, ind):
    return _HOBr_seasalt(c, ind, True, True)


def HOBrUptkBySALACL(c, ind):
    return _HOBr_seasalt(c, ind, False, False)


def HOBrUptkBySALCCL(c, ind):
    return _HOBr_seasalt(c, ind, False, True)


def HOBrUptkByHSO3m(c, ind):
    k = 0.0
    srMw = c.SR_MW[ind["HOBr"]]
    if not c.stratBox:
        gammaLiq, k_tot, _, _, k_HSO3m, _ = Gam_HOBr_Cld(c, ind)
        brLiq = k_HSO3m / k_tot
        k = k + CloudHet(c, srMw, gammaLiq, 0.0, brLiq, 0.0)
    return kIIR1Ltd(c.C[ind["HOBr"]], c.C[ind["SO2"]], k)


def Gam_HOCl_Cld(c, ind):
    INV_AB = 1.0 / 0.8
    D_l = 2.0e-5
    k_Cl = 1.5e+4 * c.H_conc_LCl * c.Cl_conc_Cld
    k_SO3 = 2.8e+5 * c.TSO3_aq
    k_tot = k_Cl + k_SO3
    gamma = branchCl = branchSO3 = 0.0
    if k_tot > 0.0:
        cavg = _cavg(c, ind, "HOCl")
        H_X = _henry(c, ind, "HOCl", c.INV_TEMP)
        l_r = math.sqrt(D_l / k_tot)
        gb_tot = c.FOUR_R_T * H_X * l_r * k_tot / cavg
        gb_tot = gb_tot * ReactoDiff_Corr(c.rLiq, l_r)
        gamma = 1.0 / (INV_AB + 1.0 / gb_tot)
        branchCl = k_Cl / k_tot
        branchSO3 = k_SO3 / k_tot
    return gamma, branchCl, branchSO3


def Gam_HOCl_Aer(c, ind, radius, C_Hp, C_Cl):
    INV_AB = 1.0 / 0.8
    D_l = 2.0e-5
    K_TER = 1.5e+4
    if not C_Cl > 0.0:
        return 0.0
    cavg = _cavg(c, ind, "HOCl")
    H_X = _henry(c, ind, "HOCl", c.INV_TEMP)
    l_r = math.sqrt(D_l / (K_TER * C_Hp * C_Cl))
    gb = c.FOUR_R_T * H_X * l_r * K_TER * C_Hp * C_Cl / cavg
    gb = gb * ReactoDiff_Corr(radius, l_r)
    return 1.0 / (INV_AB + 1.0 / gb)


def HOClUptkByHCl(c, ind):
    k = 0.0
    srMw = c.SR_MW[ind["HOCl"]]
    if c.stratBox:
        gamma = 0.8
        k = k + c.ars_l1k(c.xArea[SUL], c.xRadi[SUL], gamma, srMw)
        k = k + c.xArea[SLA] * c.KHETI_SLA[HOCl_plus_HCl]
        gamma = 0.2
        if c.natSurface:
            gamma = 0.1
        k = k + c.ars_l1k(c.xArea[IIC], c.xRadi[IIC], gamma, srMw)
        return kIIR1Ltd(c.C[ind["HOCl"]], c.C[ind["HCl"]], k)
    gamma, branchCl, _ = Gam_HOCl_Cld(c, ind)
    branch = branchCl * c.frac_Cl_CldG
    gammaIce = 0.22 * c.HCl_theta
    brIce = 1.0
    k = k + CloudHet(c, srMw, gamma, gammaIce, branch, brIce)
    return kIIR1Ltd(c.C[ind["HOCl"]], c.C[ind["HCl"]], k)


def HOClUptkByHBr(c, ind):
    k = 0.0
    srMw = c.SR_MW[ind["HOCl"]]
    if c.stratBox:
        gamma = 0.8
        k = k + c.ars_l1k(c.xArea[SUL], c.xRadi[SUL], gamma, srMw)
        k = k + c.xArea[SLA] * c.KHETI_SLA[HOCl_plus_HBr]
        gamma = 0.3
        k = k + c.ars_l1k(c.xArea[IIC], c.xRadi[IIC], gamma, srMw)
    k = kIIR1Ltd(c.C[ind["HOCl"]], c.C[ind["HBr"]], k)
    if c.TurnOffHetRates:
        k = 0.0
    return k


def HOClUptkBySALACL(c, ind):
    k = 0.0
    srMw = c.SR_MW[ind["HOCl"]]
    if not c.stratBox:
        gamma, branchCl, _ = Gam_HOCl_Cld(c, ind)
        branch = branchCl * c.frac_Cl_CldA
        k = k + CloudHet(c, srMw, gamma, 0.0, branch, 0.0)
    if c.SSA_is_Acid:
        gamma = Gam_HOCl_Aer(c, ind, c.aClRadi, c.H_conc_SSA, c.Cl_conc_SSA)
        area = c.ClearFr * c.aClArea * c.f_Acid_SSA
        k = k + c.ars_l1k(area, c.aClRadi, gamma, srMw)
    return kIIR1Ltd(c.C[ind["HOCl"]], c.C[ind["SALACL"]], k)


def HOClUptkBySALCCL(c, ind):
    k = 0.0
    srMw = c.SR_MW[ind["HOCl"]]
    if not c.stratBox:
        gamma, branchCl, _ = Gam_HOCl_Cld(c, ind)
        branch = branchCl * c.frac_Cl_CldC
        k = k + CloudHet(c, srMw, gamma, 0.0, branch, 0.0)
    if c.SSC_is_Acid:
        gamma = Gam_HOCl_Aer(c, ind, c.xRadi[SSC], c.H_conc_SSC, c.Cl_conc_SSC)
        area = c.ClearFr * c.xArea[SSC] * c.f_Acid_SSC
        k = k + c.ars_l1k(area, c.xRadi[SSC], gamma, srMw)
    return kIIR1Ltd(c.C[ind["HOCl"]], c.C[ind["SALCCL"]], k)


def HOClUptkByHSO3m(c, ind):
    k = 0.0
    srMw = c.SR_MW[ind["HOCl"]]
    if not c.stratBox:
        gamma, _, branchSO3 = Gam_HOCl_Cld(c, ind)
        branch = branchSO3 * c.frac_HSO3_aq
        k = k + CloudHet(c, srMw, gamma, 0.0, branch, 0.0)
    return kIIR1Ltd(c.C[ind["HOCl"]], c.C[ind["SO2"]], k) * c.HSO3m


def IONO2uptkByH2O(c, ind):
    k = 0.0
    srMw = c.SR_MW[ind["IONO2"]]
    area = c.ClearFr * c.xArea[SUL]
    gamma = max((0.0021 * c.TEMP - 0.561), 0.0)
    k = k + c.ars_l1k(area, c.xRadi[SUL], gamma, srMw)
    k = k + c.xArea[SLA] * c.KHETI_SLA[BrNO3_plus_H2O]
    area = c.ClearFr * c.xArea[IIC]
    gamma = 0.3
    if c.natSurface:
        gamma = 0.001
    k = k + c.ars_l1k(area, c.xRadi[IIC], gamma, srMw)
    k = k + CloudHet(c, srMw, 0.01, 0.01, 1.0, 1.0)
    return kIIR1Ltd(c.C[ind["IONO2"]], c.C[ind["H2O"]], k)


def N2O5uptkByCloud(c, ind):
    const = 0.03 / 0.019
    gamma = const * math.exp(-25.5265 + 9283.76 / c.TEMP - 851801.0 / c.TEMP ** 2)
    return CloudHet(c, c.SR_MW[ind["N2O5"]], gamma, 0.02, 1.0, 1.0)


def N2O5uptkByStratHCl(c, ind):
    k = 0.0
    if c.stratBox:
        k = k + (c.xArea[SLA] * c.KHETI_SLA[N2O5_plus_HCl])
        gamma = 0.03
        if c.natSurface:
            gamma = 0.003
        k = k + c.ars_l1k(c.xArea[IIC], c.xRadi[IIC], gamma, c.SR_MW[ind["N2O5"]])
    return kIIR1Ltd(c.C[ind["N2O5"]], c.C[ind["HCl"]], k)


def NO2uptk1stOrdAndCloud(c, ind):
    k = 0.0
    srMw = c.SR_MW[ind["NO2"]]
    gamma = 1.0e-8
    for a in range(DU1, DU1 + 7):
        k = k + c.ars_l1k(c.xArea[a], c.xRadi[a], gamma, srMw)
    k = k + c.ars_l1k(c.xArea[SUL], c.xRadi[SUL], 5e-6, srMw)
    k = k + c.ars_l1k(c.xArea[BKC], c.xRadi[BKC], 1e-4, srMw)
    k = k + c.ars_l1k(c.xArea[ORC], c.xRadi[ORC], 1e-6, srMw)
    if c.RELHUM < 40.0:
        gamma = 1.0e-8
    elif c.RELHUM > 70.0:
        gamma = 1.0e-4
    else:
        gamma = 1.0e-8 + (1e-4 - 1e-8) * (c.RELHUM - 40.0) / 30.0
    k = k + c.ars_l1k(c.xArea[SSA], c.xRadi[SSA], gamma, srMw)
    k = k + c.ars_l1k(c.xArea[SSC], c.xRadi[SSC], gamma, srMw)
    gamma = 1.0e-4
    k = k + c.ars_l1k(c.xArea[SLA], c.xRadi[SLA], gamma, srMw)
    k = k + c.ars_l1k(c.xArea[IIC], c.xRadi[IIC], gamma, srMw)
    k = k + CloudHet(c, c.SR_MW[ind["NO2"]], 1.0e-8, 0.0, 1.0, 0.0)
    return k


def Gam_NO3(c, ind, aArea, aRadi, aWater, C_X):
    INV_AB = 1.0 / 1.3e-2
    Vol = aArea * aRadi * 1.0e-3 / 3.0
    WaterC = aWater / 18.0e+12 / Vol
    cavg = _cavg(c, ind, "NO3")
    k_tot = (2.76e+6 * C_X) + (23.0 * WaterC)
    gamma = 0.0
    if k_tot > 0.0:
        H_X = 0.6 * CON_ATM_BAR
        l_r = math.sqrt(1.0e-5 / k_tot)
        gb = c.FOUR_R_T * H_X * l_r * k_tot / cavg
        corr = ReactoDiff_Corr(aRadi, l_r)
        gb = gb * corr
        gamma = 1.0 / (INV_AB + 1.0 / gb)
    return gamma


def NO3uptk1stOrdAndCloud(c, ind):
    k = 0.0
    srMw = c.SR_MW[ind["NO3"]]
    gamma = 0.01
    for a in range(DU1, DU1 + 7):
        k = k + c.ars_l1k(c.xArea[a], c.xRadi[a], gamma, srMw)
    gamma = 2.0e-4 if c.RELHUM < 50.0 else 1.0e-3
    k = k + c.ars_l1k(c.xArea[BKC], c.xRadi[BKC], gamma, srMw)
    k = k + c.ars_l1k(c.xArea[ORC], c.xRadi[ORC], 0.005, srMw)
    gamma = 0.1
    k = k + c.ars_l1k(c.xArea[SLA], c.xRadi[SLA], gamma, srMw)
    k = k + c.ars_l1k(c.xArea[IIC], c.xRadi[IIC], gamma, srMw)
    k = k + CloudHet(c, c.SR_MW[ind["NO3"]], 0.002, 0.001, 1.0, 1.0)
    return k


def NO3hypsisClonSALA(c, ind):
    area, radi, water, conc = c.aClArea, c.aClRadi, c.aWater[1], c.Cl_conc_SSA
    gamma = Gam_NO3(c, ind, area, radi, water, conc) * 0.01
    area = c.ClearFr * area
    return c.ars_l1k(area, radi, gamma, c.SR_MW[ind["NO3"]])


def NO3hypsisClonSALC(c, ind):
    area, radi, water, conc = c.xArea[SSC], c.xRadi[SSC], c.aWater[2], c.Cl_conc_SSC
    gamma = Gam_NO3(c, ind, area, radi, water, conc) * 0.01
    area = c.ClearFr * area
    return c.ars_l1k(area, radi, gamma, c.SR_MW[ind["NO3"]])


def Gamma_O3_Br(c, ind, radius, C_Br):
    K0_O3 = 1.1e-2 * CON_ATM_BAR
    if not C_Br > 0.0:
        return 0.0
    H_X = K0_O3 * math.exp(2300.0 * (c.INV_TEMP - INV_T298))
    cavg = _cavg(c, ind, "O3")
    Nmax = 3.0e+14
    KLangC = 1.0e-13
    k_s = 1.0e-16
    C_Br_surf = min(3.41e+14 * C_Br, Nmax)
    gs = (4.0 * k_s * C_Br_surf * KLangC * Nmax) / (cavg * (1.0 + KLangC * c.C[ind["O3"]]))
    k_b = 6.3e+8 * math.exp(-4.45e+3 / c.TEMP)
    D_l = 8.9e-6
    l_r = math.sqrt(D_l / (k_b * C_Br))
    gb = c.FOUR_R_T * H_X * l_r * k_b * C_Br / cavg
    gb = gb * ReactoDiff_Corr(radius, l_r)
    return gb + gs


def O3uptkByBrInTropCloud(c, ind, Br_branch):
    if c.stratBox:
        return 0.0
    gamma = Gamma_O3_Br(c, ind, c.rLiq, c.Br_conc_Cld)
    return CloudHet(c, c.SR_MW[ind["O3"]], gamma, 0.0, Br_branch, 0.0)


def O3uptkByHBr(c, ind):
    k = O3uptkByBrInTropCloud(c, ind, c.frac_Br_CldG)
    return kIIR1Ltd(c.C[ind["O3"]], c.C[ind["HBr"]], k)


def O3uptkByBrSALA(c, ind):
    k = 0.0
    if c.stratBox:
        return k
    k = k + O3uptkByBrInTropCloud(c, ind, c.frac_Br_CldA)
    if c.SSA_is_Acid:
        area = c.ClearFr * c.aClArea * c.f_Acid_SSA
        gamma = Gamma_O3_Br(c, ind, c.aClRadi, c.Br_conc_SSA)
        k = k + c.ars_l1k(area, c.aClRadi, gamma, c.SR_MW[ind["O3"]])
    return kIIR1Ltd(c.C[ind["O3"]], c.C[ind["BrSALA"]], k)


def O3uptkByBrSALC(c, ind):
    k = 0.0
    if c.stratBox:
        return k
    k = k + O3uptkByBrInTropCloud(c, ind, c.frac_Br_CldC)
    if c.SSC_is_Acid:
        area = c.ClearFr * c.xArea[SSC] * c.f_Acid_SSC
        gamma = Gamma_O3_Br(c, ind, c.xRadi[SSC], c.Br_conc_SSC)
        k = k + c.ars_l1k(area, c.xRadi[SSC], gamma, c.SR_MW[ind["O3"]])
    return kIIR1Ltd(c.C[ind["O3"]], c.C[ind["BrSALC"]], k)


AVO = 6.022140857e+23
FOUR_RGASLATM = 4.0 * RGASLATM


def ClNO2_BT(Cl, H2O):
    k2k3 = 1.0 / 4.5e+2
    if H2O < 0.1:
        phi = 0.0
        if Cl > 1e-3:
            phi = 1.0
        return phi
    return 1.0 / (1.0 + k2k3 * safe_div(H2O, Cl, 1.0e+30))


def N2O5_InorgOrg(c, ind, volInorg, volOrg, H2Oinorg, H2Oorg, Rcore, NIT, Cl):
    """-> gamma, Y_ClNO2, Rp, areaTotal  (fullchem_RateLawFuncs.F90:2699-2858)"""
    KH, k3k2b, beta, delta = 5.1e+1, 4.0e-2, 1.15e+6, 1.3e-1
    Haq, Daq, ONE_THIRD = 5e+3, 1e-9, 1.0 / 3.0
    volTotal = volInorg + volOrg
    H2Ototal = H2Oinorg + H2Oorg
    volRatioDry = safe_div(max(volInorg - H2Oinorg, 0.0), max(volTotal - H2Ototal, 0.0), 0.0)
    Rp = safe_div(Rcore, volRatioDry ** ONE_THIRD, Rcore)
    l = Rp - Rcore
    M_N2O5 = c.MW[ind["N2O5"]] * 1.0e-3
    speed = math.sqrt(c.EIGHT_RSTARG_T / (PI * M_N2O5))
    M_H2O = H2Ototal / 18e+0 / volTotal * 1000.0
    M_NIT = NIT / volTotal / AVO * 1000.0
    M_Cl = Cl / volTotal / AVO * 1000.0
    OCratio = (((c.OMOC_POA + c.OMOC_OPOA) / 2.0) - 1.17) / 1.29
    eps = 1.5e-1 * OCratio + 1.6e-3 * c.RELHUM
    if l <= 0.0e+0:
        gamma_coat = 0.0
    else:
        gamma_coat = (c.FOUR_RGASLATM_T * 1.0e-3 * eps * Haq * Daq * Rcore / 100.0) / (speed * l / 100.0 * Rp / 100.0)
    areaTotal = 3.0 * volTotal / Rp
    if M_H2O < 0.1:
        gamma_core = 0.005
    else:
        speed = speed * 1e+2
        A = ((4.0 * volTotal) / (speed * areaTotal)) * KH
        A = min(A, 3.2e-8)
        if delta * M_H2O < 1e-2:
            k2f = beta * (delta * M_H2O)
        else:
            k2f = beta * (1e+0 - math.exp(-delta * M_H2O))
        gamma_core = A * k2f * (1.0 - 1.0 / (1.0 + safe_div(k3k2b * M_H2O, M_NIT, 1.0e+30)))
    if gamma_coat <= 0.0:
        gamma = gamma_core
    elif gamma_core <= 0.0:
        gamma = 0.0
    else:
        gamma = 1.0 / ((1.0 / gamma_core) + (1.0 / gamma_coat))
    return gamma, ClNO2_BT(M_Cl, M_H2O), Rp, areaTotal


def N2O5uptkByH2O(c, ind):
    k = 0.0
    srMw = c.SR_MW[ind["N2O5"]]
    gamma = 0.02
    for a in range(DU1, DU1 + 7):
        k = k + c.ars_l1k(c.ClearFr * c.xArea[a], c.xRadi[a], gamma, srMw)
    gamma, Y_ClNO2, Rp, SA = N2O5_InorgOrg(c, ind, c.AClVol, c.xVol_ORC, c.xH2O_SUL, c.xH2O_ORC, c.aClRadi,
                                           c.C[ind["NIT"]], c.C[ind["SALACL"]])
    ktmp = c.ars_l1k(c.ClearFr * SA, Rp, gamma, srMw)
    k = k + ktmp - (ktmp * Y_ClNO2 * 0.25)
    gamma = 0.005
    k = k + c.ars_l1k(c.ClearFr * c.xArea[BKC], c.xRadi[BKC], gamma, srMw)
    gamma, Y_ClNO2, Rp, SA = N2O5_InorgOrg(c, ind, c.xVol_SSC, 0.0, c.xH2O_SSC, 0.0, c.xRadi[SSC],
                                           c.C[ind["NITs"]], c.C[ind["SALCCL"]])
    ktmp = c.ars_l1k(c.ClearFr * SA, Rp, gamma, srMw)
    k = k + ktmp - (ktmp * Y_ClNO2)
    k = k + c.xArea[SLA] * c.KHETI_SLA[N2O5_plus_H2O]
    gamma = 0.02
    if c.natSurface:
        gamma = 4.0e-4
    k = k + c.ars_l1k(c.ClearFr * c.xArea[IIC], c.xRadi[IIC], gamma, srMw)
    return kIIR1Ltd(c.C[ind["N2O5"]], c.C[ind["H2O"]], k)


def N2O5uptkBySALACl(c, ind):
    k = 0.0
    if c.stratBox:
        return k
    gamma, Y_ClNO2, Rp, SA = N2O5_InorgOrg(c, ind, c.AClVol, c.xVol_ORC, c.xH2O_SUL, c.xH2O_ORC, c.aClRadi,
                                           c.C[ind["NIT"]], c.C[ind["SALACL"]])
    k = c.ars_l1k(c.ClearFr * SA, Rp, gamma, c.SR_MW[ind["N2O5"]])
    k = k * Y_ClNO2 * 0.25
    return kIIR1Ltd(c.C[ind["N2O5"]], c.C[ind["SALACL"]], k)


def N2O5uptkBySALCCl(c, ind):
    k = 0.0
    if c.stratBox:
        return k
    gamma, Y_ClNO2, Rp, SA = N2O5_InorgOrg(c, ind, c.xVol_SSC, 0.0, c.xH2O_SSC, 0.0, c.xRadi[SSC],
                                           c.C[ind["NITs"]], c.C[ind["SALCCL"]])
    k = c.ars_l1k(c.ClearFr * SA, Rp, gamma, c.SR_MW[ind["N2O5"]])
    k = k * Y_ClNO2
    return kIIR1Ltd(c.C[ind["N2O5"]], c.C[ind["SALCCL"]], k)


LAWS2 = {f.__name__: f for f in (N2O5uptkByH2O, N2O5uptkBySALACl, N2O5uptkBySALCCl,
    BrNO3uptkByH2O, BrNO3uptkByHCl, ClNO2uptkByBrSALA, ClNO2uptkByBrSALC, ClNO2uptkByHBr, ClNO2uptkBySALACL,
    ClNO2uptkBySALCCL, ClNO2uptkByHCl, ClNO3uptkByH2O, ClNO3uptkByHCl, ClNO3uptkByHBr, ClNO3uptkByBrSALA,
    ClNO3uptkByBrSALC, ClNO3uptkBySALACL, ClNO3uptkBySALCCL, HOBrUptkByHBr, HOBrUptkByHCl, HOBrUptkByBrSALA,
    HOBrUptkByBrSALC, HOBrUptkBySALACL, HOBrUptkBySALCCL, HOBrUptkByHSO3m, HOClUptkByHCl, HOClUptkByHBr,
    HOClUptkBySALACL, HOClUptkBySALCCL, HOClUptkByHSO3m, IONO2uptkByH2O, N2O5uptkByCloud, N2O5uptkByStratHCl,
    NO2uptk1stOrdAndCloud, NO3uptk1stOrdAndCloud, NO3hypsisClonSALA, NO3hypsisClonSALC, O3uptkByHBr, O3uptkByBrSALA,
    O3uptkByBrSALC)}


def evaluate2(rconst_exprs, ind, cell, mw, henry_k0, henry_cr):
    """{reaction index: value} for the laws of the second part (those called as Law(State_Het))"""
    cell.MW, cell.HENRY_K0, cell.HENRY_CR = mw, henry_k0, henry_cr
    _prep2(cell)
    out = {}
    for r, e in enumerate(rconst_exprs):
        if e is None:
            continue
        m = re.fullmatch(r"(\w+)\(\s*State_Het\s*\)", e.strip())
        if m and m.group(1) in LAWS2:
            out[r] = LAWS2[m.group(1)](cell, ind)
    return out
