// Heterogeneous rate laws evaluated on the device (first part): the 61 uptake constants of fullchem
// whose laws need only aerosol area / radius, sea-salt alkalinity flags, the sulfate acidity and a few species
// concentrations:
//   VOCuptk1stOrd, IEPOXuptk1stOrd (+ EpoxUptkGamma), MGLYuptk1stOrd, GLYXuptk1stOrd
//                                   KPP/fullchem/fullchem_RateLawFuncs.F90:3286-3458
//   Iuptk{BySulf,BySALA,ByAlkSALA,BySALC,ByAlkSALC}1stOrd, IbrkdnByAcid{BrSALA,BrSALC,SALACl,SALCCl}   :2371-2542
//   HO2uptk1stOrd :1461-1484, HBrUptkBySALA / SALC :1423-1455, OHuptkBySALACl / SALCCl :3244-3280
//   Ars_L1k, kIIR1Ltd, SafeDiv, Is_SafeDiv            KPP/fullchem/rateLawUtilFuncs.F90:77-140, 459-495
// The cloud / halogen / N2O5 / NO2 / NO3 laws and K_MT / K_CLD (sulfur chemistry) are NOT here: their constants keep
// arriving through khet.  Inputs: one HetCell per cell = the GCKPP_HET_* fields of include/gckpp_gpu.h (a subset of
// the reference's HetState, commonIncludeVars.H:112-210, logicals as 0/1), SR_MW and the concentrations.
#pragma once
#include <math.h>
#include "tables.h"

#define GCKPP_NHET_FIELDS 48

struct HetCell {
  double SUNCOS, stratBox, SSA_is_Alk, SSA_is_Acid, SSC_is_Alk, SSC_is_Acid;
  double f_Alk_SSA, f_Alk_SSC, f_Acid_SSA, f_Acid_SSC, ClearFr, aClArea, aClRadi, Cl_conc_SSA, Cl_conc_SSC;
  double gamma_HO2, H_PLUS, NO3_molal, SO4_molal, HSO4_molal;
  double xArea[14], xRadi[14];      // DU1..DU7, SUL, BKC, ORC, SSA, SSC, SLA, IIC (1-based in the reference)
};

enum { HA_DU1 = 0, HA_SUL = 7, HA_BKC = 8, HA_ORC = 9, HA_SSA = 10, HA_SSC = 11, HA_SLA = 12, HA_IIC = 13 };

__device__ __forceinline__ HetCell het_load(const double *__restrict__ het, size_t stride)
{
  HetCell H;
  double *p = reinterpret_cast<double *>(&H);
#pragma unroll
  for (int k = 0; k < GCKPP_NHET_FIELDS; k++) p[k] = het[(size_t)k * stride];
  return H;
}

// Fortran EXPONENT(x): e with x = f * 2**e, 0.5 <= |f| < 1; EXPONENT(0) = 0
__device__ __forceinline__ int het_exponent(double x) { int e = 0; if (x != 0.0) frexp(x, &e); return e; }

__device__ __forceinline__ double het_SafeDiv(double num, double denom, double alt)
{
  const int ediff = het_exponent(num) - het_exponent(denom);
  if (ediff > 1023 || denom == 0.0) return alt;
  if (ediff < -1020) return 0.0;
  return num / denom;
}
__device__ __forceinline__ bool het_Is_SafeDiv(double num, double denom)
{
  const int ediff = het_exponent(num) - het_exponent(denom);
  return !(ediff < -1020 || ediff > 1023 || denom == 0.0);
}

// first-order loss on an aerosol surface: gas diffusion + surface uptake in series
__device__ __forceinline__ double het_Ars_L1k(const MetCell &m, double area, double radius, double gamma, double srMw)
{
  if (gamma < 1.0e-30 || radius < 1.0e-30) return 0.0;
  const double dfkg = (9.45e+17 / m.NUMDEN) * m.SR_TEMP * sqrt(3.472e-2 + 1.0 / (srMw * srMw));
  return area / ((radius / dfkg) + 2.749064e-4 * srMw / (gamma * m.SR_TEMP));
}

// second-order constant from a first-order one, limited so that neither reactant is consumed faster than HET_MIN_LIFE
__device__ __forceinline__ double het_kIIR1Ltd(double concGas, double concEduct, double kISource)
{
  const double HET_MIN_LIFE = 1.0e-3, HET_MIN_RATE = 1.0 / HET_MIN_LIFE;
  if (concEduct < 1.0) return 0.0;
  if (!het_Is_SafeDiv(concGas * kISource, concEduct)) return 0.0;
  const double kIGas = kISource;
  const double kIEduct = kIGas * concGas / concEduct;
  double kII = kIGas / concEduct;
  if (kIGas > 0.0) {
    const double lifeA = het_SafeDiv(1.0, kIGas, 0.0), lifeB = het_SafeDiv(1.0, kIEduct, 0.0);
    if (lifeA < lifeB && lifeA < HET_MIN_LIFE) kII = het_SafeDiv(HET_MIN_RATE, concEduct, 0.0);
    else if (lifeB < HET_MIN_LIFE) kII = het_SafeDiv(HET_MIN_RATE, concGas, 0.0);
  }
  return kII;
}

#define HET_CRITRH 35.0

__device__ __forceinline__ double het_VOCuptk1stOrd(const MetCell &m, const HetCell &H, double srMw, double gamma)
{
  double k = 0.0;
  if (m.RELHUM >= HET_CRITRH) {
    k = k + het_Ars_L1k(m, H.xArea[HA_SUL], H.xRadi[HA_SUL], gamma, srMw);
    k = k + het_Ars_L1k(m, H.xArea[HA_BKC], H.xRadi[HA_BKC], gamma, srMw);
    k = k + het_Ars_L1k(m, H.xArea[HA_ORC], H.xRadi[HA_ORC], gamma, srMw);
    k = k + het_Ars_L1k(m, H.xArea[HA_SSA], H.xRadi[HA_SSA], gamma, srMw);
    k = k + het_Ars_L1k(m, H.xArea[HA_SSC], H.xRadi[HA_SSC], gamma, srMw);
    k = k + het_Ars_L1k(m, H.xArea[HA_SLA], H.xRadi[HA_SLA], gamma, srMw);
    k = k + het_Ars_L1k(m, H.xArea[HA_IIC], H.xRadi[HA_IIC], gamma, srMw);
  }
  return k;
}

__device__ __forceinline__ double het_EpoxUptkGamma(const MetCell &m, const HetCell &H, double srMw)
{
  const double DIFF_N2O5_STD = 1.0e-1, MACOEFF = 1.0e-1, K_HPLUS = 3.6e-2, K_NUC = 2.0e-4, K_HSO4 = 7.3e-4, K_HYDRO = 0.0,
               HSTAR_EPOX = 1.7e+7;
  double valTmp = 0.0;
  const double aerVol = (H.xArea[HA_SUL] * H.xRadi[HA_SUL]) / 3.0;
  const double xmms = sqrt((2.117e+8 * m.TEMP) / (srMw * srMw));
  const double kPart = (K_HPLUS * H.H_PLUS) + (K_NUC * H.H_PLUS * (H.NO3_molal + H.SO4_molal)) + (K_HSO4 * H.HSO4_molal) + (K_HYDRO);
  const double val1 = (H.xRadi[HA_SUL] * xmms) / (4.0 * DIFF_N2O5_STD);
  const double val2 = (1.0 / MACOEFF);
  if (H.xArea[HA_SUL] > 0.0 && xmms > 0.0) valTmp = (m.FOUR_RGASLATM_T * aerVol * HSTAR_EPOX * kPart) / (H.xArea[HA_SUL] * xmms);
  double val3 = 0.0;
  if (valTmp > 0.0) val3 = 1.0 / valTmp;
  double gamma = 0.0;
  if (kPart >= 1.e-8) gamma = 1.0 / (val1 + val2 + val3);
  if (gamma < 0.0) gamma = 0.0;
  return gamma;
}

__device__ __forceinline__ double het_IEPOXuptk1stOrd(const MetCell &m, const HetCell &H, double srMw, int doScale)
{
  double k = 0.0;
  if (m.RELHUM >= HET_CRITRH) {
    double gamma = het_EpoxUptkGamma(m, H, srMw);
    if (doScale && H.H_PLUS > 8.0e-5) gamma = gamma / 30.0;
    k = k + het_Ars_L1k(m, H.xArea[HA_SUL], H.xRadi[HA_SUL], gamma, srMw);
  }
  return k;
}

__device__ __forceinline__ double het_MGLYuptk1stOrd(const MetCell &m, const HetCell &H, double srMw)
{
  double k = 0.0;
  if (m.RELHUM >= HET_CRITRH) k = k + het_Ars_L1k(m, H.xArea[HA_SUL], H.xRadi[HA_SUL], 3.6e-7, srMw);
  return k;
}

__device__ __forceinline__ double het_GLYXuptk1stOrd(const MetCell &m, const HetCell &H, double srMw)
{
  double k = 0.0;
  if (m.RELHUM >= HET_CRITRH) {
    const double gamma = (H.SUNCOS > 0.0) ? 4.4e-3 : 8.0e-6;
    k = k + het_Ars_L1k(m, H.xArea[HA_SUL], H.xRadi[HA_SUL], gamma, srMw);
  }
  return k;
}

__device__ __forceinline__ double het_IuptkBySulf1stOrd(const MetCell &m, const HetCell &H, double srMw, double gamma)
{
  double k = het_Ars_L1k(m, H.xArea[HA_SUL], H.xRadi[HA_SUL], gamma, srMw);
  k = k + het_Ars_L1k(m, H.xArea[HA_SLA], H.xRadi[HA_SLA], gamma, srMw);
  return k;
}
__device__ __forceinline__ double het_IuptkBySALA1stOrd(const MetCell &m, const HetCell &H, double srMw, double gamma)
{
  if (H.stratBox != 0.0) return 0.0;
  return het_Ars_L1k(m, H.xArea[HA_SSA], H.xRadi[HA_SSA], gamma, srMw);
}
__device__ __forceinline__ double het_IuptkByAlkSALA1stOrd(const MetCell &m, const HetCell &H, double srMw, double gamma)
{
  if (H.stratBox != 0.0) return 0.0;
  if (H.SSA_is_Alk != 0.0) return het_Ars_L1k(m, H.f_Alk_SSA * H.xArea[HA_SSA], H.xRadi[HA_SSA], gamma, srMw);
  return 0.0;
}
__device__ __forceinline__ double het_IuptkBySALC1stOrd(const MetCell &m, const HetCell &H, double srMw, double gamma)
{
  if (H.stratBox != 0.0) return 0.0;
  return het_Ars_L1k(m, H.xArea[HA_SSC], H.xRadi[HA_SSC], gamma, srMw);
}
__device__ __forceinline__ double het_IuptkByAlkSALC1stOrd(const MetCell &m, const HetCell &H, double srMw, double gamma)
{
  if (H.stratBox != 0.0) return 0.0;
  if (H.SSC_is_Alk != 0.0) return het_Ars_L1k(m, H.f_Alk_SSC * H.xArea[HA_SSC], H.xRadi[HA_SSC], gamma, srMw);
  return 0.0;
}
// breakdown of an iodine species on acidic sea salt, limited by the halide it releases (cBr / cCl = its concentration)
__device__ __forceinline__ double het_IbrkdnByAcidBrSALA(const MetCell &m, const HetCell &H, double srMw, double conc, double gamma, double cBrSALA)
{
  if (H.stratBox != 0.0 || H.SSA_is_Acid == 0.0) return 0.0;
  const double k = 0.15 * het_Ars_L1k(m, H.f_Acid_SSA * H.xArea[HA_SSA], H.xRadi[HA_SSA], gamma, srMw);
  return het_kIIR1Ltd(conc, cBrSALA, k);
}
__device__ __forceinline__ double het_IbrkdnByAcidBrSALC(const MetCell &m, const HetCell &H, double srMw, double conc, double gamma, double cBrSALC)
{
  if (H.stratBox != 0.0 || H.SSC_is_Acid == 0.0) return 0.0;
  const double k = 0.15 * het_Ars_L1k(m, H.f_Acid_SSC * H.xArea[HA_SSC], H.xRadi[HA_SSC], gamma, srMw);
  return het_kIIR1Ltd(conc, cBrSALC, k);
}
__device__ __forceinline__ double het_IbrkdnByAcidSALACl(const MetCell &m, const HetCell &H, double srMw, double conc, double gamma, double cSALACl)
{
  if (H.stratBox != 0.0 || H.SSA_is_Acid == 0.0) return 0.0;
  const double k = 0.85 * het_Ars_L1k(m, H.f_Acid_SSA * H.xArea[HA_SSA], H.xRadi[HA_SSA], gamma, srMw);
  return het_kIIR1Ltd(conc, cSALACl, k);
}
__device__ __forceinline__ double het_IbrkdnByAcidSALCCl(const MetCell &m, const HetCell &H, double srMw, double conc, double gamma, double cSALCCl)
{
  if (H.stratBox != 0.0 || H.SSC_is_Acid == 0.0) return 0.0;
  const double k = 0.85 * het_Ars_L1k(m, H.f_Acid_SSC * H.xArea[HA_SSC], H.xRadi[HA_SSC], gamma, srMw);
  return het_kIIR1Ltd(conc, cSALCCl, k);
}

__device__ __forceinline__ double het_HO2uptk1stOrd(const MetCell &m, const HetCell &H, double srMwHO2)
{
  double k = 0.0;
#pragma unroll
  for (int a = HA_DU1; a <= HA_SSC; a++) k = k + het_Ars_L1k(m, H.xArea[a], H.xRadi[a], H.gamma_HO2, srMwHO2);   // DU1-7, SUL, BKC, ORC, SSA, SSC
  return k;
}

__device__ __forceinline__ double het_HBrUptkBySALA(const MetCell &m, const HetCell &H, double srMwHBr)
{
  if (H.stratBox != 0.0) return 0.0;
  const double gamma = 1.3e-8 * exp(4290.0 / m.TEMP);
  return het_Ars_L1k(m, H.ClearFr * H.aClArea, H.aClRadi, gamma, srMwHBr);
}
__device__ __forceinline__ double het_HBrUptkBySALC(const MetCell &m, const HetCell &H, double srMwHBr)
{
  if (H.stratBox != 0.0) return 0.0;
  const double gamma = 1.3e-8 * exp(4290.0 / m.TEMP);
  return het_Ars_L1k(m, H.ClearFr * H.xArea[HA_SSC], H.xRadi[HA_SSC], gamma, srMwHBr);
}

__device__ __forceinline__ double het_OHuptkBySALACl(const MetCell &m, const HetCell &H, double srMwOH, double cOH, double cSALACl)
{
  if (H.stratBox != 0.0) return 0.0;
  const double k = het_Ars_L1k(m, H.aClArea, H.aClRadi, 0.04 * H.Cl_conc_SSA, srMwOH);
  return het_kIIR1Ltd(cOH, cSALACl, k);
}
__device__ __forceinline__ double het_OHuptkBySALCCl(const MetCell &m, const HetCell &H, double srMwOH, double cOH, double cSALCCl)
{
  if (H.stratBox != 0.0) return 0.0;
  const double k = het_Ars_L1k(m, H.xArea[HA_SSC], H.xRadi[HA_SSC], 0.04 * H.Cl_conc_SSC, srMwOH);
  return het_kIIR1Ltd(cOH, cSALCCl, k);
}
