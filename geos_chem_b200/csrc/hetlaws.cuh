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
#include "gen/fullchem_hetind.h"

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

// ---------------------------------------------------------------------------------------------------------------------
// Second part: cloud and halogen uptake (fullchem_RateLawFuncs.F90:803-3238).  These laws also read species data of the
// host's species database (MW, Henry's law K0 / CR) and more HetState fields; a context object carries everything.
//   CloudHet rateLawUtilFuncs.F90:142-250, ReactoDiff_Corr :434-453, Br2_Yield :1490-1503
//   BrNO3uptkByH2O / HCl :803-862; Gam_ClNO2 + ClNO2uptkBy* :868-1089; Gam_ClNO3_Aer / _Ice + ClNO3uptkBy* :1095-1417
//   Gam_HOBr_Aer / _Cld / _Ice + HOBrUptkBy* :1505-2032; Gam_HOCl_Cld / _Aer + HOClUptkBy* :2072-2335
//   IONO2uptkByH2O :2544-2577; N2O5uptkByCloud / ByStratHCl :2881-2922; NO2 / NO3uptk1stOrdAndCloud :2928-3058
//   Gam_NO3 + NO3hypsisClonSALA / SALC :2982-3100; Gamma_O3_Br + O3uptkBy* :3106-3238
//   N2O5_InorgOrg + N2O5uptkByH2O / BySALACl / BySALCCl :2583-2879 (at the end of this file)
// Not here: K_MT / K_CLD and the SRHOBr / SRHOCl sums (the sulfur module).
struct HetCell2 {
  double natSurface, TurnOffHetRates, CldFr, aIce, aLiq, rIce, rLiq, pHCloud, pHSSA[2];
  double Cl_conc_Cld, Br_conc_Cld, Br_conc_SSA, Br_conc_SSC, Br_over_Cl_Cld, Br_over_Cl_SSA, Br_over_Cl_SSC;
  double frac_Br_CldA, frac_Br_CldC, frac_Br_CldG, frac_Cl_CldA, frac_Cl_CldC, frac_Cl_CldG, frac_SALACL, frac_HSO3_aq, HSO3m;
  double HCl_theta, HBr_theta, HNO3_theta, H_conc_LCl, H_conc_SSA, H_conc_SSC, HSO3_aq, SO3_aq, TSO3_aq, aWater[2];
  double KHETI_SLA[11];
};
#define GCKPP_NHET_FIELDS2 48
static_assert(sizeof(HetCell2) == GCKPP_NHET_FIELDS2 * sizeof(double), "HetCell2 must be 48 doubles");

__device__ __forceinline__ HetCell2 het_load2(const double *__restrict__ het, size_t stride)
{
  HetCell2 H;
  double *p = reinterpret_cast<double *>(&H);
#pragma unroll
  for (int k = 0; k < GCKPP_NHET_FIELDS2; k++) p[k] = het[(size_t)(GCKPP_NHET_FIELDS + k) * stride];
  return H;
}

// KHETI_SLA slots (1-based in the reference)
enum { SLA_N2O5_H2O = 0, SLA_N2O5_HCl, SLA_ClNO3_H2O, SLA_ClNO3_HCl, SLA_ClNO3_HBr, SLA_BrNO3_H2O, SLA_BrNO3_HCl,
       SLA_HOCl_HCl, SLA_HOCl_HBr, SLA_HOBr_HCl, SLA_HOBr_HBr };

// third block: what N2O5_InorgOrg reads (wet volumes and water contents of the sulfate / organic / sea-salt aerosol,
// OM/OC ratios)
struct HetCell3 { double AClVol, xVol_ORC, xVol_SSC, xH2O_SUL, xH2O_ORC, xH2O_SSC, OMOC_POA, OMOC_OPOA; };
#define GCKPP_NHET_FIELDS3 8
static_assert(sizeof(HetCell3) == GCKPP_NHET_FIELDS3 * sizeof(double), "HetCell3 must be 8 doubles");
__device__ __forceinline__ HetCell3 het_load3(const double *__restrict__ het, size_t stride)
{
  HetCell3 K;
  double *p = reinterpret_cast<double *>(&K);
#pragma unroll
  for (int k = 0; k < GCKPP_NHET_FIELDS3; k++) p[k] = het[(size_t)(GCKPP_NHET_FIELDS + GCKPP_NHET_FIELDS2 + k) * stride];
  return K;
}

struct HetCtx {
  const MetCell &m; const HetCell &H; const HetCell2 &G; const HetCell3 &K;
  const double *srmw, *mw, *hk0, *hcr, *conc; size_t stride;
  __device__ __forceinline__ double C(int i) const { return conc[(size_t)i * stride]; }
  __device__ __forceinline__ double Ars(double area, double radius, double gamma, double srMw) const { return het_Ars_L1k(m, area, radius, gamma, srMw); }
};

#define HET_PI 3.14159265358979323
#define HET_CON_ATM_BAR (1.0 / 1.01325)
#define HET_INV_T298 (1.0 / 298.15)

__device__ __forceinline__ double het_ReactoDiff_Corr(double radius, double l)
{
  const double x = radius / l;
  if (x > 1000.0) return 1.0;
  if (x < 0.1) return x / 3.0;
  const double y = exp(-2.0 * x);            // coth as the reference writes it (rateLawUtilFuncs.F90:423-432)
  return (1.0 + y) / (1.0 - y) - (1.0 / x);
}

__device__ __forceinline__ double het_Br2_Yield(double r)
{
  double y = 0.0;
  if (r > 0.0) { y = 0.41 * log10(r) + 2.25; y = fmax(fmin(y, 0.9), 0.0); }
  return y;
}

// grid-average loss frequency in a partly cloudy box (entrainment-limited uptake)
__device__ __forceinline__ double het_CloudHet(const HetCtx &x, double srMw, double gamLiq, double gamIce, double brLiq, double brIce)
{
  const HetCell2 &G = x.G;
  const double tauc = 3600.0;
  if (G.CldFr < 0.0001 || G.aLiq + G.aIce <= 0.0) return 0.0;
  double kI = 0.0, kIb = 0.0, ktmp;
  if (brLiq > 0.0) {
    const double area = het_SafeDiv(G.aLiq, G.CldFr, 0.0);
    if (area > 0.0) { ktmp = x.Ars(area, G.rLiq, gamLiq, srMw); kI = kI + ktmp; kIb = kIb + (ktmp * brLiq); }
  }
  if (brIce > 0.0) {
    const double area = het_SafeDiv(G.aIce, G.CldFr, 0.0);
    if (area > 0.0) { ktmp = x.Ars(area, G.rIce, gamIce, srMw); kI = kI + ktmp; kIb = kIb + (ktmp * brIce); }
  }
  const double branch = het_SafeDiv(kIb, kI, 0.0);
  if (!(branch > 0.0)) return 0.0;
  const double kk = kI * tauc;
  double ff = het_SafeDiv(G.CldFr, x.H.ClearFr, 1.0e+30);
  ff = fmin(ff, 1.0e+30);
  double xx = (ff - kk - 1.0) / 2.0 + sqrt(1.0 + ff * ff + kk * kk + 2.0 * ff + 2.0 * kk - 2.0 * ff * kk) / 2.0;
  xx = fmax(xx, 0.0);
  double kHet = kI / (1.0 + het_SafeDiv(1.0, xx, 1.0e+30));
  return kHet * branch;
}

// mean molecular speed [cm/s] of a species of molecular weight mw [g/mol]
__device__ __forceinline__ double het_cavg(const HetCtx &x, double mw) { return sqrt(x.m.EIGHT_RSTARG_T / (HET_PI * (mw * 1.0e-3))) * 100.0; }

// ---- BrNO3
__device__ __forceinline__ double het2_BrNO3uptkByH2O(const HetCtx &x)
{
  const HetCell &H = x.H; const HetCell2 &G = x.G;
  double k = 0.0;
  const double gamLiq = 0.0021 * x.m.TEMP - 0.561, gamIce = 5.3e-4 * exp(1100.0 / x.m.TEMP), srMw = x.srmw[HETIND_BrNO3];
  double gamma = gamLiq;
  k = k + x.Ars(H.ClearFr * H.xArea[HA_SUL], H.xRadi[HA_SUL], gamma, srMw);
  k = k + x.Ars(H.ClearFr * H.xArea[HA_SSA], H.xRadi[HA_SSA], gamma, srMw);
  k = k + x.Ars(H.ClearFr * H.xArea[HA_SSC], H.xRadi[HA_SSC], gamma, srMw);
  k = k + H.xArea[HA_SLA] * G.KHETI_SLA[SLA_BrNO3_H2O];
  gamma = 0.3;
  if (G.natSurface != 0.0) gamma = 0.001;
  k = k + x.Ars(H.ClearFr * H.xArea[HA_IIC], H.xRadi[HA_IIC], gamma, srMw);
  k = k + het_CloudHet(x, srMw, gamLiq, gamIce, 1.0, 1.0);
  return het_kIIR1Ltd(x.C(HETIND_BrNO3), x.C(HETIND_H2O), k);
}
__device__ __forceinline__ double het2_BrNO3uptkByHCl(const HetCtx &x)
{
  const HetCell &H = x.H; const HetCell2 &G = x.G;
  double k = 0.0;
  const double srMw = x.srmw[HETIND_BrNO3];
  if (H.stratBox != 0.0) {
    k = k + x.Ars(H.xArea[HA_SUL], H.xRadi[HA_SUL], 0.9, srMw);
    k = k + H.xArea[HA_SLA] * G.KHETI_SLA[SLA_BrNO3_HCl];
    k = k + x.Ars(H.xArea[HA_IIC], H.xRadi[HA_IIC], 0.3, srMw);
  }
  k = het_kIIR1Ltd(x.C(HETIND_BrNO3), x.C(HETIND_HCl), k);
  if (G.TurnOffHetRates != 0.0) k = 0.0;
  return k;
}

// ---- ClNO2
__device__ __forceinline__ void het_Gam_ClNO2(const HetCtx &x, double radius, double pH, double C_Cl, double C_Br, double &gamma, double &branchCl, double &branchBr)
{
  const double INV_AB = 1.0 / 0.01, D_l = 1.0e-5;
  const double cavg = het_cavg(x, x.mw[HETIND_ClNO2]);
  const double H_X = 4.5e-2 * HET_CON_ATM_BAR;
  double k_Cl = 1.0e+7 * C_Cl;
  if (pH >= 2.0) k_Cl = 0.0;
  const double k_Br = (1.01e-1 / (H_X * H_X * D_l)) * C_Br;
  const double k_tot = k_Cl + k_Br;
  gamma = 0.0; branchCl = 0.0; branchBr = 0.0;
  if (k_tot > 0.0) {
    const double l_r = sqrt(D_l / k_tot);
    double gb_tot = x.m.FOUR_R_T * H_X * l_r * k_tot / cavg;
    gb_tot = gb_tot * het_ReactoDiff_Corr(radius, l_r);
    gamma = 1.0 / (INV_AB + 1.0 / gb_tot);
    branchCl = k_Cl / k_tot;
    branchBr = k_Br / k_tot;
  }
}
// which: 0 BrSALA, 1 BrSALC, 2 HBr, 3 SALACL, 4 SALCCL, 5 HCl
__device__ __forceinline__ double het_ClNO2uptk(const HetCtx &x, int which)
{
  const HetCell &H = x.H; const HetCell2 &G = x.G;
  double k = 0.0, gamma, bCl, bBr;
  const double srMw = x.srmw[HETIND_ClNO2];
  const bool br = which <= 2;
  if (H.stratBox == 0.0) {
    het_Gam_ClNO2(x, G.rLiq, G.pHCloud, G.Cl_conc_Cld, G.Br_conc_Cld, gamma, bCl, bBr);
    const double frac = which == 0 ? G.frac_Br_CldA : which == 1 ? G.frac_Br_CldC : which == 2 ? G.frac_Br_CldG
                      : which == 3 ? G.frac_Cl_CldA : which == 4 ? G.frac_Cl_CldC : G.frac_Cl_CldG;
    const double branch = (br ? bBr : bCl) * frac;
    k = k + het_CloudHet(x, srMw, gamma, 0.0, branch, 0.0);
  }
  if (which == 0 || which == 3) {
    het_Gam_ClNO2(x, H.aClRadi, G.pHSSA[0], H.Cl_conc_SSA, G.Br_conc_SSA, gamma, bCl, bBr);
    k = k + x.Ars(H.ClearFr * H.aClArea, H.aClRadi, gamma, srMw) * (br ? bBr : bCl);
  } else if (which == 1) {
    het_Gam_ClNO2(x, H.xRadi[HA_SSC], G.pHSSA[1], H.Cl_conc_SSC, G.Br_conc_SSC, gamma, bCl, bBr);
    k = k + x.Ars(H.ClearFr * H.xArea[HA_SSC], H.xRadi[HA_SSC], gamma, srMw) * bBr;
  }
  const int educt = which == 0 ? HETIND_BrSALA : which == 1 ? HETIND_BrSALC : which == 2 ? HETIND_HBr
                  : which == 3 ? HETIND_SALACL : which == 4 ? HETIND_SALCCL : HETIND_HCl;
  return het_kIIR1Ltd(x.C(HETIND_ClNO2), x.C(educt), k);
}
__device__ __forceinline__ double het2_ClNO2uptkByBrSALA(const HetCtx &x) { return het_ClNO2uptk(x, 0); }
__device__ __forceinline__ double het2_ClNO2uptkByBrSALC(const HetCtx &x) { return het_ClNO2uptk(x, 1); }
__device__ __forceinline__ double het2_ClNO2uptkByHBr(const HetCtx &x) { return het_ClNO2uptk(x, 2); }
__device__ __forceinline__ double het2_ClNO2uptkBySALACL(const HetCtx &x) { return het_ClNO2uptk(x, 3); }
__device__ __forceinline__ double het2_ClNO2uptkBySALCCL(const HetCtx &x) { return het_ClNO2uptk(x, 4); }
__device__ __forceinline__ double het2_ClNO2uptkByHCl(const HetCtx &x) { return het_ClNO2uptk(x, 5); }

// ---- ClNO3
__device__ __forceinline__ void het_Gam_ClNO3_Aer(const HetCtx &x, double C_Br, double &gamma, double &branchBr)
{
  const double INV_AB = 1.0 / 0.108, K_0 = 1.2e+5 * 1.2e+5, D_l = 5.0e-6;
  const double cavg = het_cavg(x, x.mw[HETIND_ClNO3]);
  const double k_Br = 1.0e+12 * C_Br;
  const double k_tot = K_0 + k_Br;
  const double gb_tot = x.m.FOUR_R_T * sqrt(k_tot * D_l) / cavg;
  gamma = 1.0 / (INV_AB + 1.0 / gb_tot);
  branchBr = k_Br / k_tot;
}
__device__ __forceinline__ void het_Gam_ClNO3_Ice(const HetCtx &x, double &gamma, double &brHCl, double &brHBr, double &brH2O)
{
  const double twenty = 1.0 / 0.5;
  const double g1 = 0.24 * x.G.HCl_theta, g2 = 0.56 * x.G.HBr_theta;
  const double cavg = het_cavg(x, x.mw[HETIND_ClNO3]);
  const double H2Os = 1e+15 - (3.0 * 2.7e+14 * x.G.HNO3_theta);
  const double kks = 4.0 * 5.2e-17 * exp(2032.0 / x.m.TEMP);
  const double g3 = 1.0 / (twenty + cavg / (kks * H2Os));
  gamma = g1 + g2 + g3;
  brHCl = g1 / gamma; brHBr = g2 / gamma; brH2O = g3 / gamma;
}
__device__ __forceinline__ double het2_ClNO3uptkByH2O(const HetCtx &x)
{
  const HetCell &H = x.H; const HetCell2 &G = x.G;
  double k = 0.0, gamma, branchBr, gammaIce, d1, d2, branchIce;
  const double srMw = x.srmw[HETIND_ClNO3];
  het_Gam_ClNO3_Aer(x, G.Br_conc_SSA, gamma, branchBr);
  double branchLiq = (1.0 - branchBr) * (1.0 - G.frac_SALACL);
  k = k + x.Ars(H.ClearFr * H.aClArea, H.aClRadi, gamma, srMw) * branchLiq;
  k = k + H.xArea[HA_SLA] * G.KHETI_SLA[SLA_ClNO3_H2O];
  gamma = 0.3;
  if (G.natSurface != 0.0) gamma = 0.004;
  k = k + x.Ars(H.xArea[HA_IIC], H.xRadi[HA_IIC], gamma, srMw);
  if (H.stratBox == 0.0) {
    het_Gam_ClNO3_Aer(x, G.Br_conc_Cld, gamma, branchBr);
    branchLiq = 1.0 - branchBr;
    het_Gam_ClNO3_Ice(x, gammaIce, d1, d2, branchIce);
    k = k + het_CloudHet(x, srMw, gamma, gammaIce, branchLiq, branchIce);
  }
  return het_kIIR1Ltd(x.C(HETIND_ClNO3), x.C(HETIND_H2O), k);
}
__device__ __forceinline__ double het2_ClNO3uptkByHCl(const HetCtx &x)
{
  const HetCell &H = x.H; const HetCell2 &G = x.G;
  double k = 0.0, gamma, gammaIce, branchIce, d1, d2;
  const double srMw = x.srmw[HETIND_ClNO3];
  if (H.stratBox != 0.0) {
    gamma = 0.1e-4;
    k = k + x.Ars(H.xArea[HA_SUL], H.xRadi[HA_SUL], gamma, srMw);
    k = k + H.xArea[HA_SLA] * G.KHETI_SLA[SLA_ClNO3_HCl];
    gamma = 0.3;
    if (G.natSurface != 0.0) gamma = 0.2;
    k = k + x.Ars(H.xArea[HA_IIC], H.xRadi[HA_IIC], gamma, srMw);
  } else {
    het_Gam_ClNO3_Ice(x, gammaIce, branchIce, d1, d2);
    k = k + het_CloudHet(x, srMw, 0.0, gammaIce, 0.0, branchIce);
  }
  return het_kIIR1Ltd(x.C(HETIND_ClNO3), x.C(HETIND_HCl), k);
}
__device__ __forceinline__ double het2_ClNO3uptkByHBr(const HetCtx &x)
{
  const HetCell &H = x.H; const HetCell2 &G = x.G;
  double k = 0.0, gamma, branchBr, gammaIce, branchIce, d1, d2;
  const double srMw = x.srmw[HETIND_ClNO3];
  if (H.stratBox != 0.0) {
    k = k + H.xArea[HA_SLA] * G.KHETI_SLA[SLA_ClNO3_HBr];
    k = k + x.Ars(H.xArea[HA_IIC], H.xRadi[HA_IIC], 0.3, srMw);
  } else {
    het_Gam_ClNO3_Aer(x, G.Br_conc_Cld, gamma, branchBr);
    const double branchLiq = branchBr * G.frac_Br_CldG;
    het_Gam_ClNO3_Ice(x, gammaIce, d1, branchIce, d2);
    k = het_CloudHet(x, srMw, gamma, gammaIce, branchLiq, branchIce);
  }
  k = het_kIIR1Ltd(x.C(HETIND_ClNO3), x.C(HETIND_HBr), k);
  if (G.TurnOffHetRates != 0.0) k = 0.0;
  return k;
}
__device__ __forceinline__ double het_ClNO3uptkByBrSAL(const HetCtx &x, bool coarse)
{
  const HetCell &H = x.H; const HetCell2 &G = x.G;
  double k = 0.0, gamma, branchBr;
  const double srMw = x.srmw[HETIND_ClNO3];
  if (H.stratBox == 0.0) {
    het_Gam_ClNO3_Aer(x, G.Br_conc_Cld, gamma, branchBr);
    const double branch = branchBr * (coarse ? G.frac_Br_CldC : G.frac_Br_CldA);
    k = k + het_CloudHet(x, srMw, gamma, 0.0, branch, 0.0);
  }
  het_Gam_ClNO3_Aer(x, coarse ? G.Br_conc_SSC : G.Br_conc_SSA, gamma, branchBr);
  if (coarse) k = k + x.Ars(H.ClearFr * H.xArea[HA_SSC], H.xRadi[HA_SSC], gamma, srMw) * branchBr;
  else k = k + x.Ars(H.ClearFr * H.aClArea, H.aClRadi, gamma, srMw) * branchBr;
  k = het_kIIR1Ltd(x.C(HETIND_ClNO3), x.C(coarse ? HETIND_BrSALC : HETIND_BrSALA), k);
  if (G.TurnOffHetRates != 0.0) k = 0.0;
  return k;
}
__device__ __forceinline__ double het2_ClNO3uptkByBrSALA(const HetCtx &x) { return het_ClNO3uptkByBrSAL(x, false); }
__device__ __forceinline__ double het2_ClNO3uptkByBrSALC(const HetCtx &x) { return het_ClNO3uptkByBrSAL(x, true); }
__device__ __forceinline__ double het2_ClNO3uptkBySALACL(const HetCtx &x)
{
  const HetCell &H = x.H; const HetCell2 &G = x.G;
  if (H.stratBox != 0.0) return 0.0;
  double gamma, branchBr;
  het_Gam_ClNO3_Aer(x, G.Br_conc_SSA, gamma, branchBr);
  const double branch = (1.0 - branchBr) * G.frac_SALACL;
  const double k = 0.0 + x.Ars(H.ClearFr * H.aClArea, H.aClRadi, gamma, x.srmw[HETIND_ClNO3]) * branch;
  return het_kIIR1Ltd(x.C(HETIND_ClNO3), x.C(HETIND_SALACL), k);
}
__device__ __forceinline__ double het2_ClNO3uptkBySALCCL(const HetCtx &x)
{
  const HetCell &H = x.H; const HetCell2 &G = x.G;
  if (H.stratBox != 0.0) return 0.0;
  double gamma, branchBr;
  het_Gam_ClNO3_Aer(x, G.Br_conc_SSC, gamma, branchBr);
  const double branch = 1.0 - branchBr;
  const double k = 0.0 + x.Ars(H.ClearFr * H.xArea[HA_SSC], H.xRadi[HA_SSC], gamma, x.srmw[HETIND_ClNO3]) * branch;
  return het_kIIR1Ltd(x.C(HETIND_ClNO3), x.C(HETIND_SALCCL), k);
}

// ---- HOBr
__device__ __forceinline__ double het_henry(const HetCtx &x, int ind) { return (x.hk0[ind] * HET_CON_ATM_BAR) * exp(x.hcr[ind] * (1.0 / x.m.TEMP - HET_INV_T298)); }

__device__ __forceinline__ double het_Gam_HOBr_Aer(const HetCtx &x, double radius, double C_Hp, double C_Clm, double C_Brm)
{
  const double INV_AB = 1.0 / 0.6, D_l = 1.4e-5;
  const double H_X = het_henry(x, HETIND_HOBr);
  const double cavg = het_cavg(x, x.mw[HETIND_HOBr]);
  const double C_Hp1 = fmax(fmin(C_Hp, 1.0e-6), 1.0e-9), C_Hp2 = fmax(fmin(C_Hp, 1.0e-2), 1.0e-6);
  const double k_tot = 2.3e+10 * C_Clm * C_Hp1 + 1.6e+10 * C_Brm * C_Hp2;
  double gamma = 0.0;
  if (k_tot > 0.0) {
    const double l_r = sqrt(D_l / k_tot);
    double gb_tot = x.m.FOUR_R_T * H_X * l_r * k_tot / cavg;
    gb_tot = gb_tot * het_ReactoDiff_Corr(radius, l_r);
    gamma = 1.0 / (INV_AB + 1.0 / gb_tot);
  }
  return gamma;
}
struct HOBrCld { double gamma, k_tot, k_Cl, k_Br, k_HSO3, k_SO3; };
__device__ __forceinline__ HOBrCld het_Gam_HOBr_Cld(const HetCtx &x)
{
  const HetCell2 &G = x.G;
  const double INV_AB = 1.0 / 0.6, D_l = 1.4e-5;
  HOBrCld r;
  const double H_X = het_henry(x, HETIND_HOBr);
  const double cavg = het_cavg(x, x.mw[HETIND_HOBr]);
  double C_Hp1 = fmin(G.H_conc_LCl, 1.0e-6), C_Hp2 = fmin(G.H_conc_LCl, 1.0e-2);
  C_Hp1 = fmax(C_Hp1, 1.0e-9); C_Hp2 = fmax(C_Hp2, 1.0e-6);
  r.k_Cl = 2.3e+10 * G.Cl_conc_Cld * C_Hp1;
  r.k_Br = 1.6e+10 * G.Br_conc_Cld * C_Hp2;
  r.k_HSO3 = 2.6e+7 * G.HSO3_aq;
  r.k_SO3 = 5.0e+9 * G.SO3_aq;
  r.k_tot = r.k_Cl + r.k_Br + r.k_HSO3 + r.k_SO3;
  r.gamma = 0.0;
  if (r.k_tot > 0.0) {
    const double l_r = sqrt(D_l / r.k_tot);
    double gb_tot = x.m.FOUR_R_T * H_X * l_r * r.k_tot / cavg;
    gb_tot = gb_tot * het_ReactoDiff_Corr(G.rLiq, l_r);
    r.gamma = 1.0 / (INV_AB + 1.0 / gb_tot);
  }
  return r;
}
__device__ __forceinline__ void het_Gam_HOBr_Ice(const HetCtx &x, double &gamma, double &brHCl, double &brHBr)
{
  const double gHCl = x.G.HCl_theta * 0.25, gHBr = x.G.HBr_theta * 4.8e-4 * exp(1240.0 / x.m.TEMP);
  gamma = gHCl + gHBr;
  brHCl = 0.0; brHBr = 0.0;
  if (gamma > 0.0) { brHCl = gHCl / gamma; brHBr = gHBr / gamma; }
}
// HOBr + HBr (byHBr) or + HCl in cloud, on stratospheric aerosol and on ice
__device__ __forceinline__ double het_HOBrUptkByHX(const HetCtx &x, bool byHBr)
{
  const HetCell &H = x.H; const HetCell2 &G = x.G;
  double k = 0.0, brIce = 0.0, brLiq = 0.0, gammaIce = 0.0, gammaLiq = 0.0, dummy;
  const double srMw = x.srmw[HETIND_HOBr];
  if (H.stratBox != 0.0) {
    gammaLiq = byHBr ? 0.25 : 0.2;
    k = k + x.Ars(H.xArea[HA_SUL], H.xRadi[HA_SUL], gammaLiq, srMw);
    k = k + H.xArea[HA_SLA] * G.KHETI_SLA[byHBr ? SLA_HOBr_HBr : SLA_HOBr_HCl];
    gammaIce = 0.3;
    if (G.natSurface != 0.0) gammaIce = byHBr ? 0.001 : 0.1;
    k = k + x.Ars(H.xArea[HA_IIC], H.xRadi[HA_IIC], gammaIce, srMw);
  } else {
    const HOBrCld c = het_Gam_HOBr_Cld(x);
    gammaLiq = c.gamma;
    const double branch_0 = (c.k_Cl + c.k_Br) / c.k_tot;
    double branch = branch_0 * (byHBr ? 0.9 : 0.1);
    if (G.Br_over_Cl_Cld <= 5.0e-4) branch = branch_0 * (byHBr ? het_Br2_Yield(G.Br_over_Cl_Cld) : (1.0 - het_Br2_Yield(G.Br_over_Cl_Cld)));
    brLiq = branch * (byHBr ? G.frac_Br_CldG : G.frac_Cl_CldG);
    if (byHBr) het_Gam_HOBr_Ice(x, gammaIce, dummy, brIce); else het_Gam_HOBr_Ice(x, gammaIce, brIce, dummy);
    k = k + het_CloudHet(x, srMw, gammaLiq, gammaIce, brLiq, brIce);
  }
  return het_kIIR1Ltd(x.C(HETIND_HOBr), x.C(byHBr ? HETIND_HBr : HETIND_HCl), k);
}
__device__ __forceinline__ double het2_HOBrUptkByHBr(const HetCtx &x) { return het_HOBrUptkByHX(x, true); }
__device__ __forceinline__ double het2_HOBrUptkByHCl(const HetCtx &x) { return het_HOBrUptkByHX(x, false); }
// HOBr on sea salt: bromide (toBr) or chloride channel, fine (SSA) or coarse (SSC) mode
__device__ __forceinline__ double het_HOBrUptkBySeaSalt(const HetCtx &x, bool toBr, bool coarse)
{
  const HetCell &H = x.H; const HetCell2 &G = x.G;
  double k = 0.0;
  const double srMw = x.srmw[HETIND_HOBr];
  if (H.stratBox == 0.0) {
    const HOBrCld c = het_Gam_HOBr_Cld(x);
    const double branch_0 = (c.k_Cl + c.k_Br) / c.k_tot;
    double branch = branch_0 * (toBr ? 0.9 : 0.1);
    if (G.Br_over_Cl_Cld <= 5.0e-4) branch = branch_0 * (toBr ? het_Br2_Yield(G.Br_over_Cl_Cld) : (1.0 - het_Br2_Yield(G.Br_over_Cl_Cld)));
    const double frac = toBr ? (coarse ? G.frac_Br_CldC : G.frac_Br_CldA) : (coarse ? G.frac_Cl_CldC : G.frac_Cl_CldA);
    k = k + het_CloudHet(x, srMw, c.gamma, 0.0, branch * frac, 0.0);
  }
  if ((coarse ? H.SSC_is_Acid : H.SSA_is_Acid) != 0.0) {
    const double radius = coarse ? H.xRadi[HA_SSC] : H.aClRadi;
    const double gammaAer = coarse ? het_Gam_HOBr_Aer(x, radius, G.H_conc_SSC, H.Cl_conc_SSC, G.Br_conc_SSC)
                                   : het_Gam_HOBr_Aer(x, radius, G.H_conc_SSA, H.Cl_conc_SSA, G.Br_conc_SSA);
    const double ratio = coarse ? G.Br_over_Cl_SSC : G.Br_over_Cl_SSA;
    double branch = toBr ? 0.9 : 0.1;
    if (ratio <= 5.0e-4) branch = toBr ? het_Br2_Yield(ratio) : 1.0 - het_Br2_Yield(ratio);
    const double area = coarse ? H.ClearFr * H.xArea[HA_SSC] * H.f_Acid_SSC : H.ClearFr * H.aClArea * H.f_Acid_SSA;
    k = k + x.Ars(area, radius, gammaAer, srMw) * branch;
  }
  const int educt = toBr ? (coarse ? HETIND_BrSALC : HETIND_BrSALA) : (coarse ? HETIND_SALCCL : HETIND_SALACL);
  return het_kIIR1Ltd(x.C(HETIND_HOBr), x.C(educt), k);
}
__device__ __forceinline__ double het2_HOBrUptkByBrSALA(const HetCtx &x) { return het_HOBrUptkBySeaSalt(x, true, false); }
__device__ __forceinline__ double het2_HOBrUptkByBrSALC(const HetCtx &x) { return het_HOBrUptkBySeaSalt(x, true, true); }
__device__ __forceinline__ double het2_HOBrUptkBySALACL(const HetCtx &x) { return het_HOBrUptkBySeaSalt(x, false, false); }
__device__ __forceinline__ double het2_HOBrUptkBySALCCL(const HetCtx &x) { return het_HOBrUptkBySeaSalt(x, false, true); }
__device__ __forceinline__ double het2_HOBrUptkByHSO3m(const HetCtx &x)
{
  double k = 0.0;
  if (x.H.stratBox == 0.0) {
    const HOBrCld c = het_Gam_HOBr_Cld(x);
    k = k + het_CloudHet(x, x.srmw[HETIND_HOBr], c.gamma, 0.0, c.k_HSO3 / c.k_tot, 0.0);
  }
  return het_kIIR1Ltd(x.C(HETIND_HOBr), x.C(HETIND_SO2), k);
}

// ---- HOCl
__device__ __forceinline__ void het_Gam_HOCl_Cld(const HetCtx &x, double &gamma, double &branchCl, double &branchSO3)
{
  const HetCell2 &G = x.G;
  const double INV_AB = 1.0 / 0.8, D_l = 2.0e-5;
  const double k_Cl = 1.5e+4 * G.H_conc_LCl * G.Cl_conc_Cld, k_SO3 = 2.8e+5 * G.TSO3_aq;
  const double k_tot = k_Cl + k_SO3;
  gamma = 0.0; branchCl = 0.0; branchSO3 = 0.0;
  if (k_tot > 0.0) {
    const double cavg = het_cavg(x, x.mw[HETIND_HOCl]);
    const double H_X = (x.hk0[HETIND_HOCl] * HET_CON_ATM_BAR) * exp(x.hcr[HETIND_HOCl] * (x.m.INV_TEMP - HET_INV_T298));
    const double l_r = sqrt(D_l / k_tot);
    double gb_tot = x.m.FOUR_R_T * H_X * l_r * k_tot / cavg;
    gb_tot = gb_tot * het_ReactoDiff_Corr(G.rLiq, l_r);
    gamma = 1.0 / (INV_AB + 1.0 / gb_tot);
    branchCl = k_Cl / k_tot;
    branchSO3 = k_SO3 / k_tot;
  }
}
__device__ __forceinline__ double het_Gam_HOCl_Aer(const HetCtx &x, double radius, double C_Hp, double C_Cl)
{
  const double INV_AB = 1.0 / 0.8, D_l = 2.0e-5, K_TER = 1.5e+4;
  if (!(C_Cl > 0.0)) return 0.0;
  const double cavg = het_cavg(x, x.mw[HETIND_HOCl]);
  const double H_X = (x.hk0[HETIND_HOCl] * HET_CON_ATM_BAR) * exp(x.hcr[HETIND_HOCl] * (x.m.INV_TEMP - HET_INV_T298));
  const double l_r = sqrt(D_l / (K_TER * C_Hp * C_Cl));
  double gb = x.m.FOUR_R_T * H_X * l_r * K_TER * C_Hp * C_Cl / cavg;
  gb = gb * het_ReactoDiff_Corr(radius, l_r);
  return 1.0 / (INV_AB + 1.0 / gb);
}
__device__ __forceinline__ double het2_HOClUptkByHCl(const HetCtx &x)
{
  const HetCell &H = x.H; const HetCell2 &G = x.G;
  double k = 0.0, gamma, branchCl, dummy;
  const double srMw = x.srmw[HETIND_HOCl];
  if (H.stratBox != 0.0) {
    k = k + x.Ars(H.xArea[HA_SUL], H.xRadi[HA_SUL], 0.8, srMw);
    k = k + H.xArea[HA_SLA] * G.KHETI_SLA[SLA_HOCl_HCl];
    gamma = 0.2;
    if (G.natSurface != 0.0) gamma = 0.1;
    k = k + x.Ars(H.xArea[HA_IIC], H.xRadi[HA_IIC], gamma, srMw);
    return het_kIIR1Ltd(x.C(HETIND_HOCl), x.C(HETIND_HCl), k);
  }
  het_Gam_HOCl_Cld(x, gamma, branchCl, dummy);
  const double branch = branchCl * G.frac_Cl_CldG;
  k = k + het_CloudHet(x, srMw, gamma, 0.22 * G.HCl_theta, branch, 1.0);
  return het_kIIR1Ltd(x.C(HETIND_HOCl), x.C(HETIND_HCl), k);
}
__device__ __forceinline__ double het2_HOClUptkByHBr(const HetCtx &x)
{
  const HetCell &H = x.H; const HetCell2 &G = x.G;
  double k = 0.0;
  const double srMw = x.srmw[HETIND_HOCl];
  if (H.stratBox != 0.0) {
    k = k + x.Ars(H.xArea[HA_SUL], H.xRadi[HA_SUL], 0.8, srMw);
    k = k + H.xArea[HA_SLA] * G.KHETI_SLA[SLA_HOCl_HBr];
    k = k + x.Ars(H.xArea[HA_IIC], H.xRadi[HA_IIC], 0.3, srMw);
  }
  k = het_kIIR1Ltd(x.C(HETIND_HOCl), x.C(HETIND_HBr), k);
  if (G.TurnOffHetRates != 0.0) k = 0.0;
  return k;
}
__device__ __forceinline__ double het_HOClUptkBySAL(const HetCtx &x, bool coarse)
{
  const HetCell &H = x.H; const HetCell2 &G = x.G;
  double k = 0.0, gamma, branchCl, dummy;
  const double srMw = x.srmw[HETIND_HOCl];
  if (H.stratBox == 0.0) {
    het_Gam_HOCl_Cld(x, gamma, branchCl, dummy);
    k = k + het_CloudHet(x, srMw, gamma, 0.0, branchCl * (coarse ? G.frac_Cl_CldC : G.frac_Cl_CldA), 0.0);
  }
  if ((coarse ? H.SSC_is_Acid : H.SSA_is_Acid) != 0.0) {
    if (coarse) {
      gamma = het_Gam_HOCl_Aer(x, H.xRadi[HA_SSC], G.H_conc_SSC, H.Cl_conc_SSC);
      k = k + x.Ars(H.ClearFr * H.xArea[HA_SSC] * H.f_Acid_SSC, H.xRadi[HA_SSC], gamma, srMw);
    } else {
      gamma = het_Gam_HOCl_Aer(x, H.aClRadi, G.H_conc_SSA, H.Cl_conc_SSA);
      k = k + x.Ars(H.ClearFr * H.aClArea * H.f_Acid_SSA, H.aClRadi, gamma, srMw);
    }
  }
  return het_kIIR1Ltd(x.C(HETIND_HOCl), x.C(coarse ? HETIND_SALCCL : HETIND_SALACL), k);
}
__device__ __forceinline__ double het2_HOClUptkBySALACL(const HetCtx &x) { return het_HOClUptkBySAL(x, false); }
__device__ __forceinline__ double het2_HOClUptkBySALCCL(const HetCtx &x) { return het_HOClUptkBySAL(x, true); }
__device__ __forceinline__ double het2_HOClUptkByHSO3m(const HetCtx &x)
{
  double k = 0.0, gamma, dummy, branchSO3;
  if (x.H.stratBox == 0.0) {
    het_Gam_HOCl_Cld(x, gamma, dummy, branchSO3);
    k = k + het_CloudHet(x, x.srmw[HETIND_HOCl], gamma, 0.0, branchSO3 * x.G.frac_HSO3_aq, 0.0);
  }
  return het_kIIR1Ltd(x.C(HETIND_HOCl), x.C(HETIND_SO2), k) * x.G.HSO3m;
}

// ---- IONO2, N2O5 (cloud, stratospheric HCl), NO2, NO3
__device__ __forceinline__ double het2_IONO2uptkByH2O(const HetCtx &x)
{
  const HetCell &H = x.H; const HetCell2 &G = x.G;
  double k = 0.0;
  const double srMw = x.srmw[HETIND_IONO2];
  double gamma = fmax((0.0021 * x.m.TEMP - 0.561), 0.0);
  k = k + x.Ars(H.ClearFr * H.xArea[HA_SUL], H.xRadi[HA_SUL], gamma, srMw);
  k = k + H.xArea[HA_SLA] * G.KHETI_SLA[SLA_BrNO3_H2O];
  gamma = 0.3;
  if (G.natSurface != 0.0) gamma = 0.001;
  k = k + x.Ars(H.ClearFr * H.xArea[HA_IIC], H.xRadi[HA_IIC], gamma, srMw);
  k = k + het_CloudHet(x, srMw, 0.01, 0.01, 1.0, 1.0);
  return het_kIIR1Ltd(x.C(HETIND_IONO2), x.C(HETIND_H2O), k);
}
__device__ __forceinline__ double het2_N2O5uptkByCloud(const HetCtx &x)
{
  const double cst = 0.03 / 0.019, T = x.m.TEMP;
  const double gamma = cst * exp(-25.5265 + 9283.76 / T - 851801.0 / (T * T));
  return het_CloudHet(x, x.srmw[HETIND_N2O5], gamma, 0.02, 1.0, 1.0);
}
__device__ __forceinline__ double het2_N2O5uptkByStratHCl(const HetCtx &x)
{
  const HetCell &H = x.H; const HetCell2 &G = x.G;
  double k = 0.0;
  if (H.stratBox != 0.0) {
    k = k + (H.xArea[HA_SLA] * G.KHETI_SLA[SLA_N2O5_HCl]);
    double gamma = 0.03;
    if (G.natSurface != 0.0) gamma = 0.003;
    k = k + x.Ars(H.xArea[HA_IIC], H.xRadi[HA_IIC], gamma, x.srmw[HETIND_N2O5]);
  }
  return het_kIIR1Ltd(x.C(HETIND_N2O5), x.C(HETIND_HCl), k);
}
__device__ __forceinline__ double het2_NO2uptk1stOrdAndCloud(const HetCtx &x)
{
  const HetCell &H = x.H;
  double k = 0.0, gamma;
  const double srMw = x.srmw[HETIND_NO2], relhum = x.m.RELHUM;
#pragma unroll
  for (int a = HA_DU1; a < HA_SUL; a++) k = k + x.Ars(H.xArea[a], H.xRadi[a], 1.0e-8, srMw);
  k = k + x.Ars(H.xArea[HA_SUL], H.xRadi[HA_SUL], 5e-6, srMw);
  k = k + x.Ars(H.xArea[HA_BKC], H.xRadi[HA_BKC], 1e-4, srMw);
  k = k + x.Ars(H.xArea[HA_ORC], H.xRadi[HA_ORC], 1e-6, srMw);
  if (relhum < 40.0) gamma = 1.0e-8;
  else if (relhum > 70.0) gamma = 1.0e-4;
  else gamma = 1.0e-8 + (1e-4 - 1e-8) * (relhum - 40.0) / 30.0;
  k = k + x.Ars(H.xArea[HA_SSA], H.xRadi[HA_SSA], gamma, srMw);
  k = k + x.Ars(H.xArea[HA_SSC], H.xRadi[HA_SSC], gamma, srMw);
  k = k + x.Ars(H.xArea[HA_SLA], H.xRadi[HA_SLA], 1.0e-4, srMw);
  k = k + x.Ars(H.xArea[HA_IIC], H.xRadi[HA_IIC], 1.0e-4, srMw);
  k = k + het_CloudHet(x, srMw, 1.0e-8, 0.0, 1.0, 0.0);
  return k;
}
__device__ __forceinline__ double het_Gam_NO3(const HetCtx &x, double aArea, double aRadi, double aWater, double C_X)
{
  const double INV_AB = 1.0 / 1.3e-2;
  const double Vol = aArea * aRadi * 1.0e-3 / 3.0;
  const double WaterC = aWater / 18.0e+12 / Vol;
  const double cavg = het_cavg(x, x.mw[HETIND_NO3]);
  const double k_tot = (2.76e+6 * C_X) + (23.0 * WaterC);
  double gamma = 0.0;
  if (k_tot > 0.0) {
    const double H_X = 0.6 * HET_CON_ATM_BAR;
    const double l_r = sqrt(1.0e-5 / k_tot);
    double gb = x.m.FOUR_R_T * H_X * l_r * k_tot / cavg;
    gb = gb * het_ReactoDiff_Corr(aRadi, l_r);
    gamma = 1.0 / (INV_AB + 1.0 / gb);
  }
  return gamma;
}
__device__ __forceinline__ double het2_NO3uptk1stOrdAndCloud(const HetCtx &x)
{
  const HetCell &H = x.H;
  double k = 0.0;
  const double srMw = x.srmw[HETIND_NO3];
#pragma unroll
  for (int a = HA_DU1; a < HA_SUL; a++) k = k + x.Ars(H.xArea[a], H.xRadi[a], 0.01, srMw);
  const double gamma = (x.m.RELHUM < 50.0) ? 2.0e-4 : 1.0e-3;
  k = k + x.Ars(H.xArea[HA_BKC], H.xRadi[HA_BKC], gamma, srMw);
  k = k + x.Ars(H.xArea[HA_ORC], H.xRadi[HA_ORC], 0.005, srMw);
  k = k + x.Ars(H.xArea[HA_SLA], H.xRadi[HA_SLA], 0.1, srMw);
  k = k + x.Ars(H.xArea[HA_IIC], H.xRadi[HA_IIC], 0.1, srMw);
  k = k + het_CloudHet(x, srMw, 0.002, 0.001, 1.0, 1.0);
  return k;
}
__device__ __forceinline__ double het2_NO3hypsisClonSALA(const HetCtx &x)
{
  const HetCell &H = x.H;
  const double gamma = het_Gam_NO3(x, H.aClArea, H.aClRadi, x.G.aWater[0], H.Cl_conc_SSA) * 0.01;
  return x.Ars(H.ClearFr * H.aClArea, H.aClRadi, gamma, x.srmw[HETIND_NO3]);
}
__device__ __forceinline__ double het2_NO3hypsisClonSALC(const HetCtx &x)
{
  const HetCell &H = x.H;
  const double gamma = het_Gam_NO3(x, H.xArea[HA_SSC], H.xRadi[HA_SSC], x.G.aWater[1], H.Cl_conc_SSC) * 0.01;
  return x.Ars(H.ClearFr * H.xArea[HA_SSC], H.xRadi[HA_SSC], gamma, x.srmw[HETIND_NO3]);
}

// ---- O3 + bromide
__device__ __forceinline__ double het_Gamma_O3_Br(const HetCtx &x, double Radius, double C_Br)
{
  if (!(C_Br > 0.0)) return 0.0;
  const double K0_O3 = 1.1e-2 * HET_CON_ATM_BAR;
  const double H_X = K0_O3 * exp(2300.0 * (x.m.INV_TEMP - HET_INV_T298));
  const double cavg = het_cavg(x, x.mw[HETIND_O3]);
  const double Nmax = 3.0e+14, KLangC = 1.0e-13, k_s = 1.0e-16;
  const double C_Br_surf = fmin(3.41e+14 * C_Br, Nmax);
  const double gs = (4.0 * k_s * C_Br_surf * KLangC * Nmax) / (cavg * (1.0 + KLangC * x.C(HETIND_O3)));
  const double k_b = 6.3e+8 * exp(-4.45e+3 / x.m.TEMP);
  const double D_l = 8.9e-6;
  const double l_r = sqrt(D_l / (k_b * C_Br));
  double gb = x.m.FOUR_R_T * H_X * l_r * k_b * C_Br / cavg;
  gb = gb * het_ReactoDiff_Corr(Radius, l_r);
  return gb + gs;
}
__device__ __forceinline__ double het_O3uptkByBrInTropCloud(const HetCtx &x, double Br_branch)
{
  if (x.H.stratBox != 0.0) return 0.0;
  const double gamma = het_Gamma_O3_Br(x, x.G.rLiq, x.G.Br_conc_Cld);
  return het_CloudHet(x, x.srmw[HETIND_O3], gamma, 0.0, Br_branch, 0.0);
}
__device__ __forceinline__ double het2_O3uptkByHBr(const HetCtx &x)
{
  return het_kIIR1Ltd(x.C(HETIND_O3), x.C(HETIND_HBr), het_O3uptkByBrInTropCloud(x, x.G.frac_Br_CldG));
}
__device__ __forceinline__ double het_O3uptkByBrSAL(const HetCtx &x, bool coarse)
{
  const HetCell &H = x.H; const HetCell2 &G = x.G;
  if (H.stratBox != 0.0) return 0.0;
  double k = 0.0 + het_O3uptkByBrInTropCloud(x, coarse ? G.frac_Br_CldC : G.frac_Br_CldA);
  if ((coarse ? H.SSC_is_Acid : H.SSA_is_Acid) != 0.0) {
    if (coarse) k = k + x.Ars(H.ClearFr * H.xArea[HA_SSC] * H.f_Acid_SSC, H.xRadi[HA_SSC], het_Gamma_O3_Br(x, H.xRadi[HA_SSC], G.Br_conc_SSC), x.srmw[HETIND_O3]);
    else k = k + x.Ars(H.ClearFr * H.aClArea * H.f_Acid_SSA, H.aClRadi, het_Gamma_O3_Br(x, H.aClRadi, G.Br_conc_SSA), x.srmw[HETIND_O3]);
  }
  return het_kIIR1Ltd(x.C(HETIND_O3), x.C(coarse ? HETIND_BrSALC : HETIND_BrSALA), k);
}
__device__ __forceinline__ double het2_O3uptkByBrSALA(const HetCtx &x) { return het_O3uptkByBrSAL(x, false); }
__device__ __forceinline__ double het2_O3uptkByBrSALC(const HetCtx &x) { return het_O3uptkByBrSAL(x, true); }

// ---- N2O5 on aerosol with an organic coating (McDuffie 2018 / Bertram & Thornton 2009): N2O5_InorgOrg :2699-2858,
//      ClNO2_BT :2860-2879, N2O5uptkByH2O :2583-2641, N2O5uptkBySALACl :2643-2670, N2O5uptkBySALCCl :2672-2697
#define HET_AVO 6.022140857e+23
__device__ __forceinline__ double het_ClNO2_BT(double Cl, double H2O)
{
  const double k2k3 = 1.0 / 4.5e+2;
  if (H2O < 0.1) return (Cl > 1e-3) ? 1.0 : 0.0;
  return 1.0 / (1.0 + k2k3 * het_SafeDiv(H2O, Cl, 1.0e+30));
}
struct N2O5IO { double gamma, Y_ClNO2, Rp, areaTotal; };
__device__ __forceinline__ N2O5IO het_N2O5_InorgOrg(const HetCtx &x, double volInorg, double volOrg, double H2Oinorg, double H2Oorg,
                                                    double Rcore, double NIT, double Cl)
{
  const double KH = 5.1e+1, k3k2b = 4.0e-2, beta = 1.15e+6, delta = 1.3e-1, Haq = 5e+3, Daq = 1e-9, ONE_THIRD = 1.0 / 3.0;
  N2O5IO r;
  const double volTotal = volInorg + volOrg, H2Ototal = H2Oinorg + H2Oorg;
  const double volRatioDry = het_SafeDiv(fmax(volInorg - H2Oinorg, 0.0), fmax(volTotal - H2Ototal, 0.0), 0.0);
  r.Rp = het_SafeDiv(Rcore, pow(volRatioDry, ONE_THIRD), Rcore);
  const double l = r.Rp - Rcore;
  double speed = sqrt(x.m.EIGHT_RSTARG_T / (HET_PI * (x.mw[HETIND_N2O5] * 1.0e-3)));
  const double M_H2O = H2Ototal / 18e+0 / volTotal * 1000.0;
  const double M_NIT = NIT / volTotal / HET_AVO * 1000.0;
  const double M_Cl = Cl / volTotal / HET_AVO * 1000.0;
  const double OCratio = (((x.K.OMOC_POA + x.K.OMOC_OPOA) / 2.0) - 1.17) / 1.29;
  const double eps = 1.5e-1 * OCratio + 1.6e-3 * x.m.RELHUM;
  double gamma_coat, gamma_core;
  if (l <= 0.0) gamma_coat = 0.0;
  else gamma_coat = (x.m.FOUR_RGASLATM_T * 1.0e-3 * eps * Haq * Daq * Rcore / 100.0) / (speed * l / 100.0 * r.Rp / 100.0);
  r.areaTotal = 3.0 * volTotal / r.Rp;
  if (M_H2O < 0.1) {
    gamma_core = 0.005;
  } else {
    speed = speed * 1e+2;
    double A = ((4.0 * volTotal) / (speed * r.areaTotal)) * KH;
    A = fmin(A, 3.2e-8);
    double k2f;
    if (delta * M_H2O < 1e-2) k2f = beta * (delta * M_H2O);
    else k2f = beta * (1e+0 - exp(-delta * M_H2O));
    gamma_core = A * k2f * (1.0 - 1.0 / (1.0 + het_SafeDiv(k3k2b * M_H2O, M_NIT, 1.0e+30)));
  }
  if (gamma_coat <= 0.0) r.gamma = gamma_core;
  else if (gamma_core <= 0.0) r.gamma = 0.0;
  else r.gamma = 1.0 / ((1.0 / gamma_core) + (1.0 / gamma_coat));
  r.Y_ClNO2 = het_ClNO2_BT(M_Cl, M_H2O);
  return r;
}
__device__ __forceinline__ N2O5IO het_N2O5_fine(const HetCtx &x)
{
  return het_N2O5_InorgOrg(x, x.K.AClVol, x.K.xVol_ORC, x.K.xH2O_SUL, x.K.xH2O_ORC, x.H.aClRadi, x.C(HETIND_NIT), x.C(HETIND_SALACL));
}
__device__ __forceinline__ N2O5IO het_N2O5_coarse(const HetCtx &x)
{
  return het_N2O5_InorgOrg(x, x.K.xVol_SSC, 0.0, x.K.xH2O_SSC, 0.0, x.H.xRadi[HA_SSC], x.C(HETIND_NITs), x.C(HETIND_SALCCL));
}
__device__ __forceinline__ double het2_N2O5uptkByH2O(const HetCtx &x)
{
  const HetCell &H = x.H; const HetCell2 &G = x.G;
  double k = 0.0, ktmp;
  const double srMw = x.srmw[HETIND_N2O5];
#pragma unroll
  for (int a = HA_DU1; a < HA_SUL; a++) k = k + x.Ars(H.ClearFr * H.xArea[a], H.xRadi[a], 0.02, srMw);
  N2O5IO r = het_N2O5_fine(x);
  ktmp = x.Ars(H.ClearFr * r.areaTotal, r.Rp, r.gamma, srMw);
  k = k + ktmp - (ktmp * r.Y_ClNO2 * 0.25);
  k = k + x.Ars(H.ClearFr * H.xArea[HA_BKC], H.xRadi[HA_BKC], 0.005, srMw);
  r = het_N2O5_coarse(x);
  ktmp = x.Ars(H.ClearFr * r.areaTotal, r.Rp, r.gamma, srMw);
  k = k + ktmp - (ktmp * r.Y_ClNO2);
  k = k + H.xArea[HA_SLA] * G.KHETI_SLA[SLA_N2O5_H2O];
  double gamma = 0.02;
  if (G.natSurface != 0.0) gamma = 4.0e-4;
  k = k + x.Ars(H.ClearFr * H.xArea[HA_IIC], H.xRadi[HA_IIC], gamma, srMw);
  return het_kIIR1Ltd(x.C(HETIND_N2O5), x.C(HETIND_H2O), k);
}
__device__ __forceinline__ double het2_N2O5uptkBySALACl(const HetCtx &x)
{
  if (x.H.stratBox != 0.0) return 0.0;
  const N2O5IO r = het_N2O5_fine(x);
  double k = x.Ars(x.H.ClearFr * r.areaTotal, r.Rp, r.gamma, x.srmw[HETIND_N2O5]);
  k = k * r.Y_ClNO2 * 0.25;
  return het_kIIR1Ltd(x.C(HETIND_N2O5), x.C(HETIND_SALACL), k);
}
__device__ __forceinline__ double het2_N2O5uptkBySALCCl(const HetCtx &x)
{
  if (x.H.stratBox != 0.0) return 0.0;
  const N2O5IO r = het_N2O5_coarse(x);
  double k = x.Ars(x.H.ClearFr * r.areaTotal, r.Rp, r.gamma, x.srmw[HETIND_N2O5]);
  k = k * r.Y_ClNO2;
  return het_kIIR1Ltd(x.C(HETIND_N2O5), x.C(HETIND_SALCCL), k);
}
