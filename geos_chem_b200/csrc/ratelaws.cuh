// Gas-phase rate laws on the device (K1).  Same formulas as the reference's
// KPP/fullchem/rateLawUtilFuncs.F90:41-72 (Arrhenius forms) and
// KPP/fullchem/fullchem_RateLawFuncs.F90:107-800 (fall-off, branching, equilibrium forms),
// written for one cell per thread with the per-cell scalars in MetCell.
// Integer powers are explicit products; real powers use pow().  FP64 throughout.
#pragma once
#include "tables.h"

#define RLD __device__ __forceinline__ double

RLD rl_max0(double k) { return k > 0.0 ? k : 0.0; }
RLD rl_cube(double t) { return t * t * t; }

// --- Arrhenius family (rateLawUtilFuncs.F90:41-72)
RLD GCARR_ab(const MetCell &m, double a0, double b0) { return a0 * pow(m.K300_OVER_TEMP, b0); }
RLD GCARR_ac(const MetCell &m, double a0, double c0) { return a0 * exp(c0 / m.TEMP); }
RLD GCARR_abc(const MetCell &m, double a0, double b0, double c0)
{ return a0 * exp(c0 / m.TEMP) * pow(m.K300_OVER_TEMP, b0); }

// --- temperature-polynomial prefactors (fullchem_RateLawFuncs.F90:107-141)
RLD ARRPLUS_ade(const MetCell &m, double a0, double d0, double e0)
{ return rl_max0(a0 * (d0 + (m.TEMP * e0))); }
RLD ARRPLUS_abde(const MetCell &m, double a0, double b0, double d0, double e0)
{ return rl_max0(a0 * (d0 + (m.TEMP * e0)) * exp(-b0 / m.TEMP)); }
RLD TUNPLUS_abcde(const MetCell &m, double a0, double b0, double c0, double d0, double e0)
{
  double k = a0 * (d0 + (m.TEMP * e0));
  k = k * exp(b0 / m.TEMP) * exp(c0 / rl_cube(m.TEMP));
  return rl_max0(k);
}

// --- isoprene peroxy isomerisation branches (:143-182)
RLD rl_iso_k2(const MetCell &m, double c0, double d0, double e0, double f0, double g0)
{
  double k0 = d0 * exp(e0 / m.TEMP) * exp(1.0E8 / rl_cube(m.TEMP));
  double k1 = f0 * exp(g0 / m.TEMP);
  return c0 * k0 / (k0 + k1);
}
RLD GC_ISO1(const MetCell &m, double a0, double b0, double c0, double d0, double e0, double f0, double g0)
{ double k2 = rl_iso_k2(m, c0, d0, e0, f0, g0); return a0 * exp(b0 / m.TEMP) * (1.0 - k2); }
RLD GC_ISO2(const MetCell &m, double a0, double b0, double c0, double d0, double e0, double f0, double g0)
{ double k2 = rl_iso_k2(m, c0, d0, e0, f0, g0); return a0 * exp(b0 / m.TEMP) * k2; }

// --- epoxide formation (:184-196)
RLD GC_EPO_a(const MetCell &m, double a1, double e1, double m1)
{
  double k1 = 1.0 / (m1 * m.NUMDEN + 1.0);
  return a1 * exp(e1 / m.TEMP) * k1;
}

// --- PAN-type Troe expressions (:198-246)
RLD rl_troe_pan(double k0, double k1, double cf)
{
  double lcf = log10(cf);
  double kr = k0 / k1;
  double nc = 0.75 - 1.27 * lcf;
  double q = log10(kr) / nc;
  double f = pow(10.0, lcf / (1.0 + q * q));
  return k0 * k1 * f / (k0 + k1);
}
RLD GC_PAN_abab(const MetCell &m, double a0, double b0, double a1, double b1, double cf)
{
  double k0 = a0 * exp(b0 / m.TEMP);
  double k1 = a1 * exp(b1 / m.TEMP);
  return rl_troe_pan(k0 * m.NUMDEN, k1, cf);
}
RLD GC_PAN_acac(const MetCell &m, double a0, double c0, double a1, double c1, double cf)
{
  double k0 = a0 * pow(m.TEMP_OVER_K300, c0);
  double k1 = a1 * pow(m.TEMP_OVER_K300, c1);
  return rl_troe_pan(k0 * m.NUMDEN, k1, cf);
}

// --- organic nitrate yield parameterisation (:248-298)
RLD rl_nit_k2(const MetCell &m, double n)
{
  double k0 = 2.0E-22 * exp(n);
  double t = m.TEMP / 298.0;
  double t2 = t * t, t4 = t2 * t2;
  double k1 = 4.3E-1 * (1.0 / (t4 * t4));
  k0 = k0 * m.NUMDEN;
  k1 = k0 / k1;
  double l = log10(k1);
  return (k0 / (1.0 + k1)) * pow(4.1E-1, 1.0 / (1.0 + l * l));
}
RLD GC_NIT(const MetCell &m, double a0, double b0, double c0, double n, double x0, double y0)
{
  double k2 = rl_nit_k2(m, n);
  double k3 = k2 / (k2 + c0);
  double k4 = a0 * (x0 - m.TEMP * y0);
  return rl_max0(k4 * exp(b0 / m.TEMP) * k3);
}
RLD GC_ALK(const MetCell &m, double a0, double b0, double c0, double n, double x0, double y0)
{
  double k2 = rl_nit_k2(m, n);
  double k3 = c0 / (k2 + c0);
  double k4 = a0 * (x0 - m.TEMP * y0);
  return rl_max0(k4 * exp(b0 / m.TEMP) * k3);
}

// --- HO2 self reaction, branching, RO2+HO2 (:300-352)
RLD GC_HO2HO2_acac(const MetCell &m, double a0, double c0, double a1, double c1)
{
  double k0 = a0 * exp(c0 / m.TEMP);
  double k1 = a1 * exp(c1 / m.TEMP);
  return (k0 + k1 * m.NUMDEN) * (1.0 + 1.4E-21 * m.H2O * exp(2200.0 / m.TEMP));
}
RLD GC_TBRANCH_1_acac(const MetCell &m, double a0, double c0, double a1, double c1)
{
  double k0 = a0 * exp(c0 / m.TEMP);
  double k1 = a1 * exp(c1 / m.TEMP);
  return k0 / (1.0 + k1);
}
RLD GC_RO2HO2_aca(const MetCell &m, double a0, double c0, double a1)
{
  double k = a0 * exp(c0 / m.TEMP);
  return k * (1.0 - exp(-0.245 * a1));
}

// --- DMS + OH addition, glyoxal + NO3 (:354-392)
RLD GC_DMSOH_acac(const MetCell &m, double a0, double c0, double a1, double c1)
{
  double k0 = a0 * exp(c0 / m.TEMP);
  double k1 = a1 * exp(c1 / m.TEMP);
  return (k0 * m.NUMDEN * 0.2095e0) / (1.0 + k1 * 0.2095e0);
}
RLD GC_GLYXNO3_ac(const MetCell &m, double a0, double c0)
{
  double O2 = m.NUMDEN * 0.2095;
  double k = a0 * exp(c0 / m.TEMP);
  return k * (O2 + 3.5E+18) / (2.0 * O2 + 3.5E+18);
}

// --- glycolaldehyde / hydroxyacetone + OH branches (:394-470)
RLD rl_frac(const MetCell &m, double pre, double tscale)
{
  double f = 1.0 - pre * exp((-1.0 / tscale) * m.TEMP);
  return rl_max0(f);
}
RLD GC_GLYCOH_A_a(const MetCell &m, double a0) { return a0 * rl_frac(m, 11.0729, 73.0); }
RLD GC_GLYCOH_B_a(const MetCell &m, double a0) { return a0 * (1.0 - rl_frac(m, 11.0729, 73.0)); }
RLD GC_HACOH_A_ac(const MetCell &m, double a0, double c0)
{ double k0 = a0 * exp(c0 / m.TEMP); return k0 * rl_frac(m, 23.7, 60.0); }
RLD GC_HACOH_B_ac(const MetCell &m, double a0, double c0)
{ double k0 = a0 * exp(c0 / m.TEMP); return k0 * (1.0 - rl_frac(m, 23.7, 60.0)); }

// --- RO2 + NO nitrate branching (:472-560)
RLD GC_RO2NO_A1_ac(const MetCell &m, double a0, double c0) { return a0 * exp(c0 / m.TEMP) * 3.0e-4; }
RLD GC_RO2NO_B1_ac(const MetCell &m, double a0, double c0)
{ return a0 * exp(c0 / m.TEMP) * (1.0 - 3.0e-4); }
RLD rl_fyrno3(const MetCell &m, double yyyn, double a1)
{
  double xxyn = 1.94e-22 * exp(0.97 * a1) * m.NUMDEN;
  double aaa = log10(xxyn / yyyn);
  double zzyn = (1.0 / (1.0 + (aaa * aaa)));
  double rarb = (xxyn / (1.0 + (xxyn / yyyn))) * (pow(0.411, zzyn));
  return (rarb / (1.0 + rarb));
}
RLD GC_RO2NO_A2_aca(const MetCell &m, double a0, double c0, double a1)
{
  double k0 = a0 * exp(c0 / m.TEMP);
  double yyyn = 0.826 * (pow(300.0 / m.TEMP, 8.1));
  return k0 * rl_fyrno3(m, yyyn, a1);
}
RLD GC_RO2NO_B2_aca(const MetCell &m, double a0, double c0, double a1)
{
  double k0 = a0 * exp(c0 / m.TEMP);
  double yyyn = 0.826 * (pow(m.K300_OVER_TEMP, 8.1));
  return k0 * (1.0 - rl_fyrno3(m, yyyn, a1));
}

// --- JPL three-body fall-off family (:562-800)
RLD rl_jpl(double rlow, double rhigh, double fv)
{
  double xyrat = rlow / rhigh;
  double blog = log10(xyrat);
  double fexp = 1.0 / (1.0 + (blog * blog));
  return rlow * (pow(fv, fexp)) / (1.0 + xyrat);
}
RLD GCJPLPR_aa(const MetCell &m, double a1, double a2, double fv) { return rl_jpl(a1 * m.NUMDEN, a2, fv); }
RLD GCJPLPR_aba(const MetCell &m, double a1, double b1, double a2, double fv)
{ return rl_jpl(a1 * (pow(m.K300_OVER_TEMP, b1)) * m.NUMDEN, a2, fv); }
RLD GCJPLPR_abab(const MetCell &m, double a1, double b1, double a2, double b2, double fv)
{
  double rlow = a1 * (pow(m.K300_OVER_TEMP, b1)) * m.NUMDEN;
  double rhigh = a2 * (pow(m.K300_OVER_TEMP, b2));
  return rl_jpl(rlow, rhigh, fv);
}
RLD GCJPLPR_abcabc(const MetCell &m, double a1, double b1, double c1, double a2, double b2, double c2, double fv)
{
  double rlow = a1 * (pow(m.K300_OVER_TEMP, b1)) * exp(c1 / m.TEMP) * m.NUMDEN;
  double rhigh = a2 * (pow(m.K300_OVER_TEMP, b2)) * exp(c2 / m.TEMP);
  return rl_jpl(rlow, rhigh, fv);
}
RLD GCJPLEQ_acabab(const MetCell &m, double a0, double c0, double a1, double b1, double a2, double b2, double fv)
{
  double k0 = a0 * exp(c0 / m.TEMP);
  double k1 = GCJPLPR_abab(m, a1, b1, a2, b2, fv);
  return k1 / k0;
}
RLD GCJPLAC_ababac(const MetCell &m, double a1, double b1, double a2, double b2, double a3, double c3, double fv)
{
  double rlow = a1 * (pow(m.K300_OVER_TEMP, b1)) * m.NUMDEN;
  double rhigh = a2 * (pow(m.K300_OVER_TEMP, b2));
  double k1 = rl_jpl(rlow, rhigh, fv);
  double k2 = a3 * exp(c3 / m.TEMP);
  return k2 * (1.0 - (k1 / rhigh));
}
#undef RLD
