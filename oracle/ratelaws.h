/* CPU ORACLE -- TEST INFRASTRUCTURE ONLY.
 * Gas-phase rate laws, restated from KPP/fullchem/rateLawUtilFuncs.F90:41-72 (GCARR_*) and
 * KPP/fullchem/fullchem_RateLawFuncs.F90:107-800.  Each function keeps the reference's
 * operation order; Fortran `x**y` with real y -> pow(x,y), integer powers are written out
 * the way gfortran lowers them (x**2 = x*x, x**3 = x*x*x, x**(-8) = 1/(((x*x)^2)^2)).
 * The Fortran module variables TEMP, NUMDEN, H2O, ... travel in kpp_met_t. */
#ifndef KPP_RATELAWS_H
#define KPP_RATELAWS_H
#include <math.h>
#include "kpp_oracle.h"

#define RL static inline double
#define MAX0(x) ((x) > 0.0 ? (x) : 0.0)

/* rateLawUtilFuncs.F90:41-72 */
RL GCARR_ab(const kpp_met_t *m, double a0, double b0) { return a0 * pow(m->K300_OVER_TEMP, b0); }
RL GCARR_ac(const kpp_met_t *m, double a0, double c0) { return a0 * exp(c0 / m->TEMP); }
RL GCARR_abc(const kpp_met_t *m, double a0, double b0, double c0)
{ return a0 * exp(c0 / m->TEMP) * pow(m->K300_OVER_TEMP, b0); }

/* fullchem_RateLawFuncs.F90:107-141 */
RL ARRPLUS_ade(const kpp_met_t *m, double a0, double d0, double e0)
{ double k = a0 * (d0 + (m->TEMP * e0)); return MAX0(k); }
RL ARRPLUS_abde(const kpp_met_t *m, double a0, double b0, double d0, double e0)
{ double k = a0 * (d0 + (m->TEMP * e0)) * exp(-b0 / m->TEMP); return MAX0(k); }
RL TUNPLUS_abcde(const kpp_met_t *m, double a0, double b0, double c0, double d0, double e0)
{
  double T = m->TEMP;
  double k = a0 * (d0 + (T * e0));
  k = k * exp(b0 / T) * exp(c0 / (T * T * T));
  return MAX0(k);
}
/* :143-182 */
RL GC_ISO1(const kpp_met_t *m, double a0, double b0, double c0, double d0, double e0, double f0, double g0)
{
  double T = m->TEMP;
  double k0 = d0 * exp(e0 / T) * exp(1.0E8 / (T * T * T));
  double k1 = f0 * exp(g0 / T);
  double k2 = c0 * k0 / (k0 + k1);
  return a0 * exp(b0 / T) * (1.0 - k2);
}
RL GC_ISO2(const kpp_met_t *m, double a0, double b0, double c0, double d0, double e0, double f0, double g0)
{
  double T = m->TEMP;
  double k0 = d0 * exp(e0 / T) * exp(1.0E8 / (T * T * T));
  double k1 = f0 * exp(g0 / T);
  double k2 = c0 * k0 / (k0 + k1);
  return a0 * exp(b0 / T) * k2;
}
/* :184-196 */
RL GC_EPO_a(const kpp_met_t *m, double a1, double e1, double m1)
{
  double k1 = 1.0 / (m1 * m->NUMDEN + 1.0);
  return a1 * exp(e1 / m->TEMP) * k1;
}
/* :198-246 */
RL gc_pan_tail(double k0, double k1, double cf)
{
  double kr = k0 / k1;
  double nc = 0.75 - 1.27 * (log10(cf));
  double q = log10(kr) / nc;
  double f = pow(10.0, log10(cf) / (1.0 + q * q));
  return k0 * k1 * f / (k0 + k1);
}
RL GC_PAN_abab(const kpp_met_t *m, double a0, double b0, double a1, double b1, double cf)
{
  double k0 = a0 * exp(b0 / m->TEMP);
  double k1 = a1 * exp(b1 / m->TEMP);
  k0 = k0 * m->NUMDEN;
  return gc_pan_tail(k0, k1, cf);
}
RL GC_PAN_acac(const kpp_met_t *m, double a0, double c0, double a1, double c1, double cf)
{
  double k0 = a0 * pow(m->TEMP_OVER_K300, c0);
  double k1 = a1 * pow(m->TEMP_OVER_K300, c1);
  k0 = k0 * m->NUMDEN;
  return gc_pan_tail(k0, k1, cf);
}
/* :248-298 ; (TEMP/298)**(-8): gfortran lowers an integer power to repeated squaring and a reciprocal */
RL pow_m8(double x) { double x2 = x * x, x4 = x2 * x2, x8 = x4 * x4; return 1.0 / x8; }
RL GC_NIT(const kpp_met_t *m, double a0, double b0, double c0, double n, double x0, double y0)
{
  double T = m->TEMP;
  double k0 = 2.0E-22 * exp(n);
  double k1 = 4.3E-1 * pow_m8(T / 298.0);
  k0 = k0 * m->NUMDEN;
  k1 = k0 / k1;
  double l = log10(k1);
  double k2 = (k0 / (1.0 + k1)) * pow(4.1E-1, 1.0 / (1.0 + l * l));
  double k3 = k2 / (k2 + c0);
  double k4 = a0 * (x0 - T * y0);
  double k = k4 * exp(b0 / T) * k3;
  return MAX0(k);
}
RL GC_ALK(const kpp_met_t *m, double a0, double b0, double c0, double n, double x0, double y0)
{
  double T = m->TEMP;
  double k0 = 2.0E-22 * exp(n);
  double k1 = 4.3E-1 * pow_m8(T / 298.0);
  k0 = k0 * m->NUMDEN;
  k1 = k0 / k1;
  double l = log10(k1);
  double k2 = (k0 / (1.0 + k1)) * pow(4.1E-1, 1.0 / (1.0 + l * l));
  double k3 = c0 / (k2 + c0);
  double k4 = a0 * (x0 - T * y0);
  double k = k4 * exp(b0 / T) * k3;
  return MAX0(k);
}
/* :300-352 */
RL GC_HO2HO2_acac(const kpp_met_t *m, double a0, double c0, double a1, double c1)
{
  double k0 = a0 * exp(c0 / m->TEMP);
  double k1 = a1 * exp(c1 / m->TEMP);
  return (k0 + k1 * m->NUMDEN) * (1.0 + 1.4E-21 * m->H2O * exp(2200.0 / m->TEMP));
}
RL GC_TBRANCH_1_acac(const kpp_met_t *m, double a0, double c0, double a1, double c1)
{
  double k0 = a0 * exp(c0 / m->TEMP);
  double k1 = a1 * exp(c1 / m->TEMP);
  return k0 / (1.0 + k1);
}
RL GC_RO2HO2_aca(const kpp_met_t *m, double a0, double c0, double a1)
{
  double k = a0 * exp(c0 / m->TEMP);
  return k * (1.0 - exp(-0.245 * a1));
}
/* :354-392 */
RL GC_DMSOH_acac(const kpp_met_t *m, double a0, double c0, double a1, double c1)
{
  double k0 = a0 * exp(c0 / m->TEMP);
  double k1 = a1 * exp(c1 / m->TEMP);
  return (k0 * m->NUMDEN * 0.2095e0) / (1.0 + k1 * 0.2095e0);
}
RL GC_GLYXNO3_ac(const kpp_met_t *m, double a0, double c0)
{
  double O2 = m->NUMDEN * 0.2095;
  double k = a0 * exp(c0 / m->TEMP);
  return k * (O2 + 3.5E+18) / (2.0 * O2 + 3.5E+18);
}
/* :394-470 */
RL GC_GLYCOH_A_a(const kpp_met_t *m, double a0)
{
  const double exp_arg = -1.0 / 73.0;
  double f = 1.0 - 11.0729 * exp(exp_arg * m->TEMP);
  f = MAX0(f);
  return a0 * f;
}
RL GC_GLYCOH_B_a(const kpp_met_t *m, double a0)
{
  const double exp_arg = -1.0 / 73.0;
  double f = 1.0 - 11.0729 * exp(exp_arg * m->TEMP);
  f = MAX0(f);
  return a0 * (1.0 - f);
}
RL GC_HACOH_A_ac(const kpp_met_t *m, double a0, double c0)
{
  const double exp_arg = -1.0 / 60.0;
  double k0 = a0 * exp(c0 / m->TEMP);
  double f = 1.0 - 23.7 * exp(exp_arg * m->TEMP);
  f = MAX0(f);
  return k0 * f;
}
RL GC_HACOH_B_ac(const kpp_met_t *m, double a0, double c0)
{
  const double exp_arg = -1.0 / 60.0;
  double k0 = a0 * exp(c0 / m->TEMP);
  double f = 1.0 - 23.7 * exp(exp_arg * m->TEMP);
  f = MAX0(f);
  return k0 * (1.0 - f);
}
/* :472-560 */
RL GC_RO2NO_A1_ac(const kpp_met_t *m, double a0, double c0) { return a0 * exp(c0 / m->TEMP) * 3.0e-4; }
RL GC_RO2NO_B1_ac(const kpp_met_t *m, double a0, double c0)
{
  const double one_minus_fyrno3 = 1.0 - 3.0e-4;
  return a0 * exp(c0 / m->TEMP) * one_minus_fyrno3;
}
RL gc_ro2no_fyrno3(double numden, double yyyn, double a1)
{
  double xxyn = 1.94e-22 * exp(0.97 * a1) * numden;
  double aaa = log10(xxyn / yyyn);
  double zzyn = (1.0 / (1.0 + (aaa * aaa)));
  double rarb = (xxyn / (1.0 + (xxyn / yyyn))) * (pow(0.411, zzyn));
  return (rarb / (1.0 + rarb));
}
RL GC_RO2NO_A2_aca(const kpp_met_t *m, double a0, double c0, double a1)
{
  double k0 = a0 * exp(c0 / m->TEMP);
  double yyyn = 0.826 * (pow(300.0 / m->TEMP, 8.1));
  return k0 * gc_ro2no_fyrno3(m->NUMDEN, yyyn, a1);
}
RL GC_RO2NO_B2_aca(const kpp_met_t *m, double a0, double c0, double a1)
{
  double k0 = a0 * exp(c0 / m->TEMP);
  double yyyn = 0.826 * (pow(m->K300_OVER_TEMP, 8.1));
  return k0 * (1.0 - gc_ro2no_fyrno3(m->NUMDEN, yyyn, a1));
}
/* :562-800  JPL fall-off family */
RL jpl_falloff(double rlow, double rhigh, double fv)
{
  double xyrat = rlow / rhigh;
  double blog = log10(xyrat);
  double fexp = 1.0 / (1.0 + (blog * blog));
  return rlow * (pow(fv, fexp)) / (1.0 + xyrat);
}
RL GCJPLPR_aa(const kpp_met_t *m, double a1, double a2, double fv)
{ return jpl_falloff(a1 * m->NUMDEN, a2, fv); }
RL GCJPLPR_aba(const kpp_met_t *m, double a1, double b1, double a2, double fv)
{ return jpl_falloff(a1 * (pow(m->K300_OVER_TEMP, b1)) * m->NUMDEN, a2, fv); }
RL GCJPLPR_abab(const kpp_met_t *m, double a1, double b1, double a2, double b2, double fv)
{
  double rlow = a1 * (pow(m->K300_OVER_TEMP, b1)) * m->NUMDEN;
  double rhigh = a2 * (pow(m->K300_OVER_TEMP, b2));
  return jpl_falloff(rlow, rhigh, fv);
}
RL GCJPLPR_abcabc(const kpp_met_t *m, double a1, double b1, double c1, double a2, double b2, double c2, double fv)
{
  double rlow = a1 * (pow(m->K300_OVER_TEMP, b1)) * exp(c1 / m->TEMP) * m->NUMDEN;
  double rhigh = a2 * (pow(m->K300_OVER_TEMP, b2)) * exp(c2 / m->TEMP);
  return jpl_falloff(rlow, rhigh, fv);
}
RL GCJPLEQ_acabab(const kpp_met_t *m, double a0, double c0, double a1, double b1, double a2, double b2, double fv)
{
  double k0 = a0 * exp(c0 / m->TEMP);
  double k1 = GCJPLPR_abab(m, a1, b1, a2, b2, fv);
  return k1 / k0;
}
RL GCJPLAC_ababac(const kpp_met_t *m, double a1, double b1, double a2, double b2, double a3, double c3, double fv)
{
  double rlow = a1 * (pow(m->K300_OVER_TEMP, b1)) * m->NUMDEN;
  double rhigh = a2 * (pow(m->K300_OVER_TEMP, b2));
  double k1 = jpl_falloff(rlow, rhigh, fv);
  double k2 = a3 * exp(c3 / m->TEMP);
  return k2 * (1.0 - (k1 / rhigh));
}
#undef RL
#endif
