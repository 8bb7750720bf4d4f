/* CPU ORACLE -- TEST INFRASTRUCTURE ONLY (see kpp_oracle.h).
 * Auto-reduce Rosenbrock integrators (gckpp_Integrator.F90:789-1700): placeholder until the
 * restatement lands -- returns -99, which the reference itself treats as "fall back to the
 * standard solver" (gckpp_Integrator.F90:523). */
#include "kpp_oracle.h"
struct ros_ctx; struct ros_method;
int kpp_oracle_ar_integrate(void *c, const void *ros, double *Y, double Tstart, double Tend,
                            double *Tout, const double *AbsTol, const double *RelTol, int Autonomous,
                            int VectorTol, int Max_no_steps, double Roundoff, double Hmin, double Hmax,
                            double Hstart, double FacMin, double FacMax, double FacRej, double FacSafe,
                            double threshold, int target_spc, double thr_ratio, int append)
{
  (void)c; (void)ros; (void)Y; (void)Tstart; (void)Tend; (void)Tout; (void)AbsTol; (void)RelTol;
  (void)Autonomous; (void)VectorTol; (void)Max_no_steps; (void)Roundoff; (void)Hmin; (void)Hmax;
  (void)Hstart; (void)FacMin; (void)FacMax; (void)FacRej; (void)FacSafe; (void)threshold;
  (void)target_spc; (void)thr_ratio; (void)append;
  return -99;
}
