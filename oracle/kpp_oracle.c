/* CPU ORACLE -- TEST INFRASTRUCTURE ONLY (see kpp_oracle.h).
 *
 * Restates, in plain C and in the reference's operation order:
 *   Integrate          KPP/fullchem/gckpp_Integrator.F90:80-162
 *   Rosenbrock         :165-531   (option decoding, defaults, tolerance checks)
 *   ros_Integrator     :578-786   (adaptive Rosenbrock time loop)
 *   ros_ErrorNorm      :1715-1745
 *   ros_FunTimeDerivative :1749-1769
 *   ros_PrepareMatrix  :1921-1999
 *   ros_Decomp/Solve   :2003-2057
 *   Ros2..Rang3 tables :2062-2476  (Rodas3 = :2239-2303 is what GEOS-Chem selects)
 *   KppDecomp          KPP/fullchem/gckpp_LinearAlgebra.F90:46-83
 *   WLAMCH             :3861-3895
 * and the cell loop of Do_FullChem (GeosCore/fullchem_mod.F90:528-546) for batches.
 * The Hg integrator text is identical apart from which Fun it calls
 * (KPP/Hg/gckpp_Integrator.F90), so one implementation serves both; the carbon mechanism
 * uses forward Euler (KPP/carbon/gckpp_Integrator.F90:155-215).
 * The auto-reduce integrators (:789-1700) live in kpp_oracle_ar.c.
 */
#include <float.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif
#include "kpp_oracle.h"

enum { Nfun = 0, Njac, Nstp, Nacc, Nrej, Ndec, Nsol, Nsng };
enum { Ntexit = 0, Nhexit, Nhnew, NARthr };

const kpp_mech_t *kpp_oracle_mech(int id)
{
  switch (id) {
  case 0: return &kpp_mech_fullchem;
  case 1: return &kpp_mech_Hg;
  case 2: return &kpp_mech_carbon;
  default: return 0;
  }
}

int kpp_oracle_dims(int id, int *d)
{
  const kpp_mech_t *m = kpp_oracle_mech(id);
  if (!m) return -1;
  d[0] = m->nvar; d[1] = m->nfix; d[2] = m->nspec; d[3] = m->nreact;
  d[4] = m->lu_nonzero; d[5] = m->nphot; d[6] = m->next;
  return 0;
}

/* ---- Rosenbrock method tables (gckpp_Integrator.F90:2062-2476) ------------------------ */
typedef struct {
  int S;
  double A[15], C[15], M[6], E[6], Alpha[6], Gamma[6], ELO;
  int NewF[6];
} ros_method_t;

static void method_Ros2(ros_method_t *r)
{ /* :2062-2105 */
  double g = 1.0 + 1.0 / sqrt(2.0);
  memset(r, 0, sizeof *r);
  r->S = 2;
  r->A[0] = (1.0) / g;
  r->C[0] = (-2.0) / g;
  r->NewF[0] = 1; r->NewF[1] = 1;
  r->M[0] = (3.0) / (2.0 * g);
  r->M[1] = (1.0) / (2.0 * g);
  r->E[0] = 1.0 / (2.0 * g);
  r->E[1] = 1.0 / (2.0 * g);
  r->ELO = 2.0;
  r->Alpha[0] = 0.0; r->Alpha[1] = 1.0;
  r->Gamma[0] = g; r->Gamma[1] = -g;
}
static void method_Ros3(ros_method_t *r)
{ /* :2108-2157 */
  memset(r, 0, sizeof *r);
  r->S = 3;
  r->A[0] = 1.0; r->A[1] = 1.0; r->A[2] = 0.0;
  r->C[0] = -0.10156171083877702091975600115545e+01;
  r->C[1] = 0.40759956452537699824805835358067e+01;
  r->C[2] = 0.92076794298330791242156818474003e+01;
  r->NewF[0] = 1; r->NewF[1] = 1; r->NewF[2] = 0;
  r->M[0] = 0.1e+01;
  r->M[1] = 0.61697947043828245592553615689730e+01;
  r->M[2] = -0.42772256543218573326238373806514;
  r->E[0] = 0.5;
  r->E[1] = -0.29079558716805469821718236208017e+01;
  r->E[2] = 0.22354069897811569627360909276199;
  r->ELO = 3.0;
  r->Alpha[0] = 0.0;
  r->Alpha[1] = 0.43586652150845899941601945119356;
  r->Alpha[2] = 0.43586652150845899941601945119356;
  r->Gamma[0] = 0.43586652150845899941601945119356;
  r->Gamma[1] = 0.24291996454816804366592249683314;
  r->Gamma[2] = 0.21851380027664058511513169485832e+01;
}
static void method_Ros4(ros_method_t *r)
{ /* :2164-2231 */
  memset(r, 0, sizeof *r);
  r->S = 4;
  r->A[0] = 0.2000000000000000e+01; r->A[1] = 0.1867943637803922e+01; r->A[2] = 0.2344449711399156;
  r->A[3] = r->A[1]; r->A[4] = r->A[2]; r->A[5] = 0.0;
  r->C[0] = -0.7137615036412310e+01; r->C[1] = 0.2580708087951457e+01; r->C[2] = 0.6515950076447975;
  r->C[3] = -0.2137148994382534e+01; r->C[4] = -0.3214669691237626; r->C[5] = -0.6949742501781779;
  r->NewF[0] = 1; r->NewF[1] = 1; r->NewF[2] = 1; r->NewF[3] = 0;
  r->M[0] = 0.2255570073418735e+01; r->M[1] = 0.2870493262186792; r->M[2] = 0.4353179431840180;
  r->M[3] = 0.1093502252409163e+01;
  r->E[0] = -0.2815431932141155; r->E[1] = -0.7276199124938920e-01; r->E[2] = -0.1082196201495311;
  r->E[3] = -0.1093502252409163e+01;
  r->ELO = 4.0;
  r->Alpha[0] = 0.0; r->Alpha[1] = 0.1145640000000000e+01; r->Alpha[2] = 0.6552168638155900;
  r->Alpha[3] = r->Alpha[2];
  r->Gamma[0] = 0.5728200000000000; r->Gamma[1] = -0.1769193891319233e+01;
  r->Gamma[2] = 0.7592633437920482; r->Gamma[3] = -0.1049021087100450;
}
static void method_Rodas3(ros_method_t *r)
{ /* :2239-2303 */
  memset(r, 0, sizeof *r);
  r->S = 4;
  r->A[0] = 0.0; r->A[1] = 2.0; r->A[2] = 0.0; r->A[3] = 2.0; r->A[4] = 0.0; r->A[5] = 1.0;
  r->C[0] = 4.0; r->C[1] = 1.0; r->C[2] = -1.0; r->C[3] = 1.0; r->C[4] = -1.0; r->C[5] = -(8.0 / 3.0);
  r->NewF[0] = 1; r->NewF[1] = 0; r->NewF[2] = 1; r->NewF[3] = 1;
  r->M[0] = 2.0; r->M[1] = 0.0; r->M[2] = 1.0; r->M[3] = 1.0;
  r->E[0] = 0.0; r->E[1] = 0.0; r->E[2] = 0.0; r->E[3] = 1.0;
  r->ELO = 3.0;
  r->Alpha[0] = 0.0; r->Alpha[1] = 0.0; r->Alpha[2] = 1.0; r->Alpha[3] = 1.0;
  r->Gamma[0] = 0.5; r->Gamma[1] = 1.5; r->Gamma[2] = 0.0; r->Gamma[3] = 0.0;
}
static void method_Rodas4(ros_method_t *r)
{ /* :2310-2405 */
  memset(r, 0, sizeof *r);
  r->S = 6;
  r->Alpha[0] = 0.000; r->Alpha[1] = 0.386; r->Alpha[2] = 0.210; r->Alpha[3] = 0.630;
  r->Alpha[4] = 1.000; r->Alpha[5] = 1.000;
  r->Gamma[0] = 0.2500000000000000; r->Gamma[1] = -0.1043000000000000; r->Gamma[2] = 0.1035000000000000;
  r->Gamma[3] = -0.3620000000000023e-01; r->Gamma[4] = 0.0; r->Gamma[5] = 0.0;
  r->A[0] = 0.1544000000000000e+01; r->A[1] = 0.9466785280815826; r->A[2] = 0.2557011698983284;
  r->A[3] = 0.3314825187068521e+01; r->A[4] = 0.2896124015972201e+01; r->A[5] = 0.9986419139977817;
  r->A[6] = 0.1221224509226641e+01; r->A[7] = 0.6019134481288629e+01; r->A[8] = 0.1253708332932087e+02;
  r->A[9] = -0.6878860361058950; r->A[10] = r->A[6]; r->A[11] = r->A[7]; r->A[12] = r->A[8];
  r->A[13] = r->A[9]; r->A[14] = 1.0;
  r->C[0] = -0.5668800000000000e+01; r->C[1] = -0.2430093356833875e+01; r->C[2] = -0.2063599157091915;
  r->C[3] = -0.1073529058151375; r->C[4] = -0.9594562251023355e+01; r->C[5] = -0.2047028614809616e+02;
  r->C[6] = 0.7496443313967647e+01; r->C[7] = -0.1024680431464352e+02; r->C[8] = -0.3399990352819905e+02;
  r->C[9] = 0.1170890893206160e+02; r->C[10] = 0.8083246795921522e+01; r->C[11] = -0.7981132988064893e+01;
  r->C[12] = -0.3152159432874371e+02; r->C[13] = 0.1631930543123136e+02; r->C[14] = -0.6058818238834054e+01;
  r->M[0] = r->A[6]; r->M[1] = r->A[7]; r->M[2] = r->A[8]; r->M[3] = r->A[9]; r->M[4] = 1.0; r->M[5] = 1.0;
  r->E[0] = 0.0; r->E[1] = 0.0; r->E[2] = 0.0; r->E[3] = 0.0; r->E[4] = 0.0; r->E[5] = 1.0;
  r->NewF[0] = 1; r->NewF[1] = 1; r->NewF[2] = 1; r->NewF[3] = 1; r->NewF[4] = 1; r->NewF[5] = 1;
  r->ELO = 4.0;
}
static void method_Rang3(ros_method_t *r)
{ /* :2412-2476 */
  memset(r, 0, sizeof *r);
  r->S = 4;
  r->A[0] = 5.09052051067020e+00; r->A[1] = 5.09052051067020e+00; r->A[2] = 0.0;
  r->A[3] = 4.97628111010787e+00; r->A[4] = 2.77268164715849e-02; r->A[5] = 2.29428036027904e-01;
  r->C[0] = -1.16790812312283e+01; r->C[1] = -1.64057326467367e+01; r->C[2] = -2.77268164715850e-01;
  r->C[3] = -8.38103960500476e+00; r->C[4] = -8.48328409199343e-01; r->C[5] = 2.87009860433106e-01;
  r->M[0] = 5.22582761233094e+00; r->M[1] = -5.56971148154165e-01; r->M[2] = 3.57979469353645e-01;
  r->M[3] = 1.72337398521064e+00;
  r->E[0] = -5.16845212784040e+00; r->E[1] = -1.26351942603842e+00; r->E[2] = -1.11022302462516e-16;
  r->E[3] = 2.22044604925031e-16;
  r->Alpha[0] = 0.0; r->Alpha[1] = 2.21878746765329e+00; r->Alpha[2] = 2.21878746765329e+00;
  r->Alpha[3] = 1.55392337535788e+00;
  r->Gamma[0] = 4.35866521508459e-01; r->Gamma[1] = -1.78292094614483e+00; r->Gamma[2] = -2.46541900496934e+00;
  r->Gamma[3] = -8.05529997906370e-01;
  r->NewF[0] = 1; r->NewF[1] = 1; r->NewF[2] = 1; r->NewF[3] = 1;
  r->ELO = 3.0;
}

/* WLAMCH('E') (gckpp_LinearAlgebra.F90:3861-3895): halves Eps from 0.5 until 1+Eps == 1, then
 * doubles it once: 2^-52 in IEEE double */
static double wlamch_eps(void)
{
  volatile double suma;
  double eps = pow(0.5, 1);
  int i;
  for (i = 1; i <= 80; i++) {
    eps = eps * 0.5;
    suma = 1.0 + eps;
    if (suma <= 1.0) break;
  }
  return eps * 2;
}

/* KppDecomp (gckpp_LinearAlgebra.F90:46-83): in-place row-wise sparse LU, no pivoting */
static int kpp_decomp(const kpp_mech_t *m, double *JVS, double *W)
{
  const int *crow = m->lu_crow, *diag = m->lu_diag, *icol = m->lu_icol;
  double a = 0.;
  int k, kk, j, jj;
  for (k = 0; k < m->nvar; k++) {
    if (fabs(JVS[diag[k]]) < DBL_MIN) return k + 1; /* TINY(a) */
    for (kk = crow[k]; kk < crow[k + 1]; kk++) W[icol[kk]] = JVS[kk];
    for (kk = crow[k]; kk < diag[k]; kk++) {
      j = icol[kk];
      a = -W[j] / JVS[diag[j]];
      W[j] = -a;
      for (jj = diag[j] + 1; jj < crow[j + 1]; jj++) W[icol[jj]] = W[icol[jj]] + a * JVS[jj];
    }
    for (kk = crow[k]; kk < crow[k + 1]; kk++) JVS[kk] = W[icol[kk]];
  }
  return 0;
}

typedef struct {
  const kpp_mech_t *m;
  const double *RCONST, *FIX;
  double *A, *P, *D, *B, *W; /* scratch */
  int *ISTATUS;
  double *RSTATUS;
} ros_ctx_t;

/* FunTemplate (gckpp_Integrator.F90:2487-2515) with ICNTRL(15) = -1 (no in-integrator rate update) */
static void fun_template(ros_ctx_t *c, const double *Y, double *Ydot)
{
  if (c->m->fun_is_split)
    c->m->fun_split(Y, c->FIX, c->RCONST, Ydot, c->P, c->D, c->A, 0);
  else
    c->m->fun(Y, c->FIX, c->RCONST, Ydot, c->A);
}

/* ros_ErrorNorm (:1715-1745) */
static double ros_error_norm(int N, const double *Y, const double *Ynew, const double *Yerr,
                             const double *AbsTol, const double *RelTol, int VectorTol)
{
  double Err = 0.0, Scale, Ymax;
  int i;
  for (i = 0; i < N; i++) {
    Ymax = fmax(fabs(Y[i]), fabs(Ynew[i]));
    if (VectorTol) Scale = AbsTol[i] + RelTol[i] * Ymax;
    else Scale = AbsTol[0] + RelTol[0] * Ymax;
    double q = Yerr[i] / Scale;
    Err = Err + q * q;
  }
  Err = sqrt(Err / N);
  return fmax(Err, 1.0e-10);
}

static void waxpy(int N, double a, const double *x, double *y)
{ /* WAXPY (gckpp_LinearAlgebra.F90:3758-3790): y = y + a*x, returns immediately when a == 0 */
  int i;
  if (a == 0.0) return;
  for (i = 0; i < N; i++) y[i] = y[i] + a * x[i];
}

/* ros_Integrator (:578-786).  Returns IERR. */
static int ros_integrator(ros_ctx_t *c, const ros_method_t *ros, double *Y, double Tstart, double Tend,
                          double *Tout, const double *AbsTol, const double *RelTol, int Autonomous,
                          int VectorTol, int Max_no_steps, double Roundoff, double Hmin, double Hmax,
                          double Hstart, double FacMin, double FacMax, double FacRej, double FacSafe,
                          int clip_negative)
{
  const kpp_mech_t *m = c->m;
  const int N = m->nvar, S = ros->S;
  const double DeltaMin = 1.0E-5;
  double Ynew[(size_t)N * (5 + S)];            /* work arrays on the stack, like the reference's automatic arrays (:590-600) */
  double *Fcn0 = Ynew + N, *Fcn = Fcn0 + N, *dFdT = Fcn + N, *Yerr = dFdT + N, *K = Yerr + N;
  double Jac0[(size_t)m->lu_nonzero * 2];
  double *Ghimj = Jac0 + m->lu_nonzero;
  double T, H, Hnew, HC, HG, Fac, Tau, Err;
  int Direction, j, istage, i, IERR = 0;
  int RejectLastH, RejectMoreH;
  int *IST = c->ISTATUS;
  double *RST = c->RSTATUS;

  T = Tstart;
  RST[Nhexit] = 0.0;
  H = fmin(fmax(fabs(Hmin), fabs(Hstart)), fabs(Hmax));
  if (fabs(H) <= 10.0 * Roundoff) H = DeltaMin;
  Direction = (Tend >= Tstart) ? +1 : -1;
  H = Direction * H;
  RejectLastH = 0;
  RejectMoreH = 0;

  while ((Direction > 0 && ((T - Tend) + Roundoff <= 0.0)) ||
         (Direction < 0 && ((Tend - T) + Roundoff <= 0.0))) {
    if (IST[Nstp] > Max_no_steps) { IERR = -6; goto done; }
    if (((T + 0.1 * H) == T) || (H <= Roundoff)) { IERR = -7; goto done; }
    H = fmin(H, fabs(Tend - T));
    fun_template(c, Y, Fcn0);
    IST[Nfun]++;
    if (!Autonomous) { /* ros_FunTimeDerivative (:1749-1769) */
      double Delta = sqrt(Roundoff) * fmax(1.0E-6, fabs(T));
      fun_template(c, Y, dFdT); /* rates do not depend on T when ICNTRL(15) = -1 */
      IST[Nfun]++;
      waxpy(N, -1.0, Fcn0, dFdT);
      for (i = 0; i < N; i++) dFdT[i] = (1.0 / Delta) * dFdT[i];
    }
    m->jac_sp(Y, c->FIX, c->RCONST, Jac0, c->B, 0);
    IST[Njac]++;

    for (;;) { /* UntilAccepted */
      /* ros_PrepareMatrix (:1921-1999) */
      int Nconsecutive = 0, Singular = 1;
      while (Singular) {
        double ghinv;
        int ising;
        for (i = 0; i < m->lu_nonzero; i++) Ghimj[i] = -Jac0[i];
        ghinv = 1.0 / (Direction * H * ros->Gamma[0]);
        for (i = 0; i < N; i++) Ghimj[m->lu_diag[i]] = Ghimj[m->lu_diag[i]] + ghinv;
        ising = kpp_decomp(m, Ghimj, c->W);
        IST[Ndec]++;
        if (ising == 0) {
          Singular = 0;
        } else {
          IST[Nsng]++;
          Nconsecutive++;
          Singular = 1;
          if (Nconsecutive <= 5) H = H * 0.5;
          else break;
        }
      }
      if (Singular) { IERR = -8; goto done; }

      for (istage = 1; istage <= S; istage++) {
        double *Ki = K + (size_t)N * (istage - 1);
        if (istage == 1) {
          memcpy(Fcn, Fcn0, sizeof(double) * N);
        } else if (ros->NewF[istage - 1]) {
          memcpy(Ynew, Y, sizeof(double) * N);
          for (j = 1; j <= istage - 1; j++)
            waxpy(N, ros->A[(istage - 1) * (istage - 2) / 2 + j - 1], K + (size_t)N * (j - 1), Ynew);
          Tau = T + ros->Alpha[istage - 1] * Direction * H;
          (void)Tau;
          fun_template(c, Ynew, Fcn);
          IST[Nfun]++;
        }
        memcpy(Ki, Fcn, sizeof(double) * N);
        for (j = 1; j <= istage - 1; j++) {
          HC = ros->C[(istage - 1) * (istage - 2) / 2 + j - 1] / (Direction * H);
          waxpy(N, HC, K + (size_t)N * (j - 1), Ki);
        }
        if (!Autonomous && ros->Gamma[istage - 1] != 0.0) {
          HG = Direction * H * ros->Gamma[istage - 1];
          waxpy(N, HG, dFdT, Ki);
        }
        m->solve(Ghimj, Ki, 0);
        IST[Nsol]++;
      }
      memcpy(Ynew, Y, sizeof(double) * N);
      for (j = 1; j <= S; j++) waxpy(N, ros->M[j - 1], K + (size_t)N * (j - 1), Ynew);
      for (i = 0; i < N; i++) Yerr[i] = 0.0;
      for (j = 1; j <= S; j++) waxpy(N, ros->E[j - 1], K + (size_t)N * (j - 1), Yerr);
      Err = ros_error_norm(N, Y, Ynew, Yerr, AbsTol, RelTol, VectorTol);

      Fac = fmin(FacMax, fmax(FacMin, FacSafe / pow(Err, 1.0 / ros->ELO)));
      Hnew = H * Fac;
      IST[Nstp]++;
      if ((Err <= 1.0) || (H <= Hmin)) {
        IST[Nacc]++;
        if (clip_negative) { for (i = 0; i < N; i++) Y[i] = fmax(Ynew[i], 0.0); }
        else memcpy(Y, Ynew, sizeof(double) * N);
        T = T + Direction * H;
        Hnew = fmax(Hmin, fmin(Hnew, Hmax));
        if (RejectLastH) Hnew = fmin(Hnew, H);
        RST[Nhexit] = H;
        RST[Nhnew] = Hnew;
        RST[Ntexit] = T;
        RejectLastH = 0;
        RejectMoreH = 0;
        H = Hnew;
        break;
      } else {
        if (RejectMoreH) Hnew = H * FacRej;
        RejectMoreH = RejectLastH;
        RejectLastH = 1;
        H = Hnew;
        if (IST[Nacc] >= 1) IST[Nrej]++;
      }
    }
  }
  IERR = 1;
done:
  *Tout = T;
  return IERR;
}

/* forward Euler "integrator" of the carbon mechanism (KPP/carbon/gckpp_Integrator.F90:155-215), literally:
 * Ynew = Y + dYdt*(Tend-Tstart); if ICNTRL(16) > 0 the entries are scanned in order and a negative one sets
 * IERR = -9 and is clipped (ICNTRL(16) = 1) or makes the routine RETURN with Y untouched (= 2; = 3 would STOP,
 * reported the same way here); otherwise Y = Ynew and IERR = 1 (which also overwrites the -9 of the clip case). */
static int feuler_integrator(ros_ctx_t *c, double *Y, double Tstart, double Tend, int icntrl16)
{
  const int N = c->m->nvar;
  double *dYdt = malloc(sizeof(double) * N), *Ynew = malloc(sizeof(double) * N);
  int i, ierr = 1, early = 0;
  fun_template(c, Y, dYdt);
  c->ISTATUS[Nfun]++;
  for (i = 0; i < N; i++) Ynew[i] = Y[i] + dYdt[i] * (Tend - Tstart);
  if (icntrl16 > 0) {
    for (i = 0; i < N && !early; i++)
      if (Ynew[i] < 0.0) {
        if (icntrl16 == 1) Ynew[i] = 0.0;
        else if (icntrl16 == 2 || icntrl16 == 3) early = 1;
      }
  }
  if (early) ierr = -9;
  else for (i = 0; i < N; i++) Y[i] = Ynew[i];
  free(Ynew);
  c->ISTATUS[Nstp]++;
  c->ISTATUS[Nacc]++;
  c->RSTATUS[Ntexit] = Tend;
  c->RSTATUS[Nhexit] = Tend - Tstart;
  c->RSTATUS[Nhnew] = Tend - Tstart;
  free(dYdt);
  return ierr;
}

#include "kpp_oracle_ar.c"   /* auto-reduce integrator: needs the static types and helpers above */

/* Integrate + Rosenbrock (:80-162, :165-531) */
int kpp_oracle_integrate_cell(int mech_id, double tin, double tout, double *C, const double *RCONST,
                              const double *ATOL, const double *RTOL, const int *ICNTRL_U,
                              const double *RCNTRL_U, int *ISTATUS, double *RSTATUS)
{
  const kpp_mech_t *m = kpp_oracle_mech(mech_id);
  int ICNTRL[20];
  double RCNTRL[20];
  int i, IERR = 0;
  ros_method_t ros;
  ros_ctx_t ctx;
  if (!m) return -100;
  for (i = 0; i < 20; i++) { ICNTRL[i] = 0; RCNTRL[i] = 0.0; ISTATUS[i] = 0; RSTATUS[i] = 0.0; }
  ICNTRL[14] = 5; /* default: update SUN + RCONST inside the integrator */
  if (ICNTRL_U) for (i = 0; i < 20; i++) if (ICNTRL_U[i] != 0) ICNTRL[i] = ICNTRL_U[i];
  if (RCNTRL_U) for (i = 0; i < 20; i++) if (RCNTRL_U[i] > 0) RCNTRL[i] = RCNTRL_U[i];
  /* The oracle restates the path GEOS-Chem uses: ICNTRL(15) = -1 (Q2).  Any other value would
   * call Update_SUN/Update_RCONST inside Fun/Jac, which needs the host's met state. */
  if (ICNTRL[14] != -1) return -101;

  double scratch[(size_t)m->nreact + 3 * (size_t)m->nvar + (size_t)m->nb + 8];
  ctx.m = m;
  ctx.RCONST = RCONST;
  ctx.FIX = C + m->nvar;
  ctx.A = scratch;
  ctx.P = ctx.A + m->nreact;
  ctx.D = ctx.P + m->nvar;
  ctx.W = ctx.D + m->nvar;
  ctx.B = ctx.W + m->nvar;
  ctx.ISTATUS = ISTATUS;
  ctx.RSTATUS = RSTATUS;

  if (!m->jac_sp) { /* carbon: forward Euler */
    IERR = feuler_integrator(&ctx, C, tin, tout, ICNTRL[15]);
    return IERR;
  }

  /* ---- Rosenbrock(): option decoding ---- */
  {
    double *Y = C;
    const int N = m->nvar;
    double Tstart = tin, Tend = tout, Texit = tin;
    double Roundoff, FacMin, FacMax, FacRej, FacSafe, Hmin, Hmax, Hstart, Redux_threshold;
    int UplimTol, Max_no_steps, Autonomous, VectorTol, Autoreduce, Autoreduce_Append, AR_target_spc;
    double AR_thr_ratio;
    const double DeltaMin = 1.0E-5;
    for (i = 0; i < 8; i++) ISTATUS[i] = 0;
    for (i = 0; i < 4; i++) RSTATUS[i] = 0.0;
    Autonomous = !(ICNTRL[0] == 0);
    if (ICNTRL[1] == 0) { VectorTol = 1; UplimTol = N; } else { VectorTol = 0; UplimTol = 1; }
    switch (ICNTRL[2]) {
    case 1: method_Ros2(&ros); break;
    case 2: method_Ros3(&ros); break;
    case 3: method_Ros4(&ros); break;
    case 0: case 4: method_Rodas3(&ros); break;
    case 5: method_Rodas4(&ros); break;
    case 6: method_Rang3(&ros); break;
    default: IERR = -2; goto out;
    }
    if (ICNTRL[3] == 0) Max_no_steps = 200000;
    else if (ICNTRL[3] > 0) Max_no_steps = ICNTRL[3];
    else { IERR = -1; goto out; }
    Autoreduce = (ICNTRL[11] == 1);
    Autoreduce_Append = (ICNTRL[12] == 1);
    AR_target_spc = ICNTRL[13];
    Roundoff = wlamch_eps();
    if (RCNTRL[0] == 0.0) Hmin = 0.0; else if (RCNTRL[0] > 0.0) Hmin = RCNTRL[0]; else { IERR = -3; goto out; }
    if (RCNTRL[1] == 0.0) Hmax = fabs(Tend - Tstart);
    else if (RCNTRL[1] > 0.0) Hmax = fmin(fabs(RCNTRL[1]), fabs(Tend - Tstart));
    else { IERR = -3; goto out; }
    if (RCNTRL[2] == 0.0) Hstart = fmax(Hmin, DeltaMin);
    else if (RCNTRL[2] > 0.0) Hstart = fmin(fabs(RCNTRL[2]), fabs(Tend - Tstart));
    else { IERR = -3; goto out; }
    if (RCNTRL[3] == 0.0) FacMin = 0.2; else if (RCNTRL[3] > 0.0) FacMin = RCNTRL[3]; else { IERR = -4; goto out; }
    if (RCNTRL[4] == 0.0) FacMax = 6.0; else if (RCNTRL[4] > 0.0) FacMax = RCNTRL[4]; else { IERR = -4; goto out; }
    if (RCNTRL[5] == 0.0) FacRej = 0.1; else if (RCNTRL[5] > 0.0) FacRej = RCNTRL[5]; else { IERR = -4; goto out; }
    if (RCNTRL[6] == 0.0) FacSafe = 0.9; else if (RCNTRL[6] > 0.0) FacSafe = RCNTRL[6]; else { IERR = -4; goto out; }
    for (i = 0; i < UplimTol; i++) {
      if ((ATOL[i] <= 0.0) || (RTOL[i] <= 10.0 * Roundoff) || (RTOL[i] >= 1.0)) { IERR = -5; goto out; }
    }
    Redux_threshold = 1.e2;
    if (RCNTRL[11] > 0.0) Redux_threshold = RCNTRL[11];
    else if (RCNTRL[11] < 0.0) Autoreduce = 0; /* unreachable through Integrate: Q3 */
    AR_thr_ratio = RCNTRL[13];

    if (Autoreduce) {
      IERR = kpp_oracle_ar_integrate(&ctx, &ros, Y, Tstart, Tend, &Texit, ATOL, RTOL, Autonomous, VectorTol,
                                     Max_no_steps, Roundoff, Hmin, Hmax, Hstart, FacMin, FacMax, FacRej,
                                     FacSafe, Redux_threshold, AR_target_spc, AR_thr_ratio, Autoreduce_Append);
    }
    if (!Autoreduce || IERR == -99) {
      IERR = ros_integrator(&ctx, &ros, Y, Tstart, Tend, &Texit, ATOL, RTOL, Autonomous, VectorTol,
                            Max_no_steps, Roundoff, Hmin, Hmax, Hstart, FacMin, FacMax, FacRej, FacSafe,
                            ICNTRL[15] == 1);
    }
  }
out:
  return IERR;
}

/* Batched driver: the cell loop of Do_FullChem (GeosCore/fullchem_mod.F90:528-546), gathering
 * each cell from cell-fastest arrays like :817-821 and scattering back like :1329-1350
 * (without the MAX(C,0) clip: the raw integrator output is what parity is checked on, Q5). */
int kpp_oracle_integrate(int mech_id, int ncell, double tin, double tout, const double *conc_in,
                         const double *rconst, const double *atol, const double *rtol,
                         const int *icntrl, const double *rcntrl, const double *hstart,
                         double *conc_out, int *istatus, double *rstatus, int *ierr, int nthreads)
{
  const kpp_mech_t *m = kpp_oracle_mech(mech_id);
  if (!m) return -1;
  const int nspec = m->nspec, nreact = m->nreact;
  const size_t nc = (size_t)ncell;
#ifdef _OPENMP
  if (nthreads > 0) omp_set_num_threads(nthreads);
#else
  (void)nthreads;
#endif
#pragma omp parallel
  {
    double *C = malloc(sizeof(double) * nspec);
    double *R = malloc(sizeof(double) * nreact);
    int IST[20];
    double RST[20], RC[20];
    long cell;
    int s, r;
#pragma omp for schedule(dynamic, 24)
    for (cell = 0; cell < (long)ncell; cell++) {
      for (s = 0; s < nspec; s++) C[s] = conc_in[(size_t)s * nc + cell];
      for (r = 0; r < nreact; r++) R[r] = rconst[(size_t)r * nc + cell];
      for (s = 0; s < 20; s++) RC[s] = rcntrl ? rcntrl[s] : 0.0;
      if (hstart) RC[2] = hstart[cell];
      int e = kpp_oracle_integrate_cell(mech_id, tin, tout, C, R, atol, rtol, icntrl, RC, IST, RST);
      for (s = 0; s < nspec; s++) conc_out[(size_t)s * nc + cell] = C[s];
      if (istatus) for (s = 0; s < 8; s++) istatus[(size_t)s * nc + cell] = IST[s];
      if (rstatus) for (s = 0; s < 4; s++) rstatus[(size_t)s * nc + cell] = RST[s];
      if (ierr) ierr[cell] = e;
    }
    free(C);
    free(R);
  }
  return 0;
}

static void set_met(kpp_met_t *m, double temp, double numden, double h2o)
{ /* Set_Kpp_GridBox_Values (GeosCore/fullchem_mod.F90:2139-2150); H2O is passed already as AVGW*NUMDEN */
  m->TEMP = temp;
  m->NUMDEN = numden;
  m->H2O = h2o;
  m->PRESS = 0.0;
  m->INV_TEMP = 1.0 / temp;
  m->TEMP_OVER_K300 = temp / 300.0;
  m->K300_OVER_TEMP = 300.0 / temp;
  m->SR_TEMP = sqrt(temp);
}

int kpp_oracle_update_rconst(int mech_id, int ncell, const double *temp, const double *numden,
                             const double *h2o, const double *photol, const double *khet,
                             double *rconst, int nthreads)
{
  const kpp_mech_t *m = kpp_oracle_mech(mech_id);
  if (!m) return -1;
  const size_t nc = (size_t)ncell;
#ifdef _OPENMP
  if (nthreads > 0) omp_set_num_threads(nthreads);
#else
  (void)nthreads;
#endif
#pragma omp parallel
  {
    double *R = malloc(sizeof(double) * m->nreact);
    double *PH = malloc(sizeof(double) * (m->nphot + 1));
    double *KH = malloc(sizeof(double) * (m->next + 1));
    long cell;
    int k;
#pragma omp for schedule(static)
    for (cell = 0; cell < (long)ncell; cell++) {
      kpp_met_t met;
      set_met(&met, temp[cell], numden[cell], h2o[cell]);
      for (k = 0; k < m->nphot; k++) PH[k] = photol ? photol[(size_t)k * nc + cell] : 0.0;
      for (k = 0; k < m->next; k++) KH[k] = khet ? khet[(size_t)k * nc + cell] : 0.0;
      m->update_rconst(&met, PH, KH, R);
      for (k = 0; k < m->nreact; k++) rconst[(size_t)k * nc + cell] = R[k];
    }
    free(R); free(PH); free(KH);
  }
  return 0;
}

int kpp_oracle_fun(int mech_id, const double *C, const double *RCONST, double *Vdot, double *A)
{
  const kpp_mech_t *m = kpp_oracle_mech(mech_id);
  if (!m) return -1;
  double *P = malloc(sizeof(double) * 2 * m->nvar), *D = P + m->nvar;
  if (m->fun_is_split) m->fun_split(C, C + m->nvar, RCONST, Vdot, P, D, A, 0);
  else m->fun(C, C + m->nvar, RCONST, Vdot, A);
  free(P);
  return 0;
}

int kpp_oracle_jac(int mech_id, const double *C, const double *RCONST, double *JVS)
{
  const kpp_mech_t *m = kpp_oracle_mech(mech_id);
  if (!m || !m->jac_sp) return -1;
  double *B = malloc(sizeof(double) * (m->nb + 1));
  m->jac_sp(C, C + m->nvar, RCONST, JVS, B, 0);
  free(B);
  return 0;
}

int kpp_oracle_decomp(int mech_id, double *JVS)
{
  const kpp_mech_t *m = kpp_oracle_mech(mech_id);
  if (!m || !m->jac_sp) return -1;
  double *W = malloc(sizeof(double) * m->nvar);
  int ier = kpp_decomp(m, JVS, W);
  free(W);
  return ier;
}

int kpp_oracle_solve(int mech_id, const double *JVS, double *X)
{
  const kpp_mech_t *m = kpp_oracle_mech(mech_id);
  if (!m || !m->solve) return -1;
  m->solve(JVS, X, 0);
  return 0;
}
