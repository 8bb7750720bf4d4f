// Table-driven Rosenbrock kernel ("kernel 0"): works for any KPP mechanism from its tables
// (fullchem, Hg, and every ICNTRL(3) method), one cell per lane, persistent lanes.
// It is the general path and the in-GPU arithmetic reference for the mechanism-specialised
// kernel: this translation unit is compiled with -fmad=false and evaluates every sum in the
// generated order of the reference, so its step sequences match the CPU restatement.
//
// Reference routines covered (KPP/fullchem/...):
//   ros_Integrator      gckpp_Integrator.F90:578-786      -> ros_generic_kernel
//   ros_PrepareMatrix   gckpp_Integrator.F90:1921-1999    -> g_jac_neg + g_decomp + singular loop
//   ros_ErrorNorm       gckpp_Integrator.F90:1715-1745    -> inline in the kernel
//   Fun_SPLIT / Fun     gckpp_Function.F90:2172-5060 / 51-2152  -> g_fun
//   Jac_SP              gckpp_Jacobian.F90:48-20887       -> g_jac_neg
//   KppDecomp           gckpp_LinearAlgebra.F90:46-83     -> g_decomp
//   KppSolve            gckpp_LinearAlgebra.F90:644-2309  -> g_solve
#include <float.h>
#include <math.h>
#include "ros_common.cuh"

namespace {

__device__ __forceinline__ double term_eval(int4 t, const double *yv, const double *rc,
                                            const double *__restrict__ lit, size_t st)
{
  double x = t.x >= 0 ? rc[(size_t)t.x * st] : __ldg(lit + (~t.x));
  if (t.y != -1) x = x * (t.y >= 0 ? yv[(size_t)t.y * st] : __ldg(lit + (-2 - t.y)));
  if (t.z != -1) x = x * (t.z >= 0 ? yv[(size_t)t.z * st] : __ldg(lit + (-2 - t.z)));
  if (t.w != -1) x = x * (t.w >= 0 ? yv[(size_t)t.w * st] : __ldg(lit + (-2 - t.w)));
  return x;
}

// A(r) for all reactions
__device__ void g_rates(const MechDev &M, const double *yv, const double *rc, double *A, size_t st)
{
  for (int r = 0; r < M.nreact; r++) A[(size_t)r * st] = term_eval(__ldg(M.a_term + r), yv, rc, M.lit, st);
}

// Vdot = Fun(Y): split form (P - D*V) for fullchem, aggregate for Hg/carbon (see FunTemplate).
// pd != nullptr && wr: also store Prod = P_VAR and Loss = D_VAR (FunSplitF, gckpp_Integrator.F90:2518-2547)
__device__ void g_fun(const MechDev &M, const double *yv, const double *rc, double *A, double *out, size_t st,
                      double *prod = nullptr, double *loss = nullptr, bool wr = false)
{
  g_rates(M, yv, rc, A, st);
  if (M.fun_split) {
    for (int i = 0; i < M.nvar; i++) {
      double P = 0.0, D = 0.0;
      int e = __ldg(M.p_ptr + i + 1);
      for (int k = __ldg(M.p_ptr + i); k < e; k++) P = P + __ldg(M.p_coef + k) * A[(size_t)__ldg(M.p_rxn + k) * st];
      e = __ldg(M.d_ptr + i + 1);
      for (int k = __ldg(M.d_ptr + i); k < e; k++) D = D + term_eval(__ldg(M.d_term + k), yv, rc, M.lit, st);
      out[(size_t)i * st] = P - D * yv[(size_t)i * st];
      if (prod && wr) { prod[(size_t)i * st] = P; loss[(size_t)i * st] = D; }
    }
  } else {
    for (int i = 0; i < M.nvar; i++) {
      double s = 0.0;
      int e = __ldg(M.v_ptr + i + 1);
      for (int k = __ldg(M.v_ptr + i); k < e; k++) s = s + __ldg(M.v_coef + k) * A[(size_t)__ldg(M.v_rxn + k) * st];
      out[(size_t)i * st] = s;
    }
  }
}

// G = sign*Jac_SP(Y) (sign = -1 builds Ghimj's off-diagonal part), then G(diag) += ghinv
__device__ void g_jac(const MechDev &M, const double *yv, const double *rc, double *B, double *G,
                      double sign, double ghinv, size_t st)
{
  for (int m = 0; m < M.nb; m++) B[(size_t)m * st] = term_eval(__ldg(M.b_term + m), yv, rc, M.lit, st);
  for (int k = 0; k < M.nnz; k++) {
    double s = 0.0;
    int e = __ldg(M.j_ptr + k + 1);
    for (int q = __ldg(M.j_ptr + k); q < e; q++) s = s + __ldg(M.j_coef + q) * B[(size_t)__ldg(M.j_b + q) * st];
    G[(size_t)k * st] = sign * s;
  }
  if (ghinv != 0.0)
    for (int i = 0; i < M.nvar; i++) {
      size_t d = (size_t)__ldg(M.diag + i) * st;
      G[d] = G[d] + ghinv;
    }
}

// KppDecomp: row-wise sparse LU without pivoting, in place. Returns 0 or the 1-based singular row.
__device__ int g_decomp(const MechDev &M, double *G, double *W, size_t st)
{
  int ier = 0;
  for (int k = 0; k < M.nvar; k++) {
    int c0 = __ldg(M.crow + k), c1 = __ldg(M.crow + k + 1), dk = __ldg(M.diag + k);
    if (ier == 0 && fabs(G[(size_t)dk * st]) < DBL_MIN) ier = k + 1;
    for (int kk = c0; kk < c1; kk++) W[(size_t)__ldg(M.icol + kk) * st] = G[(size_t)kk * st];
    for (int kk = c0; kk < dk; kk++) {
      int j = __ldg(M.icol + kk);
      int dj = __ldg(M.diag + j), ej = __ldg(M.crow + j + 1);
      double a = -W[(size_t)j * st] / G[(size_t)dj * st];
      W[(size_t)j * st] = -a;
      for (int jj = dj + 1; jj < ej; jj++) {
        size_t c = (size_t)__ldg(M.icol + jj) * st;
        W[c] = W[c] + a * G[(size_t)jj * st];
      }
    }
    for (int kk = c0; kk < c1; kk++) G[(size_t)kk * st] = W[(size_t)__ldg(M.icol + kk) * st];
  }
  return ier;
}

// KppSolve: unit-lower forward sweep then upper back sweep with diagonal divide, in place on X.
__device__ void g_solve(const MechDev &M, const double *G, double *X, size_t st)
{
  for (int i = 0; i < M.nvar; i++) {
    int c0 = __ldg(M.crow + i), d = __ldg(M.diag + i);
    if (d > c0) {
      double x = X[(size_t)i * st];
      for (int k = c0; k < d; k++) x = x - G[(size_t)k * st] * X[(size_t)__ldg(M.icol + k) * st];
      X[(size_t)i * st] = x;
    }
  }
  for (int i = M.nvar - 1; i >= 0; i--) {
    int d = __ldg(M.diag + i), c1 = __ldg(M.crow + i + 1);
    double x = X[(size_t)i * st];
    for (int k = d + 1; k < c1; k++) x = x - G[(size_t)k * st] * X[(size_t)__ldg(M.icol + k) * st];
    X[(size_t)i * st] = x / G[(size_t)d * st];
  }
}

__global__ void __launch_bounds__(128) ros_generic_kernel(MechDev M, RosArgs a)
{
  const int lane = threadIdx.x & 31;
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  double *ws = a.work + (size_t)warp * a.ws_stride + lane;
  const size_t st = 32;
  const RosOpts &o = a.o;
  const WsLayout &L = a.L;
  const int N = M.nvar;
  double *Y = ws + (size_t)L.Y * st, *YN = ws + (size_t)L.YN * st, *F0 = ws + (size_t)L.F0 * st;
  double *FC = ws + (size_t)L.FC * st, *K = ws + (size_t)L.K * st, *G = ws + (size_t)L.G * st;
  double *RC = ws + (size_t)L.RC * st, *AB = ws + (size_t)L.AB * st, *W = ws + (size_t)L.W * st;
  double *PR = ws + (size_t)L.PR * st, *LS = ws + (size_t)L.LS * st, *MK = ws + (size_t)L.MK * st;   // auto-reduce only
  const double Dir = (double)o.Direction;

  // per-lane integration state
  bool have = false, exhausted = false, newstep = false;
  bool RejectLastH = false, RejectMoreH = false, reduced = false;
  double arthr = 0.0;
  int cell = -1, nconsec = 0, ierr_cell = 0;
  int ist[8];
  double T = 0.0, H = 0.0, Hexit = 0.0, Hnew_out = 0.0, Texit = 0.0;
  unsigned long long acc_stp = 0, acc_acc = 0, acc_fail = 0, acc_done = 0;

  for (;;) {
    // ---- retire finished cells and refill idle lanes -------------------------------------
    for (;;) {
      int w = fetch_work(a.next, !have && !exhausted, lane);
      if (!have && !exhausted) {
        if (w >= a.nwork) {
          exhausted = true;
        } else {
          cell = a.cell_list ? a.cell_list[w] : w;
          for (int s = 0; s < M.nspec; s++) {
            double v = a.conc_in[(size_t)s * a.ncell + cell];
            Y[(size_t)s * st] = v;
            if (s >= N) YN[(size_t)s * st] = v;      // fixed species ride along in both state vectors
          }
          for (int r = 0; r < M.nreact; r++) RC[(size_t)r * st] = a.rconst[(size_t)r * a.rc_stride + (cell - a.rc_cell0)];
#pragma unroll
          for (int q = 0; q < 8; q++) ist[q] = 0;
          // Integrate's merge (RCNTRL_U > 0 overrides) and Rosenbrock's Hstart rule (:420-428)
          double hs = a.hstart ? a.hstart[cell] : o.Hstart_rcntrl;
          double Hstart = (hs > 0.0) ? fmin(fabs(hs), fabs(o.Tend - o.Tstart)) : fmax(o.Hmin, 1.0E-5);
          T = o.Tstart;
          Hexit = 0.0; Hnew_out = 0.0; Texit = 0.0;
          H = fmin(fmax(fabs(o.Hmin), fabs(Hstart)), fabs(o.Hmax));      // :637
          if (fabs(H) <= 10.0 * o.Roundoff) H = 1.0E-5;
          H = Dir * H;
          RejectLastH = false; RejectMoreH = false; reduced = false; arthr = 0.0;
          have = true; newstep = true; nconsec = 0; ierr_cell = 0;
        }
      }
      if (have && newstep) {
        // TimeLoop condition and the two guards at its top (:652-665)
        bool inloop = (o.Direction > 0) ? ((T - o.Tend) + o.Roundoff <= 0.0) : ((o.Tend - T) + o.Roundoff <= 0.0);
        if (!inloop) ierr_cell = 1;
        else if (ist[Nstp] > o.Max_no_steps) ierr_cell = -6;
        else if (((T + 0.1 * H) == T) || (H <= o.Roundoff)) ierr_cell = -7;
        else H = fmin(H, fabs(o.Tend - T));
      }
      if (have && ierr_cell != 0) {
        // ---- finalize this cell
        if (a.ar_on && ierr_cell == 1 && reduced) {
          // 1st order solution for the removed species (AutoReduce_1stOrder, :1702-1712), initial Prod/Loss
          for (int i = 0; i < N; i++) {
            if (MK[(size_t)i * st] == 0.0) {
              const double P = PR[(size_t)i * st], k = LS[(size_t)i * st], y = Y[(size_t)i * st];
              if (k > 1.e-30 && y > 1.e-30) {
                const double term = P / k;
                Y[(size_t)i * st] = term + (y - term) * exp(-k * (o.Tend - o.Tstart));
              }
            }
          }
        }
        for (int s = 0; s < M.nspec; s++) a.conc_out[(size_t)s * a.ncell + cell] = Y[(size_t)s * st];
        if (a.istatus)
#pragma unroll
          for (int q = 0; q < 8; q++) a.istatus[(size_t)q * a.ncell + cell] = ist[q];
        if (a.rstatus) {
          a.rstatus[cell] = Texit;
          a.rstatus[(size_t)a.ncell + cell] = Hexit;
          a.rstatus[(size_t)2 * a.ncell + cell] = Hnew_out;
          a.rstatus[(size_t)3 * a.ncell + cell] = arthr;
        }
        if (a.ierr) a.ierr[cell] = ierr_cell;
        acc_stp += ist[Nstp]; acc_acc += ist[Nacc]; acc_done++;
        if (ierr_cell < 0) acc_fail++;
        have = false; ierr_cell = 0;
      }
      if (!__any_sync(FULLMASK, !have && !exhausted)) break;
    }
    if (!__any_sync(FULLMASK, have)) break;

    // ---- one Rosenbrock attempt for every lane of the warp ---------------------------------
    // Fcn0 = Fun(Y) at the start of a step (:668); lanes repeating a rejected step recompute
    // the identical values, so no predicate is needed on the data.
    if (__any_sync(FULLMASK, have && newstep))
      g_fun(M, Y, RC, AB, F0, st, a.ar_on ? PR : nullptr, LS, have && newstep && !reduced && T == o.Tstart);
    if (a.ar_on && have && newstep && !reduced) {
      // species whose production and loss are both below the threshold leave the implicit system (:918-962)
      double thr = a.ar_threshold;
      if (a.ar_target > 0) {
        const size_t t = (size_t)(a.ar_target - 1) * st;
        thr = a.ar_ratio * fmax(LS[t] * Y[t], PR[t]);
        arthr = thr;
      }
      for (int i = 0; i < N; i++) {
        const bool keep = a.ar_keep_active && a.ar_keep_spc && a.ar_keep_spc[i];
        const bool rmv = !keep && fabs(LS[(size_t)i * st] * Y[(size_t)i * st]) < thr && fabs(PR[(size_t)i * st]) < thr;
        MK[(size_t)i * st] = rmv ? 0.0 : 1.0;
      }
      reduced = true;
    }
    if (have && newstep) {
      ist[Nfun]++;
      if (!o.Autonomous) ist[Nfun]++;   // ros_FunTimeDerivative: with ICNTRL(15)=-1 Fun does not depend on T, dFdT == 0
      ist[Njac]++;
      nconsec = 0;
    }
    // Ghimj = 1/(H*gamma) - Jac0 (:1973-1977). Jac0 is recomputed from Y instead of being kept:
    // same values, and 45 KB less state per cell.
    double ghinv = 1.0 / (Dir * H * o.Gamma[0]);
    g_jac(M, Y, RC, AB, G, -1.0, ghinv, st);
    if (a.ar_on && have && reduced) {
      // the compressed system of ros_cPrepareMatrix, expressed on the full pattern: rows and columns of the
      // removed species become identity rows/columns; kept rows then see exactly the reference's operations
      for (int i = 0; i < N; i++) {
        const bool ki = MK[(size_t)i * st] != 0.0;
        const int c1 = __ldg(M.crow + i + 1);
        for (int k = __ldg(M.crow + i); k < c1; k++)
          if (!ki || MK[(size_t)__ldg(M.icol + k) * st] == 0.0) G[(size_t)k * st] = 0.0;
        if (!ki) G[(size_t)__ldg(M.diag + i) * st] = 1.0;
      }
    }
    int ising = g_decomp(M, G, W, st);
    bool skip = false;
    if (have) {
      ist[Ndec]++;
      if (ising != 0) {              // :1985-1995
        ist[Nsng]++;
        nconsec++;
        if (nconsec <= 5) { H = H * 0.5; skip = true; newstep = false; }
        else { ierr_cell = -8; skip = true; }
      } else {
        nconsec = 0;
      }
    }
    if (!__any_sync(FULLMASK, have && !skip)) continue;

    const double *src = F0;
    for (int is = 1; is <= o.S; is++) {
      double *Ki = K + (size_t)N * (is - 1) * st;
      if (is > 1 && o.NewF[is - 1]) {
        for (int i = 0; i < N; i++) {
          double v = Y[(size_t)i * st];
          for (int j = 1; j < is; j++) {
            double aj = o.A[(is - 1) * (is - 2) / 2 + j - 1];
            if (aj != 0.0) v = v + aj * K[((size_t)N * (j - 1) + i) * st];
          }
          YN[(size_t)i * st] = v;
        }
        g_fun(M, YN, RC, AB, FC, st);
        if (have && !skip) ist[Nfun]++;
        src = FC;
      }
      for (int i = 0; i < N; i++) {
        double v = src[(size_t)i * st];
        for (int j = 1; j < is; j++) {
          double HC = o.C[(is - 1) * (is - 2) / 2 + j - 1] / (Dir * H);
          if (HC != 0.0) v = v + HC * K[((size_t)N * (j - 1) + i) * st];
        }
        if (a.ar_on && have && reduced && MK[(size_t)i * st] == 0.0) v = 0.0;     // K of a removed species stays 0
        Ki[(size_t)i * st] = v;
      }
      g_solve(M, G, Ki, st);
      if (have && !skip) ist[Nsol]++;
    }
    // new solution, error estimate and its scaled norm (:729-740, :1715-1745)
    double Err = 0.0;
    for (int i = 0; i < N; i++) {
      double y = Y[(size_t)i * st];
      double yn = y, ye = 0.0;
      for (int j = 1; j <= o.S; j++) {
        double kj = K[((size_t)N * (j - 1) + i) * st];
        if (o.M[j - 1] != 0.0) yn = yn + o.M[j - 1] * kj;
        if (o.E[j - 1] != 0.0) ye = ye + o.E[j - 1] * kj;
      }
      YN[(size_t)i * st] = yn;
      double Ymax = fmax(fabs(y), fabs(yn));
      double Scale = o.VectorTol ? (__ldg(a.atol + i) + __ldg(a.rtol + i) * Ymax) : (__ldg(a.atol) + __ldg(a.rtol) * Ymax);
      double q = ye / Scale;
      Err = Err + q * q;
    }
    Err = fmax(sqrt(Err / (double)N), 1.0e-10);

    if (have && !skip) {
      double Fac = fmin(o.FacMax, fmax(o.FacMin, o.FacSafe / pow(Err, 1.0 / o.ELO)));   // :743
      double Hnew = H * Fac;
      ist[Nstp]++;
      if ((Err <= 1.0) || (H <= o.Hmin)) {       // accept (:748-768)
        ist[Nacc]++;
        for (int i = 0; i < N; i++) {
          double v = YN[(size_t)i * st];
          Y[(size_t)i * st] = o.ClipNegative ? fmax(v, 0.0) : v;
        }
        T = T + Dir * H;
        Hnew = fmax(o.Hmin, fmin(Hnew, o.Hmax));
        if (RejectLastH) Hnew = fmin(Hnew, H);
        Hexit = H; Hnew_out = Hnew; Texit = T;
        RejectLastH = false; RejectMoreH = false;
        H = Hnew;
        newstep = true;
      } else {                                   // reject (:769-777)
        if (RejectMoreH) Hnew = H * o.FacRej;
        RejectMoreH = RejectLastH;
        RejectLastH = true;
        H = Hnew;
        if (ist[Nacc] >= 1) ist[Nrej]++;
        newstep = false;
      }
    }
  }
  // ---- per-launch totals (diagnostics only)
  for (int off = 16; off > 0; off >>= 1) {
    acc_stp += __shfl_down_sync(FULLMASK, acc_stp, off);
    acc_acc += __shfl_down_sync(FULLMASK, acc_acc, off);
    acc_fail += __shfl_down_sync(FULLMASK, acc_fail, off);
    acc_done += __shfl_down_sync(FULLMASK, acc_done, off);
  }
  if (lane == 0 && a.sums) {
    atomicAdd(a.sums + 0, acc_stp);
    atomicAdd(a.sums + 1, acc_acc);
    atomicAdd(a.sums + 2, acc_fail);
    atomicAdd(a.sums + 3, acc_done);
  }
}

// ---- single-routine kernels over user arrays (stride = ncell), used by the diagnostics entry
// points (Fun(...,Aout)) and the parity tests of the pieces.
__global__ void fun_cells_kernel(MechDev M, int ncell, const double *conc, const double *rconst,
                                 double *vdot, double *aout)
{
  int cell = blockIdx.x * blockDim.x + threadIdx.x;
  if (cell >= ncell) return;
  g_fun(M, conc + cell, rconst + cell, aout + cell, vdot + cell, (size_t)ncell);
}
__global__ void jac_cells_kernel(MechDev M, int ncell, const double *conc, const double *rconst,
                                 double *bwork, double *jvs)
{
  int cell = blockIdx.x * blockDim.x + threadIdx.x;
  if (cell >= ncell) return;
  g_jac(M, conc + cell, rconst + cell, bwork + cell, jvs + cell, 1.0, 0.0, (size_t)ncell);
}
__global__ void decomp_cells_kernel(MechDev M, int ncell, double *jvs, double *wwork, int *ier)
{
  int cell = blockIdx.x * blockDim.x + threadIdx.x;
  if (cell >= ncell) return;
  ier[cell] = g_decomp(M, jvs + cell, wwork + cell, (size_t)ncell);
}
__global__ void solve_cells_kernel(MechDev M, int ncell, const double *jvs, double *x)
{
  int cell = blockIdx.x * blockDim.x + threadIdx.x;
  if (cell >= ncell) return;
  g_solve(M, jvs + cell, x + cell, (size_t)ncell);
}

// Forward Euler "integrator" of the carbon mechanism (KPP/carbon/gckpp_Integrator.F90:155-215):
// Ynew = Y + dYdt*(Tend-Tstart); ICNTRL(16): 0 keep negatives, 1 clip to zero, 2 flag IERR=-1.
#define FE_MAX 64
__global__ void feuler_kernel(MechDev M, RosArgs a, int icntrl16)
{
  int w = blockIdx.x * blockDim.x + threadIdx.x;
  if (w >= a.nwork) return;
  int cell = a.cell_list ? a.cell_list[w] : w;
  double y[FE_MAX], rc[FE_MAX], A[FE_MAX], vd[FE_MAX];
  for (int s = 0; s < M.nspec; s++) y[s] = a.conc_in[(size_t)s * a.ncell + cell];
  for (int r = 0; r < M.nreact; r++) rc[r] = a.rconst[(size_t)r * a.rc_stride + (cell - a.rc_cell0)];
  g_fun(M, y, rc, A, vd, 1);
  // KPP/carbon/gckpp_Integrator.F90:187-211: a negative entry is clipped (ICNTRL(16) = 1) or makes the routine
  // return IERR = -9 with Y untouched (= 2; = 3 STOPs in the reference, reported the same way here)
  int ierr = 1;
  bool early = false;
  double dt = a.o.Tend - a.o.Tstart;
  for (int i = 0; i < M.nvar; i++) vd[i] = y[i] + vd[i] * dt;
  if (icntrl16 > 0)
    for (int i = 0; i < M.nvar && !early; i++)
      if (vd[i] < 0.0) {
        if (icntrl16 == 1) vd[i] = 0.0;
        else if (icntrl16 == 2 || icntrl16 == 3) early = true;
      }
  if (early) ierr = -9;
  else for (int i = 0; i < M.nvar; i++) y[i] = vd[i];
  for (int s = 0; s < M.nspec; s++) a.conc_out[(size_t)s * a.ncell + cell] = y[s];
  if (a.istatus) {
    for (int q = 0; q < 8; q++) a.istatus[(size_t)q * a.ncell + cell] = 0;
    a.istatus[(size_t)Nfun * a.ncell + cell] = 1;
    a.istatus[(size_t)Nstp * a.ncell + cell] = 1;
    a.istatus[(size_t)Nacc * a.ncell + cell] = 1;
  }
  if (a.rstatus) {
    a.rstatus[cell] = a.o.Tend;
    a.rstatus[(size_t)a.ncell + cell] = dt;
    a.rstatus[(size_t)2 * a.ncell + cell] = dt;
    a.rstatus[(size_t)3 * a.ncell + cell] = 0.0;
  }
  if (a.ierr) a.ierr[cell] = ierr;
  if (a.sums) {
    atomicAdd(a.sums + 0, 1ull); atomicAdd(a.sums + 1, 1ull); atomicAdd(a.sums + 3, 1ull);
    if (ierr < 0) atomicAdd(a.sums + 2, 1ull);
  }
}

}  // namespace

cudaError_t launch_feuler(const MechDev &M, const RosArgs &a, int icntrl16, cudaStream_t s)
{
  if (M.nspec > FE_MAX || M.nreact > FE_MAX) return cudaErrorInvalidValue;
  feuler_kernel<<<(a.nwork + 127) / 128, 128, 0, s>>>(M, a, icntrl16);
  return cudaGetLastError();
}
// Auto-reduce decision of ros_yIntegrator (gckpp_Integrator.F90:904-962) as a pass of its own, for the kernels that
// integrate on the full pattern with a per-cell keep mask: one cell per thread, Prod = P_VAR and LossY = D_VAR * V from
// FunSplitF at (Tstart, the initial concentrations), in the reference's summation order (this unit has no FMA
// contraction); a rate A(r) is re-evaluated where the reference reads it from the array -- same operations, same bits.
// mask[i][cell] = 0 when species i leaves the implicit system.  rstatus(NARthr) receives the threshold used.
__device__ __forceinline__ double term_eval2(int4 t, const double *yv, size_t sy, const double *rc, size_t sr, const double *__restrict__ lit)
{
  double x = t.x >= 0 ? rc[(size_t)t.x * sr] : __ldg(lit + (~t.x));
  if (t.y != -1) x = x * (t.y >= 0 ? yv[(size_t)t.y * sy] : __ldg(lit + (-2 - t.y)));
  if (t.z != -1) x = x * (t.z >= 0 ? yv[(size_t)t.z * sy] : __ldg(lit + (-2 - t.z)));
  if (t.w != -1) x = x * (t.w >= 0 ? yv[(size_t)t.w * sy] : __ldg(lit + (-2 - t.w)));
  return x;
}
__device__ __forceinline__ void prod_loss(const MechDev &M, int i, const double *yv, size_t sy, const double *rc, size_t sr, double &P, double &D)
{
  P = 0.0; D = 0.0;
  int e = __ldg(M.p_ptr + i + 1);
  for (int k = __ldg(M.p_ptr + i); k < e; k++)
    P = P + __ldg(M.p_coef + k) * term_eval2(__ldg(M.a_term + __ldg(M.p_rxn + k)), yv, sy, rc, sr, M.lit);
  e = __ldg(M.d_ptr + i + 1);
  for (int k = __ldg(M.d_ptr + i); k < e; k++) D = D + term_eval2(__ldg(M.d_term + k), yv, sy, rc, sr, M.lit);
}
__global__ void __launch_bounds__(128) ar_mask_kernel(MechDev M, RosArgs a, unsigned char *__restrict__ mask)
{
  const int w = blockIdx.x * blockDim.x + threadIdx.x;
  if (w >= a.nwork) return;
  const int cell = a.cell_list ? a.cell_list[w] : w;
  const double *yv = a.conc_in + cell, *rc = a.rconst + (cell - a.rc_cell0);
  const size_t sy = (size_t)a.ncell, sr = (size_t)a.rc_stride;
  double thr = a.ar_threshold, arthr = 0.0, P, D;
  if (a.ar_target > 0) {
    const int t = a.ar_target - 1;
    prod_loss(M, t, yv, sy, rc, sr, P, D);
    thr = a.ar_ratio * fmax(D * yv[(size_t)t * sy], P);
    arthr = thr;
  }
  for (int i = 0; i < M.nvar; i++) {
    prod_loss(M, i, yv, sy, rc, sr, P, D);
    const bool keep = a.ar_keep_active && a.ar_keep_spc && a.ar_keep_spc[i];
    const bool rmv = !keep && fabs(D * yv[(size_t)i * sy]) < thr && fabs(P) < thr;
    mask[(size_t)i * sy + cell] = rmv ? 0 : 1;
  }
  if (a.rstatus) a.rstatus[(size_t)3 * a.ncell + cell] = arthr;
}
// The closing step of the auto-reduce solver for those kernels: the removed species of every cell that finished with
// IERR = 1 get their first-order solution (AutoReduce_1stOrder, gckpp_Integrator.F90:1702-1712, called at :1232-1236)
// from Prod / Loss at the initial state -- re-evaluated here from conc_in exactly as ar_mask_kernel evaluated them
// (conc_in must not alias conc_out).
__global__ void __launch_bounds__(128) ar_first_order_kernel(MechDev M, RosArgs a, const unsigned char *__restrict__ mask)
{
  const int w = blockIdx.x * blockDim.x + threadIdx.x;
  if (w >= a.nwork) return;
  const int cell = a.cell_list ? a.cell_list[w] : w;
  if (a.ierr[cell] != 1) return;
  const double *yv = a.conc_in + cell, *rc = a.rconst + (cell - a.rc_cell0);
  const size_t sy = (size_t)a.ncell, sr = (size_t)a.rc_stride;
  for (int i = 0; i < M.nvar; i++) {
    if (mask[(size_t)i * sy + cell]) continue;
    double P, k;
    prod_loss(M, i, yv, sy, rc, sr, P, k);
    const double y = a.conc_out[(size_t)i * sy + cell];
    if (k > 1.e-30 && y > 1.e-30) {
      const double term = P / k;
      a.conc_out[(size_t)i * sy + cell] = term + (y - term) * exp(-k * (a.o.Tend - a.o.Tstart));
    }
  }
}
cudaError_t launch_ar_first_order(const MechDev &M, const RosArgs &a, const unsigned char *mask, cudaStream_t s)
{
  if (!M.fun_split || !a.ierr || a.conc_in == a.conc_out) return cudaErrorInvalidValue;
  ar_first_order_kernel<<<(a.nwork + 127) / 128, 128, 0, s>>>(M, a, mask);
  return cudaGetLastError();
}
cudaError_t launch_ar_mask(const MechDev &M, const RosArgs &a, unsigned char *mask, cudaStream_t s)
{
  if (!M.fun_split) return cudaErrorInvalidValue;
  ar_mask_kernel<<<(a.nwork + 127) / 128, 128, 0, s>>>(M, a, mask);
  return cudaGetLastError();
}
cudaError_t launch_ros_generic(const MechDev &M, const RosArgs &a, int blocks, int threads, cudaStream_t s)
{
  ros_generic_kernel<<<blocks, threads, 0, s>>>(M, a);
  return cudaGetLastError();
}
cudaError_t launch_fun_cells(const MechDev &M, int ncell, const double *conc, const double *rconst,
                             double *vdot, double *aout, cudaStream_t s)
{
  fun_cells_kernel<<<(ncell + 127) / 128, 128, 0, s>>>(M, ncell, conc, rconst, vdot, aout);
  return cudaGetLastError();
}
cudaError_t launch_jac_cells(const MechDev &M, int ncell, const double *conc, const double *rconst,
                             double *bwork, double *jvs, cudaStream_t s)
{
  jac_cells_kernel<<<(ncell + 127) / 128, 128, 0, s>>>(M, ncell, conc, rconst, bwork, jvs);
  return cudaGetLastError();
}
cudaError_t launch_decomp_cells(const MechDev &M, int ncell, double *jvs, double *wwork, int *ier, cudaStream_t s)
{
  decomp_cells_kernel<<<(ncell + 127) / 128, 128, 0, s>>>(M, ncell, jvs, wwork, ier);
  return cudaGetLastError();
}
cudaError_t launch_solve_cells(const MechDev &M, int ncell, const double *jvs, double *x, cudaStream_t s)
{
  solve_cells_kernel<<<(ncell + 127) / 128, 128, 0, s>>>(M, ncell, jvs, x);
  return cudaGetLastError();
}
