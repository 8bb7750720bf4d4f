// Unrolled kernel ("kernel 4") for SMALL mechanisms: one grid cell per THREAD, the whole state of the cell in
// registers / thread-local memory, Fun / Jac_SP / KppDecomp / KppSolve as generated straight-line code with
// compile-time indices (kppgen/emit_unrolled.py -> gen/<mech>_unrolled.cuh).  No tables, no shared memory, no
// communication; lanes are persistent (a lane whose cell is finished stores it and takes the next one).
// Hg: 32 variable species, 161 matrix entries -- about 4.9 KB of state per cell, 8 warps per SM.
//
// Reference routines covered (KPP/Hg/...): ros_Integrator gckpp_Integrator.F90:578-786 (stage loop, every ICNTRL(3)
// method), ros_PrepareMatrix :1921-1999, ros_ErrorNorm :1715-1745, Fun, Jac_SP, KppDecomp, KppSolve (generated).
#include <float.h>
#include <math.h>
#include "ros_common.cuh"
#include "kernels.h"
// Where the big per-cell arrays -- matrix G, rate constants RC, stage vectors K -- live is a build-time choice
// (UNR_SMEM); every index is a compile-time constant either way:
//   0 (default)  thread-local: what does not fit the registers is local memory (L2-resident); the kernel is bound by
//                that latency, so occupancy wins over registers: UNR_MINB blocks of 128 threads per SM = 2 / 3 / 4 / 6 / 8
//                (255 / 168 / 128 / 80 / 64 registers) -> 9.6 / 10.1 / 11.8 / 12.9 / 13.4 M cells/s for Hg
//   1            shared memory, [index][thread]: conflict-free LDS/STS with immediate offsets, but only 64 cells per
//                SM fit (Hg: (161 + 94 + 6*32) x 64 x 8 B), two warps cannot hide their own latencies: 8.9 M cells/s
// (profiles/r02y_hg_bench*.log)
#ifndef UNR_SMEM
#define UNR_SMEM 0
#endif
#ifndef UNR_MINB
#define UNR_MINB 8
#endif
#if UNR_SMEM
#define UNR_BLOCK 64
#define UNR_STRIDE UNR_BLOCK
#define UNR_BOUNDS __launch_bounds__(UNR_BLOCK, 1)
#else
#define UNR_BLOCK 128
#define UNR_STRIDE 1
#define UNR_BOUNDS __launch_bounds__(UNR_BLOCK, UNR_MINB)
#endif
#define U_CTX_ARG double *__restrict__ sm_
#define U_G(k) sm_[(k) * UNR_STRIDE]
#define U_RC(r) sm_[(NNZ + (r)) * UNR_STRIDE]
#define U_K(j, i) sm_[(NNZ + NREACT + (j) * NVAR + (i)) * UNR_STRIDE]      // NNZ, NREACT, NVAR: the mechanism's, in scope at the use
#include "gen/Hg_unrolled.cuh"

namespace {

template <class U>
__global__ void UNR_BOUNDS ros_unrolled_kernel(RosArgs a)
{
  constexpr int N = U::NVAR, NS = U::NSPEC, NR = U::NREACT, NNZ = U::NNZ, NREACT = U::NREACT, NVAR = U::NVAR;
#if UNR_SMEM
  extern __shared__ __align__(16) double smem_[];
  double *sm_ = smem_ + threadIdx.x;
#else
  double loc_[NNZ + NREACT + 6 * NVAR];
  double *sm_ = loc_;
#endif
  const int lane = threadIdx.x & 31;
  const RosOpts &o = a.o;
  const double Dir = (double)o.Direction;
  double Y[NS], YS[NS], RINV[N], F0[N], FC[N], X[N];

  bool have = false, exhausted = false, newstep = false;
  bool RejectLastH = false, RejectMoreH = false;
  int cell = -1, nconsec = 0, ierr_cell = 0;
  int ist[8];
  double T = 0.0, H = 0.0, Hexit = 0.0, Hnew_out = 0.0, Texit = 0.0;
  unsigned long long acc_stp = 0, acc_acc = 0, acc_fail = 0, acc_done = 0;
#pragma unroll
  for (int q = 0; q < 8; q++) ist[q] = 0;
#pragma unroll
  for (int s = 0; s < NS; s++) { Y[s] = 0.0; YS[s] = 0.0; }
#pragma unroll
  for (int r = 0; r < NR; r++) U_RC(r) = 0.0;

  for (;;) {
    // ---- retire finished cells and refill idle lanes (same protocol as ros_generic_kernel)
    for (;;) {
      const int w = fetch_work(a.next, !have && !exhausted, lane);
      if (!have && !exhausted) {
        if (w >= a.nwork) {
          exhausted = true;
        } else {
          cell = a.cell_list ? a.cell_list[w] : w;
#pragma unroll
          for (int s = 0; s < NS; s++) { Y[s] = a.conc_in[(size_t)s * a.ncell + cell]; YS[s] = Y[s]; }
#pragma unroll
          for (int r = 0; r < NR; r++) U_RC(r) = a.rconst[(size_t)r * a.rc_stride + (cell - a.rc_cell0)];
#pragma unroll
          for (int q = 0; q < 8; q++) ist[q] = 0;
          const double hs = a.hstart ? a.hstart[cell] : o.Hstart_rcntrl;
          const double Hstart = (hs > 0.0) ? fmin(fabs(hs), fabs(o.Tend - o.Tstart)) : fmax(o.Hmin, 1.0E-5);
          T = o.Tstart;
          Hexit = 0.0; Hnew_out = 0.0; Texit = 0.0;
          H = fmin(fmax(fabs(o.Hmin), fabs(Hstart)), fabs(o.Hmax));      // :637
          if (fabs(H) <= 10.0 * o.Roundoff) H = 1.0E-5;
          H = Dir * H;
          RejectLastH = false; RejectMoreH = false;
          have = true; newstep = true; nconsec = 0; ierr_cell = 0;
        }
      }
      if (have && newstep) {
        const bool inloop = (o.Direction > 0) ? ((T - o.Tend) + o.Roundoff <= 0.0) : ((o.Tend - T) + o.Roundoff <= 0.0);
        if (!inloop) ierr_cell = 1;
        else if (ist[Nstp] > o.Max_no_steps) ierr_cell = -6;
        else if (((T + 0.1 * H) == T) || (H <= o.Roundoff)) ierr_cell = -7;
        else H = fmin(H, fabs(o.Tend - T));
      }
      if (have && ierr_cell != 0) {
#pragma unroll
        for (int s = 0; s < NS; s++) a.conc_out[(size_t)s * a.ncell + cell] = Y[s];
        if (a.istatus)
#pragma unroll
          for (int q = 0; q < 8; q++) a.istatus[(size_t)q * a.ncell + cell] = ist[q];
        if (a.rstatus) {
          a.rstatus[cell] = Texit;
          a.rstatus[(size_t)a.ncell + cell] = Hexit;
          a.rstatus[(size_t)2 * a.ncell + cell] = Hnew_out;
          a.rstatus[(size_t)3 * a.ncell + cell] = 0.0;
        }
        if (a.ierr) a.ierr[cell] = ierr_cell;
        acc_stp += ist[Nstp]; acc_acc += ist[Nacc]; acc_done++;
        if (ierr_cell < 0) acc_fail++;
        have = false; ierr_cell = 0;
      }
      if (!__any_sync(FULLMASK, !have && !exhausted)) break;
    }
    if (!__any_sync(FULLMASK, have)) break;

    // ---- one Rosenbrock attempt for every lane of the warp; idle lanes compute on their last (finite) data
    U::fun(Y, sm_, F0);
    if (have && newstep) {
      ist[Nfun]++;
      if (!o.Autonomous) ist[Nfun]++;
      ist[Njac]++;
      nconsec = 0;
    }
    U::jac(Y, sm_, 1.0 / (Dir * H * o.Gamma[0]));
    const int ising = U::lu(sm_, RINV);
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

    const double rH = 1.0 / (Dir * H);
    bool useFC = false;
#pragma unroll 1
    for (int is = 1; is <= o.S; is++) {
      const int base = (is - 1) * (is - 2) / 2;
      if (is > 1 && o.NewF[is - 1]) {
#pragma unroll
        for (int i = 0; i < N; i++) {
          double v = Y[i];
          for (int j = 1; j < is; j++) v = fma(o.A[base + j - 1], U_K(j - 1, i), v);
          YS[i] = v;
        }
        U::fun(YS, sm_, FC);
        if (have && !skip) ist[Nfun]++;
        useFC = true;
      }
#pragma unroll
      for (int i = 0; i < N; i++) {
        double v = useFC ? FC[i] : F0[i];
        for (int j = 1; j < is; j++) v = fma(o.C[base + j - 1] * rH, U_K(j - 1, i), v);
        X[i] = v;
      }
      U::solve(sm_, RINV, X);
#pragma unroll
      for (int i = 0; i < N; i++) U_K(is - 1, i) = X[i];
      if (have && !skip) ist[Nsol]++;
    }
    // new solution, error estimate and its scaled norm (:729-740, :1715-1745)
    double Err = 0.0;
#pragma unroll
    for (int i = 0; i < N; i++) {
      double yn = Y[i], ye = 0.0;
      for (int j = 0; j < o.S; j++) {
        yn = fma(o.M[j], U_K(j, i), yn);
        ye = fma(o.E[j], U_K(j, i), ye);
      }
      X[i] = yn;
      const double Ymax = fmax(fabs(Y[i]), fabs(yn));
      const double Scale = o.VectorTol ? fma(__ldg(a.rtol + i), Ymax, __ldg(a.atol + i)) : fma(__ldg(a.rtol), Ymax, __ldg(a.atol));
      const double q = ye / Scale;
      Err = fma(q, q, Err);
    }
    Err = fmax(sqrt(Err / (double)N), 1.0e-10);

    if (have && !skip) {
      const double Fac = fmin(o.FacMax, fmax(o.FacMin, o.FacSafe / pow(Err, 1.0 / o.ELO)));   // :743
      double Hnew = H * Fac;
      ist[Nstp]++;
      if ((Err <= 1.0) || (H <= o.Hmin)) {       // accept (:748-768)
        ist[Nacc]++;
#pragma unroll
        for (int i = 0; i < N; i++) Y[i] = o.ClipNegative ? fmax(X[i], 0.0) : X[i];
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

}  // namespace

bool unrolled_kernel_supports(int mech_id) { return mech_id == 1 /* GCKPP_MECH_HG */; }
int unrolled_block_threads() { return UNR_BLOCK; }
int unrolled_blocks_per_sm() { return UNR_SMEM ? 1 : UNR_MINB; }

cudaError_t launch_ros_unrolled(int mech_id, const RosArgs &a, int blocks, cudaStream_t s)
{
  if (mech_id != 1) return cudaErrorInvalidConfiguration;
  using U = Hg_unrolled;
  const size_t smem = UNR_SMEM ? sizeof(double) * UNR_BLOCK * (U::NNZ + U::NREACT + 6 * U::NVAR) : 0;
  if (smem) {
    cudaError_t e = cudaFuncSetAttribute(ros_unrolled_kernel<U>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
  }
  ros_unrolled_kernel<U><<<blocks, UNR_BLOCK, smem, s>>>(a);
  return cudaGetLastError();
}
