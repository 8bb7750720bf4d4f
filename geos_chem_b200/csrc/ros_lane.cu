// Lane kernel ("kernel 3"): one grid cell per LANE, one warp per block, two blocks per SM for fullchem.
//
// The 32 cells of a warp run the same stream of control words (kppgen/lsched.py) on their own columns of an
// [element][32 lanes] workspace in HBM, so every global access of the warp is one 256-byte row, every shared-memory
// access is conflict-free, and nothing is ever exchanged between lanes: no shuffles, no barriers, no atomics on the
// data path.  What makes it fast is that the workspace operands arrive through a cp.async ring 16 batches ahead
// (csrc/lane_engine.cuh) while the randomly accessed vector of each phase -- the state under evaluation, the working
// row of the LU, the right-hand side of a solve -- sits in shared memory.  It is bound by HBM bandwidth (the LU
// factors of 32 cells, 1.45 MB per warp, stream through once per solve), not by dependent-instruction latency.
// Lanes are persistent: a lane whose cell is finished stores it and takes the next cell (warp-ballot refill), so
// the per-cell adaptive step counts idle nothing until the grid runs out.
//
// Any KPP mechanism with a Jacobian and every ICNTRL(3) method (the stage loop is generic); FMA contraction and the
// re-associations documented in lsched.py put it at rounding distance from the reference order.
//
// Reference routines covered (KPP/fullchem/...):
//   ros_Integrator      gckpp_Integrator.F90:578-786      -> ros_lane_kernel
//   ros_PrepareMatrix   gckpp_Integrator.F90:1921-1999    -> jac streams + LuF + singular loop
//   ros_ErrorNorm       gckpp_Integrator.F90:1715-1745    -> per lane, inline
//   Fun                 gckpp_Function.F90:51-2152        -> RatesF + SumsF
//   Jac_SP              gckpp_Jacobian.F90:48-20887       -> RatesF + SumsF (negated, diagonal shifted)
//   KppDecomp           gckpp_LinearAlgebra.F90:46-83     -> LuF
//   KppSolve            gckpp_LinearAlgebra.F90:644-2309  -> SolveF
#include <float.h>
#include <math.h>
#include "lane_engine.cuh"
#include "ros_common.cuh"
#include "ros_lane.h"

namespace {

#define LK_NOP 0
#define LK_LOAD 1
#define LK_ELIM 2
#define LK_FINL 3
#define LK_FIND 4
#define LK_FINU 5
#define LU_VALID (1u << 23)
#define LU_PF (1u << 24)
#define SV_VALID (1u << 23)
#define SV_LAST (1u << 24)
#define SV_RINV (1u << 25)
#define SV_PF (1u << 26)
#define SM_LAST (1u << 16)
#define SM_DIAG (1u << 17)
#define REC_FRESH (1u << 13)

// all pointers below already point at this lane's column: element k is p[32 * k]
#define EL(p, k) (p)[(size_t)(k) << 5]

// T(4t+e) = RCX(rcx) * VEC(v1) * VEC(v2) * VEC(v3)
struct RatesF {
  const double *rcx; const double *V; double *out; int t;
  __device__ __forceinline__ void issue(const uint4 *rec, unsigned dst)
  {
    const uint4 a = rec[0], b = rec[1];
    cp_async8(dst, &EL(rcx, a.x & 0xffffu));
    cp_async8(dst + 256, &EL(rcx, a.z & 0xffffu));
    cp_async8(dst + 512, &EL(rcx, b.x & 0xffffu));
    cp_async8(dst + 768, &EL(rcx, b.z & 0xffffu));
  }
  __device__ __forceinline__ void first(const uint4 *, const double *) { t = 0; }
  __device__ __forceinline__ void consume(const uint4 *rec, const double *g, const uint4 *, const double *)
  {
    const uint4 a = rec[0], b = rec[1];
    const double g0 = g[0], g1 = g[32], g2 = g[64], g3 = g[96];
    const double x0 = EL(V, a.x >> 16), y0 = EL(V, a.y & 0xffffu), z0 = EL(V, a.y >> 16);
    const double x1 = EL(V, a.z >> 16), y1 = EL(V, a.w & 0xffffu), z1 = EL(V, a.w >> 16);
    const double x2 = EL(V, b.x >> 16), y2 = EL(V, b.y & 0xffffu), z2 = EL(V, b.y >> 16);
    const double x3 = EL(V, b.z >> 16), y3 = EL(V, b.w & 0xffffu), z3 = EL(V, b.w >> 16);
    double *o = out + ((size_t)t << 7);
    __stcg(o, g0 * x0 * y0 * z0);
    __stcg(o + 32, g1 * x1 * y1 * z1);
    __stcg(o + 64, g2 * x2 * y2 * z2);
    __stcg(o + 96, g3 * x3 * y3 * z3);
    t++;
  }
};

// s += coef * T(src); LAST: out(n++) = s   (NEG: out = [ghinv on the diagonal] - s)
template <bool NEG>
struct SumsF {
  const double *src; double *out; double ghinv; double s; int n;
  __device__ __forceinline__ void issue(const uint4 *rec, unsigned dst)
  {
    const uint4 a = rec[0];
    cp_async8(dst, &EL(src, a.x & 0xffffu));
    cp_async8(dst + 256, &EL(src, a.y & 0xffffu));
    cp_async8(dst + 512, &EL(src, a.z & 0xffffu));
    cp_async8(dst + 768, &EL(src, a.w & 0xffffu));
  }
  __device__ __forceinline__ void first(const uint4 *, const double *) { s = 0.0; n = 0; }
  __device__ __forceinline__ void one(unsigned w, double c, double g)
  {
    s = fma(c, g, s);
    if (w & SM_LAST) {
      __stcg(&EL(out, n), NEG ? ((w & SM_DIAG) ? ghinv : 0.0) - s : s);
      n++;
      s = 0.0;
    }
  }
  __device__ __forceinline__ void consume(const uint4 *rec, const double *g, const uint4 *, const double *)
  {
    const uint4 a = rec[0];
    const double2 c01 = *reinterpret_cast<const double2 *>(rec + 2), c23 = *reinterpret_cast<const double2 *>(rec + 3);
    const double g0 = g[0], g1 = g[32], g2 = g[64], g3 = g[96];
    one(a.x, c01.x, g0); one(a.y, c01.y, g1); one(a.z, c23.x, g2); one(a.w, c23.y, g3);
  }
};

// KppDecomp, row by row: LOAD the row into W, ELIMinate with the finished rows, FINalise (see lsched.py)
struct LuF {
  double *ga; double *W; double rinv; int sing;
  // operands of the batch about to be consumed
  unsigned hdr; uint4 ew; double wj, g[4], wc[4];
  __device__ __forceinline__ void issue(const uint4 *rec, unsigned dst)
  {
    const uint4 a = rec[0];
    if (a.x & LU_PF) cp_async8(dst, &EL(ga, a.x & 0x1fffu));
    if (a.y & LU_PF) cp_async8(dst + 256, &EL(ga, a.y & 0x1fffu));
    if (a.z & LU_PF) cp_async8(dst + 512, &EL(ga, a.z & 0x1fffu));
    if (a.w & LU_PF) cp_async8(dst + 768, &EL(ga, a.w & 0x1fffu));
  }
  __device__ __forceinline__ double opnd(unsigned w, double ring, int kind)
  {
    if (w & LU_PF) return ring;
    if ((w & LU_VALID) && (kind == LK_LOAD || kind == LK_ELIM)) return __ldcg(&EL(ga, w & 0x1fffu));
    return 0.0;
  }
  __device__ __forceinline__ void load(const uint4 *rec, const double *r)
  {
    ew = rec[0];
    hdr = rec[1].x;
    const int kind = hdr & 7;
    wj = EL(W, (hdr >> 3) & 1023);
    wc[0] = EL(W, (ew.x >> 13) & 1023); wc[1] = EL(W, (ew.y >> 13) & 1023);
    wc[2] = EL(W, (ew.z >> 13) & 1023); wc[3] = EL(W, (ew.w >> 13) & 1023);
    g[0] = opnd(ew.x, r[0], kind); g[1] = opnd(ew.y, r[32], kind); g[2] = opnd(ew.z, r[64], kind); g[3] = opnd(ew.w, r[96], kind);
  }
  __device__ __forceinline__ void first(const uint4 *rec, const double *r) { rinv = 0.0; sing = 0; load(rec, r); }
  __device__ __forceinline__ void consume(const uint4 *, const double *, const uint4 *rec1, const double *r1)
  {
    // this batch's operands are in registers; take them out before the next batch's overwrite them
    const unsigned h = hdr;
    const uint4 e = ew;
    const double m = wj, g0 = g[0], g1 = g[1], g2 = g[2], g3 = g[3], w0 = wc[0], w1 = wc[1], w2 = wc[2], w3 = wc[3];
    const bool fresh = (rec1[1].x & REC_FRESH) != 0;
    if (!fresh) load(rec1, r1);
    const int kind = h & 7;
    if (kind == LK_ELIM) {
      EL(W, (e.x >> 13) & 1023) = fma(-m, g0, w0);
      EL(W, (e.y >> 13) & 1023) = fma(-m, g1, w1);
      EL(W, (e.z >> 13) & 1023) = fma(-m, g2, w2);
      EL(W, (e.w >> 13) & 1023) = fma(-m, g3, w3);
    } else if (kind == LK_LOAD) {
      EL(W, (e.x >> 13) & 1023) = g0; EL(W, (e.y >> 13) & 1023) = g1;
      EL(W, (e.z >> 13) & 1023) = g2; EL(W, (e.w >> 13) & 1023) = g3;
    } else if (kind == LK_FINL || kind == LK_FINU) {
      const double sc = (kind == LK_FINU) ? rinv : 1.0;
      if (e.x & LU_VALID) __stcg(&EL(ga, e.x & 0x1fffu), w0 * sc);
      if (e.y & LU_VALID) __stcg(&EL(ga, e.y & 0x1fffu), w1 * sc);
      if (e.z & LU_VALID) __stcg(&EL(ga, e.z & 0x1fffu), w2 * sc);
      if (e.w & LU_VALID) __stcg(&EL(ga, e.w & 0x1fffu), w3 * sc);
    } else if (kind == LK_FIND) {
      if (!(fabs(m) >= DBL_MIN)) sing = 1;              // singular test of ros_PrepareMatrix (:1985), also catches NaN
      rinv = 1.0 / m;
      __stcg(&EL(ga, e.x & 0x1fffu), rinv);
    }
    if (fresh) load(rec1, r1);
  }
};

// one triangular sweep of KppSolve on x (= VEC) in place
struct SolveF {
  const double *ga; double *X; double s;
  uint4 ew; unsigned r01, r23; double g[4], xc[4], xi[4];
  __device__ __forceinline__ void issue(const uint4 *rec, unsigned dst)
  {
    const uint4 a = rec[0];
    if (a.x & SV_PF) cp_async8(dst, &EL(ga, a.x & 0x1fffu));
    if (a.y & SV_PF) cp_async8(dst + 256, &EL(ga, a.y & 0x1fffu));
    if (a.z & SV_PF) cp_async8(dst + 512, &EL(ga, a.z & 0x1fffu));
    if (a.w & SV_PF) cp_async8(dst + 768, &EL(ga, a.w & 0x1fffu));
  }
  __device__ __forceinline__ void load(const uint4 *rec, const double *r)
  {
    ew = rec[0];
    const uint4 b = rec[1];
    r01 = b.x; r23 = b.y;
    xc[0] = EL(X, (ew.x >> 13) & 1023); xc[1] = EL(X, (ew.y >> 13) & 1023);
    xc[2] = EL(X, (ew.z >> 13) & 1023); xc[3] = EL(X, (ew.w >> 13) & 1023);
    xi[0] = EL(X, r01 & 0xffffu); xi[1] = EL(X, r01 >> 16); xi[2] = EL(X, r23 & 0xffffu); xi[3] = EL(X, r23 >> 16);
    g[0] = (ew.x & SV_PF) ? r[0] : 0.0; g[1] = (ew.y & SV_PF) ? r[32] : 0.0;
    g[2] = (ew.z & SV_PF) ? r[64] : 0.0; g[3] = (ew.w & SV_PF) ? r[96] : 0.0;
  }
  __device__ __forceinline__ void first(const uint4 *rec, const double *r) { s = 0.0; load(rec, r); }
  __device__ __forceinline__ void one(unsigned w, double gg, double x, int row, double xrow)
  {
    if (!(w & SV_RINV)) s = fma(gg, x, s);
    if (w & SV_LAST) {
      double v = xrow - s;
      if (w & SV_RINV) v *= gg;
      EL(X, row) = v;
      s = 0.0;
    }
  }
  __device__ __forceinline__ void consume(const uint4 *, const double *, const uint4 *rec1, const double *r1)
  {
    const uint4 e = ew;
    const unsigned a01 = r01, a23 = r23;
    const double g0 = g[0], g1 = g[1], g2 = g[2], g3 = g[3], x0 = xc[0], x1 = xc[1], x2 = xc[2], x3 = xc[3];
    const double i0 = xi[0], i1 = xi[1], i2 = xi[2], i3 = xi[3];
    const bool fresh = (rec1[1].z & REC_FRESH) != 0;
    if (!fresh) load(rec1, r1);
    one(e.x, g0, x0, a01 & 0xffffu, i0);
    one(e.y, g1, x1, a01 >> 16, i1);
    one(e.z, g2, x2, a23 & 0xffffu, i2);
    one(e.w, g3, x3, a23 >> 16, i3);
    if (fresh) load(rec1, r1);
  }
};

__global__ void __launch_bounds__(32, 1) ros_lane_kernel(LaneArgs P, RosArgs a)
{
  extern __shared__ __align__(128) unsigned char sm[];
  const int lane = threadIdx.x;
  const LaneDims &D = P.d;
  double *V = reinterpret_cast<double *>(sm) + lane;                                  // VEC, this lane's column
  double *dring = reinterpret_cast<double *>(sm + (size_t)D.nvec * 256);
  uint4 *tring = reinterpret_cast<uint4 *>(sm + (size_t)D.nvec * 256 + LANE_DRING_BYTES);
  double *ws = P.ws + (size_t)blockIdx.x * P.ws_stride + lane;
  double *Y = ws + ((size_t)P.oY << 5), *YN = ws + ((size_t)P.oYN << 5), *F0 = ws + ((size_t)P.oF0 << 5);
  double *FC = ws + ((size_t)P.oFC << 5), *K = ws + ((size_t)P.oK << 5), *GA = ws + ((size_t)P.oGA << 5);
  double *RCX = ws + ((size_t)P.oRCX << 5), *AB = ws + ((size_t)P.oAB << 5);
  const RosOpts &o = a.o;
  const int N = D.nvar;
  const double Dir = (double)o.Direction;

  for (int l = 0; l < D.nlit; l++) {
    const double v = __ldg(P.lit + l);
    EL(V, D.nspec + l) = v;
    EL(RCX, D.nreact + l) = v;
  }
  EL(V, D.one) = 1.0; EL(V, D.one + 1) = 0.0; EL(V, D.one + 2) = 0.0;
  for (int i = 0; i < D.nspec; i++) EL(V, i) = 0.0;
  for (int i = 0; i < D.ab_len; i++) EL(AB, i) = 0.0;       // padding entries of the rate streams read it
  for (int r = 0; r < D.nreact; r++) EL(RCX, r) = 0.0;

  // Fun(VEC) -> out
  auto fun = [&](double *out) {
    RatesF rf{RCX, V, AB, 0};
    run_stream<2>(P.rates_a, P.nchunk_ra, tring, dring, rf);
    SumsF<false> sf{AB, out, 0.0, 0.0, 0};
    run_stream<4>(P.sums_v, P.nchunk_sv, tring, dring, sf);
  };

#ifdef LANE_PROFILE
  long long pacc_[12] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0}, pt_ = clock64();
#define LPROF(i) do { const long long n_ = clock64(); pacc_[i] += n_ - pt_; pt_ = n_; } while (0)
#else
#define LPROF(i)
#endif
  // v(i) = base(i) + sum_{j<nk} cf[j] * K_j(i), eight elements at a time with every load of the block in flight
  // before the first use (the vectors live in HBM/L2; one exposed latency per block instead of one per load)
  auto lincomb = [&](const double *base, const double (&cf)[6], int nk, auto &&sink) {
#pragma unroll 1
    for (int i0 = 0; i0 < N; i0 += 8) {
      double v[8], b[8], k[6][8];
#pragma unroll
      for (int u = 0; u < 8; u++) v[u] = b[u] = (i0 + u < N) ? __ldcg(&EL(base, i0 + u)) : 0.0;
#pragma unroll
      for (int j = 0; j < 6; j++)
#pragma unroll
        for (int u = 0; u < 8; u++) k[j][u] = (j < nk && i0 + u < N) ? __ldcg(&EL(K, N * j + i0 + u)) : 0.0;
#pragma unroll
      for (int j = 0; j < 6; j++)
#pragma unroll
        for (int u = 0; u < 8; u++) v[u] = fma(cf[j], k[j][u], v[u]);
#pragma unroll
      for (int u = 0; u < 8; u++) if (i0 + u < N) sink(i0 + u, v[u], b[u], k[0][u], k[1][u], k[2][u], k[3][u], k[4][u], k[5][u]);
    }
  };
  // per-lane integration state (same variables as ros_generic_kernel)
  bool have = false, exhausted = false, newstep = false;
  bool RejectLastH = false, RejectMoreH = false;
  int cell = -1, nconsec = 0, ierr_cell = 0;
  int ist[8];
  double T = 0.0, H = 0.0, Hexit = 0.0, Hnew_out = 0.0, Texit = 0.0;
  unsigned long long acc_stp = 0, acc_acc = 0, acc_fail = 0, acc_done = 0;
#pragma unroll
  for (int q = 0; q < 8; q++) ist[q] = 0;

  for (;;) {
    // ---- retire finished cells and refill idle lanes -------------------------------------
    for (;;) {
      const int w = fetch_work(a.next, !have && !exhausted, lane);
      if (!have && !exhausted) {
        if (w >= a.nwork) {
          exhausted = true;
        } else {
          cell = a.cell_list ? a.cell_list[w] : w;
#pragma unroll 8
          for (int s = 0; s < D.nspec; s++) {
            const double v = __ldcs(a.conc_in + (size_t)s * a.ncell + cell);
            EL(Y, s) = v;
            if (s >= N) EL(V, s) = v;                  // fixed species stay in VEC for the whole integration
          }
#pragma unroll 8
          for (int r = 0; r < D.nreact; r++) EL(RCX, r) = __ldcs(a.rconst + (size_t)r * a.rc_stride + (cell - a.rc_cell0));
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
        // TimeLoop condition and the two guards at its top (:652-665)
        const bool inloop = (o.Direction > 0) ? ((T - o.Tend) + o.Roundoff <= 0.0) : ((o.Tend - T) + o.Roundoff <= 0.0);
        if (!inloop) ierr_cell = 1;
        else if (ist[Nstp] > o.Max_no_steps) ierr_cell = -6;
        else if (((T + 0.1 * H) == T) || (H <= o.Roundoff)) ierr_cell = -7;
        else H = fmin(H, fabs(o.Tend - T));
      }
      if (have && ierr_cell != 0) {
#pragma unroll 8
        for (int s = 0; s < D.nspec; s++) a.conc_out[(size_t)s * a.ncell + cell] = EL(Y, s);
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
    LPROF(0);

    // ---- one Rosenbrock attempt for every lane of the warp ---------------------------------
    // Fcn0 = Fun(Y) and Jac0 are recomputed by lanes repeating a rejected step (identical values); (:668-672)
#pragma unroll 8
    for (int i = 0; i < N; i++) EL(V, i) = __ldcg(&EL(Y, i));
    LPROF(1);
    fun(F0);
    LPROF(2);
    if (have && newstep) {
      ist[Nfun]++;
      if (!o.Autonomous) ist[Nfun]++;   // ros_FunTimeDerivative: with ICNTRL(15)=-1 Fun does not depend on T, dFdT == 0
      ist[Njac]++;
      nconsec = 0;
    }
    // Ghimj = 1/(H*gamma) - Jac0 (:1973-1977), written in the stream order of the solves
    {
      RatesF rf{RCX, V, AB, 0};
      run_stream<2>(P.rates_b, P.nchunk_rb, tring, dring, rf);
      LPROF(3);
      SumsF<true> sf{AB, GA, 1.0 / (Dir * H * o.Gamma[0]), 0.0, 0};
      run_stream<4>(P.sums_j, P.nchunk_sj, tring, dring, sf);
      LPROF(4);
    }
    int ising;
    {
      LuF lf;
      lf.ga = GA; lf.W = V;
      run_stream<2>(P.lu, P.nchunk_lu, tring, dring, lf);
      ising = lf.sing;
      LPROF(5);
    }
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
    const double rH = 1.0 / (Dir * H);
    for (int is = 1; is <= o.S; is++) {
      double *Ki = K + ((size_t)N * (is - 1) << 5);
      const int base = (is - 1) * (is - 2) / 2;
      if (is > 1 && o.NewF[is - 1]) {
        double cf[6] = {0, 0, 0, 0, 0, 0};
        for (int j = 1; j < is; j++) cf[j - 1] = o.A[base + j - 1];
        lincomb(Y, cf, is - 1, [&](int i, double v, double, double, double, double, double, double, double) { EL(V, i) = v; });
        LPROF(6);
        fun(FC);
        LPROF(2);
        if (have && !skip) ist[Nfun]++;
        src = FC;
      }
      {
        double cf[6] = {0, 0, 0, 0, 0, 0};
        for (int j = 1; j < is; j++) cf[j - 1] = o.C[base + j - 1] * rH;
        lincomb(src, cf, is - 1, [&](int i, double v, double, double, double, double, double, double, double) { EL(V, i) = v; });
      }
      LPROF(6);
      {
        SolveF sf;
        sf.ga = GA; sf.X = V;
        run_stream<2>(P.fwd, P.nchunk_fwd, tring, dring, sf);
        run_stream<2>(P.bwd, P.nchunk_bwd, tring, dring, sf);
      }
      LPROF(7);
#pragma unroll 8
      for (int i = 0; i < N; i++) __stcg(&EL(Ki, i), EL(V, i));
      if (have && !skip) ist[Nsol]++;
      LPROF(6);
    }
    // new solution, error estimate and its scaled norm (:729-740, :1715-1745)
    double Err = 0.0;
    {
      double cm[6] = {0, 0, 0, 0, 0, 0};
      for (int j = 0; j < o.S; j++) cm[j] = o.M[j];
      double ce[6] = {0, 0, 0, 0, 0, 0};
      for (int j = 0; j < o.S; j++) ce[j] = o.E[j];
      lincomb(Y, cm, o.S, [&](int i, double yn, double y0, double k0, double k1, double k2, double k3, double k4, double k5) {
        const double ye = fma(ce[5], k5, fma(ce[4], k4, fma(ce[3], k3, fma(ce[2], k2, fma(ce[1], k1, ce[0] * k0)))));
        __stcg(&EL(YN, i), yn);
        const double Ymax = fmax(fabs(y0), fabs(yn));
        const double Scale = o.VectorTol ? fma(__ldg(a.rtol + i), Ymax, __ldg(a.atol + i)) : fma(__ldg(a.rtol), Ymax, __ldg(a.atol));
        const double q = ye / Scale;
        Err = fma(q, q, Err);
      });
    }
    Err = fmax(sqrt(Err / (double)N), 1.0e-10);

    if (have && !skip) {
      const double Fac = fmin(o.FacMax, fmax(o.FacMin, o.FacSafe / pow(Err, 1.0 / o.ELO)));   // :743
      double Hnew = H * Fac;
      ist[Nstp]++;
      if ((Err <= 1.0) || (H <= o.Hmin)) {       // accept (:748-768)
        ist[Nacc]++;
#pragma unroll 8
        for (int i = 0; i < N; i++) {
          const double v = __ldcg(&EL(YN, i));
          __stcg(&EL(Y, i), o.ClipNegative ? fmax(v, 0.0) : v);
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
    LPROF(8);
  }
#ifdef LANE_PROFILE
  if (blockIdx.x == 0 && lane == 0 && a.sums) for (int i = 0; i < 12; i++) a.sums[8 + i] = (unsigned long long)pacc_[i];
#endif
  {
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
}

}  // namespace

size_t lane_smem_bytes(const LaneDims &d) { return (size_t)d.nvec * 256 + LANE_DRING_BYTES + LANE_TRING_BYTES(4); }

cudaError_t launch_ros_lane(const LaneArgs &P, const RosArgs &a, int blocks, cudaStream_t s)
{
  const size_t smem = lane_smem_bytes(P.d);
  cudaError_t e = cudaFuncSetAttribute(ros_lane_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  ros_lane_kernel<<<blocks, 32, smem, s>>>(P, a);
  return cudaGetLastError();
}
