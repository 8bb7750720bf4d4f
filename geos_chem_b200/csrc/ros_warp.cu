// Warp-per-cell Rosenbrock (Rodas3) kernel -- the production path on B200.
//
// Reference routines covered (KPP/fullchem/..., identical structure for KPP/Hg):
//   ros_Integrator      gckpp_Integrator.F90:578-786      -> integrate_cell (stage loop, error control)
//   ros_PrepareMatrix   gckpp_Integrator.F90:1921-1999    -> "jvs" + "lu" bundle streams + tail_lu, singular test
//   ros_ErrorNorm       gckpp_Integrator.F90:1715-1745    -> warp all-reduce
//   Fun                 gckpp_Function.F90:51-2152        -> rate phase + "vdot" stream (aggregate form)
//   Jac_SP              gckpp_Jacobian.F90:48-20887       -> partials phase + "jvs" stream
//   KppDecomp           gckpp_LinearAlgebra.F90:46-83     -> "lu" stream (head pivots, pull form) + tail_lu
//   KppSolve            gckpp_LinearAlgebra.F90:644-2309  -> "fwd" / "bwd" streams + tail_solve
//
// Execution model
//   * ONE WARP INTEGRATES ONE CELL, from its first step to Tend, with its own adaptive step sequence; a
//     warp that finishes a cell pulls the next one from a global counter (replaces OpenMP SCHEDULE(DYNAMIC),
//     fullchem_mod.F90:541-542).  Warps never wait for each other: there is no block barrier after start-up,
//     the only ordering primitive is __syncwarp().  Retirement is therefore per cell, and cells with many
//     internal steps never idle cells with few.
//   * Everything an attempt touches more than once lives in the warp's slice of shared memory: the sparse
//     matrix G (LU_NONZERO doubles, KPP's LU_ICOL order), the state under evaluation YG, the right-hand side
//     X and the rate / partial-derivative scratch SCR.  A block is just as many warps as slices fit (3 for
//     fullchem, 16 for Hg).  Lane l owns elements l, l+32, ... of the state vectors (Y, Fcn0, K1..K3) in
//     registers.
//   * The sparse kernels are table driven (kppgen/wsched.py): a linear stream of bundles of 32 lane items in
//     pull form (every target once, all of its terms), executed with exactly T terms per lane.  The stream of
//     one attempt is cyclic and is prefetched with 16-byte cp.async into a per-warp shared-memory ring.
//   * The last 32 rows/columns (where KPP's ordering concentrates the fill-in) are a dense chain: their Schur
//     complement is factorised in registers (lane i = row i, pivot row broadcast by shuffles) and the
//     triangular sweeps carry x in a register per lane.
//
// Arithmetic: FP64 throughout, FMA contraction allowed, sums re-associated (see wsched.py); pivots are
// stored as reciprocals, so the four solves of an attempt contain no division.
#include <float.h>
#include <math.h>
#include <string.h>
#include <vector>
#include "ros_common.cuh"
#include "ros_warp.h"
#include "gen/fullchem_dims.h"
#include "gen/Hg_dims.h"
#include "../../include/gckpp_gpu.h"

#ifdef WARP_PROFILE
#define WPROF_DECL long long pt_ = clock64(), pacc_[12] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
#define WPROF(i) do { long long t_ = clock64(); pacc_[i] += t_ - pt_; pt_ = t_; } while (0)
#else
#define WPROF_DECL
#define WPROF(i) do { } while (0)
#endif

namespace {

constexpr int RS = WARP_RS;
constexpr unsigned TNONE = 0xFFFFu;
constexpr unsigned F_SYNC = 1u << 9, F_WRITE = 1u << 10, F_MUL = 1u << 11, F_DIAG = 1u << 12;
constexpr int SMEM_LIMIT = 232448;       // 227 KB opt-in maximum per block on sm_100

constexpr int cmax(int a, int b) { return a > b ? a : b; }
constexpr int cmin(int a, int b) { return a < b ? a : b; }
constexpr int a128(int x) { return (x + 127) & ~127; }

template <class M, int MAXW>
struct WLayT {
  static constexpr int NYG = M::NSPEC + M::NLIT + 1;       // [VAR, FIX, literals, 1.0]
  static constexpr int NSCR = cmax(M::NREACT, M::NB);
  // block-shared tables
  static constexpr int oCOEF = 0;
  static constexpr int oTPOS = a128(M::NCOEF * 8);
  static constexpr int oDIAG = oTPOS + 32 * 32 * 2;
  static constexpr int oWARP = a128(oDIAG + M::NVAR * 2);
  // one warp's slice (every array 128-byte aligned: the bank analysis of wsched.py is relative to that)
  static constexpr int wG = 0;
  static constexpr int wYG = a128((M::NNZ + 1) * 8);
  static constexpr int wX = wYG + a128(NYG * 8);
  static constexpr int wSCR = wX + a128(M::NVAR * 8);
  static constexpr int wRING = wSCR + a128(NSCR * 8);
  static constexpr int WSZ = wRING + RS * 512;
  static constexpr int NWB = cmin(MAXW, (SMEM_LIMIT - oWARP) / WSZ);
  static constexpr int TOTAL = oWARP + NWB * WSZ;
  static constexpr int NQ = (M::NSPEC + 31) / 32;
};
template <class M> struct WLay : WLayT<M, 16> {};

enum { K_VDOT, K_JVS, K_LU, K_SOLVE };

// ---- streamed tables: 16 bytes per lane per row, cp.async ring ----------------------------------------
struct WReader {
  const uint4 *gsrc;        // this lane's column of the stream
  uint32_t ring;            // shared-space byte address of this lane's column of the warp's ring
  int L, irow, islot, cslot, pos;
  __device__ __forceinline__ void issue()
  {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n\tcp.async.commit_group;"
                 :: "r"(ring + islot * 512), "l"(gsrc + (size_t)irow * 32) : "memory");
    if (++irow == L) irow = 0;
    if (++islot == RS) islot = 0;
  }
  __device__ __forceinline__ void seek(int row)
  {
    asm volatile("cp.async.wait_all;" ::: "memory");
    __syncwarp();
    pos = irow = row; islot = 0; cslot = 0;
#pragma unroll 1
    for (int i = 0; i < RS - 1; i++) issue();
  }
  // the stream is consumed in the fixed order of an accepted attempt; anything else (rejected step,
  // singular matrix, failed cell) re-positions the ring
  __device__ __forceinline__ void at(int row) { if (pos != row) seek(row); }
  __device__ __forceinline__ uint4 next()
  {
    uint4 v;
    asm volatile("cp.async.wait_group %0;" :: "n"(RS - 2) : "memory");
    asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(ring + cslot * 512) : "memory");
    if (++cslot == RS) cslot = 0;
    if (++pos == L) pos = 0;
    issue();
    return v;
  }
};

struct WCtx {
  unsigned char *Gb, *Xb, *Sb;
  const unsigned char *Cb;
  double ghinv;
  bool sing;
};

__device__ __forceinline__ double ldb(const unsigned char *base, unsigned off) { return *reinterpret_cast<const double *>(base + off); }
__device__ __forceinline__ void stb(unsigned char *base, unsigned off, double v) { *reinterpret_cast<double *>(base + off) = v; }

// nb bundles of one phase: every lane applies exactly T terms, partial sums of a split row are combined by a
// segmented shuffle, the lane that owns the target finishes it.
template <int KIND>
__device__ __forceinline__ void run_phase(WReader &rd, int nb, WCtx &c)
{
  const unsigned char *Hb = (KIND == K_VDOT || KIND == K_JVS) ? c.Cb : c.Gb;
  const unsigned char *Lb = (KIND == K_VDOT || KIND == K_JVS) ? c.Sb : (KIND == K_LU ? c.Gb : c.Xb);
  unsigned char *Tb = (KIND == K_VDOT || KIND == K_SOLVE) ? c.Xb : c.Gb;
#pragma unroll 1
  for (int b = 0; b < nb; b++) {
    uint4 r = rd.next();
    const unsigned hdr = r.x, meta = r.y;
    const int T = meta & 63, lg = (meta >> 6) & 7;
    if (meta & F_SYNC) __syncwarp();
    const bool wr = (meta & F_WRITE) != 0;
    double old = 0.0, mul = 1.0;
    if (KIND == K_LU || KIND == K_SOLVE) {
      if (wr) old = ldb(Tb, hdr & 0xffffu);
      if (wr && (meta & F_MUL)) mul = ldb(c.Gb, hdr >> 16);
    }
    double a0 = 0.0, a1 = 0.0;
    if (T > 0) a0 = fma(ldb(Hb, r.z >> 16), ldb(Lb, r.z & 0xffffu), a0);
    if (T > 1) a1 = fma(ldb(Hb, r.w >> 16), ldb(Lb, r.w & 0xffffu), a1);
#pragma unroll 1
    for (int k = 2; k < T; k += 4) {
      r = rd.next();
      a0 = fma(ldb(Hb, r.x >> 16), ldb(Lb, r.x & 0xffffu), a0);
      if (k + 1 < T) a1 = fma(ldb(Hb, r.y >> 16), ldb(Lb, r.y & 0xffffu), a1);
      if (k + 2 < T) a0 = fma(ldb(Hb, r.z >> 16), ldb(Lb, r.z & 0xffffu), a0);
      if (k + 3 < T) a1 = fma(ldb(Hb, r.w >> 16), ldb(Lb, r.w & 0xffffu), a1);
    }
    double acc = a0 + a1;
    for (int s = 0; s < lg; s++) acc += __shfl_down_sync(FULLMASK, acc, 1 << s);
    if (wr) {
      const unsigned t = hdr & 0xffffu;
      if (KIND == K_VDOT) {
        stb(Tb, t, acc);
      } else if (KIND == K_JVS) {
        stb(Tb, t, ((meta & F_DIAG) ? c.ghinv : 0.0) - acc);
      } else {
        double v = (old - acc) * mul;
        if (KIND == K_LU && (meta & F_DIAG)) {
          if (!(fabs(v) >= DBL_MIN)) c.sing = true;      // singular test of ros_PrepareMatrix, also catches NaN
          v = 1.0 / v;
        }
        stb(Tb, t, v);
      }
    }
  }
  __syncwarp();
}

// ---- tail block ------------------------------------------------------------------------------------------
// Dense right-looking LU of the Schur complement: lane i holds row i in registers, the pivot row is
// broadcast with shuffles.  The row registers ROTATE by one column per pivot (r[0] is always the pivot
// column), so the pivot loop stays rolled.  Entries outside the LU pattern are exact zeros and stay zero
// (fill-in closure).  The factors are stored in their final form: L multipliers, the reciprocal diagonal,
// and U entries scaled by the reciprocal diagonal of their row.
template <class M>
__device__ __forceinline__ bool tail_lu(double *Gc, const uint16_t *tposT, int lane)
{
  constexpr int m = M::TAIL;
  double r[m];
#pragma unroll
  for (int k = 0; k < m; k++) {
    const unsigned p = tposT[k * 32 + lane];
    r[k] = (p != TNONE) ? Gc[p] : 0.0;
  }
  double rinv = 1.0 / __shfl_sync(FULLMASK, r[0], 0);
  double myrd = 0.0;
  bool sing = false;
#pragma unroll 1
  for (int j = 0; j < m; j++) {
    const double l = (lane > j) ? r[0] * rinv : 0.0;
    if (lane == j) {
      myrd = rinv;
      sing = !(fabs(r[0]) >= DBL_MIN);
    }
    const unsigned p = tposT[j * 32 + lane];
    if (p != TNONE) Gc[p] = (lane > j) ? l : (lane == j ? rinv : r[0] * myrd);
#pragma unroll
    for (int k0 = 1; k0 < m; k0 += 8) {
      int hi[8], lo[8];
#pragma unroll
      for (int q = 0; q < 8; q++)
        if (k0 + q < m) {
          hi[q] = __shfl_sync(FULLMASK, __double2hiint(r[k0 + q]), j);
          lo[q] = __shfl_sync(FULLMASK, __double2loint(r[k0 + q]), j);
        }
#pragma unroll
      for (int q = 0; q < 8; q++)
        if (k0 + q < m) r[k0 + q - 1] = fma(-l, __hiloint2double(hi[q], lo[q]), r[k0 + q]);
      if (k0 == 1) rinv = 1.0 / __shfl_sync(FULLMASK, r[0], (j + 1) & 31);
    }
    r[m - 1] = 0.0;
  }
  return __any_sync(FULLMASK, sing && lane < m);
}

// the tail rows of one solve: forward chain x_i -= L(i,j) x_j (j ascending), scaling by the reciprocal
// pivot, backward chain x_i -= U'(i,j) x_j (j descending); x stays in a register per lane
template <class M>
__device__ __forceinline__ void tail_solve(const double *Gc, double *Xc, const uint16_t *tposT, const uint16_t *diag, int lane)
{
  constexpr int m = M::TAIL;
  double x = (lane < m) ? Xc[M::HEAD + lane] : 0.0;
#pragma unroll 1
  for (int j0 = 0; j0 < m - 1; j0 += 8) {
    double g[8];
#pragma unroll
    for (int q = 0; q < 8; q++) {
      const int j = j0 + q;
      const unsigned p = (j < m - 1) ? tposT[j * 32 + lane] : TNONE;
      g[q] = (lane > j && p != TNONE) ? Gc[p] : 0.0;
    }
#pragma unroll
    for (int q = 0; q < 8; q++) {
      const double xj = __shfl_sync(FULLMASK, x, (j0 + q) & 31);
      x = fma(-g[q], xj, x);
    }
  }
  if (lane < m) x *= Gc[diag[M::HEAD + lane]];
#pragma unroll 1
  for (int j0 = m - 1; j0 >= 1; j0 -= 8) {
    double g[8];
#pragma unroll
    for (int q = 0; q < 8; q++) {
      const int j = j0 - q;
      const unsigned p = (j >= 1) ? tposT[j * 32 + lane] : TNONE;
      g[q] = (lane < j && p != TNONE) ? Gc[p] : 0.0;
    }
#pragma unroll
    for (int q = 0; q < 8; q++) {
      const double xj = __shfl_sync(FULLMASK, x, (j0 - q) & 31);
      x = fma(-g[q], xj, x);
    }
  }
  if (lane < m) Xc[M::HEAD + lane] = x;
}

template <class M>
__global__ void __launch_bounds__(WLay<M>::NWB * 32, 1) ros_warp_kernel(WarpArgs P, RosArgs a)
{
  using L = WLay<M>;
  constexpr int N = M::NVAR, NQ = L::NQ;
  extern __shared__ __align__(128) unsigned char smem[];
  double *COEF = reinterpret_cast<double *>(smem + L::oCOEF);
  uint16_t *tposT = reinterpret_cast<uint16_t *>(smem + L::oTPOS);
  uint16_t *diag = reinterpret_cast<uint16_t *>(smem + L::oDIAG);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (int i = tid; i < M::NCOEF; i += blockDim.x) COEF[i] = P.coefs[i];
  for (int i = tid; i < 32 * 32; i += blockDim.x) tposT[i] = P.tpos[i];
  for (int i = tid; i < N; i += blockDim.x) diag[i] = P.diag[i];
  unsigned char *wb = smem + L::oWARP + warp * L::WSZ;
  double *G = reinterpret_cast<double *>(wb + L::wG);        // [NNZ+1]
  double *YG = reinterpret_cast<double *>(wb + L::wYG);      // [NYG] state under evaluation + literals + 1.0
  double *X = reinterpret_cast<double *>(wb + L::wX);        // [NVAR] right-hand side / solution
  double *SCR = reinterpret_cast<double *>(wb + L::wSCR);    // [NSCR] A(r) or B(m)
  for (int k = lane; k < L::NYG; k += 32)
    YG[k] = (k < M::NSPEC) ? 1.0 : (k < M::NSPEC + M::NLIT ? P.lit[k - M::NSPEC] : 1.0);
  for (int k = lane; k <= M::NNZ; k += 32) G[k] = 0.0;
  __syncthreads();           // the only block barrier: the shared tables are in place

  const RosOpts &o = a.o;
  const double Dir = (double)o.Direction;
  const int gwarp = blockIdx.x * L::NWB + warp;
  double *rcsA = P.rcs + (size_t)gwarp * (M::NREACT + M::NB);     // rate constants of the cell in A(r) order
  double *rcsB = rcsA + M::NREACT;                                  // ... and in B(m) order
  const uint2 *awt = reinterpret_cast<const uint2 *>(P.aw), *bwt = reinterpret_cast<const uint2 *>(P.bw);

  WCtx c;
  c.Gb = wb + L::wG; c.Xb = wb + L::wX; c.Sb = wb + L::wSCR; c.Cb = smem + L::oCOEF;
  c.ghinv = 0.0; c.sing = false;
  WReader rd;
  rd.gsrc = P.stream + lane;
  rd.ring = (uint32_t)__cvta_generic_to_shared(wb + L::wRING + lane * 16);
  rd.L = P.rows_total;
  rd.seek(0);

  unsigned long long acc_stp = 0, acc_acc = 0, acc_fail = 0, acc_done = 0;
  WPROF_DECL

  // X = Fun(YG): A(r) = RCT(r) * prod(V) by the lane that owns reaction r, then the vdot stream
  auto fun = [&](int which) {
#pragma unroll 2
    for (int r = lane; r < M::NREACT; r += 32) {
      const uint2 w = __ldg(awt + r);
      const double rc = __ldcg(rcsA + r);
      SCR[r] = rc * YG[w.x >> 16] * YG[w.y & 0xffff] * YG[w.y >> 16];
    }
    rd.at(P.off_vdot[which]);
    run_phase<K_VDOT>(rd, P.nb[WP_VDOT], c);
  };
  // KppSolve on X in place
  auto solve = [&](int which) {
#pragma unroll 1
    for (int half = 0; half < 2; half++) {
      rd.at(half ? P.off_bwd[which] : P.off_fwd[which]);
      run_phase<K_SOLVE>(rd, P.nb[half ? WP_BWD : WP_FWD], c);
      if (half == 0) tail_solve<M>(G, X, tposT, diag, lane);
    }
  };

  for (;;) {
    int w = 0;
    if (lane == 0) w = atomicAdd(a.next, 1);
    w = __shfl_sync(FULLMASK, w, 0);
    if (w >= a.nwork) break;
    const int cell = a.cell_list ? a.cell_list[w] : w;

    // ---- load the cell: concentrations into the owner registers and YG, rate constants in item order
    double Y[NQ], F0[NQ], K1[NQ], K2[NQ], K3[NQ];
#pragma unroll
    for (int q = 0; q < NQ; q++) {
      const int e = lane + 32 * q;
      Y[q] = (e < M::NSPEC) ? a.conc_in[(size_t)e * a.ncell + cell] : 0.0;
      F0[q] = K1[q] = K2[q] = K3[q] = 0.0;
      if (e < M::NSPEC) YG[e] = Y[q];
    }
#pragma unroll 4
    for (int r = lane; r < M::NREACT; r += 32) {
      const int i0 = __ldg(awt + r).x & 0xffff;
      __stcg(rcsA + r, i0 < M::NREACT ? a.rconst[(size_t)i0 * a.ncell + cell] : P.lit[i0 - M::NREACT]);
    }
#pragma unroll 4
    for (int m = lane; m < M::NB; m += 32) {
      const int i0 = __ldg(bwt + m).x & 0xffff;
      __stcg(rcsB + m, i0 < M::NREACT ? a.rconst[(size_t)i0 * a.ncell + cell] : P.lit[i0 - M::NREACT]);
    }
    __syncwarp();

    // ---- Rosenbrock() start-up (gckpp_Integrator.F90:420-428, :637)
    int ist[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    const double hs = a.hstart ? a.hstart[cell] : o.Hstart_rcntrl;
    const double Hstart = (hs > 0.0) ? fmin(fabs(hs), fabs(o.Tend - o.Tstart)) : fmax(o.Hmin, 1.0E-5);
    double T = o.Tstart, Hexit = 0.0, Hnewx = 0.0, Texit = 0.0;
    double H = fmin(fmax(fabs(o.Hmin), fabs(Hstart)), fabs(o.Hmax));
    if (fabs(H) <= 10.0 * o.Roundoff) H = 1.0E-5;
    H = Dir * H;
    bool rejLast = false, rejMore = false;
    int ierr = 0;
    WPROF(0);

    // ---- TimeLoop (:652)
    for (;;) {
      const bool inloop = (o.Direction > 0) ? ((T - o.Tend) + o.Roundoff <= 0.0) : ((o.Tend - T) + o.Roundoff <= 0.0);
      if (!inloop) { ierr = 1; break; }
      if (ist[Nstp] > o.Max_no_steps) { ierr = -6; break; }
      if (((T + 0.1 * H) == T) || (H <= o.Roundoff)) { ierr = -7; break; }
      H = fmin(H, fabs(o.Tend - T));
      // Fcn0 = Fun(Y); YG already holds Y (cell load / accepted step)
      fun(0);
#pragma unroll
      for (int q = 0; q < NQ; q++) {
        const int e = lane + 32 * q;
        F0[q] = (e < N) ? X[e] : 0.0;
      }
      ist[Nfun]++;
      if (!o.Autonomous) ist[Nfun]++;
      ist[Njac]++;
      WPROF(1);
      int nconsec = 0;
      // ---- UntilAccepted (:681)
      for (;;) {
        // after a rejected attempt YG holds a stage state: (re)store Y for the Jacobian
#pragma unroll
        for (int q = 0; q < NQ; q++) {
          const int e = lane + 32 * q;
          if (e < N) YG[e] = Y[q];
        }
        // ---- Ghimj = 1/(H*gamma) - Jac0 (:1973-1977); Jac0 is recomputed per attempt
        for (int k = lane; k < M::NNZ; k += 32) G[k] = 0.0;         // structural zeros / fill-in slots
        __syncwarp();
#pragma unroll 2
        for (int m = lane; m < M::NB; m += 32) {
          const uint2 w2 = __ldg(bwt + m);
          const double rc = __ldcg(rcsB + m);
          SCR[m] = rc * YG[w2.x >> 16] * YG[w2.y & 0xffff] * YG[w2.y >> 16];
        }
        c.ghinv = 1.0 / (Dir * H * o.Gamma[0]);
        c.sing = false;
        rd.at(P.off_jvs);
        run_phase<K_JVS>(rd, P.nb[WP_JVS], c);
        WPROF(2);
        // ---- sparse LU (KppDecomp): head pivots from the stream, then the tail block in registers
        run_phase<K_LU>(rd, P.nb[WP_LU], c);
        WPROF(3);
        bool sing = tail_lu<M>(G, tposT, lane);
        sing = __any_sync(FULLMASK, sing || c.sing);
        __syncwarp();
        WPROF(4);
        ist[Ndec]++;
        if (sing) {                                  // ros_PrepareMatrix :1985-1995
          ist[Nsng]++;
          nconsec++;
          if (nconsec <= 5) { H *= 0.5; continue; }
          ierr = -8;
          break;
        }
        const double dh = Dir * H;
        // ---- the four stages (Rodas3: NewF = T,F,T,T; :691-724), one copy of the Fun and solve code
#pragma unroll 1
        for (int st = 0; st < 4; st++) {
          if (st >= 2) {           // the state to evaluate: Y + sum_j A(st,j) K_j
#pragma unroll
            for (int q = 0; q < NQ; q++) {
              const int e = lane + 32 * q;
              if (e < N)
                YG[e] = (st == 2) ? fma(o.A[2], K2[q], fma(o.A[1], K1[q], Y[q]))
                                  : fma(o.A[5], K3[q], fma(o.A[4], K2[q], fma(o.A[3], K1[q], Y[q])));
            }
            __syncwarp();
            fun(st - 1);
          }
          // right-hand side K_st = Fcn + sum_j C(st,j)/H K_j
#pragma unroll
          for (int q = 0; q < NQ; q++) {
            const int e = lane + 32 * q;
            if (e < N) {
              double v;
              if (st == 0) v = F0[q];
              else if (st == 1) v = fma(o.C[0] / dh, K1[q], F0[q]);
              else if (st == 2) v = fma(o.C[2] / dh, K2[q], fma(o.C[1] / dh, K1[q], X[e]));
              else v = fma(o.C[5] / dh, K3[q], fma(o.C[4] / dh, K2[q], fma(o.C[3] / dh, K1[q], X[e])));
              X[e] = v;
            }
          }
          __syncwarp();
          WPROF(6);
          solve(st);
          if (st < 3) {
#pragma unroll
            for (int q = 0; q < NQ; q++) {
              const int e = lane + 32 * q;
              const double v = (e < N) ? X[e] : 0.0;
              if (st == 0) K1[q] = v;
              else if (st == 1) K2[q] = v;
              else K3[q] = v;
            }
          }
          WPROF(7);
        }
        // ---- new solution, error estimate and norm (:729-740, :1715-1745); K4 is read from X
        double yn[NQ], e2 = 0.0;
#pragma unroll
        for (int q = 0; q < NQ; q++) {
          const int e = lane + 32 * q;
          yn[q] = Y[q];
          if (e < N) {
            const double k4 = X[e];
            yn[q] = fma(o.M[3], k4, fma(o.M[2], K3[q], fma(o.M[1], K2[q], fma(o.M[0], K1[q], Y[q]))));
            const double ye = fma(o.E[3], k4, fma(o.E[2], K3[q], fma(o.E[1], K2[q], o.E[0] * K1[q])));
            const double at = __ldg(a.atol + (o.VectorTol ? e : 0)), rt = __ldg(a.rtol + (o.VectorTol ? e : 0));
            const double sc = at + rt * fmax(fabs(Y[q]), fabs(yn[q]));
            const double qq = ye / sc;
            e2 = fma(qq, qq, e2);
          }
        }
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) e2 += __shfl_xor_sync(FULLMASK, e2, off);
        const double Err = fmax(sqrt(e2 / (double)N), 1.0e-10);
        ist[Nfun] += 2; ist[Nsol] += 4;
        const double Fac = fmin(o.FacMax, fmax(o.FacMin, o.FacSafe / pow(Err, 1.0 / o.ELO)));
        double Hnew = H * Fac;
        ist[Nstp]++;
        if ((Err <= 1.0) || (H <= o.Hmin)) {           // accept (:748-768)
          ist[Nacc]++;
#pragma unroll
          for (int q = 0; q < NQ; q++) {
            const int e = lane + 32 * q;
            if (e < N) {
              Y[q] = o.ClipNegative ? fmax(yn[q], 0.0) : yn[q];
              YG[e] = Y[q];
            }
          }
          __syncwarp();
          T = T + Dir * H;
          Hnew = fmax(o.Hmin, fmin(Hnew, o.Hmax));
          if (rejLast) Hnew = fmin(Hnew, H);
          Hexit = H; Hnewx = Hnew; Texit = T;
          rejLast = false; rejMore = false;
          H = Hnew;
          WPROF(8);
          break;
        }
        if (rejMore) Hnew = H * o.FacRej;              // reject (:770-777)
        rejMore = rejLast;
        rejLast = true;
        H = Hnew;
        if (ist[Nacc] >= 1) ist[Nrej]++;
        WPROF(8);
      }
      if (ierr != 0) break;
    }

    // ---- retire the cell
#pragma unroll
    for (int q = 0; q < NQ; q++) {
      const int e = lane + 32 * q;
      if (e < M::NSPEC) a.conc_out[(size_t)e * a.ncell + cell] = Y[q];
    }
    if (a.istatus && lane < 8) {
      int v = 0;
#pragma unroll
      for (int q = 0; q < 8; q++) v = (lane == q) ? ist[q] : v;
      a.istatus[(size_t)lane * a.ncell + cell] = v;
    }
    if (a.rstatus && lane < 4)
      a.rstatus[(size_t)lane * a.ncell + cell] = lane == 0 ? Texit : (lane == 1 ? Hexit : (lane == 2 ? Hnewx : 0.0));
    if (a.ierr && lane == 0) a.ierr[cell] = ierr;
    acc_stp += ist[Nstp]; acc_acc += ist[Nacc]; acc_done++;
    if (ierr < 0) acc_fail++;
    WPROF(9);
  }
#ifdef WARP_PROFILE
  if (tid == 0 && blockIdx.x == 0 && a.sums)
    for (int i = 0; i < 12; i++) a.sums[8 + i] = (unsigned long long)pacc_[i];
#endif
  if (lane == 0 && a.sums) {
    atomicAdd(a.sums + 0, acc_stp);
    atomicAdd(a.sums + 1, acc_acc);
    atomicAdd(a.sums + 2, acc_fail);
    atomicAdd(a.sums + 3, acc_done);
  }
  asm volatile("cp.async.wait_all;" ::: "memory");
}

}  // namespace

// ---- host side ---------------------------------------------------------------------------------------

// Encode a rate/partial term (tables.h: rate, f0, f1, f2) as two words of 16-bit indices:
//   rate index into [RCONST, literals], factor indices into YG = [VAR, FIX, literals, 1.0].
static void encode_term(const int *t, int nreact, int nspec, int nlit, uint32_t *out)
{
  auto fac = [&](int f) -> uint32_t {
    if (f >= 0) return (uint32_t)f;
    if (f == -1) return (uint32_t)(nspec + nlit);
    return (uint32_t)(nspec + (-2 - f));
  };
  uint32_t i0 = t[0] >= 0 ? (uint32_t)t[0] : (uint32_t)(nreact + (~t[0]));
  out[0] = i0 | (fac(t[1]) << 16);
  out[1] = fac(t[2]) | (fac(t[3]) << 16);
}

template <class M> static bool dims_match(const gckpp_host_tables_t *T, const gckpp_wsched_tables_t *S)
{
  return T->nvar == M::NVAR && T->nspec == M::NSPEC && T->nreact == M::NREACT && T->nnz == M::NNZ && T->nb == M::NB &&
         T->nlit == M::NLIT && S->ncoef == M::NCOEF && S->tail == M::TAIL && S->head == M::HEAD;
}

bool warp_kernel_supports(int mech_id) { return mech_id == GCKPP_MECH_FULLCHEM || mech_id == GCKPP_MECH_HG; }
int warp_cells_per_block(int mech_id) { return mech_id == GCKPP_MECH_FULLCHEM ? WLay<fullchem_dims>::NWB : WLay<Hg_dims>::NWB; }
int warp_smem_bytes(int mech_id) { return mech_id == GCKPP_MECH_FULLCHEM ? WLay<fullchem_dims>::TOTAL : WLay<Hg_dims>::TOTAL; }
size_t warp_rcs_doubles_per_warp(int mech_id)
{
  return mech_id == GCKPP_MECH_FULLCHEM ? (size_t)fullchem_dims::NREACT + fullchem_dims::NB : (size_t)Hg_dims::NREACT + Hg_dims::NB;
}

int warp_plan_build(int mech_id, const gckpp_host_tables_t *T, const gckpp_wsched_tables_t *S, WarpHostPlan &hp)
{
  if (!warp_kernel_supports(mech_id)) return -1;
  if (mech_id == GCKPP_MECH_FULLCHEM ? !dims_match<fullchem_dims>(T, S) : !dims_match<Hg_dims>(T, S)) return -2;
  if (T->nspec + T->nlit + 1 >= 65536 || T->nreact + T->nlit >= 65535 || (T->nnz + 1) * 8 >= 65536) return -4;
  hp.stream.clear();
  auto push = [&](int ph) {
    int off = (int)(hp.stream.size() / 128);
    hp.stream.insert(hp.stream.end(), S->rows[ph], S->rows[ph] + (size_t)S->nrows[ph] * 128);
    return off;
  };
  // the order one Rodas3 attempt consumes the tables in (NewF = T,F,T,T)
  hp.off_vdot[0] = push(WP_VDOT);
  hp.off_jvs = push(WP_JVS);
  hp.off_lu = push(WP_LU);
  hp.off_fwd[0] = push(WP_FWD); hp.off_bwd[0] = push(WP_BWD);
  hp.off_fwd[1] = push(WP_FWD); hp.off_bwd[1] = push(WP_BWD);
  hp.off_vdot[1] = push(WP_VDOT);
  hp.off_fwd[2] = push(WP_FWD); hp.off_bwd[2] = push(WP_BWD);
  hp.off_vdot[2] = push(WP_VDOT);
  hp.off_fwd[3] = push(WP_FWD); hp.off_bwd[3] = push(WP_BWD);
  hp.rows_total = (int)(hp.stream.size() / 128);
  if (hp.rows_total < 2 * RS) return -5;
  for (int p = 0; p < 5; p++) hp.nb[p] = S->nbundles[p];
  hp.aw.resize(2 * (size_t)T->nreact);
  for (int r = 0; r < T->nreact; r++) encode_term(T->a_term + 4 * r, T->nreact, T->nspec, T->nlit, &hp.aw[2 * r]);
  hp.bw.resize(2 * (size_t)(T->nb > 0 ? T->nb : 1));
  for (int m = 0; m < T->nb; m++) encode_term(T->b_term + 4 * m, T->nreact, T->nspec, T->nlit, &hp.bw[2 * m]);
  hp.diag.resize(T->nvar);
  for (int i = 0; i < T->nvar; i++) hp.diag[i] = (uint16_t)T->diag[i];
  return 0;
}

template <class M>
static cudaError_t launch_t(const WarpArgs &P, const RosArgs &a, int blocks, cudaStream_t s)
{
  auto k = ros_warp_kernel<M>;
  // the opt-in is a per-device attribute of the function: set it on every launch (cheap) rather than cache it per process
  cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, WLay<M>::TOTAL);
  if (e != cudaSuccess) return e;
  k<<<blocks, WLay<M>::NWB * 32, WLay<M>::TOTAL, s>>>(P, a);
  return cudaGetLastError();
}

cudaError_t launch_ros_warp(int mech_id, const WarpArgs &P, const RosArgs &a, int blocks, cudaStream_t s)
{
  if (mech_id == GCKPP_MECH_FULLCHEM) return launch_t<fullchem_dims>(P, a, blocks, s);
  if (mech_id == GCKPP_MECH_HG) return launch_t<Hg_dims>(P, a, blocks, s);
  return cudaErrorInvalidConfiguration;
}
