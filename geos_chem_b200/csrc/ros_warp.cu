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
// bundle-level profile (WARP_PROFILE): the dependent "sink" makes the in-order issue wait for the value first
#ifdef WARP_PROFILE
#define BSINK32(v) asm volatile("xor.b32 %0, %0, %1;" : "+r"(bsink_) : "r"(v))
#define BSINKD(v) asm volatile("xor.b32 %0, %0, %1;" : "+r"(bsink_) : "r"(__double2loint(v)))
#define BPROF(i) do { if (bp) { long long t_; asm volatile("mov.u64 %0, %%clock64;" : "=l"(t_)); bp[i] += t_ - bt_; bt_ = t_; } } while (0)
#else
#define BSINK32(v) do { } while (0)
#define BSINKD(v) do { } while (0)
#define BPROF(i) do { } while (0)
#endif
#ifndef WARP_FAST_RCP
#define WARP_FAST_RCP 1       // pivot reciprocals: MUFU approximation + 2 Newton steps (1) or IEEE division (0)
#endif

namespace {

constexpr int RS = WARP_RS;
constexpr unsigned TNONE = 0xFFFFu;
constexpr unsigned F_SYNC = 1u << 9, F_WRITE = 1u << 10, F_MUL = 1u << 11, F_DIAG = 1u << 12, F_UDIAG = 1u << 18;
constexpr int PRE_SHIFT = 13, PRE_MAX = 31;      // meta bits 13-17: group barriers of levels without a bundle of this warp, executed before the bundle
constexpr int SMEM_LIMIT = 232448;       // 227 KB opt-in maximum per block on sm_100

__host__ __device__ constexpr int cmax(int a, int b) { return a > b ? a : b; }
__host__ __device__ constexpr int cmin(int a, int b) { return a < b ? a : b; }
constexpr int a128(int x) { return (x + 127) & ~127; }

// warps per cell: the big mechanism is spread over a group of warps, the small one (everything is "tail") is not
template <class M> struct GroupOf { static constexpr int WG = WARP_WG; };
template <> struct GroupOf<Hg_dims> { static constexpr int WG = 1; };

struct GCtl {                 // per-group control words in shared memory
  double red[8];
  int cell, sing, pad_[2];
};

template <class M>
struct WLay {
  static constexpr int WG = GroupOf<M>::WG, GT = WG * 32;
  static constexpr int NYG = M::NSPEC + M::NLIT + 1;       // [VAR, FIX, literals, 1.0]
  static constexpr int NSCR = M::NREACT;                   // A(r), or one half of B(m)
  static_assert(M::NB <= 2 * M::NREACT, "B(m) is evaluated in two halves");
  // block-shared tables
  static constexpr int oCOEF = 0;
  static constexpr int oTPOS = a128(M::NCOEF * 8);
  static constexpr int oDIAG = oTPOS + 32 * 32 * 2;
  static constexpr int oGRP = a128(oDIAG + M::NVAR * 2);
  // one group's slice (every array 128-byte aligned: the bank analysis of wsched.py is relative to that)
  static constexpr int gG = 0;
  static constexpr int gYG = a128((M::NNZ + 2) * 8);          // G[NNZ] = 0.0, G[NNZ+1] = 1.0
  static constexpr int gX = gYG + a128(NYG * 8);
  static constexpr int XPAD = cmax(M::NVAR, 64);           // X[XPAD] = 0.0, the operand of padding terms (never written)
  static constexpr int gSCR = gX + a128((XPAD + 1) * 8);   // X[0..63] doubles as the pivot-row buffer of tail_lu
  static constexpr int gCTL = gSCR + a128(cmax(NSCR * 8, 16 * 32 * 8));   // SCR doubles as the L panel of tail_lu
  static constexpr int GSZ = gCTL + a128((int)sizeof(GCtl));
#ifndef WARP_MAXG
#define WARP_MAXG 15
#endif
  static constexpr int NGRP = cmin(cmin(1024 / GT, WG == 1 ? 16 : WARP_MAXG), (SMEM_LIMIT - oGRP) / GSZ);
  static constexpr int TOTAL = oGRP + NGRP * GSZ;
  static constexpr int NQ = (M::NSPEC + GT - 1) / GT;
};

enum { K_VDOT, K_JVS, K_LU, K_SOLVE };

template <int WG>
__device__ __forceinline__ void gsync(int group)
{
  if (WG == 1) __syncwarp();
  else asm volatile("bar.sync %0, %1;" :: "r"(group + 1), "n"(WG * 32) : "memory");
}

// ---- streamed tables: 16 bytes per lane per row, read straight from L2 into a three-row register window ----
// (A cp.async ring through shared memory costs 8 shared-memory wavefronts per row -- a quarter of the kernel's LSU
// traffic -- and is limited to one row per ~100-165 cycles per warp; ld.global.cg costs 4 and pipelines.)
// The window always holds rows pos, pos+1, pos+2: the rows of the NEXT bundle are requested while the current one
// is being processed.  The stream of a warp holds the rows of one attempt TWICE (+ a few rows), so `pos` never has
// to wrap inside a phase: it is folded back by L at the start of a phase only.
struct WReader {
  const uint4 *gsrc;        // this lane's column of the warp's stream
  int L, pos;               // rows of one attempt, row held in w0
  uint4 w0, w1, w2;
  __device__ __forceinline__ uint4 ld(int row) const { return __ldcg(gsrc + (size_t)row * 32); }
  __device__ __forceinline__ void seek(int row) { pos = row; w0 = ld(row); w1 = ld(row + 1); w2 = ld(row + 2); }
  // the stream is consumed in the fixed order of an accepted attempt; anything else (rejected step,
  // singular matrix, failed cell) re-positions the window
  __device__ __forceinline__ void at(int row)
  {
    if (pos >= L) pos -= L;
    if (pos != row) seek(row);
  }
  __device__ __forceinline__ void advance1() { w0 = w1; w1 = w2; w2 = ld(pos + 3); pos += 1; }
  __device__ __forceinline__ void advance2() { w0 = w2; w1 = ld(pos + 3); w2 = ld(pos + 4); pos += 2; }
  __device__ __forceinline__ void advance3() { w0 = ld(pos + 3); w1 = ld(pos + 4); w2 = ld(pos + 5); pos += 3; }
};

struct WCtx {
  unsigned char *Gb, *Xb, *Sb;
  const unsigned char *Cb;
  GCtl *ctl;
  double ghinv;
  int group;
};

__device__ __forceinline__ double ldb(const unsigned char *base, unsigned off) { return *reinterpret_cast<const double *>(base + off); }
__device__ __forceinline__ void stb(unsigned char *base, unsigned off, double v) { *reinterpret_cast<double *>(base + off) = v; }

// reciprocal for the pivot chains: hardware approximation + two Newton steps (full double accuracy up to the last
// ulp; ~55 cycles instead of ~90 for the IEEE division).  A zero pivot gives Inf/NaN, which the singular test catches.
__device__ __forceinline__ double rcp_fast(double x)
{
#if WARP_FAST_RCP
  double y;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
  double e = fma(-x, y, 1.0);
  y = fma(y, e, y);
  e = fma(-x, y, 1.0);
  return fma(y, e, y);
#else
  return 1.0 / x;
#endif
}

// nb bundles of one phase from this warp's stream: every lane applies exactly T terms (all operand loads of
// a bundle are issued before its FMAs), partial sums of a split row are combined by a segmented shuffle, the
// lane that owns the target finishes it; a bundle that ends a dependency level is followed by the group barrier.
template <int KIND, int WG>
#ifdef WARP_PROFILE
__device__ __forceinline__ void run_phase(WReader &rd, int nb, int nlev, WCtx &c, long long *bp = nullptr)
#else
__device__ __forceinline__ void run_phase(WReader &rd, int nb, int nlev, WCtx &c)
#endif
{
  int done = 0;          // group barriers executed in this phase; every warp of the group executes nlev of them
#ifdef WARP_PROFILE
  unsigned bsink_ = 0;
  long long bt_;
  asm volatile("mov.u64 %0, %%clock64;" : "=l"(bt_));
#endif
  const unsigned char *Hb = (KIND == K_VDOT || KIND == K_JVS) ? c.Cb : c.Gb;
  const unsigned char *Lb = (KIND == K_VDOT || KIND == K_JVS) ? c.Sb : (KIND == K_LU ? c.Gb : c.Xb);
  unsigned char *Tb = (KIND == K_VDOT || KIND == K_SOLVE) ? c.Xb : c.Gb;
#pragma unroll 1
  for (int b = 0; b < nb; b++) {
    const uint4 r0 = rd.w0;
    const unsigned hdr = r0.x, meta = r0.y;
    const int T = meta & 63, lg = (meta >> 6) & 7;
    const bool wr = (meta & F_WRITE) != 0;
    BSINK32(meta); BPROF(0);
    for (int i = (meta >> PRE_SHIFT) & PRE_MAX; i > 0; i--) { gsync<WG>(c.group); done++; }
    BPROF(4);
    // target and scale are loaded by every lane (lanes that write nothing point at the 0.0 / 1.0 slots)
    double old = 0.0, mul = 1.0;
    if (KIND != K_VDOT) old = ldb(Tb, hdr & 0xffffu);
    if (KIND == K_LU || KIND == K_SOLVE) mul = ldb(c.Gb, hdr >> 16);
    double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
#define LDT(w, hv, lv) const double hv = ldb(Hb, (w) >> 16), lv = ldb(Lb, (w) & 0xffffu)
#define LDP(p, w, hv, lv) double hv = 0.0, lv = 0.0; if (p) { hv = ldb(Hb, (w) >> 16); lv = ldb(Lb, (w) & 0xffffu); }
    if (T <= 2) {                       // one table row
      rd.advance1();
      LDP(T > 0, r0.z, h0, l0) LDP(T > 1, r0.w, h1, l1)
      a0 = h0 * l0; a1 = h1 * l1;
    } else if (T <= 6) {                // two table rows
      const uint4 r1 = rd.w1;
      rd.advance2();
      LDT(r0.z, h0, l0); LDT(r0.w, h1, l1); LDT(r1.x, h2, l2);
      LDP(T > 3, r1.y, h3, l3) LDP(T > 4, r1.z, h4, l4) LDP(T > 5, r1.w, h5, l5)
      a0 = h0 * l0; a1 = h1 * l1; a2 = h2 * l2; a3 = h3 * l3;
      a0 = fma(h4, l4, a0); a1 = fma(h5, l5, a1);
    } else {                            // three table rows, and a loop for the rare longer lane
      const uint4 r1 = rd.w1, r2 = rd.w2;
      rd.advance3();
      {
        LDT(r0.z, h0, l0); LDT(r0.w, h1, l1); LDT(r1.x, h2, l2); LDT(r1.y, h3, l3); LDT(r1.z, h4, l4); LDT(r1.w, h5, l5);
        a0 = h0 * l0; a1 = h1 * l1; a2 = h2 * l2; a3 = h3 * l3;
        a0 = fma(h4, l4, a0); a1 = fma(h5, l5, a1);
      }
      {
        LDT(r2.x, h6, l6);
        LDP(T > 7, r2.y, h7, l7) LDP(T > 8, r2.z, h8, l8) LDP(T > 9, r2.w, h9, l9)
        a2 = fma(h6, l6, a2); a3 = fma(h7, l7, a3);
        a0 = fma(h8, l8, a0); a1 = fma(h9, l9, a1);
      }
#pragma unroll 1
      for (int k = 10; k < T; k += 4) {
        const uint4 r3 = rd.w0;
        rd.advance1();
        LDT(r3.x, g0, m0);
        LDP(k + 1 < T, r3.y, g1, m1) LDP(k + 2 < T, r3.z, g2, m2) LDP(k + 3 < T, r3.w, g3, m3)
        a0 = fma(g0, m0, a0); a1 = fma(g1, m1, a1); a2 = fma(g2, m2, a2); a3 = fma(g3, m3, a3);
      }
    }
#undef LDT
#undef LDP
    double acc = (a0 + a1) + (a2 + a3);
    BSINKD(acc); BPROF(1);
    for (int s = 0; s < lg; s++) acc += __shfl_down_sync(FULLMASK, acc, 1 << s);
    BSINKD(acc); BPROF(2);
    {
      const unsigned t = hdr & 0xffffu;
      double v;
      if (KIND == K_VDOT) v = acc;
      else if (KIND == K_JVS) v = (old - acc) + ((meta & F_DIAG) ? c.ghinv : 0.0);
      else v = (old - acc) * mul;
      if (KIND == K_LU && (meta & F_UDIAG)) {              // uniform: some lane of the bundle finishes a pivot
        if (meta & F_DIAG) {
          if (!(fabs(v) >= DBL_MIN)) c.ctl->sing = 1;      // singular test of ros_PrepareMatrix, also catches NaN
          v = rcp_fast(v);
        }
      }
      if (wr) stb(Tb, t, v);
    }
    BPROF(3);
    if (meta & F_SYNC) { gsync<WG>(c.group); done++; }
    BPROF(4);
  }
  for (; done < nlev; done++) gsync<WG>(c.group);
#ifdef WARP_PROFILE
  if (bp) { bp[5] += nb; if (bsink_ == 0x12345678u) bp[5]++; }
#endif
}

// ---- tail block ------------------------------------------------------------------------------------------
// Dense right-looking LU of the m x m Schur complement (m = 32) by ONE warp, lane i = row i, in two column panels
// of 16 so that a lane holds 16 doubles, not 32:
//   A  factor the 32 x 16 panel of columns 0..15 (pivots 0..15); the multipliers go to a dense L panel in shared memory
//   B  apply those 16 pivots to columns 16..31 (row j of U is published by lane j, multipliers come from the L panel)
//   C  factor the trailing 16 x 16 block (pivots 16..31, lanes 16..31)
// Pivot loops are fully unrolled (static register indices); a pivot row is published by its owner lane through a
// double-buffered shared-memory row and read back as broadcasts; the reciprocal of the NEXT pivot is started as soon
// as its entry has been updated.  Entries outside the LU pattern are exact zeros and stay zero (fill-in closure).
// The factors are stored in their final form: L multipliers, the reciprocal diagonal, and U entries scaled by the
// reciprocal diagonal of their row.
#ifndef WARP_TAIL_INLINE
#define WARP_TAIL_INLINE __forceinline__
#endif
template <class M>
__device__ WARP_TAIL_INLINE bool tail_lu(double *Gc, double *buf /* 2 x 16 */, double *lpan /* 16 x 32 */, const uint16_t *tposT, int lane)
{
  constexpr int m = M::TAIL, hp = 16;
  static_assert(m == 2 * hp, "two panels of 16 columns");
  double myrd = 0.0;
  bool sing = false;
  // one panel factorisation: pivots j0..j0+15 on the 16 columns held in r[]
  auto factor = [&](double (&r)[hp], int j0) {
    double rinv = rcp_fast(__shfl_sync(FULLMASK, r[0], j0));
#pragma unroll
    for (int j = 0; j < hp; j++) {
      double *pb = buf + (j & 1) * hp;
      if (j < hp - 1) {
        if (lane == j0 + j) {
#pragma unroll
          for (int k = (j + 1) & ~1; k < hp; k += 2) reinterpret_cast<double2 *>(pb)[k >> 1] = make_double2(r[k], r[k + 1]);
        }
        __syncwarp();
      }
      if (lane == j0 + j) {
        myrd = rinv;
        sing = !(fabs(r[j]) >= DBL_MIN);
      }
      const double l = (lane > j0 + j) ? r[j] * rinv : 0.0;
      if (lane > j0 + j) r[j] = l;
      if (j < hp - 1) {
        r[j + 1] = fma(-l, pb[j + 1], r[j + 1]);
        rinv = rcp_fast(__shfl_sync(FULLMASK, r[j + 1], j0 + j + 1));
#pragma unroll
        for (int k = (j + 2) & ~1; k < hp; k += 2) {
          const double2 u = reinterpret_cast<const double2 *>(pb)[k >> 1];
          if (k >= j + 2) r[k] = fma(-l, u.x, r[k]);
          r[k + 1] = fma(-l, u.y, r[k + 1]);
        }
      }
    }
  };
  auto store = [&](const double (&r)[hp], int c0) {
#pragma unroll
    for (int k = 0; k < hp; k++) {
      const unsigned p = tposT[(c0 + k) * 32 + lane];
      if (p != TNONE) Gc[p] = (c0 + k < lane) ? r[k] : (c0 + k == lane ? myrd : r[k] * myrd);
    }
  };
  {   // ---- A
    double r[hp];
#pragma unroll
    for (int k = 0; k < hp; k++) {
      const unsigned p = tposT[k * 32 + lane];
      r[k] = (p != TNONE) ? Gc[p] : 0.0;
    }
    factor(r, 0);
#pragma unroll
    for (int k = 0; k < hp; k++) lpan[k * 32 + lane] = (lane > k) ? r[k] : 0.0;
    store(r, 0);
  }
  __syncwarp();
  {   // ---- B, C
    double s[hp];
#pragma unroll
    for (int k = 0; k < hp; k++) {
      const unsigned p = tposT[(hp + k) * 32 + lane];
      s[k] = (p != TNONE) ? Gc[p] : 0.0;
    }
#pragma unroll
    for (int j = 0; j < hp; j++) {
      double *pb = buf + (j & 1) * hp;
      if (lane == j) {
#pragma unroll
        for (int k = 0; k < hp; k += 2) reinterpret_cast<double2 *>(pb)[k >> 1] = make_double2(s[k], s[k + 1]);
      }
      __syncwarp();
      const double nl = -lpan[j * 32 + lane];
#pragma unroll
      for (int k = 0; k < hp; k += 2) {
        const double2 u = reinterpret_cast<const double2 *>(pb)[k >> 1];
        s[k] = fma(nl, u.x, s[k]);
        s[k + 1] = fma(nl, u.y, s[k + 1]);
      }
    }
    __syncwarp();
    factor(s, hp);
    store(s, hp);
  }
  return __any_sync(FULLMASK, sing && lane < m);
}

// the tail rows of one solve: forward chain x_i -= L(i,j) x_j (j ascending), scaling by the reciprocal
// pivot, backward chain x_i -= U'(i,j) x_j (j descending); x stays in a register per lane
template <class M>
__device__ WARP_TAIL_INLINE void tail_solve(const double *Gc, double *Xc, const uint16_t *tposT, const uint16_t *diag, int lane)
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

// SCR[i - i0] = rate(i) * YG[f1] * YG[f2] * YG[f3] for the items i0 <= i < i1 owned by this thread (i = i0 + gtid + k*GT).
// The table words and rate constants of up to CH items are loaded before the first product (L2 latency paid once).
template <int GT, int CH>
__device__ __forceinline__ void eval_terms(const uint2 *__restrict__ wt, const double *__restrict__ rcs, const double *YG, double *SCR,
                                           int i0, int i1, int gtid)
{
#pragma unroll 1
  for (int base = i0 + gtid; base < i1; base += CH * GT) {
    uint2 w[CH];
    double rc[CH];
#pragma unroll
    for (int k = 0; k < CH; k++) {
      const int i = base + k * GT;
      if (i < i1) { w[k] = __ldg(wt + i); rc[k] = __ldcg(rcs + i); }
    }
#pragma unroll
    for (int k = 0; k < CH; k++) {
      const int i = base + k * GT;
      if (i < i1) SCR[i - i0] = rc[k] * YG[w[k].x >> 16] * YG[w[k].y & 0xffff] * YG[w[k].y >> 16];
    }
  }
}

template <class M>
__global__ void __launch_bounds__(WLay<M>::NGRP * WLay<M>::GT, 1) ros_warp_kernel(WarpArgs P, RosArgs a)
{
  using L = WLay<M>;
  constexpr int N = M::NVAR, NQ = L::NQ, WG = L::WG, GT = L::GT;
  extern __shared__ __align__(128) unsigned char smem[];
  double *COEF = reinterpret_cast<double *>(smem + L::oCOEF);
  uint16_t *tposT = reinterpret_cast<uint16_t *>(smem + L::oTPOS);
  uint16_t *diag = reinterpret_cast<uint16_t *>(smem + L::oDIAG);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int group = warp / WG, wig = warp % WG, gtid = tid - group * GT;
  // stream index of this warp: group g's LEAD warp (stream 0: narrow levels, tail chains) is its warp g % WG,
  // so the lead warps of the groups of a block sit on different SM sub-partitions
  const int sw = (wig + WG - group % WG) % WG;
  const bool lead = sw == 0;
  for (int i = tid; i < M::NCOEF; i += blockDim.x) COEF[i] = P.coefs[i];
  for (int i = tid; i < 32 * 32; i += blockDim.x) tposT[i] = P.tpos[i];
  for (int i = tid; i < N; i += blockDim.x) diag[i] = P.diag[i];
  unsigned char *gb = smem + L::oGRP + group * L::GSZ;
  double *G = reinterpret_cast<double *>(gb + L::gG);        // [NNZ+1]
  double *YG = reinterpret_cast<double *>(gb + L::gYG);      // [NYG] state under evaluation + literals + 1.0
  double *X = reinterpret_cast<double *>(gb + L::gX);        // [NVAR] right-hand side / solution
  double *SCR = reinterpret_cast<double *>(gb + L::gSCR);    // [NSCR] A(r) or half of B(m)
  GCtl *ctl = reinterpret_cast<GCtl *>(gb + L::gCTL);
  for (int k = gtid; k < L::NYG; k += GT)
    YG[k] = (k < M::NSPEC) ? 1.0 : (k < M::NSPEC + M::NLIT ? P.lit[k - M::NSPEC] : 1.0);
  for (int k = gtid; k <= M::NNZ; k += GT) G[k] = 0.0;
  if (gtid == 0) { ctl->cell = -1; ctl->sing = 0; X[L::XPAD] = 0.0; G[M::NNZ + 1] = 1.0; }
  __syncthreads();           // the only block barrier: the shared tables are in place

  const RosOpts &o = a.o;
  const double Dir = (double)o.Direction;
  const int ggroup = blockIdx.x * L::NGRP + group;
  double *rcsA = P.rcs + (size_t)ggroup * (M::NREACT + M::NB);    // rate constants of the cell in A(r) order
  double *rcsB = rcsA + M::NREACT;                                  // ... and in B(m) order
  const uint2 *awt = reinterpret_cast<const uint2 *>(P.aw), *bwt = reinterpret_cast<const uint2 *>(P.bw);

  WCtx c;
  c.Gb = gb + L::gG; c.Xb = gb + L::gX; c.Sb = gb + L::gSCR; c.Cb = smem + L::oCOEF;
  c.ctl = ctl; c.ghinv = 0.0; c.group = group;
  WReader rd;
  rd.gsrc = P.stream + (size_t)P.w_off[sw] * 32 + lane;
  rd.L = P.w_rows[sw];
  rd.seek(0);
  int nbp[WARP_NPH], sof[WARP_NSEG];
#pragma unroll
  for (int i = 0; i < WARP_NPH; i++) nbp[i] = P.nb[sw][i];
#pragma unroll
  for (int i = 0; i < WARP_NSEG; i++) sof[i] = P.seg_off[sw][i];

  unsigned long long acc_stp = 0, acc_acc = 0, acc_fail = 0, acc_done = 0;
  WPROF_DECL
#ifdef WARP_PROFILE
  long long bacc_[4][6] = {};
  const bool prof_ = (tid == 0 && blockIdx.x == 0);
#define BP(k) , (prof_ ? bacc_[k] : nullptr)
#else
#define BP(k)
#endif

  // X = Fun(YG): A(r) = RCT(r) * prod(V) by the thread that owns reaction r, then the vdot stream.
  // YG must be visible to the group on entry; X is visible on return.
  auto fun = [&](int seg) {
    eval_terms<GT, 9>(awt, rcsA, YG, SCR, 0, M::NREACT, gtid);
    gsync<WG>(group);
    WPROF(10);
    rd.at(sof[seg]);
    run_phase<K_VDOT, WG>(rd, nbp[WP_VDOT], P.nlev[WP_VDOT], c BP(0));
  };
  // KppSolve on X in place (X visible to the group on entry and on return)
  auto solve = [&](int st) {
#pragma unroll 1
    for (int half = 0; half < 2; half++) {
      rd.at(sof[WS_FWD0 + (st < 2 ? 2 * st : 3 * st - 1) + half]);
      run_phase<K_SOLVE, WG>(rd, nbp[half ? WP_BWD : WP_FWD], P.nlev[half ? WP_BWD : WP_FWD], c BP(3));
      if (half == 0) {
        WPROF(7);
        if (lead) tail_solve<M>(G, X, tposT, diag, lane);
        gsync<WG>(group);
        WPROF(5);
      }
    }
  };

  for (;;) {
    if (gtid == 0) ctl->cell = atomicAdd(a.next, 1);
    gsync<WG>(group);
    const int w = ctl->cell;
    if (w >= a.nwork) break;
    const int cell = a.cell_list ? a.cell_list[w] : w;

    // ---- load the cell: concentrations into the owner registers and YG, rate constants in item order
    double Y[NQ], F0[NQ], K1[NQ], K2[NQ], K3[NQ];
#pragma unroll
    for (int q = 0; q < NQ; q++) {
      const int e = gtid + GT * q;
      Y[q] = (e < M::NSPEC) ? a.conc_in[(size_t)e * a.ncell + cell] : 0.0;
      F0[q] = K1[q] = K2[q] = K3[q] = 0.0;
      if (e < M::NSPEC) YG[e] = Y[q];
    }
#pragma unroll 4
    for (int r = gtid; r < M::NREACT; r += GT) {
      const int i0 = __ldg(awt + r).x & 0xffff;
      __stcg(rcsA + r, i0 < M::NREACT ? a.rconst[(size_t)i0 * a.rc_stride + (cell - a.rc_cell0)] : P.lit[i0 - M::NREACT]);
    }
#pragma unroll 4
    for (int m = gtid; m < M::NB; m += GT) {
      const int i0 = __ldg(bwt + m).x & 0xffff;
      __stcg(rcsB + m, i0 < M::NREACT ? a.rconst[(size_t)i0 * a.rc_stride + (cell - a.rc_cell0)] : P.lit[i0 - M::NREACT]);
    }
    gsync<WG>(group);

    // ---- Rosenbrock() start-up (gckpp_Integrator.F90:420-428, :637); every thread of the group carries the
    // control state redundantly (identical arithmetic, identical decisions)
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
      fun(WS_VDOT0);
#pragma unroll
      for (int q = 0; q < NQ; q++) {
        const int e = gtid + GT * q;
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
          const int e = gtid + GT * q;
          if (e < N) YG[e] = Y[q];
        }
        // ---- Ghimj = 1/(H*gamma) - Jac0 (:1973-1977); Jac0 is recomputed per attempt, B(m) in two halves
        for (int k = gtid; k < M::NNZ; k += GT) G[k] = 0.0;         // structural zeros / fill-in slots
        if (gtid == 0) ctl->sing = 0;
        gsync<WG>(group);
        c.ghinv = 1.0 / (Dir * H * o.Gamma[0]);
#pragma unroll 1
        for (int half = 0; half < 2; half++) {
          const int m0 = half * L::NSCR, m1 = cmin(M::NB, m0 + L::NSCR);
          if (m0 >= m1) break;
          eval_terms<GT, 9>(bwt, rcsB, YG, SCR, m0, m1, gtid);
          gsync<WG>(group);
          rd.at(sof[half ? WS_JVS2 : WS_JVS]);
          run_phase<K_JVS, WG>(rd, nbp[half ? WP_JVS2 : WP_JVS], P.nlev[half ? WP_JVS2 : WP_JVS], c BP(1));
        }
        WPROF(2);
        // ---- sparse LU (KppDecomp): head pivots from the stream, then the tail block in registers
        rd.at(sof[WS_LU]);
        run_phase<K_LU, WG>(rd, nbp[WP_LU], P.nlev[WP_LU], c BP(2));
        WPROF(3);
        if (lead && tail_lu<M>(G, X, SCR, tposT, lane)) ctl->sing = 1;
        gsync<WG>(group);
        const bool sing = ctl->sing != 0;
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
              const int e = gtid + GT * q;
              if (e < N)
                YG[e] = (st == 2) ? fma(o.A[2], K2[q], fma(o.A[1], K1[q], Y[q]))
                                  : fma(o.A[5], K3[q], fma(o.A[4], K2[q], fma(o.A[3], K1[q], Y[q])));
            }
            gsync<WG>(group);
            fun(st == 2 ? WS_VDOT1 : WS_VDOT2);
          }
          // right-hand side K_st = Fcn + sum_j C(st,j)/H K_j
#pragma unroll
          for (int q = 0; q < NQ; q++) {
            const int e = gtid + GT * q;
            if (e < N) {
              double v;
              if (st == 0) v = F0[q];
              else if (st == 1) v = fma(o.C[0] / dh, K1[q], F0[q]);
              else if (st == 2) v = fma(o.C[2] / dh, K2[q], fma(o.C[1] / dh, K1[q], X[e]));
              else v = fma(o.C[5] / dh, K3[q], fma(o.C[4] / dh, K2[q], fma(o.C[3] / dh, K1[q], X[e])));
              X[e] = v;
            }
          }
          gsync<WG>(group);
          WPROF(6);
          solve(st);
          if (st < 3) {
#pragma unroll
            for (int q = 0; q < NQ; q++) {
              const int e = gtid + GT * q;
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
          const int e = gtid + GT * q;
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
        if (WG > 1) {
          if (lane == 0) ctl->red[sw] = e2;
          gsync<WG>(group);
          e2 = 0.0;
#pragma unroll
          for (int i = 0; i < WG; i++) e2 += ctl->red[i];
        }
        const double Err = fmax(sqrt(e2 / (double)N), 1.0e-10);
        ist[Nfun] += 2; ist[Nsol] += 4;
        const double Fac = fmin(o.FacMax, fmax(o.FacMin, o.FacSafe / pow(Err, 1.0 / o.ELO)));
        double Hnew = H * Fac;
        ist[Nstp]++;
        if ((Err <= 1.0) || (H <= o.Hmin)) {           // accept (:748-768)
          ist[Nacc]++;
#pragma unroll
          for (int q = 0; q < NQ; q++) {
            const int e = gtid + GT * q;
            if (e < N) {
              Y[q] = o.ClipNegative ? fmax(yn[q], 0.0) : yn[q];
              YG[e] = Y[q];
            }
          }
          gsync<WG>(group);
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
      const int e = gtid + GT * q;
      if (e < M::NSPEC) a.conc_out[(size_t)e * a.ncell + cell] = Y[q];
    }
    if (a.istatus && gtid < 8) {
      int v = 0;
#pragma unroll
      for (int q = 0; q < 8; q++) v = (gtid == q) ? ist[q] : v;
      a.istatus[(size_t)gtid * a.ncell + cell] = v;
    }
    if (a.rstatus && gtid < 4)
      a.rstatus[(size_t)gtid * a.ncell + cell] = gtid == 0 ? Texit : (gtid == 1 ? Hexit : (gtid == 2 ? Hnewx : 0.0));
    if (a.ierr && gtid == 0) a.ierr[cell] = ierr;
    acc_stp += ist[Nstp]; acc_acc += ist[Nacc]; acc_done++;
    if (ierr < 0) acc_fail++;
    WPROF(9);
  }
#ifdef WARP_PROFILE
  if (tid == 0 && blockIdx.x == 0 && a.sums)
  {
    for (int i = 0; i < 12; i++) a.sums[8 + i] = (unsigned long long)pacc_[i];
    for (int k = 0; k < 4; k++)
      for (int i = 0; i < 6; i++) a.sums[20 + 6 * k + i] = (unsigned long long)bacc_[k][i];
  }
#endif
  if (gtid == 0 && a.sums) {
    atomicAdd(a.sums + 0, acc_stp);
    atomicAdd(a.sums + 1, acc_acc);
    atomicAdd(a.sums + 2, acc_fail);
    atomicAdd(a.sums + 3, acc_done);
  }
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

static int group_warps(int mech_id) { return mech_id == GCKPP_MECH_FULLCHEM ? WLay<fullchem_dims>::WG : WLay<Hg_dims>::WG; }
bool warp_kernel_supports(int mech_id) { return mech_id == GCKPP_MECH_FULLCHEM || mech_id == GCKPP_MECH_HG; }
int warp_cells_per_block(int mech_id) { return mech_id == GCKPP_MECH_FULLCHEM ? WLay<fullchem_dims>::NGRP : WLay<Hg_dims>::NGRP; }
int warp_smem_bytes(int mech_id) { return mech_id == GCKPP_MECH_FULLCHEM ? WLay<fullchem_dims>::TOTAL : WLay<Hg_dims>::TOTAL; }
size_t warp_rcs_doubles_per_group(int mech_id)
{
  return mech_id == GCKPP_MECH_FULLCHEM ? (size_t)fullchem_dims::NREACT + fullchem_dims::NB : (size_t)Hg_dims::NREACT + Hg_dims::NB;
}

// One table phase split into per-warp row sequences: bundle k of dependency level l (wsched.py marks the first
// bundle of a level with SYNC) goes to warp-stream (k + l) % WG, so narrow levels rotate over the warps and every
// warp streams about the same number of table rows.  In every stream the last bundle of a level gets the SYNC
// flag (= group barrier after it); levels in which a stream has no bundle are counted in the PRE field of its
// next bundle (that many barriers before it), trailing ones are made up by the kernel (it knows the level count).
static int deal_phase(const uint32_t *rows, int nrows, int wg, std::vector<std::vector<uint32_t>> &out, std::vector<int> &nb, int &nlev)
{
  out.assign(wg, {});
  nb.assign(wg, 0);
  std::vector<std::vector<std::pair<int, int>>> levels;         // (first row, rows) of the bundles of a level
  for (int r = 0; r < nrows;) {
    const uint32_t meta = rows[(size_t)r * 128 + 1];
    const int T = meta & 63;
    const int n = 1 + (T > 2 ? (T - 2 + 3) / 4 : 0);
    if ((meta & F_SYNC) || levels.empty()) levels.push_back({});
    levels.back().push_back({r, n});
    r += n;
  }
  nlev = (int)levels.size();
  std::vector<int> pre(wg, 0);
  for (int l = 0; l < nlev; l++) {
    std::vector<std::vector<std::pair<int, int>>> mine(wg);
    for (size_t k = 0; k < levels[l].size(); k++) mine[(k + l) % wg].push_back(levels[l][k]);
    for (int w = 0; w < wg; w++) {
      if (mine[w].empty()) { pre[w]++; continue; }
      if (pre[w] > PRE_MAX) return -6;
      for (size_t i = 0; i < mine[w].size(); i++) {
        size_t at = out[w].size();
        out[w].insert(out[w].end(), rows + (size_t)mine[w][i].first * 128, rows + (size_t)(mine[w][i].first + mine[w][i].second) * 128);
        for (int ln = 0; ln < 32; ln++) {
          uint32_t &meta = out[w][at + 4 * ln + 1];
          meta &= ~F_SYNC;
          if (i == 0) meta |= (uint32_t)pre[w] << PRE_SHIFT;
          if (i + 1 == mine[w].size()) meta |= F_SYNC;
        }
        nb[w]++;
      }
      pre[w] = 0;
    }
  }
  return 0;
}

int warp_plan_build(int mech_id, const gckpp_host_tables_t *T, const gckpp_wsched_tables_t *S, WarpHostPlan &hp)
{
  if (!warp_kernel_supports(mech_id)) return -1;
  if (mech_id == GCKPP_MECH_FULLCHEM ? !dims_match<fullchem_dims>(T, S) : !dims_match<Hg_dims>(T, S)) return -2;
  if (T->nspec + T->nlit + 1 >= 65536 || T->nreact + T->nlit >= 65535 || (T->nnz + 1) * 8 >= 65536) return -4;
  const int wg = group_warps(mech_id);
  std::vector<std::vector<uint32_t>> ph[WARP_NPH];
  std::vector<int> nb[WARP_NPH];
  for (int p = 0; p < WARP_NPH; p++)
    if (int rc = deal_phase(S->rows[p], S->nrows[p], wg, ph[p], nb[p], hp.nlev[p])) return rc;
  // the order one Rodas3 attempt consumes the tables in (NewF = T,F,T,T)
  static const int seg_phase[WARP_NSEG] = {WP_VDOT, WP_JVS, WP_JVS2, WP_LU, WP_FWD, WP_BWD, WP_FWD, WP_BWD, WP_VDOT, WP_FWD, WP_BWD,
                                           WP_VDOT, WP_FWD, WP_BWD};
  hp.stream.clear();
  for (int w = 0; w < WARP_WG; w++) { hp.w_off[w] = 0; hp.w_rows[w] = 0; }
  for (int w = 0; w < wg; w++) {
    hp.w_off[w] = (int)(hp.stream.size() / 128);
    std::vector<uint32_t> ws;
    for (int sg = 0; sg < WARP_NSEG; sg++) {
      hp.seg_off[w][sg] = (int)(ws.size() / 128);
      ws.insert(ws.end(), ph[seg_phase[sg]][w].begin(), ph[seg_phase[sg]][w].end());
    }
    // a segment that ends the stream wraps to row 0
    const int rows = (int)(ws.size() / 128);
    for (int sg = 0; sg < WARP_NSEG; sg++) if (hp.seg_off[w][sg] >= rows) hp.seg_off[w][sg] = 0;
    if (rows < 2 * RS) return -5;
    hp.w_rows[w] = rows;
    // the rows of one attempt twice, plus the prefetch distance: the ring never wraps inside an attempt
    hp.stream.insert(hp.stream.end(), ws.begin(), ws.end());
    hp.stream.insert(hp.stream.end(), ws.begin(), ws.end());
    hp.stream.insert(hp.stream.end(), ws.begin(), ws.begin() + (size_t)RS * 128);
    for (int p = 0; p < WARP_NPH; p++) hp.nb[w][p] = nb[p][w];
  }
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
  k<<<blocks, WLay<M>::NGRP * WLay<M>::GT, WLay<M>::TOTAL, s>>>(P, a);
  return cudaGetLastError();
}

cudaError_t launch_ros_warp(int mech_id, const WarpArgs &P, const RosArgs &a, int blocks, cudaStream_t s)
{
  if (mech_id == GCKPP_MECH_FULLCHEM) return launch_t<fullchem_dims>(P, a, blocks, s);
  if (mech_id == GCKPP_MECH_HG) return launch_t<Hg_dims>(P, a, blocks, s);
  return cudaErrorInvalidConfiguration;
}
