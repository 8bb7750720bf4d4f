// Shared-memory-resident Rosenbrock (Rodas3) kernel -- the production path on B200.
//
// Reference routines covered (KPP/fullchem/..., identical structure for KPP/Hg):
//   ros_Integrator      gckpp_Integrator.F90:578-786      -> ros_smem_kernel (stage loop, error control)
//   ros_PrepareMatrix   gckpp_Integrator.F90:1921-1999    -> Jacobian round + LU rounds + tail LU + singular test
//   ros_ErrorNorm       gckpp_Integrator.F90:1715-1745    -> block reduction per cell
//   Fun                 gckpp_Function.F90:51-2152        -> rate phase + "vdot" round (aggregate form)
//   Jac_SP              gckpp_Jacobian.F90:48-20887       -> partials phase + "jvs" round
//   KppDecomp           gckpp_LinearAlgebra.F90:46-83     -> "lu" rounds (head pivots, DAG levels) + tail_lu
//   KppSolve            gckpp_LinearAlgebra.F90:644-2309  -> "fwd"/"bwd" rounds (push form) + tail chains
//
// Execution model
//   * One persistent thread block (NW warps) per SM integrates NC cells in LOCK STEP: every block
//     iteration is one Rosenbrock attempt for each of its NC cell slots.  A slot whose cell reaches
//     Tend (or fails) stores its result and pulls the next cell from a global counter, so the
//     per-cell adaptive step counts never idle a slot until the grid is exhausted.
//   * Everything an attempt touches more than once lives on chip.  Shared memory (per cell): the
//     sparse matrix G (LU_NONZERO doubles), the state being evaluated Yg, the right-hand side X,
//     the rate / partial-derivative scratch SCR; plus the tables of the triangular sweeps, which are
//     used four times per attempt.  Registers: thread i owns species i of every slot (Y, Fcn0, K1..K3).
//   * The sparse kernels are table driven (kppgen/sched.py): a round is a set of bundles of 32 lane
//     items; each table word is applied to all NC cells by the thread that fetched it.  The tables
//     of Fun, Jac and the LU rounds are laid out per warp in exact consumption order ("stream"),
//     cyclic over attempts, and prefetched with 16-byte cp.async into a per-warp shared-memory ring.
//   * The last 32 rows/columns (where KPP's ordering concentrates the fill-in) form a sequential
//     chain in the elimination DAG.  They are handled by ONE WARP PER CELL: the Schur complement is
//     factorised in registers (lane i = row i, pivot row broadcast by shuffles) and the triangular
//     sweeps carry x in a register per lane.
//   * Barriers: a round that keeps P < NW warps busy synchronises only those warps (named barrier P);
//     single-bundle rounds use __syncwarp only.
//   * Code size matters (one warp runs long dependent chains; the instruction cache must hold an
//     attempt): the three function evaluations and four solves of an attempt share one copy of the code.
//
// Arithmetic: FP64 throughout, FMA contraction allowed, sums re-associated (see sched.py).  The
// diagonal of the factors is stored as its reciprocal and U rows are pre-scaled by it, so the four
// solves of an attempt contain no division.
#include <float.h>
#include <math.h>
#include <string.h>
#include <type_traits>
#include <vector>
#include "ros_common.cuh"
#include "ros_smem.h"
#include "gen/fullchem_dims.h"
#include "gen/Hg_dims.h"
#include "../../include/gckpp_gpu.h"

// -DSMEM_PROFILE: thread 0 of block 0 accumulates clock64() per phase into sums[8..]
#ifdef SMEM_PROFILE
#define PROF_DECL long long pt_ = clock64(), pacc_[16] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0}; long long prl_[24], prs_[24]; for (int i_ = 0; i_ < 24; i_++) prl_[i_] = prs_[i_] = 0;
#define PROF(i) do { long long t_ = clock64(); pacc_[i] += t_ - pt_; pt_ = t_; } while (0)
#else
#define PROF_DECL
#define PROF(i) do { } while (0)
#endif

// The round directory in constant memory (one slot per mechanism): a constant load with a warp-uniform index lands in a
// uniform register, so everything decoded from a directory entry (bundles, warps, barrier size, flags) is warp-uniform for
// the compiler and the round-level branches are uniform branches.  Set per device before a launch (smem_set_directory).
#if SMEM_CONST_DIR
__constant__ uint32_t c_round_dir[2][128];
#endif

namespace {

constexpr int NC = SMEM_NC, NW = SMEM_NW, RS = SMEM_RS, NT = NW * 32;
constexpr unsigned TNONE = 0xFFFFu;

enum { OP_VDOT, OP_JVS, OP_LUDIV, OP_LUUPD, OP_SOLVE };

struct Slot {
  double T, H, Hexit, Hnew, Texit, ghinv;
  int cell, have, newstep, rejLast, rejMore, nconsec, ierr, skip, sing, accept;
  int out_cell, in_cell, out_ierr, pad_;
  int ist[8], out_ist[8];
  double out_r[3];
  double hc[6];        // C(k) / (Direction * H) of this attempt (the stage right-hand sides, :704-709), computed once per slot
};

constexpr int cmax(int a, int b) { return a > b ? a : b; }
constexpr int align16(int x) { return (x + 15) & ~15; }

// compile-time shared-memory layout of the hot arrays (byte offsets)
template <class M>
struct Lay {
  static constexpr int NYG = M::NSPEC + M::NLIT + 1;       // [VAR, FIX, literals, 1.0]
  static constexpr int NSCR = cmax(M::NREACT, M::NB);
  static constexpr int GS = M::NNZ + 1;                    // per-cell stride of G: slot NNZ holds 0.0 (padding terms)
  static constexpr int oG = 0;
  static constexpr int oYG = oG + NC * GS * 8;
  static constexpr int oX = oYG + NC * NYG * 8;
  static constexpr int oSCR = oX + NC * M::NVAR * 8;
  static constexpr int oCOEF = oSCR + (SMEM_SCR_GLOBAL ? 0 : NC * NSCR * 8);
  static constexpr int oRED = oCOEF + M::NCOEF * 8;
  static constexpr int oSLOT = oRED + NW * NC * 8;
  static constexpr int oRING = align16(oSLOT + NC * (int)sizeof(Slot));
  static constexpr int oDYN = oRING + NW * RS * 512;        // runtime-sized regions start here
  static constexpr int NA_IT = (M::NREACT + NT - 1) / NT;
  static constexpr int NB_IT = (M::NB + NT - 1) / NT;
};

// ---- per-warp streamed tables: 16 bytes per lane per chunk row, cp.async ring ------------------------
struct Reader {
  const uint4 *gsrc;        // this lane's column of the warp's stream
  uint32_t ring;            // shared-space byte address of this lane's column of the warp's ring
  int L, irow, islot, cslot;
#if SMEM_RING_PRELOAD
  uint4 pre;                // the next chunk row, already in registers (see next())
#endif
  __device__ __forceinline__ void issue()
  {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n\tcp.async.commit_group;"
                 :: "r"(ring + islot * 512), "l"(gsrc + (size_t)irow * 32));
    if (++irow == L) irow = 0;
    if (++islot == RS) islot = 0;
  }
  __device__ __forceinline__ uint4 lds()
  {
    uint4 v;
    asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(ring + cslot * 512));
    if (++cslot == RS) cslot = 0;
    return v;
  }
  __device__ __forceinline__ void prime()
  {
    for (int i = 0; i < RS - 1; i++) issue();
#if SMEM_RING_PRELOAD
    asm volatile("cp.async.wait_group %0;" :: "n"(RS - 2));
    pre = lds();
#endif
  }
  // SMEM_RING_PRELOAD: the ring read is software-pipelined -- next() hands out the row loaded by the previous call and starts
  // the load of the following one, so neither the first chunk of a bundle (right after a round barrier) nor its second chunk
  // waits for an LDS.128.  Row k+3 is copied into the slot of row k-1, which was read two calls ago: the look-ahead stays
  // three rows (one in registers, two in flight).
  __device__ __forceinline__ uint4 next()
  {
#if SMEM_RING_PRELOAD
    const uint4 v = pre;
    issue();
    asm volatile("cp.async.wait_group %0;" :: "n"(RS - 2));
    pre = lds();
    return v;
#else
    asm volatile("cp.async.wait_group %0;" :: "n"(RS - 2));
    const uint4 v = lds();
    issue();
    return v;
#endif
  }
};

// resident tables: plain shared-memory reads
struct ResReader {
  const uint4 *p;
  __device__ __forceinline__ uint4 next() { uint4 v = *p; p += 32; return v; }
};

// a resident bundle whose first two chunks were fetched before the barrier
struct PreReader {
  uint4 c0, c1;
  const uint4 *p;          // third chunk onwards
  int n;
  __device__ __forceinline__ uint4 next()
  {
    uint4 v;
    if (n == 0) v = c0;
    else if (n == 1) v = c1;
    else { v = *p; p += 32; }
    n++;
    return v;
  }
};

__device__ __forceinline__ void round_barrier(int P, int warp)
{
  if (P >= NW) __syncthreads();
  else if (P <= 1) { if (warp == 0) __syncwarp(); }
  else if (warp < P) asm volatile("bar.sync %0, %1;" :: "r"(P), "r"(P * 32) : "memory");
}

// One bundle: every lane applies its terms to all NC cells.
#ifdef SMEM_PROFILE
#define BPROF(i) do { if (bp) { long long t_ = clock64(); bp[i] += t_ - bt_; bt_ = t_; } } while (0)
#else
#define BPROF(i) do { } while (0)
#endif
template <class M, int OP, class RD>
#ifdef SMEM_PROFILE
__device__ __forceinline__ void run_bundle(RD &rd, unsigned char *smem, const Slot *slot, const double *scrg, long long *bp = nullptr)
#else
__device__ __forceinline__ void run_bundle(RD &rd, unsigned char *smem, const Slot *slot, const double *scrg)
#endif
{
  using L = Lay<M>;
#ifdef SMEM_PROFILE
  long long bt_ = clock64();
#endif
  double *G = reinterpret_cast<double *>(smem + L::oG);
  double *X = reinterpret_cast<double *>(smem + L::oX);
  const uint4 c0 = rd.next();
  const unsigned lw = c0.x;
#if SMEM_UNIFORM_LW
  // maxlen and log2(g) are the same in every lane of a bundle: taken from lane 0 through a broadcast, the chunk loop and
  // the shuffle loop become warp-uniform for the compiler (no divergence checks around the shuffles)
  const unsigned lwu = __shfl_sync(FULLMASK, lw, 0);
  const int row = lw & 0x1fff, maxlen = (lwu >> 19) & 63, lg = (lwu >> 25) & 7;
#else
  const int row = lw & 0x1fff, maxlen = (lw >> 19) & 63, lg = (lw >> 25) & 7;
#endif
  double acc[NC];
#pragma unroll
  for (int c = 0; c < NC; c++) acc[c] = 0.0;
  // Term words carry BYTE offsets; lanes with fewer terms than the bundle are padded with terms whose
  // product is an exact zero, so every term is loaded and accumulated unconditionally and the loads of
  // a whole chunk issue back to back.
  const unsigned char *Gb = smem + L::oG, *Xb = smem + L::oX, *Sb = smem + L::oSCR, *Cb = smem + L::oCOEF;
  auto ld = [](const unsigned char *base, unsigned off) { return *reinterpret_cast<const double *>(base + off); };
  auto term = [&](unsigned w) {
    const unsigned hi = w >> 16, lo = w & 0xffffu;
    if (OP == OP_VDOT || OP == OP_JVS) {
      const double cf = ld(Cb, hi);
#if SMEM_SCR_GLOBAL
#pragma unroll
      for (int c = 0; c < NC; c++) acc[c] = fma(cf, __ldcg(scrg + c * L::NSCR + (lo >> 3)), acc[c]);
#else
#pragma unroll
      for (int c = 0; c < NC; c++) acc[c] = fma(cf, ld(Sb + c * L::NSCR * 8, lo), acc[c]);
#endif
    } else if (OP == OP_LUUPD) {
#pragma unroll
      for (int c = 0; c < NC; c++) acc[c] = fma(ld(Gb + c * L::GS * 8, hi), ld(Gb + c * L::GS * 8, lo), acc[c]);
    } else if (OP == OP_SOLVE) {
#pragma unroll
      for (int c = 0; c < NC; c++) acc[c] = fma(ld(Gb + c * L::GS * 8, hi), ld(Xb + c * M::NVAR * 8, lo), acc[c]);
    }
  };
  if (OP == OP_LUDIV) {
    if ((lw >> 28) & 1) {
      const unsigned dp = c0.y & 0xffffu;
      // RECIP_DIAG tables: the pivot's diagonal already holds its reciprocal (written with its final update)
#pragma unroll
      for (int c = 0; c < NC; c++)
        G[c * L::GS + row] = M::RECIP_DIAG ? G[c * L::GS + row] * ld(Gb + c * L::GS * 8, dp) : G[c * L::GS + row] / ld(Gb + c * L::GS * 8, dp);
    }
    return;
  }
  BPROF(12);
  if ((SMEM_EXACT_STEPS >> (OP == OP_VDOT ? 0 : OP == OP_JVS ? 1 : OP == OP_LUUPD ? 2 : 3)) & 1) {
  // only the bundle's maxlen term steps are applied (warp-uniform switches, one straight-line block per count so the
  // loads of a chunk still issue back to back): most bundles of the LU and sweep rounds carry one or two terms per
  // lane, and a pad step costs as many instructions as a real one
  switch (maxlen < 3 ? maxlen : 3) {
    case 3: term(c0.y); term(c0.z); term(c0.w); break;
    case 2: term(c0.y); term(c0.z); break;
    case 1: term(c0.y); break;
    default: break;
  }
  for (int k0 = 3; k0 < maxlen; k0 += 4) {
    const uint4 cc = rd.next();
    const int rem = maxlen - k0;
    switch (rem < 4 ? rem : 4) {
      case 4: term(cc.x); term(cc.y); term(cc.z); term(cc.w); break;
      case 3: term(cc.x); term(cc.y); term(cc.z); break;
      case 2: term(cc.x); term(cc.y); break;
      default: term(cc.x); break;
    }
  }
  } else {
  term(c0.y); term(c0.z); term(c0.w);
  for (int k0 = 3; k0 < maxlen; k0 += 4) {
    const uint4 cc = rd.next();
    term(cc.x); term(cc.y); term(cc.z); term(cc.w);
  }
  }
  BPROF(13);
  for (int s = 0; s < lg; s++) {
#pragma unroll
    for (int c = 0; c < NC; c++) acc[c] += __shfl_down_sync(FULLMASK, acc[c], 1 << s);
  }
  BPROF(14);
  if ((lw >> 28) & 1) {
    if (OP == OP_VDOT) {
#pragma unroll
      for (int c = 0; c < NC; c++) X[c * M::NVAR + row] = acc[c];
    } else if (OP == OP_JVS) {
      if (M::RECIP_DIAG && ((lw >> 30) & 1)) {       // a head pivot no LU update touches: final here, store 1/d (singular test :1985)
#pragma unroll
        for (int c = 0; c < NC; c++) {
          const double d = slot[c].ghinv - acc[c];
          if (!(fabs(d) >= DBL_MIN)) const_cast<Slot *>(slot)[c].sing = 1;
          G[c * L::GS + row] = 1.0 / d;
        }
      } else {
#pragma unroll
        for (int c = 0; c < NC; c++) G[c * L::GS + row] = (((lw >> 29) & 1) ? slot[c].ghinv : 0.0) - acc[c];
      }
    } else if (OP == OP_LUUPD) {
      if (M::RECIP_DIAG && ((lw >> 30) & 1)) {       // the last update of a head pivot's diagonal
#pragma unroll
        for (int c = 0; c < NC; c++) {
          const double d = G[c * L::GS + row] - acc[c];
          if (!(fabs(d) >= DBL_MIN)) const_cast<Slot *>(slot)[c].sing = 1;
          G[c * L::GS + row] = 1.0 / d;
        }
      } else {
#pragma unroll
        for (int c = 0; c < NC; c++) G[c * L::GS + row] -= acc[c];
      }
    } else if (OP == OP_SOLVE) {
#pragma unroll
      for (int c = 0; c < NC; c++) X[c * M::NVAR + row] -= acc[c];
    }
  }
  BPROF(15);
}

// The same bundle for ONE cell (cell-split rounds, see ros_smem.h): the warp reads the same table words and touches only
// cell `cc`.  Whole chunks, like run_bundle.
template <class M, int OP, class RD>
__device__ __forceinline__ void run_bundle_cell(RD &rd, unsigned char *smem, const Slot *slot, int cc)
{
  using L = Lay<M>;
  const uint4 c0 = rd.next();
  const unsigned lw = c0.x;
#if SMEM_UNIFORM_LW
  const unsigned lwu = __shfl_sync(FULLMASK, lw, 0);
  const int row = lw & 0x1fff, maxlen = (lwu >> 19) & 63, lg = (lwu >> 25) & 7;
#else
  const int row = lw & 0x1fff, maxlen = (lw >> 19) & 63, lg = (lw >> 25) & 7;
#endif
  const unsigned char *Gb = smem + L::oG + cc * L::GS * 8, *Xb = smem + L::oX + cc * M::NVAR * 8;
  double *Gc = reinterpret_cast<double *>(smem + L::oG) + cc * L::GS;
  double *Xc = reinterpret_cast<double *>(smem + L::oX) + cc * M::NVAR;
  auto ld = [](const unsigned char *base, unsigned off) { return *reinterpret_cast<const double *>(base + off); };
  if (OP == OP_LUDIV) {
    if ((lw >> 28) & 1) Gc[row] = M::RECIP_DIAG ? Gc[row] * ld(Gb, c0.y & 0xffffu) : Gc[row] / ld(Gb, c0.y & 0xffffu);
    return;
  }
  double acc = 0.0;
  auto term = [&](unsigned w) {
    const unsigned hi = w >> 16, lo = w & 0xffffu;
    if (OP == OP_LUUPD) acc = fma(ld(Gb, hi), ld(Gb, lo), acc);
    else acc = fma(ld(Gb, hi), ld(Xb, lo), acc);
  };
  term(c0.y); term(c0.z); term(c0.w);
  for (int k0 = 3; k0 < maxlen; k0 += 4) {
    const uint4 c1 = rd.next();
    term(c1.x); term(c1.y); term(c1.z); term(c1.w);
  }
  for (int s = 0; s < lg; s++) acc += __shfl_down_sync(FULLMASK, acc, 1 << s);
  if ((lw >> 28) & 1) {
    if (OP == OP_LUUPD) {
      if (M::RECIP_DIAG && ((lw >> 30) & 1)) {
        const double d = Gc[row] - acc;
        if (!(fabs(d) >= DBL_MIN)) const_cast<Slot *>(slot)[cc].sing = 1;
        Gc[row] = 1.0 / d;
      } else Gc[row] -= acc;
    } else Xc[row] -= acc;
  }
}

// ---- tail block: one warp per cell ---------------------------------------------------------------------
// Dense right-looking LU of the Schur complement: lane i holds row i in registers, the pivot row is
// broadcast with shuffles.  The row registers ROTATE by one column per pivot (r[0] is always the pivot
// column), so the pivot loop stays rolled and the code small enough for the instruction cache.
// Entries outside the LU pattern are exact zeros and stay zero (fill-in closure).
// Pivots j0..j1-1 with KMAX live column slots: at pivot j only the slots 0..m-1-j can be non-zero, so the later
// spans of the pivot loop update (and shuffle) fewer columns -- 608 column updates instead of 992 for m = 32.
template <class M, int KMAX>
__device__ __forceinline__ void tail_lu_span(double (&r)[M::TAIL], double &rinv, double &myrd, bool &sing, double *Gc,
                                             const uint16_t *tposT, int lane, int j0, int j1)
{
#pragma unroll 1
  for (int j = j0; j < j1; j++) {
    const double l = (lane > j) ? r[0] * rinv : 0.0;
    if (lane == j) {
      myrd = rinv;
      sing = !(fabs(r[0]) >= DBL_MIN);          // singular test of ros_PrepareMatrix, also catches NaN
    }
    const unsigned p = tposT[j * 32 + lane];
    if (p != TNONE) Gc[p] = (lane > j) ? l : (lane == j ? rinv : r[0] * myrd);
    // update column j+k and rotate it to slot k-1; shuffles issued in batches so their latency overlaps
#pragma unroll
    for (int k0 = 1; k0 < KMAX; k0 += 8) {
      int hi[8], lo[8];
#pragma unroll
      for (int q = 0; q < 8; q++)
        if (k0 + q < KMAX) {
          hi[q] = __shfl_sync(FULLMASK, __double2hiint(r[k0 + q]), j);
          lo[q] = __shfl_sync(FULLMASK, __double2loint(r[k0 + q]), j);
        }
#pragma unroll
      for (int q = 0; q < 8; q++)
        if (k0 + q < KMAX) r[k0 + q - 1] = fma(-l, __hiloint2double(hi[q], lo[q]), r[k0 + q]);
      if (k0 == 1) rinv = 1.0 / __shfl_sync(FULLMASK, r[0], (j + 1) & 31);
    }
    if (KMAX == 1) rinv = 1.0 / __shfl_sync(FULLMASK, r[0], (j + 1) & 31);
    r[KMAX - 1] = 0.0;
  }
}

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
  // The reciprocal of the next pivot is started as soon as its column has been updated (first batch),
  // so its latency overlaps the rest of the row update.  The factors are stored in their final form:
  // L multipliers, the reciprocal diagonal, and U entries scaled by the reciprocal diagonal of their row
  // (lane i keeps 1/d_i from the step at which column i was the pivot).
  double rinv = 1.0 / __shfl_sync(FULLMASK, r[0], 0);
  double myrd = 0.0;
  bool sing = false;
#if SMEM_TAIL_SPANS
  constexpr int Q = (m + 3) / 4;                 // four spans of the pivot loop with 4Q, 3Q, 2Q, Q (clipped to m) live slots
  constexpr int K0 = m, K1 = m - Q > 1 ? m - Q : 1, K2 = m - 2 * Q > 1 ? m - 2 * Q : 1, K3 = m - 3 * Q > 1 ? m - 3 * Q : 1;
  tail_lu_span<M, K0>(r, rinv, myrd, sing, Gc, tposT, lane, 0, Q < m ? Q : m);
  if (Q < m) tail_lu_span<M, K1>(r, rinv, myrd, sing, Gc, tposT, lane, Q, 2 * Q < m ? 2 * Q : m);
  if (2 * Q < m) tail_lu_span<M, K2>(r, rinv, myrd, sing, Gc, tposT, lane, 2 * Q, 3 * Q < m ? 3 * Q : m);
  if (3 * Q < m) tail_lu_span<M, K3>(r, rinv, myrd, sing, Gc, tposT, lane, 3 * Q, m);
#else
  tail_lu_span<M, m>(r, rinv, myrd, sing, Gc, tposT, lane, 0, m);
#endif
  return __any_sync(FULLMASK, sing && lane < m);
}

// forward chain on the tail rows: x_i -= L(i,j) x_j, j ascending.  The matrix entries do not depend on
// the chain, so they are loaded eight columns at a time ahead of the shuffle/FMA chain.
template <class M>
__device__ __forceinline__ void tail_fwd(const double *Gc, double *Xc, const uint16_t *tposT, int lane)
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
  if (lane < m) Xc[M::HEAD + lane] = x;
}

// backward chain on the tail rows with the scaled U: x_i = b_i/d_i - sum U'(i,j) x_j, j descending
template <class M>
__device__ __forceinline__ void tail_bwd(const double *Gc, double *Xc, const uint16_t *tposT, const uint16_t *diag, int lane)
{
  constexpr int m = M::TAIL;
  double x = (lane < m) ? Xc[M::HEAD + lane] * Gc[diag[M::HEAD + (lane < m ? lane : 0)]] : 0.0;
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

// AR = auto-reduce on the full pattern (ros_yIntegrator, gckpp_Integrator.F90:789-1237): the keep mask of every cell
// comes from ar_mask_kernel; rows and columns of removed species are turned into identity rows / zero columns after
// the Jacobian round and their right-hand sides are zero, so the kept rows see exactly the operations of the reference's
// compressed system (cKppDecomp / cKppSolve, :2624-2720) and K of a removed species is 0.  The non-AR instance is the
// unchanged production kernel; to keep its parameter block (and with it its register allocation) as it was, the AR
// instance receives its two extra pointers in fields of RosArgs this kernel does not otherwise read:
//   a.work = keep mask [NVAR][ncell] (bytes), a.ar_keep_spc = (row | col << 16) of every matrix entry (uint32).
template <class M, bool AR>
__global__ void __launch_bounds__(NT, 1) ros_smem_kernel(SmemArgs P, RosArgs a)
{
  using L = Lay<M>;
  static_assert(M::NSPEC <= NT, "one owner thread per species");
  constexpr int N = M::NVAR, NA_IT = L::NA_IT, NB_IT = L::NB_IT;
  extern __shared__ __align__(16) unsigned char smem[];
  double *G = reinterpret_cast<double *>(smem + L::oG);            // [NC][NNZ+1]
  double *Yg = reinterpret_cast<double *>(smem + L::oYG);          // [NC][NYG] state under evaluation + literals + 1.0
  double *X = reinterpret_cast<double *>(smem + L::oX);            // [NC][NVAR] right-hand side / solution
#if SMEM_SCR_GLOBAL
  double *SCR = P.scr + (size_t)blockIdx.x * NC * L::NSCR;         // [NC][NSCR] A(r) or B(m), per-block global scratch
#else
  double *SCR = reinterpret_cast<double *>(smem + L::oSCR);        // [NC][NSCR] A(r) or B(m)
#endif
  double *COEF = reinterpret_cast<double *>(smem + L::oCOEF);
  double *RED = reinterpret_cast<double *>(smem + L::oRED);        // [NW][NC]
  Slot *slot = reinterpret_cast<Slot *>(smem + L::oSLOT);          // [NC]
  uint4 *RES = reinterpret_cast<uint4 *>(smem + P.s_res);          // resident chunk rows
  uint16_t *tposT = reinterpret_cast<uint16_t *>(smem + P.s_tpos);
  uint16_t *boff = reinterpret_cast<uint16_t *>(smem + P.s_boff);
  uint32_t *dir = reinterpret_cast<uint32_t *>(smem + P.s_dir);
  uint16_t *diag = reinterpret_cast<uint16_t *>(smem + P.s_diag);
  uint16_t *crow = reinterpret_cast<uint16_t *>(smem + P.s_crow);
  unsigned char *MK = smem + P.s_crow + (((M::NVAR + 1) * 2 + 15) & ~15);      // [NC][NVAR] keep flags (AR only)
  const unsigned char *ar_mask = reinterpret_cast<const unsigned char *>(a.work);
  const uint32_t *ar_rowcol = reinterpret_cast<const uint32_t *>(a.ar_keep_spc);
  unsigned mk = 0xffffffffu;                                                 // bit c: species tid of slot c is kept

  // the warp index through a broadcast: the compiler then knows it is warp-uniform and drops the divergence checks
  // (BRA.DIV / WARPSYNC) around the shuffles and barriers of the warp-conditional regions
  const int tid = threadIdx.x, lane = tid & 31, warp = __shfl_sync(FULLMASK, tid >> 5, 0);
  const RosOpts &o = a.o;
  const double Dir = (double)o.Direction;

  for (int i = tid; i < M::NCOEF; i += NT) COEF[i] = P.coefs[i];
  for (int i = tid; i < P.res_rows * 32; i += NT) RES[i] = P.resident[i];
  for (int i = tid; i < 32 * 32; i += NT) tposT[i] = P.tpos[i];
  for (int i = tid; i < P.nresb; i += NT) boff[i] = P.boff[i];
#if SMEM_CONST_DIR
  const uint32_t *dirc = c_round_dir[std::is_same<M, fullchem_dims>::value ? 0 : 1];
#else
  for (int i = tid; i < P.ndir; i += NT) dir[i] = P.dir[i];
  const uint32_t *dirc = dir;
#endif
  for (int i = tid; i < N; i += NT) diag[i] = P.diag[i];
  for (int i = tid; i <= N; i += NT) crow[i] = P.crow[i];
  for (int i = tid; i < NC * L::NYG; i += NT) {
    int k = i % L::NYG;
    Yg[i] = (k < M::NSPEC) ? 1.0 : (k < M::NSPEC + M::NLIT ? P.lit[k - M::NSPEC] : 1.0);
  }
  if (tid < NC) {
    Slot &s = slot[tid];
    s.have = 0; s.cell = -1; s.H = 1.0; s.T = 0.0; s.ghinv = 2.0; s.newstep = 0; s.skip = 1; s.sing = 0;
    s.out_cell = -1; s.in_cell = -1; s.ierr = 0; s.accept = 0;
  }

  // The rate phases are thread-private: thread t evaluates reactions t, t+NT, ... (and the partial
  // derivatives t, t+NT, ... of the Jacobian).  Their rate constants are kept in ITEM order in a
  // per-block global scratch (written by the same thread when a cell is loaded, L2 resident), and
  // the encoded terms are re-read from the table: one coalesced load latency per phase instead of
  // ~50 registers per thread.
  double *rcsA = P.rcs + (size_t)blockIdx.x * NC * (NA_IT + NB_IT) * NT;     // [NC][NA_IT][NT]
  double *rcsB = rcsA + (size_t)NC * NA_IT * NT;                               // [NC][NB_IT][NT]
  const uint2 *awt = reinterpret_cast<const uint2 *>(P.aw), *bwt = reinterpret_cast<const uint2 *>(P.bw);
  // owner registers: thread i holds species i of every slot
  double Y[NC], F0[NC], K1[NC], K2[NC], K3[NC];
#pragma unroll
  for (int c = 0; c < NC; c++) { Y[c] = 1.0; F0[c] = K1[c] = K2[c] = K3[c] = 0.0; }
  double atol_i = 1.0, rtol_i = 1.0;
  if (tid < N) {
    atol_i = o.VectorTol ? a.atol[tid] : a.atol[0];
    rtol_i = o.VectorTol ? a.rtol[tid] : a.rtol[0];
  }
  unsigned long long acc_stp = 0, acc_acc = 0, acc_fail = 0, acc_done = 0;   // meaningful in threads < NC
  bool exhausted = false;      // control threads: the work counter ran past nwork

  Reader rd;
  rd.gsrc = P.stream + (size_t)P.warp_off[warp] * 32 + lane;
  rd.ring = (uint32_t)__cvta_generic_to_shared(smem + L::oRING + warp * RS * 512 + lane * 16);
  rd.L = P.warp_rows[warp]; rd.irow = 0; rd.islot = 0; rd.cslot = 0;
  rd.prime();
  __syncthreads();
  PROF_DECL

  // a streamed round: this warp's bundles arrive in order through its ring
  auto stream_round = [&](auto opc, unsigned d) {
    const int nb = DIR_NB(d), W = DIR_W(d);
    constexpr int OPV = decltype(opc)::value;
    if (SMEM_CELL_SPLIT && (OPV == OP_SOLVE || OPV == OP_LUUPD || OPV == OP_LUDIV) && DIR_SPLIT(d)) {
      // cell-split round: W = nb * NC warps, warp b * NC + c applies bundle b to cell c
      if (warp < W) run_bundle_cell<M, (OPV == OP_SOLVE || OPV == OP_LUUPD || OPV == OP_LUDIV) ? OPV : OP_SOLVE>(rd, smem, slot, warp % NC);
      return;
    }
    if (warp < W)
      for (int b = warp; b < nb; b += W) run_bundle<M, OPV>(rd, smem, slot, SCR);
  };

  for (;;) {
    // ---- control: TimeLoop tests, retire finished cells, hand out new ones (gckpp_Integrator.F90:652-665)
    if (tid < NC) {
      Slot &s = slot[tid];
      s.out_cell = -1; s.in_cell = -1;
      auto checks = [&]() {
        bool inloop = (o.Direction > 0) ? ((s.T - o.Tend) + o.Roundoff <= 0.0) : ((o.Tend - s.T) + o.Roundoff <= 0.0);
        if (!inloop) s.ierr = 1;
        else if (s.ist[Nstp] > o.Max_no_steps) s.ierr = -6;
        else if (((s.T + 0.1 * s.H) == s.T) || (s.H <= o.Roundoff)) s.ierr = -7;
        else s.H = fmin(s.H, fabs(o.Tend - s.T));
      };
      if (s.have && s.newstep && s.ierr == 0) checks();
      if (s.have && s.ierr != 0) {
        s.out_cell = s.cell; s.out_ierr = s.ierr;
        for (int q = 0; q < 8; q++) s.out_ist[q] = s.ist[q];
        s.out_r[0] = s.Texit; s.out_r[1] = s.Hexit; s.out_r[2] = s.Hnew;
        acc_stp += s.ist[Nstp]; acc_acc += s.ist[Nacc]; acc_done++;
        if (s.ierr < 0) acc_fail++;
        s.have = 0; s.ierr = 0; s.cell = -1; s.H = 1.0; s.T = 0.0;
      }
      if (!s.have && !exhausted) {
        int w = atomicAdd(a.next, 1);
        if (w >= a.nwork) {
          exhausted = true;
        } else {
          int cell = a.cell_list ? a.cell_list[w] : w;
          s.cell = cell; s.in_cell = cell;
          for (int q = 0; q < 8; q++) s.ist[q] = 0;
          double hs = a.hstart ? a.hstart[cell] : o.Hstart_rcntrl;     // Integrate's merge + Rosenbrock :420-428
          double Hstart = (hs > 0.0) ? fmin(fabs(hs), fabs(o.Tend - o.Tstart)) : fmax(o.Hmin, 1.0E-5);
          s.T = o.Tstart; s.Hexit = 0.0; s.Hnew = 0.0; s.Texit = 0.0;
          double H = fmin(fmax(fabs(o.Hmin), fabs(Hstart)), fabs(o.Hmax));   // :637
          if (fabs(H) <= 10.0 * o.Roundoff) H = 1.0E-5;
          s.H = Dir * H;
          s.rejLast = 0; s.rejMore = 0; s.have = 1; s.newstep = 1; s.nconsec = 0; s.ierr = 0;
          checks();             // a cell that fails here idles through this attempt and is retired next
        }
      }
      s.skip = !s.have || s.ierr != 0;
      s.sing = 0;
      s.accept = 0;
      s.ghinv = 1.0 / (Dir * s.H * o.Gamma[0]);
#pragma unroll
      for (int k = 0; k < 6; k++) s.hc[k] = o.C[k] / (Dir * s.H);
      if (!s.skip && s.newstep) {
        s.ist[Nfun]++;
        if (!o.Autonomous) s.ist[Nfun]++;
        s.ist[Njac]++;
        s.nconsec = 0;
      }
    }
    __syncthreads();
    bool any = false;
#pragma unroll
    for (int c = 0; c < NC; c++) {
      const int oc = slot[c].out_cell, ic = slot[c].in_cell;
      if (oc >= 0) {
        if (tid < M::NSPEC) a.conc_out[(size_t)tid * a.ncell + oc] = Y[c];
        if (tid < 8 && a.istatus) a.istatus[(size_t)tid * a.ncell + oc] = slot[c].out_ist[tid];
        if (tid >= 32 && tid < 35 && a.rstatus) a.rstatus[(size_t)(tid - 32) * a.ncell + oc] = slot[c].out_r[tid - 32];
        if (!AR && tid == 35 && a.rstatus) a.rstatus[(size_t)3 * a.ncell + oc] = 0.0;     // AR: NARthr was stored by ar_mask_kernel
        if (tid == 36 && a.ierr) a.ierr[oc] = slot[c].out_ierr;
      }
      if (ic >= 0) {
        if (tid < M::NSPEC) Y[c] = a.conc_in[(size_t)tid * a.ncell + ic];
        if (AR && tid < N) {
          const unsigned char v = ar_mask[(size_t)tid * a.ncell + ic];
          MK[c * N + tid] = v;
          mk = (mk & ~(1u << c)) | (v ? (1u << c) : 0u);
        }
#pragma unroll
        for (int k = 0; k < NA_IT; k++) {
          const int r = k * NT + tid;
          if (r < M::NREACT) {
            const int i0 = awt[r].x & 0xffff;
            rcsA[(c * NA_IT + k) * NT + tid] = i0 < M::NREACT ? a.rconst[(size_t)i0 * a.rc_stride + (ic - a.rc_cell0)] : P.lit[i0 - M::NREACT];
          }
        }
#pragma unroll
        for (int k = 0; k < NB_IT; k++) {
          const int m = k * NT + tid;
          if (m < M::NB) {
            const int i0 = bwt[m].x & 0xffff;
            rcsB[(c * NB_IT + k) * NT + tid] = i0 < M::NREACT ? a.rconst[(size_t)i0 * a.rc_stride + (ic - a.rc_cell0)] : P.lit[i0 - M::NREACT];
          }
        }
      } else if (!slot[c].have) {
        if (tid < M::NSPEC) Y[c] = 1.0;      // idle slot: benign numbers
      }
      any |= (slot[c].have != 0);
    }
    if (!any) break;
    PROF(0);
    // Three function evaluations per attempt (Rodas3: NewF = T,F,T,T; :691-724): Fcn0 (followed by the
    // Jacobian, the LU and stages 1-2), stage 3, stage 4.  One copy of the Fun and solve code.
#pragma unroll 1
    for (int ip = 0; ip < 3; ip++) {
      // ---- the state to evaluate
      if (ip == 0) {
        if (tid < M::NSPEC) {
#pragma unroll
          for (int c = 0; c < NC; c++) Yg[c * L::NYG + tid] = Y[c];
        }
      } else if (tid < N) {
#pragma unroll
        for (int c = 0; c < NC; c++)
          Yg[c * L::NYG + tid] = (ip == 1) ? fma(o.A[2], K2[c], fma(o.A[1], K1[c], Y[c]))
                                           : fma(o.A[5], K3[c], fma(o.A[4], K2[c], fma(o.A[3], K1[c], Y[c])));
      }
      __syncthreads();
      // ---- X = Fun(Yg): A(r) = RCT(r) * prod(V) by the thread that owns reaction r, then the vdot round
      {
        uint2 w[NA_IT];
        double rc[NA_IT][NC];
#pragma unroll
        for (int k = 0; k < NA_IT; k++) {
          const int r = k * NT + tid;
          w[k] = r < M::NREACT ? awt[r] : make_uint2(0, 0);
#pragma unroll
          for (int c = 0; c < NC; c++) rc[k][c] = r < M::NREACT ? rcsA[(c * NA_IT + k) * NT + tid] : 0.0;
        }
#pragma unroll
        for (int k = 0; k < NA_IT; k++) {
          const int r = k * NT + tid;
          if (r < M::NREACT) {
            const int i1 = w[k].x >> 16, i2 = w[k].y & 0xffff, i3 = w[k].y >> 16;
#pragma unroll
            for (int c = 0; c < NC; c++)
              SCR[c * L::NSCR + r] = rc[k][c] * Yg[c * L::NYG + i1] * Yg[c * L::NYG + i2] * Yg[c * L::NYG + i3];
          }
        }
      }
      __syncthreads();
      stream_round(std::integral_constant<int, OP_VDOT>(), dirc[0]);
      __syncthreads();
      PROF(1);
      if (ip == 0) {
        if (tid < N) {
#pragma unroll
          for (int c = 0; c < NC; c++) F0[c] = X[c * N + tid];
        }
        // ---- Ghimj = 1/(H*gamma) - Jac0  (:1973-1977); Jac0 is recomputed per attempt
        {
          uint2 w[NB_IT];
          double rc[NB_IT][NC];
#pragma unroll
          for (int k = 0; k < NB_IT; k++) {
            const int m = k * NT + tid;
            w[k] = m < M::NB ? bwt[m] : make_uint2(0, 0);
#pragma unroll
            for (int c = 0; c < NC; c++) rc[k][c] = m < M::NB ? rcsB[(c * NB_IT + k) * NT + tid] : 0.0;
          }
#pragma unroll
          for (int k = 0; k < NB_IT; k++) {
            const int m = k * NT + tid;
            if (m < M::NB) {
              const int i1 = w[k].x >> 16, i2 = w[k].y & 0xffff, i3 = w[k].y >> 16;
#pragma unroll
              for (int c = 0; c < NC; c++)
                SCR[c * L::NSCR + m] = rc[k][c] * Yg[c * L::NYG + i1] * Yg[c * L::NYG + i2] * Yg[c * L::NYG + i3];
            }
          }
        }
        for (int i = tid; i < NC * L::GS; i += NT) G[i] = 0.0;        // structural zeros / fill-in slots / zero slot
        __syncthreads();
        stream_round(std::integral_constant<int, OP_JVS>(), dirc[1]);
        __syncthreads();
        if (AR) {
          for (int k = tid; k < M::NNZ; k += NT) {
            const uint32_t rcw = __ldg(ar_rowcol + k);
            const int row = rcw & 0xffff, col = rcw >> 16;
#pragma unroll
            for (int c = 0; c < NC; c++) {
              const bool kr = MK[c * N + row] != 0, kc = MK[c * N + col] != 0;
              if (!kr || !kc) G[c * L::GS + k] = (!kr && row == col) ? 1.0 : 0.0;
            }
          }
          __syncthreads();
        }
        PROF(2);
        // ---- sparse LU  (KppDecomp): head pivots by DAG level, then the tail block
#pragma unroll 1
        for (int r = 0; r < P.n_lu; r++) {
          const unsigned d = dirc[P.o_lu + r];
          if (DIR_DIV(d)) stream_round(std::integral_constant<int, OP_LUDIV>(), d);
          else stream_round(std::integral_constant<int, OP_LUUPD>(), d);
          round_barrier(DIR_P(d), warp);
#ifdef SMEM_PROFILE
          { long long t_ = clock64(); if (r < 24) prl_[r] += t_ - pt_; pt_ = t_; }
#endif
        }
        PROF(3);
        auto post_lu_row = [&](int i) {     // singular test (:1985), reciprocal diagonal, U row scaled by it
          const int dp = diag[i], e = crow[i + 1];
#pragma unroll
          for (int c = 0; c < NC; c++) {
            double rdv = G[c * L::GS + dp];
            if (!M::RECIP_DIAG) {
              if (!(fabs(rdv) >= DBL_MIN)) slot[c].sing = 1;      // also catches NaN from an earlier zero pivot
              rdv = 1.0 / rdv;
              G[c * L::GS + dp] = rdv;
            }
            for (int p = dp + 1; p < e; p++) G[c * L::GS + p] *= rdv;
          }
        };
#if SMEM_OVERLAP
        // Warps 0..NC-1 factorise the tail block; meanwhile the other warps scale the head rows and run
        // the forward head rounds of stage 1 (right-hand side Fcn0): those only read L entries of head
        // columns, which are final, and never touch the tail block.
        if (tid < N) {
#pragma unroll
          for (int c = 0; c < NC; c++) X[c * N + tid] = (AR && !((mk >> c) & 1u)) ? 0.0 : F0[c];
        }
        __syncthreads();
        if (warp < NC) {
          if (tail_lu<M>(G + warp * L::GS, tposT, lane)) slot[warp].sing = 1;
        } else {
          // head rows: reciprocal diagonals first, then every U entry of the head rows scaled by it, spread
          // evenly over the threads (a thread per row would wait for the longest row)
          if (!M::RECIP_DIAG) {
            for (int i = tid - NC * 32; i < M::HEAD; i += NT - NC * 32) {
              const int dp = diag[i];
#pragma unroll
              for (int c = 0; c < NC; c++) {
                const double d = G[c * L::GS + dp];
                if (!(fabs(d) >= DBL_MIN)) slot[c].sing = 1;
                G[c * L::GS + dp] = 1.0 / d;
              }
            }
            asm volatile("bar.sync 1, %0;" :: "n"((NW - NC) * 32) : "memory");
          }
          for (int q = tid - NC * 32; q < P.nuscale; q += NT - NC * 32) {
            const unsigned w = __ldg(P.uscale + q);
            const int p = w >> 16, dp = w & 0xffff;
#pragma unroll
            for (int c = 0; c < NC; c++) G[c * L::GS + p] *= G[c * L::GS + dp];
          }
#pragma unroll 1
          for (int r = 0; r < P.n_fwd; r++) {
            const unsigned d = dirc[P.o_fwd1 + r];
            const int nb = DIR_NB(d), W = DIR_W(d), wv = warp - NC;
            if (wv < W)
              for (int b = wv; b < nb; b += W) run_bundle<M, OP_SOLVE>(rd, smem, slot, SCR);
            asm volatile("bar.sync 1, %0;" :: "n"((NW - NC) * 32) : "memory");   // id 1: classes use ids 2..NW-1
          }
        }
        PROF(4);        // the tail rows leave tail_lu in their final form; the stage loop's barrier orders everything
#else
        if (warp < NC && tail_lu<M>(G + warp * L::GS, tposT, lane)) slot[warp].sing = 1;
        __syncthreads();
        PROF(4);
        if (tid < M::HEAD) post_lu_row(tid);
#endif
        PROF(5);
      }
      // ---- the stages that use this evaluation: ip 0 -> stages 1 and 2, ip 1 -> stage 3, ip 2 -> stage 4
      const int nsolve = (ip == 0) ? 2 : 1;
#pragma unroll 1
      for (int q = 0; q < nsolve; q++) {
        const int st = (ip == 0) ? q : ip + 1;
        const bool fwd_done = SMEM_OVERLAP && st == 0;       // stage 1: forward head rounds already ran beside tail_lu
        if (tid < N && !fwd_done) {          // right-hand side K_st = Fcn + sum_j C(st,j)/H K_j
#pragma unroll
          for (int c = 0; c < NC; c++) {
            double v;
            if (st == 0) v = F0[c];
            else if (st == 1) v = fma(slot[c].hc[0], K1[c], F0[c]);
            else if (st == 2) v = fma(slot[c].hc[2], K2[c], fma(slot[c].hc[1], K1[c], X[c * N + tid]));
            else v = fma(slot[c].hc[5], K3[c], fma(slot[c].hc[4], K2[c], fma(slot[c].hc[3], K1[c], X[c * N + tid])));
            if (AR && !((mk >> c) & 1u)) v = 0.0;
            X[c * N + tid] = v;
          }
        }
        __syncthreads();
        // ---- KppSolve on X in place: head of L (rounds), tail chains, head of U (rounds).
        // The tables do not depend on the data, so the directory entry and the first two chunks of
        // the NEXT round are fetched before waiting on the barrier of the current one.
        const int n_tot = P.n_fwd + P.n_bwd;
        auto tails = [&]() {
          if (warp < NC) tail_fwd<M>(G + warp * L::GS, X + warp * N, tposT, lane);
          __syncthreads();
          if (warp < NC) {
            tail_bwd<M>(G + warp * L::GS, X + warp * N, tposT, diag, lane);
          } else {
            for (int i = tid - NC * 32; i < M::HEAD; i += NT - NC * 32) {
              const int dp = diag[i];
#pragma unroll
              for (int c = 0; c < NC; c++) X[c * N + i] *= G[c * L::GS + dp];
            }
          }
          __syncthreads();
          PROF(11);
        };
        PROF(6);
#if SMEM_SWEEP_RESIDENT
        auto prefetch = [&](int r, unsigned &d, PreReader &pr) {
          d = 0; pr.n = 0; pr.p = RES + lane;
          if (r < n_tot) {
            d = dirc[P.o_fwd + r];
            if (warp < DIR_W(d)) {
              const uint4 *p = RES + (size_t)boff[DIR_BF(d) + warp] * 32 + lane;
              pr.c0 = p[0]; pr.c1 = p[32]; pr.p = p + 64;
            }
          }
        };
        unsigned dn; PreReader pn;
        prefetch(0, dn, pn);
#pragma unroll 1
        for (int r = 0;; r++) {
          if (r == P.n_fwd) tails();
          if (r == n_tot) break;
          const unsigned d = dn;
          const int nb = DIR_NB(d), W = DIR_W(d), bf = DIR_BF(d);
          if (warp < W) {
            run_bundle<M, OP_SOLVE>(pn, smem, slot, SCR);
            for (int b = warp + W; b < nb; b += W) {       // only if a round has more bundles than warps
              ResReader rr;
              rr.p = RES + (size_t)boff[bf + b] * 32 + lane;
              run_bundle<M, OP_SOLVE>(rr, smem, slot, SCR);
            }
          }
          PROF(8);
          prefetch(r + 1, dn, pn);
          PROF(9);
          round_barrier(DIR_P(d), warp);
          PROF(10);
        }
#else
#pragma unroll 1
        for (int r = fwd_done ? P.n_fwd : 0;; r++) {
          if (r == P.n_fwd) tails();
          if (r == n_tot) break;
          const unsigned d = dirc[P.o_fwd + r];
#ifdef SMEM_PROFILE
          if (!DIR_SPLIT(d)) {
            const int nb = DIR_NB(d), W = DIR_W(d);
            if (warp < W)
              for (int b = warp; b < nb; b += W) run_bundle<M, OP_SOLVE>(rd, smem, slot, SCR, (tid == 0 && blockIdx.x == 0) ? pacc_ : nullptr);
          } else stream_round(std::integral_constant<int, OP_SOLVE>(), d);
#else
          stream_round(std::integral_constant<int, OP_SOLVE>(), d);
#endif
          PROF(8);
          round_barrier(DIR_P(d), warp);
#ifdef SMEM_PROFILE
          { long long t_ = clock64(); if (r < 24) prs_[r] += t_ - pt_; }
#endif
          PROF(10);
        }
#endif
        if (tid < N) {
#pragma unroll
          for (int c = 0; c < NC; c++) {
            const double v = X[c * N + tid];
            if (st == 0) K1[c] = v;
            else if (st == 1) K2[c] = v;
            else if (st == 2) K3[c] = v;
          }
        }
      }
    }
    // ---- new solution, error estimate and norm  (:729-740, :1715-1745); K4 is read from X
    double yn[NC], e2[NC];
#pragma unroll
    for (int c = 0; c < NC; c++) { yn[c] = 0.0; e2[c] = 0.0; }
    if (tid < N) {
#pragma unroll
      for (int c = 0; c < NC; c++) {
        const double k4 = X[c * N + tid];
        yn[c] = fma(o.M[3], k4, fma(o.M[2], K3[c], fma(o.M[1], K2[c], fma(o.M[0], K1[c], Y[c]))));
        const double ye = fma(o.E[3], k4, fma(o.E[2], K3[c], fma(o.E[1], K2[c], o.E[0] * K1[c])));
        const double sc = atol_i + rtol_i * fmax(fabs(Y[c]), fabs(yn[c]));
        const double q = ye / sc;
        e2[c] = q * q;
      }
    }
#pragma unroll
    for (int c = 0; c < NC; c++) {
      for (int off = 16; off > 0; off >>= 1) e2[c] += __shfl_down_sync(FULLMASK, e2[c], off);
      if (lane == 0) RED[warp * NC + c] = e2[c];
    }
    __syncthreads();
    // ---- accept / reject (:743-777), one control thread per slot
    if (tid < NC) {
      Slot &s = slot[tid];
      if (!s.skip) {
        s.ist[Ndec]++;
        if (s.sing) {                           // ros_PrepareMatrix :1985-1995
          s.ist[Nsng]++;
          s.nconsec++;
          if (s.nconsec <= 5) { s.H *= 0.5; s.newstep = 0; }
          else s.ierr = -8;
        } else {
          s.nconsec = 0;
          double e = 0.0;
          for (int w = 0; w < NW; w++) e += RED[w * NC + tid];
          const double Err = fmax(sqrt(e / (double)N), 1.0e-10);
          s.ist[Nfun] += 2; s.ist[Nsol] += 4;
          const double H = s.H;
          const double Fac = fmin(o.FacMax, fmax(o.FacMin, o.FacSafe / pow(Err, 1.0 / o.ELO)));
          double Hnew = H * Fac;
          s.ist[Nstp]++;
          if ((Err <= 1.0) || (H <= o.Hmin)) {
            s.ist[Nacc]++;
            s.accept = 1;
            s.T = s.T + Dir * H;
            Hnew = fmax(o.Hmin, fmin(Hnew, o.Hmax));
            if (s.rejLast) Hnew = fmin(Hnew, H);
            s.Hexit = H; s.Hnew = Hnew; s.Texit = s.T;
            s.rejLast = 0; s.rejMore = 0;
            s.H = Hnew;
            s.newstep = 1;
          } else {
            if (s.rejMore) Hnew = H * o.FacRej;
            s.rejMore = s.rejLast;
            s.rejLast = 1;
            s.H = Hnew;
            if (s.ist[Nacc] >= 1) s.ist[Nrej]++;
            s.newstep = 0;
          }
        }
      }
    }
    __syncthreads();
    if (tid < N) {
#pragma unroll
      for (int c = 0; c < NC; c++)
        if (slot[c].accept) Y[c] = o.ClipNegative ? fmax(yn[c], 0.0) : yn[c];
    }
    __syncthreads();      // the control threads rewrite slot[] at the top of the loop
    PROF(7);
  }
#ifdef SMEM_PROFILE
  if (tid == 0 && blockIdx.x == 0 && a.sums)
  {
    for (int i = 0; i < 16; i++) a.sums[8 + i] = (unsigned long long)pacc_[i];
    for (int i = 0; i < 24; i++) { a.sums[24 + i] = (unsigned long long)prl_[i]; a.sums[48 + i] = (unsigned long long)prs_[i]; }
  }
#endif
  if (tid < NC && a.sums) {
    atomicAdd(a.sums + 0, acc_stp);
    atomicAdd(a.sums + 1, acc_acc);
    atomicAdd(a.sums + 2, acc_fail);
    atomicAdd(a.sums + 3, acc_done);
  }
  asm volatile("cp.async.wait_all;");
}

}  // namespace

// ---- host side ---------------------------------------------------------------------------------------

// Encode a rate/partial term (tables.h: rate, f0, f1, f2) as two words of 16-bit indices:
//   rate index into [RCONST, literals], factor indices into Yg = [VAR, FIX, literals, 1.0].
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

template <class M> static int dyn_offset() { return Lay<M>::oDYN; }
template <class M> static bool dims_match(const gckpp_host_tables_t *T, const gckpp_sched_tables_t *S)
{
  return T->nvar == M::NVAR && T->nspec == M::NSPEC && T->nreact == M::NREACT && T->nnz == M::NNZ && T->nb == M::NB &&
         T->nlit == M::NLIT && S->ncoef == M::NCOEF && S->tail == M::TAIL && S->head == M::HEAD;
}

template <class M> static size_t scr_doubles() { return (size_t)NC * Lay<M>::NSCR; }
size_t smem_scr_doubles_per_block(int mech_id)
{
  return mech_id == GCKPP_MECH_FULLCHEM ? scr_doubles<fullchem_dims>() : scr_doubles<Hg_dims>();
}
template <class M> static size_t rcs_doubles() { return (size_t)NC * (Lay<M>::NA_IT + Lay<M>::NB_IT) * NT; }
size_t smem_rcs_doubles_per_block(int mech_id)
{
  return mech_id == GCKPP_MECH_FULLCHEM ? rcs_doubles<fullchem_dims>() : rcs_doubles<Hg_dims>();
}

bool smem_kernel_supports(int mech_id) { return mech_id == GCKPP_MECH_FULLCHEM || mech_id == GCKPP_MECH_HG; }

int smem_plan_build(int mech_id, const gckpp_host_tables_t *T, const gckpp_sched_tables_t *S, SmemHostPlan &hp)
{
  if (!smem_kernel_supports(mech_id)) return -1;
  if (mech_id == GCKPP_MECH_FULLCHEM ? !dims_match<fullchem_dims>(T, S) : !dims_match<Hg_dims>(T, S)) return -2;
  const int *ph = S->phase;
  auto r0 = [&](int p) { return ph[2 * p]; };
  auto r1 = [&](int p) { return ph[2 * p + 1]; };
  if (r1(0) - r0(0) != 1 || r1(1) - r0(1) != 1) return -3;
  auto nbundles = [&](int r) { return (int)(S->rounds[3 * r + 1] - S->rounds[3 * r]); };
  // cell-split rounds (SMEM_CELL_SPLIT): at most NW / NC bundles, sweep rounds (phases 4, 5) and / or LU rounds (phase 2)
  auto split = [&](int r) {
    if (!SMEM_CELL_SPLIT || SMEM_SWEEP_RESIDENT || nbundles(r) * NC > NW) return false;
    const bool sweep = (r >= r0(4) && r < r1(4)) || (r >= r0(5) && r < r1(5)), lu = r >= r0(2) && r < r1(2);
    return (sweep && (SMEM_CELL_SPLIT & 1)) || (lu && (SMEM_CELL_SPLIT & 2));
  };
  auto nwarps = [&](int r) { int nb = nbundles(r); return split(r) ? nb * NC : (nb < NW ? nb : NW); };
  // directory: vdot, jvs, lu..., fwd..., bwd...
  std::vector<int> dr;
  dr.push_back(r0(0)); dr.push_back(r0(1));
  hp.o_lu = (int)dr.size(); hp.n_lu = r1(2) - r0(2);
  for (int r = r0(2); r < r1(2); r++) dr.push_back(r);
  hp.o_fwd = (int)dr.size(); hp.n_fwd = r1(4) - r0(4);
  for (int r = r0(4); r < r1(4); r++) dr.push_back(r);
  hp.o_bwd = (int)dr.size(); hp.n_bwd = r1(5) - r0(5);
  for (int r = r0(5); r < r1(5); r++) dr.push_back(r);
  // second copy of the forward rounds, scheduled on warps NC..NW-1 (they run beside the tail factorisation)
  hp.o_fwd1 = (int)dr.size();
  const int n_shift = SMEM_OVERLAP ? hp.n_fwd : 0;
  for (int r = r0(4); r < r0(4) + n_shift; r++) dr.push_back(r);
  // resident bundles: all bundles of the fwd and bwd rounds, in directory order
  hp.resident.clear(); hp.boff.clear();
  std::vector<int> bfirst(dr.size(), 0);
  for (size_t i = (size_t)hp.o_fwd; SMEM_SWEEP_RESIDENT && i < dr.size(); i++) {
    int r = dr[i];
    bfirst[i] = (int)hp.boff.size();
    for (uint32_t b = S->rounds[3 * r]; b < S->rounds[3 * r + 1]; b++) {
      hp.boff.push_back((uint16_t)(hp.resident.size() / 128));
      for (uint32_t row = S->brow[b]; row < S->brow[b + 1]; row++)
        hp.resident.insert(hp.resident.end(), S->chunks + (size_t)row * 128, S->chunks + (size_t)(row + 1) * 128);
    }
  }
  hp.resident.insert(hp.resident.end(), 128, 0u);      // the kernel prefetches one chunk row past a bundle
  if (hp.boff.size() >= 1024 || hp.resident.size() / 128 >= 65536) return -4;
  hp.dir.assign(dr.size(), 0);
  for (size_t i = 0; i < dr.size(); i++) {
    int r = dr[i], nb = nbundles(r), W = nwarps(r);
    if (nb >= 2048) return -4;
    if ((int)i >= hp.o_fwd1) {            // warp-shifted rounds: at most NW-NC warps, their own barrier
      int Ws = nb < NW - NC ? nb : NW - NC;
      hp.dir[i] = DIR_PACK(nb, Ws, 0, 0, 0);
      continue;
    }
    bool last = ((int)i + 1 == hp.o_fwd1) || (int)i + 1 == hp.o_lu || (int)i + 1 == hp.o_fwd || (int)i + 1 == hp.o_bwd || i < 2;
    int Wn = last ? NW : nwarps(dr[i + 1]);
    int Pb = last ? NW : (W > Wn ? W : Wn);
    uint32_t div = (S->rounds[3 * r + 2] & 0x10) ? 1u : 0u;
    hp.dir[i] = DIR_PACK(nb, W, Pb, div, bfirst[i]) | (split(r) ? 0x80000000u : 0u);
  }
  // per-warp streams in the order one Rodas3 attempt consumes them: vdot, jvs, lu rounds, [sweeps x2], vdot,
  // [sweeps], vdot, [sweeps]  (the sweeps only when their tables are not resident)
  std::vector<std::pair<int, int>> order;        // (round, first warp)
  auto push_sweeps = [&](bool shifted_fwd) {
    if (SMEM_SWEEP_RESIDENT) return;
    for (int r = r0(4); r < r1(4); r++) order.push_back({r, shifted_fwd ? NC : 0});
    for (int r = r0(5); r < r1(5); r++) order.push_back({r, 0});
  };
  order.push_back({r0(0), 0}); order.push_back({r0(1), 0});
  for (int r = r0(2); r < r1(2); r++) order.push_back({r, 0});
  push_sweeps(SMEM_OVERLAP); push_sweeps(false);
  order.push_back({r0(0), 0}); push_sweeps(false);
  order.push_back({r0(0), 0}); push_sweeps(false);
  std::vector<std::vector<uint32_t>> ws(NW);
  for (auto &rw : order) {
    int r = rw.first, w0 = rw.second;
    uint32_t b0 = S->rounds[3 * r], b1 = S->rounds[3 * r + 1];
    if (w0 == 0 && split(r)) {           // every bundle goes to NC warps (the shifted copy of the forward rounds is never split)
      for (uint32_t b = b0; b < b1; b++)
        for (int c = 0; c < NC; c++)
          for (uint32_t row = S->brow[b]; row < S->brow[b + 1]; row++)
            ws[(b - b0) * NC + c].insert(ws[(b - b0) * NC + c].end(), S->chunks + (size_t)row * 128, S->chunks + (size_t)(row + 1) * 128);
      continue;
    }
    int W = split(r) ? (nbundles(r) < NW ? nbundles(r) : NW) : nwarps(r);
    if (w0 > 0 && W > NW - w0) W = NW - w0;
    for (uint32_t b = b0; b < b1; b++) {
      int w = w0 + (int)((b - b0) % (uint32_t)W);
      for (uint32_t row = S->brow[b]; row < S->brow[b + 1]; row++)
        ws[w].insert(ws[w].end(), S->chunks + (size_t)row * 128, S->chunks + (size_t)(row + 1) * 128);
    }
  }
  hp.stream.clear();
  for (int w = 0; w < NW; w++) {
    // The ring needs a stream of at least 2*RS rows.  A short stream is repeated whole (it is cyclic, so
    // a multiple of the period reads the same); a warp without any work only ever prefetches zeros.
    if (ws[w].empty()) ws[w].assign((size_t)128 * 2 * RS, 0u);
    const std::vector<uint32_t> period = ws[w];
    while ((int)ws[w].size() < 128 * 2 * RS) ws[w].insert(ws[w].end(), period.begin(), period.end());
    hp.warp_off[w] = (int)(hp.stream.size() / 128);
    hp.warp_rows[w] = (int)(ws[w].size() / 128);
    hp.stream.insert(hp.stream.end(), ws[w].begin(), ws[w].end());
  }
  hp.aw.resize(2 * (size_t)T->nreact);
  for (int r = 0; r < T->nreact; r++) encode_term(T->a_term + 4 * r, T->nreact, T->nspec, T->nlit, &hp.aw[2 * r]);
  hp.bw.resize(2 * (size_t)(T->nb > 0 ? T->nb : 1));
  for (int m = 0; m < T->nb; m++) encode_term(T->b_term + 4 * m, T->nreact, T->nspec, T->nlit, &hp.bw[2 * m]);
  hp.uscale.clear();
  for (int i = 0; i < S->head; i++)
    for (int p = T->diag[i] + 1; p < T->crow[i + 1]; p++) hp.uscale.push_back(((uint32_t)p << 16) | (uint32_t)T->diag[i]);
  hp.diag.resize(T->nvar); hp.crow.resize(T->nvar + 1);
  for (int i = 0; i < T->nvar; i++) hp.diag[i] = (uint16_t)T->diag[i];
  for (int i = 0; i <= T->nvar; i++) hp.crow[i] = (uint16_t)T->crow[i];
  if (T->nspec + T->nlit + 1 >= 65536 || T->nreact + T->nlit >= 65535 || (T->nnz + 1) * 8 >= 65536) return -4;
  int off = mech_id == GCKPP_MECH_FULLCHEM ? dyn_offset<fullchem_dims>() : dyn_offset<Hg_dims>();
  auto take = [&](size_t bytes) { int o = off; off = (int)((off + bytes + 15) & ~(size_t)15); return o; };
  hp.s_res = take(hp.resident.size() * 4);
  hp.s_tpos = take(32 * 32 * 2);
  hp.s_boff = take(hp.boff.size() * 2 + 2);
  hp.s_dir = take(hp.dir.size() * 4);
  hp.s_diag = take(hp.diag.size() * 2);
  hp.s_crow = take(hp.crow.size() * 2);
  take((size_t)NC * T->nvar);                  // keep flags of the auto-reduce instance, right behind crow
  hp.s_total = off;
  hp.rowcol.resize(T->nnz);
  for (int i = 0; i < T->nvar; i++)
    for (int k = T->crow[i]; k < T->crow[i + 1]; k++) hp.rowcol[k] = (uint32_t)i | ((uint32_t)T->icol[k] << 16);
  return 0;
}

template <class M, bool AR>
static cudaError_t launch_t(const SmemArgs &P, const RosArgs &a, int blocks, cudaStream_t s)
{
  auto k = ros_smem_kernel<M, AR>;
  // the opt-in is a per-DEVICE attribute of the function (one process may hold handles on several GPUs): set it on
  // every launch -- it is cheap -- rather than cache it per process
  cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, P.s_total);
  if (e != cudaSuccess) return e;
  k<<<blocks, NT, P.s_total, s>>>(P, a);
  return cudaGetLastError();
}

cudaError_t smem_set_directory(int mech_id, const uint32_t *dir, int n, cudaStream_t s)
{
#if SMEM_CONST_DIR
  if (n > 128 || (mech_id != GCKPP_MECH_FULLCHEM && mech_id != GCKPP_MECH_HG)) return cudaErrorInvalidValue;
  return cudaMemcpyToSymbolAsync(c_round_dir, dir, sizeof(uint32_t) * (size_t)n, sizeof(uint32_t) * 128 * (mech_id == GCKPP_MECH_FULLCHEM ? 0 : 1),
                                 cudaMemcpyHostToDevice, s);
#else
  (void)mech_id; (void)dir; (void)n; (void)s;
  return cudaSuccess;
#endif
}

cudaError_t launch_ros_smem(int mech_id, const SmemArgs &P, const RosArgs &a, int blocks, cudaStream_t s, bool autoreduce)
{
  if (mech_id == GCKPP_MECH_FULLCHEM && autoreduce) return launch_t<fullchem_dims, true>(P, a, blocks, s);
  if (autoreduce) return cudaErrorInvalidConfiguration;
  if (mech_id == GCKPP_MECH_FULLCHEM) return launch_t<fullchem_dims, false>(P, a, blocks, s);
  if (mech_id == GCKPP_MECH_HG) return launch_t<Hg_dims, false>(P, a, blocks, s);
  return cudaErrorInvalidConfiguration;
}
