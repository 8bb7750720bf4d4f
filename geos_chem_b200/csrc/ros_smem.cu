// Shared-memory-resident Rosenbrock (Rodas3) kernel -- the production path on B200.
//
// Reference routines covered (KPP/fullchem/..., identical structure for KPP/Hg):
//   ros_Integrator      gckpp_Integrator.F90:578-786      -> ros_smem_kernel (stage loop, error control)
//   ros_PrepareMatrix   gckpp_Integrator.F90:1921-1999    -> Jacobian rounds + LU rounds + singular test
//   ros_ErrorNorm       gckpp_Integrator.F90:1715-1745    -> block reduction per cell
//   Fun                 gckpp_Function.F90:51-2152        -> rate phase + "vdot" round (aggregate form)
//   Jac_SP              gckpp_Jacobian.F90:48-20887       -> partials phase + "jvs" round
//   KppDecomp           gckpp_LinearAlgebra.F90:46-83     -> "lu" rounds (right-looking, DAG levels)
//   KppSolve            gckpp_LinearAlgebra.F90:644-2309  -> "fwd"/"bwd" rounds (push form)
//
// Execution model
//   * One persistent thread block (NW warps) per SM integrates NC cells in LOCK STEP: every block
//     iteration is one Rosenbrock attempt for each of its NC cell slots.  A slot whose cell reaches
//     Tend (or fails) stores its result and pulls the next cell from a global counter, so the
//     per-cell adaptive step counts never idle a slot until the grid is exhausted.
//   * Everything an attempt touches more than once lives on chip.  Shared memory (per cell): the
//     sparse matrix G (LU_NONZERO doubles), the state being evaluated Yg, the right-hand side X,
//     the rate / partial-derivative scratch SCR.  Registers: thread i owns species i of every slot
//     (Y, Fcn0, K1..K4) and the rate constants of "its" reactions (the rate phase is thread-private).
//   * The sparse kernels are table driven (kppgen/sched.py): a round is a set of bundles of 32 lane
//     items; each table word is applied to all NC cells by the thread that fetched it, which is what
//     amortises the index traffic.  The tables are laid out per warp in exact consumption order
//     ("stream"), cyclic over attempts, and prefetched with cp.async into a per-warp shared-memory
//     ring RING rows ahead -- table latency never sits on the dependency chain of a round.
//   * Barriers: a round that keeps P < NW warps busy synchronises only those warps (named barrier P);
//     runs of single-bundle rounds (the dense tail of the elimination DAG) use __syncwarp only.
//
// Arithmetic: FP64 throughout, FMA contraction allowed, sums re-associated (see sched.py).  The
// diagonal of the factors is stored as its reciprocal and U rows are pre-scaled by it, so the four
// solves of an attempt contain no division.
#include <float.h>
#include <math.h>
#include <string.h>
#include <vector>
#include "ros_common.cuh"
#include "ros_smem.h"

// -DSMEM_PROFILE: thread 0 of block 0 accumulates clock64() per phase into sums[8..]
#ifdef SMEM_PROFILE
#define PROF_DECL long long pt_ = clock64(), pacc_[12] = {0,0,0,0,0,0,0,0,0,0,0,0};
#define PROF(i) do { long long t_ = clock64(); pacc_[i] += t_ - pt_; pt_ = t_; } while (0)
#else
#define PROF_DECL
#define PROF(i) do { } while (0)
#endif

namespace {

constexpr int RING = 16;             // rows of 32 words per warp in the prefetch ring
constexpr int KCH = 4;               // table rows consumed per fetch

enum { OP_VDOT, OP_JVS, OP_LUDIV, OP_LUUPD, OP_SCALE, OP_SOLVE };

struct Slot {
  double T, H, Hexit, Hnew, Texit, ghinv, Err;
  int cell, have, newstep, rejLast, rejMore, nconsec, ierr, skip, sing, accept;
  int out_cell, in_cell, out_ierr;
  int ist[8], out_ist[8];
  double out_r[3];
};

// ---- per-warp table stream ------------------------------------------------------------------------
struct Reader {
  const uint32_t *gsrc;     // this lane's column of the warp's stream
  uint32_t ring;            // shared-space byte address of this lane's column of the warp's ring
  int L, irow, islot, cslot;
  __device__ __forceinline__ void issue()
  {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;\n\tcp.async.commit_group;"
                 :: "r"(ring + islot * 128), "l"(gsrc + (size_t)irow * 32));
    if (++irow == L) irow = 0;
    if (++islot == RING) islot = 0;
  }
  __device__ __forceinline__ void prime()
  {
    for (int i = 0; i < RING - 1; i++) issue();
  }
  // m (1..KCH, warp uniform) rows
  __device__ __forceinline__ void fetch(uint32_t (&w)[KCH], int m)
  {
    asm volatile("cp.async.wait_group %0;" :: "n"(RING - 1 - KCH));
#pragma unroll
    for (int j = 0; j < KCH; j++)
      if (j < m) {
        asm volatile("ld.shared.u32 %0, [%1];" : "=r"(w[j]) : "r"(ring + cslot * 128));
        if (++cslot == RING) cslot = 0;
      }
#pragma unroll
    for (int j = 0; j < KCH; j++)
      if (j < m) issue();
  }
  __device__ __forceinline__ uint32_t fetch1()
  {
    uint32_t w[KCH];
    fetch(w, 1);
    return w[0];
  }
};

template <int NW>
__device__ __forceinline__ void round_barrier(int P, int warp)
{
  if (P >= NW) __syncthreads();
  else if (P <= 1) { if (warp == 0) __syncwarp(); }
  else if (warp < P) asm volatile("bar.sync %0, %1;" :: "r"(P), "r"(P * 32) : "memory");
}

// One round of the bundle engine for this warp.  nb = bundles of the round (all warps).
template <int OP, int NC, int NW>
__device__ __forceinline__ void run_round(Reader &rd, int nb, int warp, const SmemDims &D,
                                          double *__restrict__ G, double *__restrict__ X,
                                          const double *__restrict__ SCR, const double *__restrict__ COEF,
                                          const Slot *slot)
{
  for (int b = warp; b < nb; b += NW) {
    const uint32_t lw = rd.fetch1();
    const int row = lw >> 16, len = (lw >> 8) & 0xff, lg = (lw >> 2) & 7;
    const int maxlen = __reduce_max_sync(FULLMASK, len);
    double acc[NC];
#pragma unroll
    for (int c = 0; c < NC; c++) acc[c] = 0.0;
    int first = 0;
    for (int k0 = 0; k0 < maxlen; k0 += KCH) {
      uint32_t w[KCH];
      const int m = min(KCH, maxlen - k0);
      rd.fetch(w, m);
      if (k0 == 0) first = w[0] & 0xffff;
      if (OP == OP_VDOT || OP == OP_JVS || OP == OP_LUUPD || OP == OP_SOLVE) {
#pragma unroll
        for (int j = 0; j < KCH; j++) {
          if (j < m && k0 + j < len) {
            const int hi = w[j] >> 16, lo = w[j] & 0xffff;
            if (OP == OP_VDOT || OP == OP_JVS) {
              const double cf = COEF[hi];
#pragma unroll
              for (int c = 0; c < NC; c++) acc[c] = fma(cf, SCR[c * D.nscr + lo], acc[c]);
            } else if (OP == OP_LUUPD) {
#pragma unroll
              for (int c = 0; c < NC; c++) acc[c] = fma(G[c * D.nnz + hi], G[c * D.nnz + lo], acc[c]);
            } else {
#pragma unroll
              for (int c = 0; c < NC; c++) acc[c] = fma(G[c * D.nnz + hi], X[c * D.nvar + lo], acc[c]);
            }
          }
        }
      }
    }
    if (OP == OP_VDOT || OP == OP_JVS || OP == OP_LUUPD || OP == OP_SOLVE) {
      for (int s = 0; s < lg; s++) {
#pragma unroll
        for (int c = 0; c < NC; c++) acc[c] += __shfl_down_sync(FULLMASK, acc[c], 1 << s);
      }
    }
    if (lw & 1) {
      if (OP == OP_VDOT) {
#pragma unroll
        for (int c = 0; c < NC; c++) X[c * D.nvar + row] = acc[c];
      } else if (OP == OP_JVS) {
#pragma unroll
        for (int c = 0; c < NC; c++) G[c * D.nnz + row] = ((lw & 2) ? slot[c].ghinv : 0.0) - acc[c];
      } else if (OP == OP_LUUPD) {
#pragma unroll
        for (int c = 0; c < NC; c++) G[c * D.nnz + row] -= acc[c];
      } else if (OP == OP_SOLVE) {
#pragma unroll
        for (int c = 0; c < NC; c++) X[c * D.nvar + row] -= acc[c];
      } else if (OP == OP_LUDIV) {
        if (len) {
#pragma unroll
          for (int c = 0; c < NC; c++) G[c * D.nnz + row] = G[c * D.nnz + row] / G[c * D.nnz + first];
        }
      } else if (OP == OP_SCALE) {
        if (len) {
#pragma unroll
          for (int c = 0; c < NC; c++) G[c * D.nnz + row] *= G[c * D.nnz + first];
        }
      }
    }
  }
}

template <int NC, int NW, int NA_IT, int NB_IT>
__global__ void __launch_bounds__(NW * 32, 1) ros_smem_kernel(SmemArgs P, RosArgs a)
{
  constexpr int NT = NW * 32;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const SmemDims D = P.D;
  double *G = reinterpret_cast<double *>(smem_raw);          // [NC][nnz]
  double *Yg = G + NC * D.nnz;                               // [NC][nyg]  state under evaluation + literals + 1.0
  double *X = Yg + NC * D.nyg;                               // [NC][nvar] right-hand side / solution
  double *SCR = X + NC * D.nvar;                             // [NC][nscr] A(r) or B(m)
  double *COEF = SCR + NC * D.nscr;                          // [ncoef]
  double *RED = COEF + D.ncoef;                              // [NW][NC]
  Slot *slot = reinterpret_cast<Slot *>(RED + NW * NC);      // [NC]
  uint32_t *ringbuf = reinterpret_cast<uint32_t *>(slot + NC);   // [NW][RING][32]
  uint16_t *prog = reinterpret_cast<uint16_t *>(ringbuf + NW * RING * 32);   // [nprog] bundles | P << 8 ... see host
  uint16_t *progP = prog + D.nprog_pad;                      // [nprog] barrier class after the round
  uint16_t *diag = progP + D.nprog_pad;                      // [nvar]
  __shared__ int s_exhausted;

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const RosOpts &o = a.o;
  const double Dir = (double)o.Direction;
  const int N = D.nvar;

  for (int i = tid; i < D.ncoef; i += NT) COEF[i] = P.coefs[i];
  for (int i = tid; i < D.nprog; i += NT) { prog[i] = P.prog_nb[i]; progP[i] = P.prog_P[i]; }
  for (int i = tid; i < N; i += NT) diag[i] = P.diag[i];
  for (int i = tid; i < NC * D.nyg; i += NT) {
    int k = i % D.nyg;
    Yg[i] = (k < D.nspec) ? 1.0 : (k < D.nspec + D.nlit ? P.lit[k - D.nspec] : 1.0);
  }
  if (tid < NC) {
    Slot &s = slot[tid];
    s.have = 0; s.cell = -1; s.H = 1.0; s.T = 0.0; s.ghinv = 2.0; s.newstep = 0; s.skip = 1; s.sing = 0;
    s.out_cell = -1; s.in_cell = -1; s.ierr = 0; s.accept = 0;
  }
  if (tid == 0) s_exhausted = 0;

  // thread-private tables: the reactions (rate phase) and partial derivatives (Jacobian phase) of this thread
  uint32_t aw0[NA_IT], aw1[NA_IT], bw0[NB_IT], bw1[NB_IT];
  double rcA[NA_IT][NC], rcB[NB_IT][NC];
#pragma unroll
  for (int k = 0; k < NA_IT; k++) {
    int r = k * NT + tid;
    aw0[k] = r < D.nreact ? P.aw[2 * r] : 0xffffffffu;
    aw1[k] = r < D.nreact ? P.aw[2 * r + 1] : 0;
#pragma unroll
    for (int c = 0; c < NC; c++) rcA[k][c] = 0.0;
  }
#pragma unroll
  for (int k = 0; k < NB_IT; k++) {
    int m = k * NT + tid;
    bw0[k] = m < D.nb ? P.bw[2 * m] : 0xffffffffu;
    bw1[k] = m < D.nb ? P.bw[2 * m + 1] : 0;
#pragma unroll
    for (int c = 0; c < NC; c++) rcB[k][c] = 0.0;
  }
  // owner registers: thread i holds species i of every slot
  double Y[NC], F0[NC], K1[NC], K2[NC], K3[NC], K4[NC];
#pragma unroll
  for (int c = 0; c < NC; c++) { Y[c] = 1.0; F0[c] = K1[c] = K2[c] = K3[c] = K4[c] = 0.0; }
  double atol_i = 1.0, rtol_i = 1.0;
  if (tid < N) {
    atol_i = o.VectorTol ? a.atol[tid] : a.atol[0];
    rtol_i = o.VectorTol ? a.rtol[tid] : a.rtol[0];
  }
  unsigned long long acc_stp = 0, acc_acc = 0, acc_fail = 0, acc_done = 0;   // meaningful in threads < NC

  Reader rd;
  rd.gsrc = P.stream + (size_t)P.warp_off[warp] * 32 + lane;
  rd.ring = (uint32_t)__cvta_generic_to_shared(ringbuf + (warp * RING) * 32 + lane);
  rd.L = P.warp_rows[warp]; rd.irow = 0; rd.islot = 0; rd.cslot = 0;
  rd.prime();
  __syncthreads();
  PROF_DECL

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
      if (!s.have && !s_exhausted) {
        int w = atomicAdd(a.next, 1);
        if (w >= a.nwork) {
          s_exhausted = 1;      // benign race: every writer stores 1
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
        if (tid < D.nspec) a.conc_out[(size_t)tid * a.ncell + oc] = Y[c];
        if (tid < 8 && a.istatus) a.istatus[(size_t)tid * a.ncell + oc] = slot[c].out_ist[tid];
        if (tid >= 32 && tid < 35 && a.rstatus) a.rstatus[(size_t)(tid - 32) * a.ncell + oc] = slot[c].out_r[tid - 32];
        if (tid == 35 && a.rstatus) a.rstatus[(size_t)3 * a.ncell + oc] = 0.0;
        if (tid == 36 && a.ierr) a.ierr[oc] = slot[c].out_ierr;
      }
      if (ic >= 0) {
        if (tid < D.nspec) Y[c] = a.conc_in[(size_t)tid * a.ncell + ic];
#pragma unroll
        for (int k = 0; k < NA_IT; k++) {
          int i0 = aw0[k] & 0xffff;
          if (aw0[k] != 0xffffffffu) rcA[k][c] = i0 < D.nreact ? a.rconst[(size_t)i0 * a.ncell + ic] : P.lit[i0 - D.nreact];
        }
#pragma unroll
        for (int k = 0; k < NB_IT; k++) {
          int i0 = bw0[k] & 0xffff;
          if (bw0[k] != 0xffffffffu) rcB[k][c] = i0 < D.nreact ? a.rconst[(size_t)i0 * a.ncell + ic] : P.lit[i0 - D.nreact];
        }
      } else if (!slot[c].have) {
        if (tid < D.nspec) Y[c] = 1.0;      // idle slot: benign numbers
      }
      any |= (slot[c].have != 0);
    }
    if (!any) break;
    PROF(0);

    int rp = 0;       // program round pointer
    // ---- helpers as lambdas ----------------------------------------------------------------------
    auto rate_phase = [&]() {       // A(r) = RCT(r) * prod(V)  (Fun, first half)
#pragma unroll
      for (int k = 0; k < NA_IT; k++) {
        if (aw0[k] != 0xffffffffu) {
          const int r = k * NT + tid, i1 = aw0[k] >> 16, i2 = aw1[k] & 0xffff, i3 = aw1[k] >> 16;
#pragma unroll
          for (int c = 0; c < NC; c++)
            SCR[c * D.nscr + r] = rcA[k][c] * Yg[c * D.nyg + i1] * Yg[c * D.nyg + i2] * Yg[c * D.nyg + i3];
        }
      }
    };
    auto solve = [&]() {            // KppSolve on X in place
      for (int r = 0; r < D.n_fwd; r++, rp++) {
        run_round<OP_SOLVE, NC, NW>(rd, prog[rp], warp, D, G, X, SCR, COEF, slot);
        round_barrier<NW>(progP[rp], warp);
      }
      for (int i = tid; i < N; i += NT) {
        const int dp = diag[i];
#pragma unroll
        for (int c = 0; c < NC; c++) X[c * N + i] *= G[c * D.nnz + dp];
      }
      __syncthreads();
      for (int r = 0; r < D.n_bwd; r++, rp++) {
        run_round<OP_SOLVE, NC, NW>(rd, prog[rp], warp, D, G, X, SCR, COEF, slot);
        round_barrier<NW>(progP[rp], warp);
      }
    };
    auto fun_to_X = [&]() {         // X = Fun(Yg)
      rate_phase();
      __syncthreads();
      run_round<OP_VDOT, NC, NW>(rd, prog[rp], warp, D, G, X, SCR, COEF, slot);
      rp++;
      __syncthreads();
    };

    // ---- Fcn0 = Fun(Y)  (:668)
    if (tid < D.nspec) {
#pragma unroll
      for (int c = 0; c < NC; c++) Yg[c * D.nyg + tid] = Y[c];
    }
    __syncthreads();
    fun_to_X();
    if (tid < N) {
#pragma unroll
      for (int c = 0; c < NC; c++) F0[c] = X[c * N + tid];
    }
    PROF(1);
    // ---- Ghimj = 1/(H*gamma) - Jac0  (:1973-1977); Jac0 is recomputed per attempt
#pragma unroll
    for (int k = 0; k < NB_IT; k++) {
      if (bw0[k] != 0xffffffffu) {
        const int m = k * NT + tid, i1 = bw0[k] >> 16, i2 = bw1[k] & 0xffff, i3 = bw1[k] >> 16;
#pragma unroll
        for (int c = 0; c < NC; c++)
          SCR[c * D.nscr + m] = rcB[k][c] * Yg[c * D.nyg + i1] * Yg[c * D.nyg + i2] * Yg[c * D.nyg + i3];
      }
    }
    __syncthreads();
    run_round<OP_JVS, NC, NW>(rd, prog[rp], warp, D, G, X, SCR, COEF, slot);
    rp++;
    __syncthreads();
    PROF(2);
    // ---- sparse LU  (KppDecomp)
    for (int r = 0; r < D.n_lu; r++, rp++) {
      const int nbk = prog[rp];
      if (nbk & 0x8000) run_round<OP_LUDIV, NC, NW>(rd, nbk & 0x7fff, warp, D, G, X, SCR, COEF, slot);
      else run_round<OP_LUUPD, NC, NW>(rd, nbk, warp, D, G, X, SCR, COEF, slot);
      round_barrier<NW>(progP[rp], warp);
    }
    PROF(3);
    for (int i = tid; i < N; i += NT) {
      const int dp = diag[i];
#pragma unroll
      for (int c = 0; c < NC; c++) {
        const double d = G[c * D.nnz + dp];
        if (!(fabs(d) >= DBL_MIN)) slot[c].sing = 1;          // also catches NaN from an earlier zero pivot
        G[c * D.nnz + dp] = 1.0 / d;
      }
    }
    __syncthreads();
    run_round<OP_SCALE, NC, NW>(rd, prog[rp], warp, D, G, X, SCR, COEF, slot);
    rp++;
    __syncthreads();

    PROF(4);
    // ---- stages (Rodas3: NewF = T,F,T,T; :691-724)
    double dh[NC];
#pragma unroll
    for (int c = 0; c < NC; c++) dh[c] = Dir * slot[c].H;
    // stage 1: K1 = Fcn0
    if (tid < N) {
#pragma unroll
      for (int c = 0; c < NC; c++) X[c * N + tid] = F0[c];
    }
    __syncthreads();
    solve();
    PROF(5);
    if (tid < N) {
#pragma unroll
      for (int c = 0; c < NC; c++) {
        K1[c] = X[c * N + tid];
        // stage 2: K2 = Fcn0 + C21/H K1
        X[c * N + tid] = fma(o.C[0] / dh[c], K1[c], F0[c]);
      }
    }
    __syncthreads();
    solve();
    PROF(5);
    if (tid < N) {
#pragma unroll
      for (int c = 0; c < NC; c++) {
        K2[c] = X[c * N + tid];
        // stage 3: Ynew = Y + A31 K1 + A32 K2
        Yg[c * D.nyg + tid] = fma(o.A[2], K2[c], fma(o.A[1], K1[c], Y[c]));
      }
    }
    __syncthreads();
    fun_to_X();
    PROF(6);
    if (tid < N) {
#pragma unroll
      for (int c = 0; c < NC; c++)
        X[c * N + tid] = fma(o.C[2] / dh[c], K2[c], fma(o.C[1] / dh[c], K1[c], X[c * N + tid]));
    }
    __syncthreads();
    solve();
    PROF(5);
    if (tid < N) {
#pragma unroll
      for (int c = 0; c < NC; c++) {
        K3[c] = X[c * N + tid];
        // stage 4: Ynew = Y + A41 K1 + A42 K2 + A43 K3
        Yg[c * D.nyg + tid] = fma(o.A[5], K3[c], fma(o.A[4], K2[c], fma(o.A[3], K1[c], Y[c])));
      }
    }
    __syncthreads();
    fun_to_X();
    PROF(6);
    if (tid < N) {
#pragma unroll
      for (int c = 0; c < NC; c++)
        X[c * N + tid] = fma(o.C[5] / dh[c], K3[c], fma(o.C[4] / dh[c], K2[c], fma(o.C[3] / dh[c], K1[c], X[c * N + tid])));
    }
    __syncthreads();
    solve();
    PROF(5);
    // ---- new solution, error estimate and norm  (:729-740, :1715-1745)
    double yn[NC], e2[NC];
#pragma unroll
    for (int c = 0; c < NC; c++) { yn[c] = 0.0; e2[c] = 0.0; }
    if (tid < N) {
#pragma unroll
      for (int c = 0; c < NC; c++) {
        K4[c] = X[c * N + tid];
        yn[c] = fma(o.M[3], K4[c], fma(o.M[2], K3[c], fma(o.M[1], K2[c], fma(o.M[0], K1[c], Y[c]))));
        const double ye = fma(o.E[3], K4[c], fma(o.E[2], K3[c], fma(o.E[1], K2[c], o.E[0] * K1[c])));
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
        if (slot[c].have && slot[c].accept) Y[c] = o.ClipNegative ? fmax(yn[c], 0.0) : yn[c];
    }
    __syncthreads();      // the control threads rewrite slot[] at the top of the loop
    PROF(7);
  }
#ifdef SMEM_PROFILE
  if (tid == 0 && blockIdx.x == 0 && a.sums)
    for (int i = 0; i < 8; i++) a.sums[8 + i] = (unsigned long long)pacc_[i];
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

int smem_plan_build(const gckpp_host_tables_t *T, const gckpp_sched_tables_t *S, int NW, SmemHostPlan &hp)
{
  const int NC = SMEM_NC;
  SmemDims &D = hp.D;
  D.nvar = T->nvar; D.nspec = T->nspec; D.nreact = T->nreact; D.nnz = T->nnz; D.nb = T->nb;
  D.nlit = T->nlit; D.nyg = T->nspec + T->nlit + 1; D.nscr = T->nreact > T->nb ? T->nreact : T->nb;
  D.ncoef = S->ncoef;
  const int *ph = S->phase;
  auto nrounds = [&](int p) { return ph[2 * p + 1] - ph[2 * p]; };
  D.n_lu = nrounds(2); D.n_fwd = nrounds(4); D.n_bwd = nrounds(5);
  if (nrounds(0) != 1 || nrounds(1) != 1 || nrounds(3) != 1) return -1;
  // program = sequence of schedule rounds of one Rodas3 attempt
  std::vector<int> prg;
  auto push_phase = [&](int p) { for (int r = ph[2 * p]; r < ph[2 * p + 1]; r++) prg.push_back(r); };
  auto push_solve = [&]() { push_phase(4); push_phase(5); };
  std::vector<int> phase_end;   // program indices after which every warp must synchronise
  auto mark_end = [&]() { phase_end.push_back((int)prg.size() - 1); };
  push_phase(0); mark_end();               // Fcn0
  push_phase(1); mark_end();               // Jacobian
  push_phase(2); mark_end();               // LU
  push_phase(3); mark_end();               // scale
  for (int st = 0; st < 4; st++) {
    if (st >= 2) { push_phase(0); mark_end(); }
    push_phase(4); mark_end();
    push_phase(5); mark_end();
  }
  (void)push_solve;
  const int np = (int)prg.size();
  D.nprog = np; D.nprog_pad = (np + 7) & ~7;
  hp.prog_nb.assign(np, 0); hp.prog_P.assign(np, 0);
  std::vector<char> is_end(np, 0);
  for (int e : phase_end) is_end[e] = 1;
  for (int i = 0; i < np; i++) {
    const uint32_t *rr = S->rounds + 3 * prg[i];
    int nb = (int)(rr[1] - rr[0]);
    if (nb >= 0x8000) return -1;
    hp.prog_nb[i] = (uint16_t)(nb | ((rr[2] & 0x10) ? 0x8000 : 0));
  }
  for (int i = 0; i < np; i++) {
    int nb = hp.prog_nb[i] & 0x7fff;
    int nn = (i + 1 < np) ? (hp.prog_nb[i + 1] & 0x7fff) : NW;
    int Pb = nb > nn ? nb : nn;
    if (Pb > NW || is_end[i]) Pb = NW;
    hp.prog_P[i] = (uint16_t)Pb;
  }
  // per-warp streams in consumption order
  std::vector<std::vector<uint32_t>> ws(NW);
  for (int i = 0; i < np; i++) {
    const uint32_t *rr = S->rounds + 3 * prg[i];
    for (uint32_t b = rr[0]; b < rr[1]; b++) {
      int w = (int)((b - rr[0]) % NW);
      uint32_t base = S->bundles[2 * b], ml = S->bundles[2 * b + 1];
      uint32_t maxlen = ml & 0xff, lg = ml >> 8;
      for (int l = 0; l < 32; l++) {
        uint32_t lw = S->lanes[b * 32 + l];
        ws[w].push_back((lw & 0xffffff03u) | (lg << 2));
      }
      for (uint32_t k = 0; k < maxlen; k++)
        for (int l = 0; l < 32; l++) ws[w].push_back(S->terms[base + k * 32 + l]);
    }
  }
  hp.stream.clear();
  for (int w = 0; w < NW; w++) {
    while ((int)ws[w].size() < 32 * 2 * RING) ws[w].insert(ws[w].end(), 32, 0u);   // never shorter than the ring (idle warps)
    hp.warp_off[w] = (int)(hp.stream.size() / 32);
    hp.warp_rows[w] = (int)(ws[w].size() / 32);
    hp.stream.insert(hp.stream.end(), ws[w].begin(), ws[w].end());
  }
  // a warp with no work at all never consumes: its padded stream is only prefetched
  hp.aw.resize(2 * (size_t)T->nreact);
  for (int r = 0; r < T->nreact; r++) encode_term(T->a_term + 4 * r, T->nreact, T->nspec, T->nlit, &hp.aw[2 * r]);
  hp.bw.resize(2 * (size_t)(T->nb > 0 ? T->nb : 1));
  for (int m = 0; m < T->nb; m++) encode_term(T->b_term + 4 * m, T->nreact, T->nspec, T->nlit, &hp.bw[2 * m]);
  hp.diag.resize(T->nvar);
  for (int i = 0; i < T->nvar; i++) hp.diag[i] = (uint16_t)T->diag[i];
  if (T->nspec + T->nlit + 1 >= 65536 || T->nreact + T->nlit >= 65535 || T->nnz >= 65536) return -1;
  size_t bytes = sizeof(double) * ((size_t)NC * (D.nnz + D.nyg + D.nvar + D.nscr) + D.ncoef + (size_t)NW * NC);
  bytes += sizeof(Slot) * NC;
  bytes += sizeof(uint32_t) * (size_t)NW * RING * 32;
  bytes += sizeof(uint16_t) * (2 * (size_t)D.nprog_pad + D.nvar);
  hp.smem_bytes = (bytes + 15) & ~(size_t)15;
  hp.NW = NW;
  return 0;
}

template <int NW, int NA_IT, int NB_IT>
static cudaError_t launch_t(const SmemArgs &P, const RosArgs &a, int blocks, size_t smem, cudaStream_t s)
{
  auto k = ros_smem_kernel<SMEM_NC, NW, NA_IT, NB_IT>;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    configured = true;
  }
  k<<<blocks, NW * 32, smem, s>>>(P, a);
  return cudaGetLastError();
}

bool smem_kernel_supports(const gckpp_host_tables_t *T, int NW)
{
  const int NT = NW * 32;
  if (T->nnz <= 0 || T->nspec > NT) return false;
  int na = (T->nreact + NT - 1) / NT, nb = (T->nb + NT - 1) / NT;
  if (NW == 12) return (na == 3 && nb == 5) || (na == 1 && nb == 1);
  if (NW == 8) return (na == 5 && nb == 8) || (na == 1 && nb == 1);
  return false;
}

cudaError_t launch_ros_smem(const SmemArgs &P, const RosArgs &a, int NW, int blocks, size_t smem, cudaStream_t s)
{
  const int NT = NW * 32;
  int na = (P.D.nreact + NT - 1) / NT, nb = (P.D.nb + NT - 1) / NT;
  if (NW == 12 && na == 3 && nb == 5) return launch_t<12, 3, 5>(P, a, blocks, smem, s);
  if (NW == 12 && na == 1 && nb == 1) return launch_t<12, 1, 1>(P, a, blocks, smem, s);
  if (NW == 8 && na == 5 && nb == 8) return launch_t<8, 5, 8>(P, a, blocks, smem, s);
  if (NW == 8 && na == 1 && nb == 1) return launch_t<8, 1, 1>(P, a, blocks, smem, s);
  return cudaErrorInvalidConfiguration;
}
