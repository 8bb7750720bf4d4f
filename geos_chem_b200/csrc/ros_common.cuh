// Shared definitions of the Rosenbrock kernels: decoded options, launch arguments, the
// warp-private workspace layout and the per-lane cell scheduler.
//
// Execution model (both kernels):
//   * one grid cell per LANE; a warp integrates 32 cells in lock step (same sparse pattern,
//     same instruction stream, different data);
//   * scratch vectors live in a warp-private workspace in HBM laid out [element][32 lanes]
//     ("AoSoA-32"): element k of a vector at ws[(off+k)*32 + lane], so every access of a
//     warp is one fully coalesced 256-byte row and all offsets are compile-time/uniform;
//   * lanes are PERSISTENT: when a lane's cell reaches Tend (or fails) it writes its results
//     and pulls the next cell index from a global counter (warp-aggregated atomicAdd), so
//     the per-cell adaptive step counts do not idle lanes until the grid is exhausted.
//     This replaces the reference's OpenMP SCHEDULE(DYNAMIC,24) (fullchem_mod.F90:541-542).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "tables.h"

// Decoded integrator options = what Rosenbrock() (gckpp_Integrator.F90:165-531) derives from
// ICNTRL/RCNTRL before calling ros_Integrator.  Decoding is done once on the host.
struct RosOpts {
  int S;                       // stages
  double A[15], C[15], M[6], E[6], Alpha[6], Gamma[6], ELO;
  int NewF[6];
  int Autonomous, VectorTol, Max_no_steps, ClipNegative;
  double Roundoff, Hmin, Hmax, FacMin, FacMax, FacRej, FacSafe;
  double Hstart_rcntrl;        // RCNTRL(3) after Integrate's merge (0 = default); used when hstart == NULL
  double Tstart, Tend;
  int Direction;
};

struct WsLayout {              // offsets in doubles per lane
  int Y, YN, F0, FC, K, G, RC, AB, W, PR, LS, MK, total;    // PR/LS/MK: auto-reduce Prod, Loss, keep mask
};

struct RosArgs {
  int ncell;                   // stride of the cell-fastest arrays
  int nwork;                   // cells to integrate in this launch
  const int *cell_list;        // [nwork] cell indices, or NULL for 0..nwork-1
  const double *conc_in, *rconst, *hstart, *atol, *rtol;
  int rc_stride, rc_cell0;     // rconst is [NREACT][rc_stride] and its column 0 is cell rc_cell0 (a wave's own rate constants)
  double *conc_out;
  int *istatus;                // [8][ncell] or NULL
  double *rstatus;             // [4][ncell] or NULL
  int *ierr;                   // [ncell] or NULL
  double *work;                // workspace: nwarps * ws_stride doubles
  size_t ws_stride;            // doubles per warp (= layout.total * 32)
  int *next;                   // work counter
  unsigned long long *sums;    // [0] sum Nstp [1] sum Nacc [2] cells with ierr<0 [3] cells done
  WsLayout L;
  RosOpts o;
  // auto-reduce (ros_yIntegrator, gckpp_Integrator.F90:789-1237): ICNTRL(12), RCNTRL(12), ICNTRL(14), RCNTRL(14)
  int ar_on, ar_target, ar_keep_active;
  double ar_threshold, ar_ratio;
  const unsigned char *ar_keep_spc;        // [nvar] keepSpcActive, or NULL
};

enum { Nfun = 0, Njac, Nstp, Nacc, Nrej, Ndec, Nsol, Nsng };

#define FULLMASK 0xffffffffu

// Hand out work indices to the lanes that need one (warp-aggregated atomic).
__device__ __forceinline__ int fetch_work(int *next, bool need, int lane)
{
  unsigned mask = __ballot_sync(FULLMASK, need);
  if (mask == 0) return -1;
  int leader = __ffs(mask) - 1;
  int base = 0;
  if (lane == leader) base = atomicAdd(next, __popc(mask));
  base = __shfl_sync(FULLMASK, base, leader);
  return need ? base + __popc(mask & ((1u << lane) - 1u)) : -1;
}
