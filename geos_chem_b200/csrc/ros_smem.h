// Interface of the shared-memory-resident Rosenbrock kernel (ros_smem.cu).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <vector>
#include "ros_common.cuh"

#ifndef SMEM_NC
#define SMEM_NC 3          // cells integrated in lock step by one thread block (measured on B200: 2 -> 156k, 3 -> 199k, 4 -> 160k cells/s)
#endif
#ifndef SMEM_NW
#define SMEM_NW 12         // warps per block
#endif
// With 2 cells per block the tables of the triangular sweeps (used 4x per attempt) stay resident in shared
// memory and the ring is 6 slots deep; with 3 cells the space goes to the third matrix and the sweep tables
// are streamed like the others through a 4-slot ring.
#if SMEM_NC <= 2
#define SMEM_SWEEP_RESIDENT 1
#define SMEM_RS 6          // ring slots (512-byte chunk rows) per warp for the streamed tables
#elif SMEM_NC == 3
#define SMEM_SWEEP_RESIDENT 0
#ifndef SMEM_RS
#define SMEM_RS 4
#endif
#else
// 4 cells: the rate / partial-derivative scratch moves to a per-block global buffer (L2 resident, read with
// ld.global.cg by the vdot / jvs rounds) and the ring shrinks to 3 slots
#define SMEM_SWEEP_RESIDENT 0
#define SMEM_RS 3
#endif
// Run stage 1's forward head rounds and the scaling of the head rows on warps NC..NW-1 while warps 0..NC-1
// factorise the tail block (needs the streamed sweep tables: the schedule of those rounds is warp-shifted).
#ifndef SMEM_OVERLAP
#define SMEM_OVERLAP (!SMEM_SWEEP_RESIDENT)
#endif
// the pivot loop of the register tail LU in four spans with shrinking column counts (0 = one span over all columns)
#ifndef SMEM_TAIL_SPANS
#define SMEM_TAIL_SPANS 1
#endif
// apply exactly maxlen term steps per bundle instead of whole chunks; bit mask over the ops (1 vdot, 2 jvs, 4 lu, 8 sweeps);
// needs tables whose terms sit in the first maxlen steps for those ops: kppgen/sched.py GCKPP_EXACT_STEPS=<same mask>.  Measured: 229 k cells/s against 239 k with whole chunks --
// a pad step is a broadcast load and costs less than the switches and the extra code
#ifndef SMEM_EXACT_STEPS
#define SMEM_EXACT_STEPS 0
#endif
// Cell-split rounds: a round with at most NW / NC bundles gives every bundle to NC warps, one cell each (warp b*NC + c).
// A warp's shared-memory loads return through its own scheduler's 32 B/clk port (~9.5 cycles per LDS.64, measured), so
// the 42 operand loads of a two-chunk bundle for three cells cost one warp ~400 cycles; split over three warps on three
// schedulers they cost ~130.  Bit mask: 1 = sweep rounds, 2 = LU rounds.  Measured: the sweep rounds do not get faster
// (240.0 k cells/s with or without), the small LU rounds do -- their division rounds are instruction-bound (3 IEEE
// divisions per lane): 244.3 k with both.
#ifndef SMEM_CELL_SPLIT
#define SMEM_CELL_SPLIT (SMEM_SWEEP_RESIDENT ? 0 : 3)
#endif
#ifndef SMEM_RING_PRELOAD
#define SMEM_RING_PRELOAD 0    // measured: 278.7 k vs 281.7 k cells/s -- handing out a preloaded row does not pay, kept as a variant
#endif
#ifndef SMEM_CONST_DIR
#define SMEM_CONST_DIR 0      // measured: 280.1 k vs 281.7 k cells/s from the shared-memory copy -- no gain, kept as a variant
#endif
#ifndef SMEM_UNIFORM_LW
#define SMEM_UNIFORM_LW 1
#endif
#if SMEM_NC >= 4
#define SMEM_SCR_GLOBAL 1
#else
#define SMEM_SCR_GLOBAL 0
#endif

// round directory entry: bundles of the round, warps that work on it, warps that synchronise after it,
// LU division round flag, first resident bundle
#define DIR_NB(d) ((int)((d) & 0x7ffu))
#define DIR_W(d) ((int)(((d) >> 11) & 31u))
#define DIR_P(d) ((int)(((d) >> 16) & 31u))
#define DIR_DIV(d) ((int)(((d) >> 21) & 1u))
#define DIR_BF(d) ((int)(((d) >> 22) & 0x1ffu))
#define DIR_SPLIT(d) ((int)((d) >> 31))
#define DIR_PACK(nb, W, P, div, bf) ((uint32_t)(nb) | ((uint32_t)(W) << 11) | ((uint32_t)(P) << 16) | ((uint32_t)(div) << 21) | ((uint32_t)(bf) << 22))

struct SmemArgs {
  // streamed tables (vdot, jvs, lu rounds): per-warp chunk rows in consumption order, cyclic per attempt
  const uint4 *stream;
  int warp_off[SMEM_NW], warp_rows[SMEM_NW];
  // resident tables (fwd, bwd rounds): copied to shared memory at kernel start
  const uint4 *resident; int res_rows;
  const uint16_t *boff; int nresb;        // first chunk row of every resident bundle
  const uint32_t *dir; int ndir;          // round directory (DIR_* above)
  int o_lu, n_lu, o_fwd, n_fwd, o_bwd, n_bwd, o_fwd1;
  const uint16_t *tpos;                   // [32][32]
  const uint16_t *diag, *crow;            // [nvar], [nvar+1]
  const uint32_t *uscale; int nuscale;    // head rows: (position of a U entry) << 16 | position of its row's diagonal
  const uint32_t *aw, *bw;                // [nreact][2], [nb][2] encoded rate / partial-derivative terms
  const double *coefs;                    // [ncoef] stoichiometric coefficients (signed)
  const double *lit;                      // [nlit] literal pool
  double *rcs;                            // per-block scratch: rate constants in item order, [blocks][NC][items]
  double *scr;                            // per-block A(r)/B(m) scratch when it is not in shared memory
  // byte offsets of the runtime-sized shared-memory regions
  int s_res, s_tpos, s_boff, s_dir, s_diag, s_crow, s_total;
};

struct SmemHostPlan {
  int warp_off[SMEM_NW], warp_rows[SMEM_NW];
  std::vector<uint32_t> stream, resident, dir, aw, bw, uscale, rowcol;
  std::vector<uint16_t> boff, diag, crow;
  int o_lu, n_lu, o_fwd, n_fwd, o_bwd, n_bwd, o_fwd1;
  int s_res, s_tpos, s_boff, s_dir, s_diag, s_crow, s_total;
};

bool smem_kernel_supports(int mech_id);
int smem_plan_build(int mech_id, const gckpp_host_tables_t *T, const gckpp_sched_tables_t *S, SmemHostPlan &hp);
size_t smem_rcs_doubles_per_block(int mech_id);
size_t smem_scr_doubles_per_block(int mech_id);
cudaError_t smem_set_directory(int mech_id, const uint32_t *dir, int n, cudaStream_t s);
cudaError_t launch_ros_smem(int mech_id, const SmemArgs &P, const RosArgs &a, int blocks, cudaStream_t s, bool autoreduce = false);
