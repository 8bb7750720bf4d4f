// Interface of the warp-per-cell Rosenbrock kernel (ros_warp.cu).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <vector>
#include "ros_common.cuh"

#ifndef WARP_RS
#define WARP_RS 8          // ring slots (512-byte table rows) per warp
#endif

// Bundle streams of one mechanism (kppgen/wsched.py documents the encoding); emitted into gen/<mech>_wsched.h.
struct gckpp_wsched_tables_t {
  const uint32_t *rows[5];             // vdot, jvs, lu, fwd, bwd: [nrows][32][4]
  int nrows[5], nbundles[5];
  const double *coefs; int ncoef;
  const uint16_t *tpos;                // [32][32] tposT[j][i] = position of G(h+i,h+j) or 0xFFFF
  int head, tail;
};

enum { WP_VDOT = 0, WP_JVS, WP_LU, WP_FWD, WP_BWD };

struct WarpArgs {
  // One Rodas3 attempt consumes the phases in the order
  //   vdot jvs lu fwd bwd fwd bwd vdot fwd bwd vdot fwd bwd
  // and the stream holds exactly that sequence (cyclic), so the ring prefetch never stalls on the common path.
  const uint4 *stream;                 // [rows_total][32]
  int rows_total;
  int off_vdot[3], off_jvs, off_lu, off_fwd[4], off_bwd[4];
  int nb[5];                           // bundles per phase
  const uint16_t *tpos;                // [32][32]
  const uint16_t *diag;                // [nvar] position of the diagonal
  const uint32_t *aw, *bw;             // [nreact][2], [nb][2] encoded rate / partial-derivative terms
  const double *coefs;                 // [ncoef] stoichiometric coefficients (signed), last = 0.0
  const double *lit;                   // [nlit]
  double *rcs;                         // per-warp scratch: rate constants in item order, [warps][NREACT + NB]
  int s_total;                         // dynamic shared memory per block
};

struct WarpHostPlan {
  std::vector<uint32_t> stream, aw, bw;
  std::vector<uint16_t> diag;
  int rows_total;
  int off_vdot[3], off_jvs, off_lu, off_fwd[4], off_bwd[4];
  int nb[5];
};

bool warp_kernel_supports(int mech_id);
int warp_plan_build(int mech_id, const gckpp_host_tables_t *T, const gckpp_wsched_tables_t *S, WarpHostPlan &hp);
int warp_cells_per_block(int mech_id);
int warp_smem_bytes(int mech_id);
size_t warp_rcs_doubles_per_warp(int mech_id);
cudaError_t launch_ros_warp(int mech_id, const WarpArgs &P, const RosArgs &a, int blocks, cudaStream_t s);
