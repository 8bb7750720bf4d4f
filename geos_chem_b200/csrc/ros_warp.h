// Interface of the warp-group-per-cell Rosenbrock kernel (ros_warp.cu).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <vector>
#include "ros_common.cuh"

#ifndef WARP_WG
#define WARP_WG 4          // warps that integrate one cell together (a "group")
#endif
#ifndef WARP_RS
#define WARP_RS 6          // ring slots (512-byte table rows) per warp
#endif
#define WARP_NPH 6         // table phases: vdot, jvs, jvs2, lu, fwd, bwd
#define WARP_NSEG 14       // table segments of one Rodas3 attempt (see WarpArgs)

// Bundle streams of one mechanism (kppgen/wsched.py documents the encoding); emitted into gen/<mech>_wsched.h.
struct gckpp_wsched_tables_t {
  const uint32_t *rows[WARP_NPH];      // vdot, jvs, jvs2, lu, fwd, bwd: [nrows][32][4]
  int nrows[WARP_NPH], nbundles[WARP_NPH];
  const double *coefs; int ncoef;
  const uint16_t *tpos;                // [32][32] tposT[j][i] = position of G(h+i,h+j) or 0xFFFF
  int head, tail;
};

enum { WP_VDOT = 0, WP_JVS, WP_JVS2, WP_LU, WP_FWD, WP_BWD };
// segments of an attempt, in consumption order
enum { WS_VDOT0 = 0, WS_JVS, WS_JVS2, WS_LU, WS_FWD0, WS_BWD0, WS_FWD1, WS_BWD1, WS_VDOT1, WS_FWD2, WS_BWD2, WS_VDOT2, WS_FWD3, WS_BWD3 };

struct WarpArgs {
  // One Rodas3 attempt consumes the table phases in the order
  //   vdot jvs jvs2 lu fwd bwd fwd bwd vdot fwd bwd vdot fwd bwd
  // Every warp of a group has its own stream holding exactly that sequence of ITS bundles (cyclic), so the
  // ring prefetch never stalls on the common path.  The bundles of a dependency level are dealt round-robin
  // to the warps; the last bundle of a warp in a level carries the SYNC flag (group barrier after it), the PRE field of a
  // bundle counts the barriers of the levels before it in which the warp had no bundle.
  const uint4 *stream;                 // all per-warp streams, [rows][32]
  int w_off[WARP_WG], w_rows[WARP_WG]; // first row and length of warp-stream w
  int seg_off[WARP_WG][WARP_NSEG];     // first row of a segment, relative to the warp-stream
  int nb[WARP_WG][WARP_NPH];           // bundles of a phase in warp-stream w
  int nlev[WARP_NPH];                  // dependency levels (= group barriers) of a phase
  const uint16_t *tpos;                // [32][32]
  const uint16_t *diag;                // [nvar] position of the diagonal
  const uint32_t *aw, *bw;             // [nreact][2], [nb][2] encoded rate / partial-derivative terms
  const double *coefs;                 // [ncoef] stoichiometric coefficients (signed), last = 0.0
  const double *lit;                   // [nlit]
  double *rcs;                         // per-group scratch: rate constants in item order, [groups][NREACT + NB]
  int s_total;                         // dynamic shared memory per block
};

struct WarpHostPlan {
  std::vector<uint32_t> stream, aw, bw;
  std::vector<uint16_t> diag;
  int w_off[WARP_WG], w_rows[WARP_WG];
  int seg_off[WARP_WG][WARP_NSEG];
  int nb[WARP_WG][WARP_NPH];
  int nlev[WARP_NPH];
};

bool warp_kernel_supports(int mech_id);
int warp_plan_build(int mech_id, const gckpp_host_tables_t *T, const gckpp_wsched_tables_t *S, WarpHostPlan &hp);
int warp_cells_per_block(int mech_id);
int warp_smem_bytes(int mech_id);
size_t warp_rcs_doubles_per_group(int mech_id);
cudaError_t launch_ros_warp(int mech_id, const WarpArgs &P, const RosArgs &a, int blocks, cudaStream_t s);
