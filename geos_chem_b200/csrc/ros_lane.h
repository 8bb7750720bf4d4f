// Lane kernel (ros_lane.cu): one grid cell per lane, stream tables from kppgen/lsched.py.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "ros_common.cuh"

struct LaneDims {
  int nvar, nspec, nreact, nlit, nvec, one, ng, ab_len, nrcx;
};

// Stream tables of one mechanism (host copies, emitted into gen/<mech>_lsched.h); every table is whole chunks of
// 16 records, a record is 2 (rates, lu, fwd, bwd) or 4 (sums) uint4.
struct gckpp_lsched_tables_t {
  LaneDims d;
  const uint32_t *rates_a, *sums_v, *rates_b, *sums_j, *lu, *fwd, *bwd;
  int nchunk_ra, nchunk_sv, nchunk_rb, nchunk_sj, nchunk_lu, nchunk_fwd, nchunk_bwd;
  const double *lit;
};

struct LaneArgs {
  LaneDims d;
  const uint4 *rates_a, *sums_v, *rates_b, *sums_j, *lu, *fwd, *bwd;
  int nchunk_ra, nchunk_sv, nchunk_rb, nchunk_sj, nchunk_lu, nchunk_fwd, nchunk_bwd;
  const double *lit;
  double *ws;                 // [blocks][ws_stride] doubles, [element][32 lanes] per block
  size_t ws_stride;
  int oY, oYN, oF0, oFC, oK, oGA, oRCX, oAB;      // element offsets in a block's workspace
};

size_t lane_smem_bytes(const LaneDims &d);
cudaError_t launch_ros_lane(const LaneArgs &P, const RosArgs &a, int blocks, cudaStream_t s);
