// Interface of the shared-memory-resident Rosenbrock kernel (ros_smem.cu).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <vector>
#include "ros_common.cuh"

#define SMEM_NC 3          // cells integrated in lock step by one thread block
#define SMEM_MAX_WARPS 16

struct SmemDims {
  int nvar, nspec, nreact, nnz, nb, nlit;
  int nyg;                 // nspec + nlit + 1: [VAR, FIX, literal pool, 1.0]
  int nscr;                // max(nreact, nb)
  int ncoef;
  int n_lu, n_fwd, n_bwd;  // rounds per phase
  int nprog, nprog_pad;    // rounds of one attempt
};

struct SmemArgs {
  SmemDims D;
  const uint32_t *stream;                 // per-warp table streams, rows of 32 words
  int warp_off[SMEM_MAX_WARPS];           // first row of each warp's stream
  int warp_rows[SMEM_MAX_WARPS];          // cyclic length of each warp's stream (rows per attempt)
  const uint16_t *prog_nb;                // [nprog] bundles of the round | 0x8000 for LU division rounds
  const uint16_t *prog_P;                 // [nprog] warps that synchronise after the round
  const uint16_t *diag;                   // [nvar] LU_DIAG
  const uint32_t *aw, *bw;                // [nreact][2], [nb][2] encoded rate / partial-derivative terms
  const double *coefs;                    // [ncoef] stoichiometric coefficients (signed)
  const double *lit;                      // [nlit] literal pool
};

struct SmemHostPlan {
  SmemDims D;
  int NW;
  size_t smem_bytes;
  int warp_off[SMEM_MAX_WARPS], warp_rows[SMEM_MAX_WARPS];
  std::vector<uint32_t> stream, aw, bw;
  std::vector<uint16_t> prog_nb, prog_P, diag;
};

bool smem_kernel_supports(const gckpp_host_tables_t *T, int NW);
int smem_plan_build(const gckpp_host_tables_t *T, const gckpp_sched_tables_t *S, int NW, SmemHostPlan &hp);
cudaError_t launch_ros_smem(const SmemArgs &P, const RosArgs &a, int NW, int blocks, size_t smem, cudaStream_t s);
