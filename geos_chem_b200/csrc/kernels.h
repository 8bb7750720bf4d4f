// Kernel launchers (one translation unit per kernel family, see the build recipe).
#pragma once
#include <cuda_runtime.h>
#include <stdarg.h>
#include "ros_common.cuh"

// ros_generic.cu (compiled with -fmad=false)
cudaError_t launch_ros_generic(const MechDev &M, const RosArgs &a, int blocks, int threads, cudaStream_t s);
cudaError_t launch_ar_mask(const MechDev &M, const RosArgs &a, unsigned char *mask, cudaStream_t s);
cudaError_t launch_ar_first_order(const MechDev &M, const RosArgs &a, const unsigned char *mask, cudaStream_t s);
cudaError_t launch_feuler(const MechDev &M, const RosArgs &a, int icntrl16, cudaStream_t s);
cudaError_t launch_fun_cells(const MechDev &M, int ncell, const double *conc, const double *rconst,
                             double *vdot, double *aout, cudaStream_t s);
cudaError_t launch_jac_cells(const MechDev &M, int ncell, const double *conc, const double *rconst,
                             double *bwork, double *jvs, cudaStream_t s);
cudaError_t launch_decomp_cells(const MechDev &M, int ncell, double *jvs, double *wwork, int *ier, cudaStream_t s);
cudaError_t launch_solve_cells(const MechDev &M, int ncell, const double *jvs, double *x, cudaStream_t s);

// kernels_misc.cu
cudaError_t launch_update_rconst(int mech_id, int ncell, const double *temp, const double *numden,
                                 const double *h2o, const double *photol, const double *khet,
                                 double *rconst, cudaStream_t s, int stride = 0, int ostride = 0,
                                 const double *het = nullptr, const double *conc = nullptr, const double *srmw = nullptr,
                                 int nspec_data = 0);
cudaError_t launch_iota(int *p, int n, cudaStream_t s);
cudaError_t launch_fill_int(int *p, int n, int v, cudaStream_t s);
cudaError_t launch_select_active(int ncell, int c0, int c1, const uint8_t *active, int nspec, const double *conc_in,
                                 double *conc_out, int *istatus, double *rstatus, int *ierr,
                                 int *cell_list, int *count, cudaStream_t s);
cudaError_t launch_select_failed(int c0, int c1, const int *ierr, int *cell_list, int *count, cudaStream_t s);
cudaError_t measure_fp64_peak(double *tflops, double *ms);

// post.cu (compiled with -fmad=false): the pieces of Do_FullChem around the integration
cudaError_t launch_zero_species(double *conc, int ncell, const int *ids, int n, cudaStream_t s);
cudaError_t launch_post_integrate(double *conc, int ncell, int nspec, const int *scale_ids, const double *scale_div, int nscale,
                                  const unsigned char *mask, float *negatives, cudaStream_t s);
cudaError_t launch_prod_loss(const double *conc, int ncell, double dt, const int *ids, int nslots, double *out, cudaStream_t s);
cudaError_t launch_oh_reactivity(const double *conc, const double *rconst, int ncell, const double *coef, const int *rxn,
                                 const int *spc, int nterms, double *out, cudaStream_t s);

// ros_unrolled.cu: one cell per thread, generated straight-line code (small mechanisms)
bool unrolled_kernel_supports(int mech_id);
int unrolled_block_threads();
int unrolled_blocks_per_sm();
cudaError_t launch_ros_unrolled(int mech_id, const RosArgs &a, int blocks, cudaStream_t s);
