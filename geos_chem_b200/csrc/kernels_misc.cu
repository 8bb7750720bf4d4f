// K1 (Update_RCONST) and the small bookkeeping kernels around the integrator.
//   update_rconst_kernel  <- Update_RCONST, KPP/<mech>/gckpp_Rates.F90:408-1503, one cell per thread;
//                            met scalars derived as in Set_Kpp_GridBox_Values (fullchem_mod.F90:2139-2150)
//   select_active_kernel  <- the InChemGrid skip of Do_FullChem (fullchem_mod.F90:804)
//   select_failed_kernel  <- the IERR<0 retry list (fullchem_mod.F90:1138)
#include "kernels.h"
#include "ratelaws.cuh"
#include "gen/fullchem_rconst.cuh"
#include "gen/Hg_rconst.cuh"
#include "gen/carbon_rconst.cuh"

namespace {

__device__ __forceinline__ MetCell make_met(double temp, double numden, double h2o)
{
  MetCell m;
  m.TEMP = temp; m.NUMDEN = numden; m.H2O = h2o;
  m.INV_TEMP = 1.0 / temp;
  m.TEMP_OVER_K300 = temp / 300.0;
  m.K300_OVER_TEMP = 300.0 / temp;
  m.SR_TEMP = sqrt(temp);
  return m;
}

template <int MECH>
__global__ void __launch_bounds__(128) update_rconst_kernel(int ncell, const double *__restrict__ temp,
    const double *__restrict__ numden, const double *__restrict__ h2o, const double *__restrict__ photol,
    const double *__restrict__ khet, double *__restrict__ rconst)
{
  int cell = blockIdx.x * blockDim.x + threadIdx.x;
  if (cell >= ncell) return;
  MetCell m = make_met(temp[cell], numden[cell], h2o[cell]);
  const double *ph = photol ? photol + cell : nullptr;
  const double *kh = khet ? khet + cell : nullptr;
  if (MECH == 0) fullchem_update_rconst_cell(m, ph, kh, rconst + cell, (size_t)ncell);
  else if (MECH == 1) Hg_update_rconst_cell(m, ph, kh, rconst + cell, (size_t)ncell);
  else carbon_update_rconst_cell(m, ph, kh, rconst + cell, (size_t)ncell);
}

__global__ void fill_int_kernel(int *p, int n, int v)
{
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] = v;
}

__global__ void select_active_kernel(int ncell, const uint8_t *__restrict__ active, int nspec,
    const double *__restrict__ conc_in, double *__restrict__ conc_out, int *istatus, double *rstatus,
    int *ierr, int *cell_list, int *count)
{
  int cell = blockIdx.x * blockDim.x + threadIdx.x;
  bool act = cell < ncell && active[cell] != 0;
  unsigned mask = __ballot_sync(0xffffffffu, act);
  int lane = threadIdx.x & 31, base = 0;
  if (mask) {
    int leader = __ffs(mask) - 1;
    if (lane == leader) base = atomicAdd(count, __popc(mask));
    base = __shfl_sync(0xffffffffu, base, leader);
  }
  if (act) cell_list[base + __popc(mask & ((1u << lane) - 1u))] = cell;
  if (cell < ncell && !act) {
    for (int s = 0; s < nspec; s++) conc_out[(size_t)s * ncell + cell] = conc_in[(size_t)s * ncell + cell];
    if (istatus) for (int q = 0; q < 8; q++) istatus[(size_t)q * ncell + cell] = 0;
    if (rstatus) for (int q = 0; q < 4; q++) rstatus[(size_t)q * ncell + cell] = 0.0;
    if (ierr) ierr[cell] = 0;
  }
}

__global__ void select_failed_kernel(int ncell, const int *__restrict__ ierr, int *cell_list, int *count)
{
  int cell = blockIdx.x * blockDim.x + threadIdx.x;
  if (cell < ncell && ierr[cell] < 0) cell_list[atomicAdd(count, 1)] = cell;
}

}  // namespace

cudaError_t launch_update_rconst(int mech_id, int ncell, const double *temp, const double *numden,
                                 const double *h2o, const double *photol, const double *khet,
                                 double *rconst, cudaStream_t s)
{
  int blocks = (ncell + 127) / 128;
  if (mech_id == 0) update_rconst_kernel<0><<<blocks, 128, 0, s>>>(ncell, temp, numden, h2o, photol, khet, rconst);
  else if (mech_id == 1) update_rconst_kernel<1><<<blocks, 128, 0, s>>>(ncell, temp, numden, h2o, photol, khet, rconst);
  else update_rconst_kernel<2><<<blocks, 128, 0, s>>>(ncell, temp, numden, h2o, photol, khet, rconst);
  return cudaGetLastError();
}
cudaError_t launch_fill_int(int *p, int n, int v, cudaStream_t s)
{
  fill_int_kernel<<<(n + 255) / 256, 256, 0, s>>>(p, n, v);
  return cudaGetLastError();
}
cudaError_t launch_select_active(int ncell, const uint8_t *active, int nspec, const double *conc_in,
                                 double *conc_out, int *istatus, double *rstatus, int *ierr,
                                 int *cell_list, int *count, cudaStream_t s)
{
  select_active_kernel<<<(ncell + 255) / 256, 256, 0, s>>>(ncell, active, nspec, conc_in, conc_out, istatus,
                                                           rstatus, ierr, cell_list, count);
  return cudaGetLastError();
}
cudaError_t launch_select_failed(int ncell, const int *ierr, int *cell_list, int *count, cudaStream_t s)
{
  select_failed_kernel<<<(ncell + 255) / 256, 256, 0, s>>>(ncell, ierr, cell_list, count);
  return cudaGetLastError();
}
