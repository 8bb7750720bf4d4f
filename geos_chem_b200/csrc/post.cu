// Kernels around the integration that let the chemical state stay on the device for a whole chemistry step
// (compiled with -fmad=false: plain IEEE operations in the reference's order, bit-comparable with the oracle):
//   zero_species      fullchem_mod.F90:941-946     C(PL_Kpp_Id(F)) = 0 for the prod/loss family species
//   post_integrate    fullchem_mod.F90:1284-1287   fullchem_ConvertEquivToAlk (fullchem_SulfurChemFuncs.F90:94-105)
//                     fullchem_mod.F90:1326-1348   KppNegatives count, C = MAX(C, 0) over the mapped species
//   prod_loss         fullchem_mod.F90:1463-1492   Loss / Prod(slot) = C(KppId) / DT
//   oh_reactivity     gckpp_Util.F90:983-1040      Get_OHreactivity, the generated sum in source order
// Cell-fastest arrays, one cell per thread: every access of a warp is one coalesced row.
#include "kernels.h"

namespace {

__global__ void zero_species_kernel(double *conc, int ncell, const int *ids, int n)
{
  const int cell = blockIdx.x * blockDim.x + threadIdx.x;
  if (cell >= ncell) return;
  for (int k = 0; k < n; k++) conc[(size_t)ids[k] * ncell + cell] = 0.0;
}

__global__ void post_integrate_kernel(double *conc, int ncell, int nspec, const int *scale_ids, const double *scale_div,
                                      int nscale, const unsigned char *mask, float *negatives)
{
  const int cell = blockIdx.x * blockDim.x + threadIdx.x;
  if (cell >= ncell) return;
  for (int k = 0; k < nscale; k++) {
    const size_t p = (size_t)scale_ids[k] * ncell + cell;
    conc[p] = conc[p] / scale_div[k];
  }
  float neg = 0.0f;
  for (int s = 0; s < nspec; s++) {
    if (mask && !mask[s]) continue;            // not a GEOS-Chem species (Map_KppSpc <= 0): left as it is
    const size_t p = (size_t)s * ncell + cell;
    const double v = conc[p];
    if (v < 0.0) { neg += 1.0f; conc[p] = 0.0; }
  }
  if (negatives) negatives[cell] += neg;
}

__global__ void prod_loss_kernel(const double *conc, int ncell, double dt, const int *ids, int nslots, double *out)
{
  const int cell = blockIdx.x * blockDim.x + threadIdx.x;
  if (cell >= ncell) return;
  for (int s = 0; s < nslots; s++) out[(size_t)s * ncell + cell] = conc[(size_t)ids[s] * ncell + cell] / dt;
}

__global__ void oh_reactivity_kernel(const double *conc, const double *rconst, int ncell, const double *coef,
                                     const int *rxn, const int *spc, int nterms, double *out)
{
  const int cell = blockIdx.x * blockDim.x + threadIdx.x;
  if (cell >= ncell) return;
  double acc = 0.0;
  for (int t = 0; t < nterms; t++) {
    const double c = __ldg(coef + t);
    double v = rconst[(size_t)__ldg(rxn + t) * ncell + cell];
    if (c != 1.0) v = c * v;                                        // "2*RR(18)"
    const int s = __ldg(spc + t);
    if (s >= 0) v = v * conc[(size_t)s * ncell + cell];            // "RR(12)*CC(89)"
    acc = (t == 0) ? v : acc + v;
  }
  out[cell] = acc;
}

}  // namespace

cudaError_t launch_zero_species(double *conc, int ncell, const int *ids, int n, cudaStream_t s)
{
  if (ncell > 0 && n > 0) zero_species_kernel<<<(ncell + 255) / 256, 256, 0, s>>>(conc, ncell, ids, n);
  return cudaGetLastError();
}
cudaError_t launch_post_integrate(double *conc, int ncell, int nspec, const int *scale_ids, const double *scale_div, int nscale,
                                  const unsigned char *mask, float *negatives, cudaStream_t s)
{
  if (ncell > 0) post_integrate_kernel<<<(ncell + 255) / 256, 256, 0, s>>>(conc, ncell, nspec, scale_ids, scale_div, nscale, mask, negatives);
  return cudaGetLastError();
}
cudaError_t launch_prod_loss(const double *conc, int ncell, double dt, const int *ids, int nslots, double *out, cudaStream_t s)
{
  if (ncell > 0 && nslots > 0) prod_loss_kernel<<<(ncell + 255) / 256, 256, 0, s>>>(conc, ncell, dt, ids, nslots, out);
  return cudaGetLastError();
}
cudaError_t launch_oh_reactivity(const double *conc, const double *rconst, int ncell, const double *coef, const int *rxn,
                                 const int *spc, int nterms, double *out, cudaStream_t s)
{
  if (ncell > 0) oh_reactivity_kernel<<<(ncell + 255) / 256, 256, 0, s>>>(conc, rconst, ncell, coef, rxn, spc, nterms, out);
  return cudaGetLastError();
}
