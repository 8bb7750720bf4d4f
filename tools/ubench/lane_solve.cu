// Micro-benchmark of the lane kernel's stream engine: the triangular-solve stream (sequential G from HBM,
// x in shared memory) with synthetic tables of the fullchem size.  Prints GB/s and cycles per cell-solve.
// Build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -I geos_chem_b200/csrc -o tools/ubench/lane_solve tools/ubench/lane_solve.cu
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "lane_engine.cuh"
#ifndef MODE
#define MODE 0
#endif

#define W_VALID (1u << 20)
#define W_LAST (1u << 18)
#define W_RINV (1u << 19)

struct SolveF {
  const double *ga; double *x; double s;
  int ib;
  __device__ __forceinline__ void first(const uint4 *, const double *) { }
  __device__ __forceinline__ void issue(const uint4 *rec, unsigned dst)
  {
    const int batch = ib++;
#if MODE == 1
    return;
#endif
    const uint4 w = rec[0];
    const double *p = ga + (size_t)batch * 4 * 32;
#ifdef PAIRS
    // [pair][lane][2] layout: two 16-byte copies per lane and batch, L1 bypassed
    const int lane = threadIdx.x & 31;
    const double *pp = p - lane + 2 * lane;
    const unsigned dd = dst - 8 * lane + 16 * lane;
    if (w.x & W_VALID) asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dd), "l"(pp) : "memory");
    if (w.z & W_VALID) asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dd + 512), "l"(pp + 64) : "memory");
#else
    if (w.x & W_VALID) cp_async8(dst, p);
    if (w.y & W_VALID) cp_async8(dst + 256, p + 32);
    if (w.z & W_VALID) cp_async8(dst + 512, p + 64);
    if (w.w & W_VALID) cp_async8(dst + 768, p + 96);
#endif
  }
  __device__ __forceinline__ void one(unsigned w, double g)
  {
    if (!(w & W_VALID)) return;
    if (!(w & W_RINV)) s = fma(g, x[(w & 511) * 32], s);
    if (w & W_LAST) {
      const int i = (w >> 9) & 511;
      double v = x[i * 32] - s;
      if (w & W_RINV) v *= g;
      x[i * 32] = v;
      s = 0.0;
    }
  }
  __device__ __forceinline__ void consume(const uint4 *rec, const double *g, const uint4 *, const double *)
  {
#if MODE == 2
    s += g[0] + g[32] + g[64] + g[96];
    return;
#endif
    const uint4 w = rec[0];
    one(w.x, g[0]); one(w.y, g[32]); one(w.z, g[64]); one(w.w, g[96]);
  }
};

__global__ void __launch_bounds__(32) k_solve(const uint4 *tab, int nchunk, const double *ga, size_t ga_stride, int nv, int reps, double *out, long long *cyc)
{
  extern __shared__ __align__(128) unsigned char sm[];
  double *x = (double *)sm;
  uint4 *ring = (uint4 *)(sm + (size_t)nv * 256);
  double *dring = (double *)(sm + (size_t)nv * 256 + LANE_TRING_BYTES(2));
  const int lane = threadIdx.x;
  for (int i = 0; i < nv; i++) x[i * 32 + lane] = 1.0 + 1e-3 * i;
  SolveF f{ga + (size_t)blockIdx.x * ga_stride + lane, x + lane, 0.0, 0};
  long long t0 = clock64();
  for (int r = 0; r < reps; r++) { f.ib = 0; run_stream<2>(tab, nchunk, ring, dring, f); }
  long long t1 = clock64();
  double acc = 0;
  for (int i = 0; i < nv; i++) acc += x[i * 32 + lane];
  out[blockIdx.x * 32 + lane] = acc;
  if (lane == 0) cyc[blockIdx.x] = t1 - t0;
}

int main(int argc, char **argv)
{
  const int N = 353, NG = 5684, nb = (NG + 3) / 4, nchunk = (nb + 15) / 16, reps = argc > 1 ? atoi(argv[1]) : 8;
  int bps = argc > 2 ? atoi(argv[2]) : 2;
  std::vector<uint4> tab((size_t)nchunk * 16 * 2, uint4{0, 0, 0, 0});
  srand(1);
  for (int e = 0; e < NG; e++) {
    const bool fwd = e < 3128;
    unsigned w = W_VALID | (rand() % N);
    if (e % 8 == 7) { w |= W_LAST | ((rand() % N) << 9); if (fwd) w |= W_RINV; }
    unsigned *p = (unsigned *)&tab[(size_t)(e / 4) * 2];
    p[e % 4] = w;
  }
  cudaDeviceProp prop; cudaGetDeviceProperties(&prop, 0);
  const int blocks = bps > 0 ? prop.multiProcessorCount * bps : 1;
  const size_t ga_stride = (size_t)nchunk * 64 * 32;
  double *ga, *out; uint4 *dtab; long long *cyc;
  cudaMalloc(&ga, ga_stride * blocks * 8); cudaMalloc(&out, blocks * 32 * 8); cudaMalloc(&dtab, tab.size() * 16); cudaMalloc(&cyc, blocks * 8);
  std::vector<double> h(ga_stride, 1e-4);
  for (int b = 0; b < blocks; b++) cudaMemcpy(ga + b * ga_stride, h.data(), ga_stride * 8, cudaMemcpyHostToDevice);
  cudaMemcpy(dtab, tab.data(), tab.size() * 16, cudaMemcpyHostToDevice);
  const int nv = N + 8, smem = nv * 256 + LANE_TRING_BYTES(2) + LANE_DRING_BYTES;
  cudaFuncSetAttribute(k_solve, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int it = 0; it < 3; it++) {
    cudaEventRecord(e0);
    k_solve<<<blocks, 32, smem>>>(dtab, nchunk, ga, ga_stride, nv, reps, out, cyc);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    std::vector<long long> c(blocks); cudaMemcpy(c.data(), cyc, blocks * 8, cudaMemcpyDeviceToHost);
    double bytes = (double)blocks * reps * NG * 256.0;
    printf("%s blocks %d (x%d/SM) reps %d: %.3f ms, %.1f GB/s of G, %.0f cycles per warp-solve (block 0), %.0f per cell-solve per SM\n",
           cudaGetErrorString(cudaGetLastError()), blocks, bps, reps, ms, bytes / ms * 1e-6, (double)c[0] / reps, (double)c[0] / reps / 32 / (bps > 0 ? bps : 1));
  }
  return 0;
}
