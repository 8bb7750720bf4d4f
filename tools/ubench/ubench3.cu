// Variants of the narrow-level loop, to find the cheapest structure (cycles per level, 1 warp working).
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#define GB(off) (*(const double *)(Gb + (off)))
template <int MODE>
__global__ void k_level(double *o, long long *t, int T, int lg, int levels, const uint4 *gtab)
{
  // MODE 0: group barrier (4 warps) per level, table row from shared memory
  // MODE 1: one warp, __syncwarp per level, table row from shared memory
  // MODE 2: one warp, __syncwarp, next level's table rows prefetched into registers (from shared memory)
  // MODE 3: like 2, table rows from global memory (L2), prefetched one level ahead
  // MODE 4: like 2 but 4 warps with bar.sync (prefetch + group barrier)
  extern __shared__ __align__(16) unsigned char sm[];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  double *G = (double *)sm;
  uint4 *tab = (uint4 *)(sm + 40960);     // 32 rows x 32 lanes
  for (int i = tid; i < 5100; i += blockDim.x) G[i] = 1.0 + 1e-9 * i;
  for (int i = tid; i < 32 * 32; i += blockDim.x) tab[i] = gtab[i];
  __syncthreads();
  const unsigned char *Gb = (const unsigned char *)G;
  const bool bar = (MODE == 0 || MODE == 4);
  if (!bar && warp > 0) return;
  long long t0 = clock64();
  double accsum = 0;
  uint4 w0 = tab[lane], w1 = tab[32 + lane];
  for (int lev = 0; lev < levels; lev++) {
    if (!bar || warp == 0) {
      int row = (2 * lev) & 31;
      uint4 a, b;
      if (MODE == 0 || MODE == 1) { a = tab[row * 32 + lane]; b = tab[(row + 1) * 32 + lane]; }
      else {
        a = w0; b = w1;
        int nrow = (2 * lev + 2) & 31;
        if (MODE == 3) { w0 = __ldcg(gtab + nrow * 32 + lane); w1 = __ldcg(gtab + (nrow + 1) * 32 + lane); }
        else { w0 = tab[nrow * 32 + lane]; w1 = tab[(nrow + 1) * 32 + lane]; }
      }
      double a0, a1, a2 = 0, a3 = 0;
      const double old = G[5000 + lane];
      {
        const double h0 = GB(a.x >> 16), l0 = GB(a.x & 0xffff), h1 = GB(a.y >> 16), l1 = GB(a.y & 0xffff);
        const double h2 = GB(a.z >> 16), l2 = GB(a.z & 0xffff), h3 = GB(a.w >> 16), l3 = GB(a.w & 0xffff);
        a0 = h0 * l0; a1 = h1 * l1; a2 = h2 * l2; a3 = h3 * l3;
        if (T > 4) {
          const double h4 = GB(b.x >> 16), l4 = GB(b.x & 0xffff), h5 = GB(b.y >> 16), l5 = GB(b.y & 0xffff);
          const double h6 = GB(b.z >> 16), l6 = GB(b.z & 0xffff), h7 = GB(b.w >> 16), l7 = GB(b.w & 0xffff);
          a0 = fma(h4, l4, a0); a1 = fma(h5, l5, a1); a2 = fma(h6, l6, a2); a3 = fma(h7, l7, a3);
        }
      }
      double acc = (a0 + a1) + (a2 + a3);
      for (int s = 0; s < lg; s++) acc += __shfl_down_sync(0xffffffffu, acc, 1 << s);
      G[5000 + ((lane + 1) & 31)] = old - acc * 1e-30;      // the next level reads what this one wrote
      accsum += acc;
    }
    if (bar) asm volatile("bar.sync 1, 128;" ::: "memory");
    else __syncwarp();
  }
  long long t1 = clock64();
  if (tid == 0) t[0] = t1 - t0;
  o[tid] = accsum;
}
int main()
{
  double *o; long long *t; cudaMalloc(&o, 1 << 16); cudaMalloc(&t, 64);
  uint4 *gtab; cudaMalloc(&gtab, 32 * 32 * 16);
  uint4 *h = new uint4[1024];
  for (int i = 0; i < 1024; i++) {
    int l = i & 31, r = i >> 5;
    unsigned a = ((l + r * 37) % 4990) * 8, b = ((l + r * 91 + 7) % 4990) * 8;
    // the first term of every lane reads what the previous level wrote
    h[i] = make_uint4(((5000u + l) * 8 << 16) | b, (b << 16) | a, ((a + 8) << 16) | (b + 8), ((b + 16) << 16) | (a + 16));
  }
  cudaMemcpy(gtab, h, 1024 * 16, cudaMemcpyHostToDevice);
  long long hc;
  const int levels = 4000;
#define RUN(M, thr) { cudaFuncSetAttribute(k_level<M>, cudaFuncAttributeMaxDynamicSharedMemorySize, 60000); \
    for (int T : {4, 8}) for (int lg : {0, 3, 5}) { k_level<M><<<148, thr, 60000>>>(o, t, T, lg, levels, gtab); \
      cudaError_t e = cudaDeviceSynchronize(); if (e) { printf("error %s\n", cudaGetErrorString(e)); return 1; } \
      cudaMemcpy(&hc, t, 8, cudaMemcpyDeviceToHost); printf("mode %d T %d lg %d: %7.1f cycles per level\n", M, T, lg, (double)hc / levels); } }
  RUN(0, 128) RUN(1, 128) RUN(2, 128) RUN(3, 128) RUN(4, 128)
  return 0;
}
