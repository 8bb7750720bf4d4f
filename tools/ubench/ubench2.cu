// Micro-benchmark of one "narrow dependency level" of the warp-group kernel: one warp of a 4-warp group reads a
// table row from shared memory, gathers 2*T operands, accumulates, reduces over 2^lg lanes by shuffles, stores,
// then the group barrier; the other three warps only take the barrier.  3 groups per block, like the kernel.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__global__ void k_level(double *o, long long *t, int T, int lg, int active_groups, int levels, int conflict)
{
  extern __shared__ __align__(16) unsigned char sm[];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, group = warp >> 2, wig = warp & 3;
  double *G = (double *)(sm + group * 57344);
  uint4 *tab = (uint4 *)(sm + group * 57344 + 40960);     // 32 rows x 32 lanes
  for (int i = tid & 127; i < 5100; i += 128) G[i] = 1.0 + 1e-9 * i;
  for (int i = tid & 127; i < 32 * 32; i += 128) {
    int l = i & 31, r = i >> 5;
    unsigned a = ((l * (conflict ? 16 : 1) + r * 37) % 5000) * 8, b = ((l * (conflict ? 16 : 1) + r * 91 + 7) % 5000) * 8;
    tab[i] = make_uint4((a << 16) | b, (b << 16) | a, ((a + 8) << 16) | (b + 8), ((b + 16) << 16) | (a + 16));
  }
  __syncthreads();
  if (group >= active_groups) return;
  long long t0 = clock64();
  double accsum = 0;
  for (int lev = 0; lev < levels; lev++) {
    if (wig == (lev & 3)) {
      double a0 = 0, a1 = 0;
      int row = lev & 31;
      for (int k = 0; k < T; k += 4) {
        uint4 w = tab[row * 32 + lane];
        row = (row + 1) & 31;
        const unsigned char *Gb = (const unsigned char *)G;
        a0 = fma(*(const double *)(Gb + (w.x >> 16)), *(const double *)(Gb + (w.x & 0xffff)), a0);
        a1 = fma(*(const double *)(Gb + (w.y >> 16)), *(const double *)(Gb + (w.y & 0xffff)), a1);
        a0 = fma(*(const double *)(Gb + (w.z >> 16)), *(const double *)(Gb + (w.z & 0xffff)), a0);
        a1 = fma(*(const double *)(Gb + (w.w >> 16)), *(const double *)(Gb + (w.w & 0xffff)), a1);
      }
      double acc = a0 + a1;
      for (int s = 0; s < lg; s++) acc += __shfl_down_sync(0xffffffffu, acc, 1 << s);
      G[5000 + lane] = G[5000 + lane] - acc * 1e-30;
      accsum += acc;
    }
    asm volatile("bar.sync %0, 128;" :: "r"(group + 1) : "memory");
  }
  long long t1 = clock64();
  if (tid == 0) t[0] = t1 - t0;
  o[tid] = accsum;
}
int main()
{
  double *o; long long *t; cudaMalloc(&o, 1 << 16); cudaMalloc(&t, 64);
  cudaFuncSetAttribute(k_level, cudaFuncAttributeMaxDynamicSharedMemorySize, 172032);
  long long h;
  const int levels = 2000;
  for (int ag : {1, 3}) for (int conflict : {0, 1}) for (int T : {4, 8}) for (int lg : {0, 3, 5}) {
    k_level<<<148, 384, 172032>>>(o, t, T, lg, ag, levels, conflict);
    cudaError_t e = cudaDeviceSynchronize(); if (e) { printf("error %s\n", cudaGetErrorString(e)); return 1; } cudaMemcpy(&h, t, 8, cudaMemcpyDeviceToHost);
    printf("groups %d conflict %d T %d lg %d: %7.1f cycles per level\n", ag, conflict, T, lg, (double)h / levels);
  }
  return 0;
}
