// Micro-benchmarks of the sm_100a latencies the warp-group kernel is bound by (one warp unless stated).
// Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o ubench ubench.cu ; run on the GPU box.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#define N 256
__global__ void k_dfma(double *o, long long *t, double a, double b) {
  double x = o[0];
  long long t0 = clock64();
#pragma unroll
  for (int i = 0; i < N; i++) x = fma(x, a, b);
  long long t1 = clock64();
  o[1] = x; if (threadIdx.x == 0) t[0] = t1 - t0;
}
__global__ void k_dfma4(double *o, long long *t, double a, double b) {   // 4 independent chains
  double x0 = o[0], x1 = o[1], x2 = o[2], x3 = o[3];
  long long t0 = clock64();
#pragma unroll
  for (int i = 0; i < N; i++) { x0 = fma(x0, a, b); x1 = fma(x1, a, b); x2 = fma(x2, a, b); x3 = fma(x3, a, b); }
  long long t1 = clock64();
  o[4] = x0 + x1 + x2 + x3; if (threadIdx.x == 0) t[0] = t1 - t0;
}
__global__ void k_dadd(double *o, long long *t, double a) {
  double x = o[0];
  long long t0 = clock64();
#pragma unroll
  for (int i = 0; i < N; i++) x = x + a;
  long long t1 = clock64();
  o[1] = x; if (threadIdx.x == 0) t[0] = t1 - t0;
}
__global__ void k_div(double *o, long long *t, double a) {
  double x = o[0];
  long long t0 = clock64();
#pragma unroll
  for (int i = 0; i < 64; i++) x = 1.0 / x + a;
  long long t1 = clock64();
  o[1] = x; if (threadIdx.x == 0) t[0] = t1 - t0;
}
__global__ void k_lds(double *o, long long *t) {
  __shared__ int s[1024];
  for (int i = threadIdx.x; i < 1024; i += 32) s[i] = (i * 37 + 11) & 1023;
  __syncwarp();
  int idx = threadIdx.x;
  long long t0 = clock64();
#pragma unroll
  for (int i = 0; i < N; i++) idx = s[idx];
  long long t1 = clock64();
  o[threadIdx.x] = idx; if (threadIdx.x == 0) t[0] = t1 - t0;
}
__global__ void k_lds64(double *o, long long *t) {     // dependent 64-bit loads: index from the loaded double
  __shared__ double s[1024];
  for (int i = threadIdx.x; i < 1024; i += 32) s[i] = (double)((i * 37 + 11) & 1023);
  __syncwarp();
  double v = (double)threadIdx.x;
  long long t0 = clock64();
#pragma unroll
  for (int i = 0; i < N; i++) v = s[(int)v];
  long long t1 = clock64();
  o[threadIdx.x] = v; if (threadIdx.x == 0) t[0] = t1 - t0;
}
__global__ void k_shfl(double *o, long long *t) {
  int v = threadIdx.x;
  long long t0 = clock64();
#pragma unroll
  for (int i = 0; i < N; i++) v = __shfl_sync(0xffffffffu, v, (v + 1) & 31);
  long long t1 = clock64();
  o[threadIdx.x] = v; if (threadIdx.x == 0) t[0] = t1 - t0;
}
__global__ void k_shfld(double *o, long long *t, double a) {   // the chain of the tail sweeps: shuffle a double, fma
  double x = o[threadIdx.x];
  long long t0 = clock64();
#pragma unroll
  for (int i = 0; i < N; i++) { double xj = __shfl_sync(0xffffffffu, x, i & 31); x = fma(-a, xj, x); }
  long long t1 = clock64();
  o[threadIdx.x] = x; if (threadIdx.x == 0) t[0] = t1 - t0;
}
__global__ void k_syncwarp(double *o, long long *t) {
  __shared__ double s[64];
  s[threadIdx.x] = threadIdx.x;
  long long t0 = clock64();
#pragma unroll
  for (int i = 0; i < N; i++) { s[(threadIdx.x + i) & 31] += 1.0; __syncwarp(); }
  long long t1 = clock64();
  o[threadIdx.x] = s[threadIdx.x]; if (threadIdx.x == 0) t[0] = t1 - t0;
}
__global__ void k_bar(double *o, long long *t, int per) {       // named barrier among `per` threads, blockDim/per groups
  int g = threadIdx.x / per;
  long long t0 = clock64();
  for (int i = 0; i < N; i++) asm volatile("bar.sync %0, %1;" :: "r"(g + 1), "r"(per) : "memory");
  long long t1 = clock64();
  if (threadIdx.x == 0) t[0] = t1 - t0;
}
__global__ void k_cpasync(const uint4 *tab, int rows, long long *t, int depth) {   // ring fetch of 512-byte rows from L2
  extern __shared__ uint4 ring[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  uint32_t base = (uint32_t)__cvta_generic_to_shared(ring + warp * 16 * 32 + lane);
  int ir = (blockIdx.x * 131 + warp * 977) % rows, is = 0, cs = 0;
  unsigned acc = 0;
  auto issue = [&]() {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n\tcp.async.commit_group;" :: "r"(base + is * 512), "l"(tab + (size_t)ir * 32 + lane) : "memory");
    if (++ir == rows) ir = 0; if (++is == 16) is = 0;
  };
  for (int i = 0; i < depth; i++) issue();
  long long t0 = clock64();
  for (int i = 0; i < 512; i++) {
    if (depth == 1) asm volatile("cp.async.wait_group 0;" ::: "memory");
    else if (depth == 3) asm volatile("cp.async.wait_group 2;" ::: "memory");
    else if (depth == 5) asm volatile("cp.async.wait_group 4;" ::: "memory");
    else if (depth == 7) asm volatile("cp.async.wait_group 6;" ::: "memory");
    else asm volatile("cp.async.wait_group 11;" ::: "memory");
    uint4 v;
    asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(base + cs * 512) : "memory");
    acc += v.x + v.w;
    if (++cs == 16) cs = 0;
    issue();
  }
  long long t1 = clock64();
  asm volatile("cp.async.wait_all;" ::: "memory");
  if (threadIdx.x == 0 && blockIdx.x == 0) { t[0] = t1 - t0; t[1] = acc; }
}
__global__ void k_ldsconf(double *o, long long *t, int stride) {   // 64 independent 8-byte loads per lane, stride in doubles between lanes
  __shared__ double s[4096];
  for (int i = threadIdx.x; i < 4096; i += blockDim.x) s[i] = i;
  __syncthreads();
  double a = 0;
  int b0 = (threadIdx.x & 31) * stride;
  long long t0 = clock64();
#pragma unroll
  for (int i = 0; i < 64; i++) a += s[(b0 + i * 33) & 4095];
  long long t1 = clock64();
  o[threadIdx.x] = a; if (threadIdx.x == 0) t[0] = t1 - t0;
}
__global__ void k_ldg_l2(const double *p, int n, long long *t) {    // dependent loads through L2 (ld.global.cg)
  int idx = threadIdx.x;
  long long t0 = clock64();
  for (int i = 0; i < 64; i++) idx = (int)__ldcg(p + ((idx * 977 + i * 131) % n));
  long long t1 = clock64();
  if (threadIdx.x == 0) { t[0] = t1 - t0; t[1] = idx; }
}
int main() {
  double *o; long long *t; cudaMalloc(&o, 1 << 16); cudaMalloc(&t, 64); cudaMemset(o, 0, 1 << 16);
  long long h[2];
  auto rd = [&](const char *name, double per) { cudaDeviceSynchronize(); cudaMemcpy(h, t, 16, cudaMemcpyDeviceToHost); printf("%-28s %8lld cycles  %7.1f per op\n", name, h[0], h[0] / per); };
  for (int rep = 0; rep < 2; rep++) {
    k_dfma<<<1, 32>>>(o, t, 1.0000001, 1e-9); rd("dfma dependent chain", N);
    k_dfma4<<<1, 32>>>(o, t, 1.0000001, 1e-9); rd("dfma 4 chains (per fma)", 4 * N);
    k_dfma4<<<1, 128>>>(o, t, 1.0000001, 1e-9); rd("dfma 4 chains x 4 warps", 4 * N);
    k_dadd<<<1, 32>>>(o, t, 1e-9); rd("dadd dependent chain", N);
    k_div<<<1, 32>>>(o, t, 0.5); rd("1.0/x + a chain", 64);
    k_lds<<<1, 32>>>(o, t); rd("lds.32 dependent", N);
    k_lds64<<<1, 32>>>(o, t); rd("lds.64 + cvt dependent", N);
    k_shfl<<<1, 32>>>(o, t); rd("shfl.32 dependent", N);
    k_shfld<<<1, 32>>>(o, t, 1e-3); rd("shfl double + fma chain", N);
    k_syncwarp<<<1, 32>>>(o, t); rd("lds+sts+syncwarp", N);
    k_bar<<<1, 128>>>(o, t, 128); rd("bar.sync 128 thr (1 group)", N);
    k_bar<<<1, 384>>>(o, t, 128); rd("bar.sync 128 thr (3 groups)", N);
    k_bar<<<1, 384>>>(o, t, 384); rd("bar.sync 384 thr", N);
    for (int st : {1, 2, 4, 16}) { k_ldsconf<<<1, 32>>>(o, t, st); char nm[64]; snprintf(nm, 64, "lds.64 x64 indep, stride %d", st); rd(nm, 64); }
    k_ldsconf<<<1, 128>>>(o, t, 1); rd("lds.64 x64 indep, 4 warps", 64);
  }
  const int rows = 1200; uint4 *tab; cudaMalloc(&tab, (size_t)rows * 512); cudaMemset(tab, 1, (size_t)rows * 512);
  double *pd; cudaMalloc(&pd, 1 << 22); cudaMemset(pd, 0, 1 << 22);
  k_ldg_l2<<<1, 32>>>(pd, 1 << 19, t); rd("ld.global.cg dependent (1 warp)", 64);
  k_ldg_l2<<<148, 384>>>(pd, 1 << 19, t); rd("ld.global.cg dependent (148x12 warps)", 64);
  cudaFuncSetAttribute(k_cpasync, cudaFuncAttributeMaxDynamicSharedMemorySize, 12 * 16 * 512);
  for (int depth : {1, 3, 5, 7, 12}) for (int blocks : {1, 148}) for (int warps : {1, 12}) {
    k_cpasync<<<blocks, warps * 32, 12 * 16 * 512>>>(tab, rows, t, depth);
    char nm[96]; snprintf(nm, 96, "cp.async row, depth %d, %dx%d warps", depth, blocks, warps); rd(nm, 512);
  }
  return 0;
}
