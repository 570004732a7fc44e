// micro-benchmark: latency of dependent FP64 ops / double shuffles / reciprocal on one warp (clock64 deltas)
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ double fast_rcp(double p) {
  double r = (double)__frcp_rn((float)p);
  r = fma(r, fma(-p, r, 1.0), r);
  r = fma(r, fma(-p, r, 1.0), r);
  return r;
}
__global__ void k(double* out, long long* t, double seed) {
  const int lane = threadIdx.x & 31;
  double a = seed + lane, b = 1.0000001;
  long long t0 = clock64();
#pragma unroll
  for (int i = 0; i < 128; ++i) a = fma(a, b, 0.5);
  long long t1 = clock64();
  double c = a;
#pragma unroll
  for (int i = 0; i < 64; ++i) c = __shfl_sync(0xffffffffu, c, (lane + 1) & 31) + 1.0;
  long long t2 = clock64();
  double d = c;
#pragma unroll
  for (int i = 0; i < 16; ++i) d = fast_rcp(d) + 2.0;
  long long t3 = clock64();
  double e = d;
#pragma unroll
  for (int i = 0; i < 16; ++i) e = 1.0 / e + 2.0;
  long long t4 = clock64();
  float f = (float)e;
#pragma unroll
  for (int i = 0; i < 128; ++i) f = fmaf(f, 1.0000001f, 0.5f);
  long long t5 = clock64();
  // 128 INDEPENDENT dfma (8 chains x 16)
  double g[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) g[j] = e + j;
  long long t6 = clock64();
#pragma unroll
  for (int i = 0; i < 16; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) g[j] = fma(g[j], b, 0.25);
  long long t7 = clock64();
  double gs = 0;
  for (int j = 0; j < 8; ++j) gs += g[j];
  out[threadIdx.x] = a + c + d + e + f + gs;
  if (threadIdx.x == 0) { t[0] = t1 - t0; t[1] = t2 - t1; t[2] = t3 - t2; t[3] = t4 - t3; t[4] = t5 - t4; t[5] = t7 - t6; }
}
int main() {
  double* o; long long* t;
  cudaMalloc(&o, 8 * 1024); cudaMalloc(&t, 64);
  for (int rep = 0; rep < 3; ++rep) k<<<1, 32>>>(o, t, 1.5);
  long long h[6];
  cudaMemcpy(h, t, 48, cudaMemcpyDeviceToHost);
  printf("1 warp: dep DFMA %.1f cyc/op | double shfl+add %.1f cyc/op | fast_rcp+add %.1f | 1.0/x+add %.1f | dep FFMA %.1f | indep DFMA %.1f cyc/op\n",
         h[0] / 128.0, h[1] / 64.0, h[2] / 16.0, h[3] / 16.0, h[4] / 128.0, h[5] / 128.0);
  for (int rep = 0; rep < 3; ++rep) k<<<1, 256>>>(o, t, 1.5);
  cudaMemcpy(h, t, 48, cudaMemcpyDeviceToHost);
  printf("8 warps: dep DFMA %.1f cyc/op | double shfl+add %.1f cyc/op | fast_rcp+add %.1f | 1.0/x+add %.1f | dep FFMA %.1f | indep DFMA %.1f cyc/op\n",
         h[0] / 128.0, h[1] / 64.0, h[2] / 16.0, h[3] / 16.0, h[4] / 128.0, h[5] / 128.0);
  int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0); printf("clock %d kHz\n", clk);
  return 0;
}
