// micro-benchmark: warp-level Gauss-Jordan [S | v] (6x6, 12x12) cold and warm, clock64 deltas
#include <cstdio>
#include <cuda_runtime.h>
constexpr unsigned FULL = 0xffffffffu;
__device__ __forceinline__ double fast_rcp(double p) {
  double r = (double)__frcp_rn((float)p);
  r = fma(r, fma(-p, r, 1.0), r);
  r = fma(r, fma(-p, r, 1.0), r);
  return r;
}
template <int N, int NR>
__device__ __forceinline__ void warp_gauss_jordan(double (&a)[N], double (&b)[NR], int lane) {
#pragma unroll
  for (int k = 0; k < N; ++k) {
    const double pinv = fast_rcp(__shfl_sync(FULL, a[k], k));
    const double f = (lane == k) ? 0.0 : a[k] * pinv;
#pragma unroll
    for (int c = k + 1; c < N; ++c) a[c] = fma(-f, __shfl_sync(FULL, a[c], k), a[c]);
#pragma unroll
    for (int c = 0; c < NR; ++c) b[c] = fma(-f, __shfl_sync(FULL, b[c], k), b[c]);
  }
  double d = 1.0;
#pragma unroll
  for (int i = 0; i < N; ++i)
    if (lane == i) d = a[i];
  const double dinv = fast_rcp(d);
#pragma unroll
  for (int c = 0; c < NR; ++c) b[c] *= dinv;
}
template <int N, int NR>
__global__ void k(const double* S, double* out, long long* t) {
  const int lane = threadIdx.x & 31;
  __shared__ double sS[144];
  for (int i = threadIdx.x; i < N * N; i += blockDim.x) sS[i] = S[i];
  __syncthreads();
  if (threadIdx.x >= 32) return;
  for (int rep = 0; rep < 4; ++rep) {
    double a[N], b[NR];
    long long t0 = clock64();
#pragma unroll
    for (int j = 0; j < N; ++j) a[j] = lane < N ? sS[lane * N + j] + rep : 0.0;
#pragma unroll
    for (int j = 0; j < NR; ++j) b[j] = lane + j + 1.0;
    warp_gauss_jordan<N, NR>(a, b, lane);
    long long t1 = clock64();
    double s = 0;
#pragma unroll
    for (int j = 0; j < NR; ++j) s += b[j];
    out[rep * 32 + lane] = s;
    if (lane == 0) t[rep] = t1 - t0;
  }
}
int main() {
  double h[144];
  for (int i = 0; i < 12; ++i)
    for (int j = 0; j < 12; ++j) h[i * 12 + j] = (i == j ? 20.0 : 0.0) + 1.0 / (1 + i + j);
  double *S, *o; long long* t;
  cudaMalloc(&S, sizeof(h)); cudaMalloc(&o, 8 * 4096); cudaMalloc(&t, 64);
  double h6[36];
  for (int i = 0; i < 6; ++i) for (int j = 0; j < 6; ++j) h6[i * 6 + j] = h[i * 12 + j];
  long long ht[4];
  cudaMemcpy(S, h6, sizeof(h6), cudaMemcpyHostToDevice);
  k<6, 1><<<1, 256>>>(S, o, t); cudaMemcpy(ht, t, 32, cudaMemcpyDeviceToHost);
  printf("GJ<6,1>  cycles: first %lld then %lld %lld %lld\n", ht[0], ht[1], ht[2], ht[3]);
  k<6, 6><<<1, 256>>>(S, o, t); cudaMemcpy(ht, t, 32, cudaMemcpyDeviceToHost);
  printf("GJ<6,6>  cycles: first %lld then %lld %lld %lld\n", ht[0], ht[1], ht[2], ht[3]);
  cudaMemcpy(S, h, sizeof(h), cudaMemcpyHostToDevice);
  k<12, 1><<<1, 256>>>(S, o, t); cudaMemcpy(ht, t, 32, cudaMemcpyDeviceToHost);
  printf("GJ<12,1> cycles: first %lld then %lld %lld %lld\n", ht[0], ht[1], ht[2], ht[3]);
  k<12, 12><<<1, 256>>>(S, o, t); cudaMemcpy(ht, t, 32, cudaMemcpyDeviceToHost);
  printf("GJ<12,12> cycles: first %lld then %lld %lld %lld\n", ht[0], ht[1], ht[2], ht[3]);
  printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
  return 0;
}
