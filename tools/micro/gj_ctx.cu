// micro-benchmark: the warp-level Gauss-Jordan inside a context like the solver block's (7 warps parked at a barrier,
// divergent code before the call, noinline template, loop over passes)
#include <cstdio>
#include <cuda_runtime.h>
constexpr unsigned FULL = 0xffffffffu;
template <int N, int NR>
__device__ __forceinline__ void warp_gauss_jordan(double (&a)[N], double (&b)[NR], int lane) {
#pragma unroll
  for (int k = 0; k < N; ++k) {
    const double pinv = 1.0 / __shfl_sync(FULL, a[k], k);
    const bool piv = lane == k;
    const double f = a[k];
#pragma unroll
    for (int c = k + 1; c < N; ++c) {
      const double t = a[c] * pinv;
      const double pc = __shfl_sync(FULL, t, k);
      a[c] = piv ? t : fma(-f, pc, a[c]);
    }
#pragma unroll
    for (int c = 0; c < NR; ++c) {
      const double t = b[c] * pinv;
      const double pc = __shfl_sync(FULL, t, k);
      b[c] = piv ? t : fma(-f, pc, b[c]);
    }
  }
}
struct Sm { double S[144]; double v[12]; double out[32]; long long t[8]; };
template <int N>
__device__ __noinline__ void step(Sm* sm, int lane, int pass, int mode) {
  double a[N], v[1];
  long long c0 = clock64();
#pragma unroll
  for (int j = 0; j < N; ++j) a[j] = lane < N ? sm->S[lane * N + j] + pass : 0.0;
  {
    double acc = 0.0;
    if (lane < N) {
      acc = sm->v[lane];
#pragma unroll
      for (int j = 0; j < N; ++j) acc = fma(sm->S[lane * N + j], sm->v[j], acc);
    }
    v[0] = acc;
  }
  if (mode & 1) __syncwarp();
  long long c1 = clock64();
  warp_gauss_jordan<N, 1>(a, v, lane);
  long long c2 = clock64();
  sm->out[lane] = v[0];
  if (lane == 0) { sm->t[0] = c1 - c0; sm->t[1] = c2 - c1; }
  __syncwarp();
}
__global__ void __launch_bounds__(256, 2) k(const double* S, double* out, long long* t, int passes, int mode) {
  __shared__ Sm sm;
  for (int i = threadIdx.x; i < 36; i += blockDim.x) sm.S[i] = S[i];
  if (threadIdx.x < 12) sm.v[threadIdx.x] = 1.0 + threadIdx.x;
  __syncthreads();
  for (int p = 0; p < passes; ++p) {
    if (threadIdx.x < 32) step<6>(&sm, (int)threadIdx.x, p, mode);
    __syncthreads();
    if (threadIdx.x == 0) { t[2 * p] = sm.t[0]; t[2 * p + 1] = sm.t[1]; }
    if (threadIdx.x < 32) out[p * 32 + threadIdx.x] = sm.out[threadIdx.x];
    __syncthreads();
  }
}
int main() {
  double h6[36];
  for (int i = 0; i < 6; ++i) for (int j = 0; j < 6; ++j) h6[i * 6 + j] = (i == j ? 20.0 : 0.0) + 1.0 / (1 + i + j);
  double *S, *o; long long* t;
  cudaMalloc(&S, sizeof(h6)); cudaMalloc(&o, 8 * 4096); cudaMalloc(&t, 256);
  cudaMemcpy(S, h6, sizeof(h6), cudaMemcpyHostToDevice);
  long long ht[8];
  for (int mode = 0; mode < 2; ++mode) {
    k<<<1, 256>>>(S, o, t, 4, mode); cudaMemcpy(ht, t, 64, cudaMemcpyDeviceToHost);
    printf("mode %d (syncwarp before GJ: %d): prep/GJ cycles per pass: %lld/%lld %lld/%lld %lld/%lld %lld/%lld\n", mode, mode & 1,
           ht[0], ht[1], ht[2], ht[3], ht[4], ht[5], ht[6], ht[7]);
  }
  printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
  return 0;
}
