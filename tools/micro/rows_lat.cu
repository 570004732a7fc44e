// micro-benchmark: one block reads 256 rows written by other SMs: strong 128-bit loads vs ld.cg 128-bit vs ld.cg 64-bit
#include <cstdio>
#include <cuda_runtime.h>
constexpr int ROWW = 184;  // 64-bit words per row
__global__ void writer(unsigned long long* rows, unsigned stamp) {
  const int r = blockIdx.x;
  if (threadIdx.x < 92) {
    const double v = 1.0 + r + threadIdx.x;
    const unsigned long long b = (unsigned long long)__double_as_longlong(v), hs = (unsigned long long)stamp << 32;
    asm volatile("st.relaxed.gpu.global.v2.u64 [%0], {%1, %2};" ::"l"(rows + (size_t)r * ROWW + 2 * threadIdx.x),
                 "l"(hs | (b & 0xffffffffull)), "l"(hs | (b >> 32)) : "memory");
  }
}
template <int MODE, int CH>
__global__ void reader(const unsigned long long* rows, int nb, unsigned stamp, double* out, long long* t) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int e = (lane * 3) % 92;  // scattered output positions like the real table
  double acc = 0;
  long long t0 = clock64();
  for (int base = warp; base < nb; base += 8 * CH) {
    unsigned long long lo[CH], hi[CH];
#pragma unroll
    for (int j = 0; j < CH; ++j) {
      const int b = base + 8 * j < nb ? base + 8 * j : nb - 1;
      const unsigned long long* p = rows + (size_t)b * ROWW + 2 * e;
      if (MODE == 0) asm volatile("ld.relaxed.gpu.global.v2.u64 {%0, %1}, [%2];" : "=l"(lo[j]), "=l"(hi[j]) : "l"(p) : "memory");
      if (MODE == 1) asm volatile("ld.global.cg.v2.u64 {%0, %1}, [%2];" : "=l"(lo[j]), "=l"(hi[j]) : "l"(p) : "memory");
      if (MODE == 2) { asm volatile("ld.global.cg.u64 %0, [%1];" : "=l"(lo[j]) : "l"(p) : "memory"); hi[j] = lo[j]; }
      if (MODE == 3) { lo[j] = __ldcg(p); hi[j] = __ldcg(p + 1); }
    }
#pragma unroll
    for (int j = 0; j < CH; ++j) acc += __longlong_as_double((long long)((hi[j] << 32) | (lo[j] & 0xffffffffull)));
  }
  long long t1 = clock64();
  out[threadIdx.x] = acc;
  if (threadIdx.x == 255) t[0] = t1 - t0;
}
int main() {
  const int nb = 256;
  unsigned long long* rows; double* out; long long* t;
  cudaMalloc(&rows, (size_t)nb * ROWW * 8); cudaMalloc(&out, 8 * 256); cudaMalloc(&t, 64);
  char* flush; cudaMalloc(&flush, 256u << 20);
  long long h;
  const char* names[4] = {"ld.relaxed.gpu.v2.u64", "ld.cg.v2.u64", "ld.cg.u64 (8 B)", "__ldcg x2"};
  for (int mode = 0; mode < 4; ++mode)
    for (int rep = 0; rep < 3; ++rep) {
      writer<<<nb, 128>>>(rows, 7 + rep);
      if (mode == 0) reader<0, 40><<<1, 256>>>(rows, nb, 7, out, t);
      if (mode == 1) reader<1, 40><<<1, 256>>>(rows, nb, 7, out, t);
      if (mode == 2) reader<2, 40><<<1, 256>>>(rows, nb, 7, out, t);
      if (mode == 3) reader<3, 40><<<1, 256>>>(rows, nb, 7, out, t);
      cudaMemcpy(&h, t, 8, cudaMemcpyDeviceToHost);
      if (rep == 2) printf("%-24s CH=40: %lld cycles for 256 rows x 32 lanes\n", names[mode], h);
    }
  for (int rep = 0; rep < 3; ++rep) {
    writer<<<nb, 128>>>(rows, 7 + rep);
    reader<0, 10><<<1, 256>>>(rows, nb, 7, out, t);
    cudaMemcpy(&h, t, 8, cudaMemcpyDeviceToHost);
    if (rep == 2) printf("%-24s CH=10: %lld cycles\n", names[0], h);
  }
  printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
  return 0;
}
