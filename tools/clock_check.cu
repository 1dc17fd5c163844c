// clock64 ticks per globaltimer nanosecond: (a) one spinning warp per SM, (b) every SM busy with FMA + MUFU work.
#include <cstdio>
#include <cuda_runtime.h>
__global__ void spin(long long cycles, long long* out, int heavy) {
  unsigned long long g0, g1;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(g0));
  const long long c0 = clock64();
  float x = threadIdx.x * 1e-3f, y = 1.0f;
  while (clock64() - c0 < cycles) {
    if (heavy) {
#pragma unroll
      for (int i = 0; i < 64; ++i) { x = __expf(x * 0.5f) - y; y = fmaf(y, 0.999f, x * 1e-6f); }
    }
  }
  const long long c1 = clock64();
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(g1));
  if (threadIdx.x == 0 && blockIdx.x == 0) { out[0] = c1 - c0; out[1] = (long long)(g1 - g0); out[2] = (long long)(x + y); }
}
int main() {
  long long* d; cudaMalloc(&d, 32);
  long long h[3];
  for (int heavy = 0; heavy < 2; ++heavy)
    for (int rep = 0; rep < 3; ++rep) {
      spin<<<heavy ? 148 * 4 : 148, heavy ? 512 : 32>>>(20000000ll, d, heavy);
      cudaDeviceSynchronize();
      cudaMemcpy(h, d, 24, cudaMemcpyDeviceToHost);
      printf("heavy=%d: %lld cycles in %lld ns -> %.3f GHz\n", heavy, h[0], h[1], double(h[0]) / double(h[1]));
    }
  return 0;
}
