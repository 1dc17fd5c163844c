// MUFU.EX2 throughput per SM sub-partition on sm_100a: independent ex2 chains, 1/2/4 warps per scheduler.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/mufu_bench tools/mufu_bench.cu && tools/mufu_bench
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ float ex2(float x) { float y; asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
template <int MIX>
__global__ void k(float* out, long long* cyc, int iters) {
  float a[8];
  for (int i = 0; i < 8; ++i) a[i] = -0.001f * (threadIdx.x + i);
  float f0 = 1.0f, f1 = 2.0f;
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      a[i] = ex2(a[i]);
      if (MIX >= 1) f0 = fmaf(f0, 1.0001f, a[i]);     // one dependent-free FMA per exp
      if (MIX >= 2) { f1 = fmaf(f1, 0.9999f, 0.5f); f0 = fmaf(f0, 0.9999f, 0.25f); }  // three per exp
    }
  }
  const long long t1 = clock64();
  float s = f0 + f1;
  for (int i = 0; i < 8; ++i) s += a[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
template <int MIX>
void run(int threads, int blocks, float* out, long long* cyc) {
  const int iters = 2000;
  k<MIX><<<blocks, threads>>>(out, cyc, iters);
  cudaDeviceSynchronize();
  k<MIX><<<blocks, threads>>>(out, cyc, iters);
  cudaDeviceSynchronize();
  long long h[1024];
  cudaMemcpy(h, cyc, sizeof(long long) * blocks, cudaMemcpyDeviceToHost);
  double mean = 0;
  for (int i = 0; i < blocks; ++i) mean += double(h[i]);
  mean /= blocks;
  const double warps_per_smsp = threads / 32 / 4.0;
  const double per_instr = mean / (iters * 8.0);          // cycles per MUFU instruction per warp
  printf("mix %d threads %4d blocks %4d: %.2f cycles per MUFU per warp, %.2f cycles per MUFU warp-instruction per scheduler (%.1f lanes/clk/SM)\n",
         MIX, threads, blocks, per_instr, per_instr / (warps_per_smsp < 1 ? 1 : warps_per_smsp), 128.0 / (per_instr / (warps_per_smsp < 1 ? 1 : warps_per_smsp)));
}
int main() {
  float* out; long long* cyc;
  cudaMalloc(&out, 1024 * 1024 * 4); cudaMalloc(&cyc, 1024 * 8);
  for (int blocks : {1, 148}) {
    for (int threads : {128, 256, 512, 1024}) { run<0>(threads, blocks, out, cyc); }
    for (int threads : {128, 256, 512}) { run<1>(threads, blocks, out, cyc); }
    for (int threads : {128, 256, 512}) { run<2>(threads, blocks, out, cyc); }
  }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
