// Micro-benchmark: how fast can 148 CTAs pull GEMM operand tiles from L2 into shared memory with TMA and nothing else?
//   build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I oron_tts_b200/csrc tools/tma_bw_bench.cu -o tools/tma_bw_bench -lcuda
//   run  : tools/tma_bw_bench
// Question: the 2-SM GEMM main loop runs at ~490 ns per 64-wide k-block under sustained load, each CTA
// fetching 32 KB per k-block (its 128 x 64 A rows + 128 x 64 of the B tile) = ~9.7 TB/s over 148 SMs, with producer
// AND consumer both waiting on the ring. Is that the L2 -> SM fabric? Every CTA walks the same tile order as the GEMM
// (A [2816, K] and W [4096, K], both L2-resident), a consumer thread frees each stage as soon as it lands.
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cuda.h>
#include <cudaTypedefs.h>
#include <cuda_runtime.h>

#include "ptx.cuh"

using namespace oron;

template <int STAGES>
__global__ void __launch_bounds__(64, 1) bw_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                                                   int units, int nkb, int tiles_m, long long* out) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t bar = base + STAGES * 32768;
  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) { mbar_init(bar + 8 * s, 1); mbar_init(bar + 8 * (STAGES + s), 1); }
    fence_mbar_init();
  }
  __syncthreads();
  const long long t0 = clock64();
  if (threadIdx.x == 0) {
    int stage = 0; uint32_t phase = 0;
    for (int u = 0; u < units; ++u) {
      const int tile = (blockIdx.x + (u / nkb) * gridDim.x);
      const int kb = u % nkb;
      const int m0 = (tile % tiles_m) * 128, n0 = ((tile / tiles_m) % 32) * 128;
      mbar_wait(bar + 8 * (STAGES + stage), phase ^ 1u, 1);
      mbar_arrive_expect_tx(bar + 8 * stage, 32768);
      tma_load_2d(base + stage * 32768, &tmA, bar + 8 * stage, kb * 64, m0);
      tma_load_2d(base + stage * 32768 + 16384, &tmB, bar + 8 * stage, kb * 64, n0);
      if (++stage == STAGES) { stage = 0; phase ^= 1u; }
    }
  } else if (threadIdx.x == 32) {
    int stage = 0; uint32_t phase = 0;
    for (int u = 0; u < units; ++u) {
      mbar_wait(bar + 8 * stage, phase, 2);
      mbar_arrive(bar + 8 * (STAGES + stage));
      if (++stage == STAGES) { stage = 0; phase ^= 1u; }
    }
    out[blockIdx.x] = clock64() - t0;
  }
}

// Cluster of two: every CTA REQUESTS only half of each stage (rank 0 the A box, rank 1 the B box) and multicasts it into both
// CTAs, so each SM still RECEIVES 32 KB per stage. Is the 64 B/clk limit on what an SM requests or on what it receives?
__device__ __forceinline__ void tma_load_2d_mc(uint32_t dst, const CUtensorMap* m, uint32_t bar, int c0, int c1, uint16_t mask) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, {%3, %4}], [%2], %5;"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(m)), "r"(bar), "r"(c0), "r"(c1), "h"(mask) : "memory");
}
__device__ __forceinline__ void mbar_arrive_rank(uint32_t bar, uint32_t rank) {
  uint32_t remote;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(bar), "r"(rank));
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(remote) : "memory");
}
template <int STAGES>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(64, 1)
bw_mc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, int units, int nkb, int tiles_m,
             long long* out) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t bar = base + STAGES * 32768;
  const uint32_t rank = cluster_ctarank();
  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) { mbar_init(bar + 8 * s, 1); mbar_init(bar + 8 * (STAGES + s), 2); }
    fence_mbar_init();
  }
  cluster_sync_all();
  const long long t0 = clock64();
  const int pair = blockIdx.x >> 1;
  if (threadIdx.x == 0) {
    int stage = 0; uint32_t phase = 0;
    for (int u = 0; u < units; ++u) {
      const int tile = (pair + (u / nkb) * (gridDim.x >> 1));
      const int kb = u % nkb;
      const int m0 = (tile % tiles_m) * 128, n0 = ((tile / tiles_m) % 32) * 128;
      mbar_wait(bar + 8 * (STAGES + stage), phase ^ 1u, 1);  // both CTAs have released this stage
      mbar_arrive_expect_tx(bar + 8 * stage, 32768);
      if (rank == 0) tma_load_2d_mc(base + stage * 32768, &tmA, bar + 8 * stage, kb * 64, m0, 3);
      else tma_load_2d_mc(base + stage * 32768 + 16384, &tmB, bar + 8 * stage, kb * 64, n0, 3);
      if (++stage == STAGES) { stage = 0; phase ^= 1u; }
    }
  } else if (threadIdx.x == 32) {
    int stage = 0; uint32_t phase = 0;
    for (int u = 0; u < units; ++u) {
      mbar_wait(bar + 8 * stage, phase, 2);
      mbar_arrive_rank(bar + 8 * (STAGES + stage), 0);
      mbar_arrive_rank(bar + 8 * (STAGES + stage), 1);
      if (++stage == STAGES) { stage = 0; phase ^= 1u; }
    }
    out[blockIdx.x] = clock64() - t0;
  }
  cluster_sync_all();
}

static CUtensorMap make_map(void* p, uint64_t cols, uint64_t rows) {
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
  auto enc = reinterpret_cast<PFN_cuTensorMapEncodeTiled>(fn);
  CUtensorMap m;
  cuuint64_t dims[2] = {cols, rows};
  cuuint64_t strides[1] = {cols * 2};
  cuuint32_t box[2] = {64, 128};
  cuuint32_t es[2] = {1, 1};
  CUresult r = enc(&m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, p, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { printf("encode failed %d\n", int(r)); exit(1); }
  return m;
}

template <int STAGES>
static void run(const CUtensorMap& a, const CUtensorMap& b, int K, int grid) {
  const int nkb = K / 64, units = 16 * nkb * (1024 / K > 0 ? 1024 / K : 1) * 4;
  const int smem = STAGES * 32768 + 1024 + 256;
  cudaFuncSetAttribute(bw_kernel<STAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  long long* out;
  cudaMalloc(&out, 8 * grid);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int i = 0; i < 3; ++i) bw_kernel<STAGES><<<grid, 64, smem>>>(a, b, units, nkb, 22, out);
  cudaEventRecord(e0);
  const int reps = 20;
  for (int i = 0; i < reps; ++i) bw_kernel<STAGES><<<grid, 64, smem>>>(a, b, units, nkb, 22, out);
  cudaEventRecord(e1);
  cudaError_t err = cudaDeviceSynchronize();
  float ms = 0;
  cudaEventElapsedTime(&ms, e0, e1);
  long long h[148];
  cudaMemcpy(h, out, 8 * (grid < 148 ? grid : 148), cudaMemcpyDeviceToHost);
  const double bytes = double(grid) * units * 32768.0 * reps;
  printf("K=%4d stages=%d grid=%3d: %7.2f TB/s  (%.0f ns per 32 KB stage per CTA, %lld cycles per stage; %s)\n", K, STAGES, grid,
         bytes / (ms * 1e-3) / 1e12, ms * 1e6 / reps / units, h[0] / units, cudaGetErrorString(err));
  cudaFree(out);
}

template <int STAGES>
static void run_mc(const CUtensorMap& a, const CUtensorMap& b, int K, int grid) {
  const int nkb = K / 64, units = 16 * nkb * (1024 / K > 0 ? 1024 / K : 1) * 4;
  const int smem = STAGES * 32768 + 1024 + 256;
  cudaFuncSetAttribute(bw_mc_kernel<STAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  long long* out;
  cudaMalloc(&out, 8 * grid);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int i = 0; i < 3; ++i) bw_mc_kernel<STAGES><<<grid, 64, smem>>>(a, b, units, nkb, 22, out);
  cudaEventRecord(e0);
  const int reps = 20;
  for (int i = 0; i < reps; ++i) bw_mc_kernel<STAGES><<<grid, 64, smem>>>(a, b, units, nkb, 22, out);
  cudaEventRecord(e1);
  cudaError_t err = cudaDeviceSynchronize();
  float ms = 0;
  cudaEventElapsedTime(&ms, e0, e1);
  long long h[148];
  cudaMemcpy(h, out, 8 * (grid < 148 ? grid : 148), cudaMemcpyDeviceToHost);
  const double bytes = double(grid) * units * 32768.0 * reps;
  printf("MULTICAST pair: K=%4d stages=%d grid=%3d: %7.2f TB/s received, half of it requested (%.0f ns per 32 KB stage per CTA, %lld cycles; %s)\n",
         K, STAGES, grid, bytes / (ms * 1e-3) / 1e12, ms * 1e6 / reps / units, h[0] / units, cudaGetErrorString(err));
  cudaFree(out);
}

int main() {
  const int K = 1024;
  void *A, *W;
  cudaMalloc(&A, 2816ull * 4096 * 2);
  cudaMalloc(&W, 4096ull * 4096 * 2);
  cudaMemset(A, 0, 2816ull * 4096 * 2);
  cudaMemset(W, 0, 4096ull * 4096 * 2);
  for (int k : {1024, 4096}) {
    CUtensorMap a = make_map(A, k, 2816), b = make_map(W, k, 4096);
    for (int grid : {148, 74}) {
      run<3>(a, b, k, grid);
      run<5>(a, b, k, grid);
      run<6>(a, b, k, grid);
      run_mc<5>(a, b, k, grid);
    }
  }
  (void)K;
  return 0;
}
