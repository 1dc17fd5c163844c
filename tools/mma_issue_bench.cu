// Micro-benchmark: what does one small tcgen05.mma cost?  (decides the attention tiling)
//   build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I oron_tts_b200/csrc tools/mma_issue_bench.cu -o tools/mma_issue_bench
//   run  : tools/mma_issue_bench
// One issuing thread per CTA runs REP groups of G MMAs (M=128, K=16, bf16, operands = zeros in smem), commits each
// group to an mbarrier and waits for it. Reported per MMA: cycles until the issue loop returns ("issue") and
// cycles until the group has retired ("total"), for N in {16,64,128,256}, one accumulator (dependent chain) or
// several rotating accumulators, and 1 or 2 co-resident CTAs per SM (the second CTA doubles the load on the pipe).
#include <cstdio>
#include <cstdint>
#include <cuda.h>
#include <cuda_runtime.h>

#include "ptx.cuh"

using namespace oron;

struct Res { long long issue, total; };

template <int N, int NACC>
__global__ void __launch_bounds__(64) bench_kernel(Res* out, int G, int REP) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const uint32_t base = smem_u32(smem);
  const uint32_t sA = base;                 // 128 x 64 bf16 (16 KB), SW128 K-major
  const uint32_t sB = base + 16384;         // 256 x 64 bf16 (32 KB)
  const uint32_t bar = base + 49152;
  const uint32_t slot = bar + 8;
  for (int i = threadIdx.x; i < 49152 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  if (threadIdx.x == 0) { mbar_init(bar, 1); fence_mbar_init(); }
  if (threadIdx.x < 32) { tmem_alloc(slot, 256); tmem_relinquish(); }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem) : "r"(slot));
  if (threadIdx.x == 32) {
    constexpr uint32_t idesc = make_idesc_bf16(128, N, 0, 0);
    const uint64_t ad = make_smem_desc_sw128(sA, 16, 1024);
    const uint64_t bd = make_smem_desc_sw128(sB, 16, 1024);
    long long t_issue = 0, t_total = 0;
    for (int r = 0; r < REP; ++r) {
      const long long t0 = clock64();
      for (int g = 0; g < G; ++g) {
        const uint32_t acc = tmem + uint32_t(g % NACC) * (256 / NACC);
        umma_bf16_ss(acc, ad + uint64_t(2 * (g & 3)), bd + uint64_t(2 * (g & 3)), idesc, g >= NACC ? 1u : 0u);
      }
      umma_commit(bar);
      const long long t1 = clock64();
      mbar_wait(bar, r & 1u, 1);
      const long long t2 = clock64();
      if (r > 0) { t_issue += t1 - t0; t_total += t2 - t0; }
    }
    if (blockIdx.x == 0) { out->issue = t_issue; out->total = t_total; }
  }
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) { tc_fence_after(); tmem_dealloc(tmem, 256); }
}

template <int N, int NACC>
void run(int ctas_per_sm, int G, int REP, Res* d) {
  int sms = 148;
  cudaFuncSetAttribute(bench_kernel<N, NACC>, cudaFuncAttributeMaxDynamicSharedMemorySize, 49152 + 64 + 1024);
  cudaMemset(d, 0, sizeof(Res));
  bench_kernel<N, NACC><<<sms * ctas_per_sm, 64, 49152 + 64 + 1024>>>(d, G, REP);
  cudaError_t e = cudaDeviceSynchronize();
  Res h;
  cudaMemcpy(&h, d, sizeof(Res), cudaMemcpyDeviceToHost);
  const double n = double(G) * (REP - 1);
  printf("N=%3d acc=%d ctas/SM=%d G=%2d : issue %6.1f cyc/mma   total %6.1f cyc/mma   (floor %d)%s\n", N, NACC, ctas_per_sm, G,
         h.issue / n, h.total / n, 128 * N / 256, e == cudaSuccess ? "" : cudaGetErrorString(e));
}

int main() {
  Res* d;
  cudaMalloc(&d, sizeof(Res));
  const int REP = 200;
  for (int cps = 1; cps <= 2; ++cps) {
    for (int G : {4, 8, 32}) {
      run<16, 1>(cps, G, REP, d);
      run<64, 1>(cps, G, REP, d);
      run<64, 2>(cps, G, REP, d);
      run<64, 4>(cps, G, REP, d);
      run<128, 1>(cps, G, REP, d);
      run<128, 2>(cps, G, REP, d);
      run<256, 1>(cps, G, REP, d);
    }
  }
  return 0;
}
