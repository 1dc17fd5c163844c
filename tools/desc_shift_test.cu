// Numerical probe: can a tcgen05.mma read a K-major SWIZZLE_128B operand tile starting at an arbitrary ROW of a larger tile that
// TMA wrote (start address = tile + shift * 128 B, i.e. not aligned to the 1024-byte swizzle atom)? And does the descriptor's
// "matrix base offset" field (bits 49-51) have to carry (address >> 7) & 7 for it?
//   build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I oron_tts_b200/csrc tools/desc_shift_test.cu -o tools/desc_shift_test -lcuda
// Why it matters: the k = 31 grouped conv re-fetches its A tile for each tap shifted by one row; if
// shifted views work, the 158-row window can stay resident and only the weights stream.
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <vector>
#include <cuda.h>
#include <cudaTypedefs.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include "ptx.cuh"

using namespace oron;

constexpr int NSHIFT = 32;

__global__ void __launch_bounds__(128, 1) probe_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                                                       float* out, int use_base_offset) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t sA = base, sB = base + 32768, bar = base + 32768 + 8192, bar2 = bar + 8, slot = bar + 16;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) { mbar_init(bar, 1); mbar_init(bar2, 1); fence_mbar_init(); }
  if (warp == 0) { tmem_alloc(slot, 64); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem) : "r"(slot));
  if (threadIdx.x == 0) {
    mbar_arrive_expect_tx(bar, 32768 + 8192);
    tma_load_2d(sA, &tmA, bar, 0, 0);
    tma_load_2d(sA + 16384, &tmA, bar, 0, 128);
    tma_load_2d(sB, &tmB, bar, 0, 0);
  }
  mbar_wait(bar, 0, 1);
  for (int shift = 0; shift < NSHIFT; ++shift) {
    if (threadIdx.x == 0) {
      tc_fence_after();
      constexpr uint32_t idesc = make_idesc_bf16(128, 64, 0, 0);
      const uint32_t a_addr = sA + uint32_t(shift) * 128u;
      uint64_t adesc = make_smem_desc_sw128(a_addr, 16, 1024);
      if (use_base_offset) adesc |= uint64_t((a_addr >> 7) & 7u) << 49;
      const uint64_t bdesc = make_smem_desc_sw128(sB, 16, 1024);
#pragma unroll
      for (int k = 0; k < 4; ++k) umma_bf16_ss(tmem, adesc + uint64_t(2 * k), bdesc + uint64_t(2 * k), idesc, k != 0);
      umma_commit(bar2);
    }
    mbar_wait(bar2, uint32_t(shift) & 1u, 2);
    tc_fence_after();
    uint32_t r[32];
    for (int h = 0; h < 2; ++h) {
      tmem_ld_32x32(tmem + (uint32_t(warp * 32) << 16) + 32 * h, r);
      tmem_wait_ld();
      for (int i = 0; i < 32; ++i) out[((long long)shift * 128 + warp * 32 + lane) * 64 + 32 * h + i] = __uint_as_float(r[i]);
    }
    tc_fence_before();
    __syncthreads();
  }
  // timing: REP x 64 back-to-back MMAs (M128 N64 K16) through the A descriptor at row shift s, one commit per 64; cycles per MMA
  // land in out[NSHIFT * 128 * 64 + s] (does an operand that straddles the 1024-byte swizzle atoms cost extra shared-memory reads?)
  if (use_base_offset == 0) {
    uint32_t ph = NSHIFT & 1u;
    for (int shift = 0; shift < 16; ++shift) {
      long long t0 = 0;
      if (threadIdx.x == 0) {
        tc_fence_after();
        constexpr uint32_t idesc = make_idesc_bf16(128, 64, 0, 0);
        const uint64_t adesc = make_smem_desc_sw128(sA + uint32_t(shift) * 128u, 16, 1024);
        const uint64_t bdesc = make_smem_desc_sw128(sB, 16, 1024);
        t0 = clock64();
        for (int rep = 0; rep < 16; ++rep)
#pragma unroll
          for (int k = 0; k < 4; ++k) umma_bf16_ss(tmem, adesc + uint64_t(2 * k), bdesc + uint64_t(2 * k), idesc, 1u);
        umma_commit(bar2);
      }
      mbar_wait(bar2, ph, 3);
      ph ^= 1u;
      if (threadIdx.x == 0) out[(long long)NSHIFT * 128 * 64 + shift] = float(clock64() - t0) / 64.f;
      tc_fence_before();
      __syncthreads();
    }
  }
  if (warp == 0) { tc_fence_after(); tmem_dealloc(tmem, 64); }
}

static CUtensorMap make_map(void* p, uint64_t cols, uint64_t rows, uint32_t box_rows) {
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
  auto enc = reinterpret_cast<PFN_cuTensorMapEncodeTiled>(fn);
  CUtensorMap m;
  cuuint64_t dims[2] = {cols, rows};
  cuuint64_t strides[1] = {cols * 2};
  cuuint32_t box[2] = {64, box_rows};
  cuuint32_t es[2] = {1, 1};
  CUresult r = enc(&m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, p, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { printf("encode failed %d\n", int(r)); exit(1); }
  return m;
}

int main() {
  std::vector<__nv_bfloat16> hA(256 * 64), hB(64 * 64);
  std::vector<float> fA(256 * 64), fB(64 * 64);
  srand(1);
  for (int i = 0; i < 256 * 64; ++i) { fA[i] = float(rand() % 7 - 3); hA[i] = __float2bfloat16(fA[i]); }
  for (int i = 0; i < 64 * 64; ++i) { fB[i] = float(rand() % 5 - 2); hB[i] = __float2bfloat16(fB[i]); }
  void *dA, *dB;
  float* dO;
  cudaMalloc(&dA, hA.size() * 2); cudaMalloc(&dB, hB.size() * 2); cudaMalloc(&dO, (NSHIFT * 128 * 64 + 64) * 4);
  cudaMemcpy(dA, hA.data(), hA.size() * 2, cudaMemcpyHostToDevice);
  cudaMemcpy(dB, hB.data(), hB.size() * 2, cudaMemcpyHostToDevice);
  CUtensorMap ta = make_map(dA, 64, 256, 128), tb = make_map(dB, 64, 64, 64);
  const int smem = 32768 + 8192 + 1024 + 64;
  cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  std::vector<float> hO(NSHIFT * 128 * 64 + 64);
  for (int ubo = 0; ubo < 2; ++ubo) {
    cudaMemset(dO, 0, hO.size() * 4);
    probe_kernel<<<1, 128, smem>>>(ta, tb, dO, ubo);
    cudaError_t e = cudaDeviceSynchronize();
    cudaMemcpy(hO.data(), dO, hO.size() * 4, cudaMemcpyDeviceToHost);
    printf("base_offset field %s (%s): shifts with exact results:", ubo ? "= (addr >> 7) & 7" : "= 0", cudaGetErrorString(e));
    for (int s = 0; s < NSHIFT; ++s) {
      bool ok = true;
      for (int m = 0; m < 128 && ok; ++m)
        for (int n = 0; n < 64; ++n) {
          float ref = 0.f;
          for (int k = 0; k < 64; ++k) ref += fA[(s + m) * 64 + k] * fB[n * 64 + k];
          if (hO[((long long)s * 128 + m) * 64 + n] != ref) { ok = false; break; }
        }
      if (ok) printf(" %d", s);
    }
    printf("\n");
    if (ubo == 0) {
      printf("cycles per M128 N64 K16 MMA (64 back to back) by row shift of the A descriptor:");
      for (int s2 = 0; s2 < 16; ++s2) printf(" %d:%.1f", s2, hO[NSHIFT * 128 * 64 + s2]);
      printf("\n");
    }
  }
  return 0;
}
