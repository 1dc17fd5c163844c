// Weight gradient of the grouped k-tap Conv1d of ConvPositionEmbedding (modules.py:120-141) on the tensor core:
//   dW[o, ci, tap] = sum_{b, t} dY[b, t, o] * X[b, t + tap - pad, g(o) * cg + ci]
// Per tap this is dY^T X_shifted restricted to the block diagonal. One CTA owns 128 output channels (two 64-channel
// blocks), four taps and a share of the (batch, 64-row) k-blocks: per k-block the dY tile [64 rows x 128 ch] is loaded once
// and multiplied with the four row-shifted X tiles [64 rows x 128 ch] (both operands MN-major as stored: the contraction
// runs over the rows; TMA zero-fills rows outside [0, rows_per_batch), which is the conv's zero padding), D[128 co x 128 ci]
// per tap in TMEM (4 x 128 columns). Only the two diagonal 64 x 64 blocks of each D are gradients (same 64-channel block;
// inside a block only same-group entries when cg < 64): half of the tensor work is discarded, which is still ~20x cheaper
// than the CUDA-core kernel (gconv_wgrad_kernel: 0.95 ms per conv at config 5). Partial sums over the row shares land with
// f32 reductions, like every other weight gradient of the step.
// Precondition: rows t >= seq_len of x and dy are zero (the forward masks them, act_bwd zeroes them).
#pragma once
#include "ptx.cuh"

namespace oron {

struct GconvTcArgs {
  int C, cg, taps, pad;
  int kb_per_batch;  // rows_per_batch / 64
  int kb_total;      // nbatch * kb_per_batch
  int nchunk;        // row shares (gridDim.z)
  float* dw;         // [C, cg, taps]
};

constexpr int GCW_THREADS = 192;           // warp 0 TMA, warp 1 MMA + TMEM, warps 2..5 epilogue
constexpr int GCW_TAPS = 4;                // taps per CTA = 512 TMEM columns / 128
constexpr int GCW_TILE_BYTES = 16384;      // 128 channels x 64 rows bf16 = two 8 KB SW128 boxes
constexpr int GCW_STAGE_BYTES = (1 + GCW_TAPS) * GCW_TILE_BYTES;
constexpr int GCW_STAGES = 2;
constexpr int GCW_SMEM_BYTES = GCW_STAGES * GCW_STAGE_BYTES + 1024 + 256;

__global__ void __launch_bounds__(GCW_THREADS, 1)
gconv_wgrad_tc_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmDY, const GconvTcArgs a) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t bar_base = smem_base + GCW_STAGES * GCW_STAGE_BYTES;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (GCW_STAGES + s); };
  const uint32_t tfull_bar = bar_base + 8u * (2 * GCW_STAGES);
  const uint32_t tmem_slot = bar_base + 8u * (2 * GCW_STAGES + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int c0 = blockIdx.x * 128;          // first output / input channel of this CTA
  const int tap0 = blockIdx.y * GCW_TAPS;
  const int ntap = min(GCW_TAPS, a.taps - tap0);
  const int kb_beg = int((long long)a.kb_total * blockIdx.z / a.nchunk);
  const int kb_end = int((long long)a.kb_total * (blockIdx.z + 1) / a.nchunk);

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmX);
    tma_prefetch_desc(&tmDY);
    for (int s = 0; s < GCW_STAGES; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    mbar_init(tfull_bar, 1);
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));

  if (kb_end > kb_beg) {
    if (warp == 0) {
      if (lane == 0) {
        int stage = 0;
        uint32_t phase = 0;
        for (int kbi = kb_beg; kbi < kb_end; ++kbi) {
          const int b = kbi / a.kb_per_batch;
          const int r0 = (kbi - b * a.kb_per_batch) * 64;
          mbar_wait(empty_bar(stage), phase ^ 1u, 41);
          mbar_arrive_expect_tx(full_bar(stage), uint32_t(1 + ntap) * GCW_TILE_BYTES);
          const uint32_t s0 = smem_base + stage * GCW_STAGE_BYTES;
#pragma unroll
          for (int i = 0; i < 2; ++i) tma_load_3d(s0 + i * 8192, &tmDY, full_bar(stage), c0 + 64 * i, r0, b);
          for (int j = 0; j < ntap; ++j) {
            const uint32_t sx = s0 + (1 + j) * GCW_TILE_BYTES;
            const int rs = r0 + tap0 + j - a.pad;  // may be negative / run past the sequence: zero fill
#pragma unroll
            for (int i = 0; i < 2; ++i) tma_load_3d(sx + i * 8192, &tmX, full_bar(stage), c0 + 64 * i, rs, b);
          }
          if (++stage == GCW_STAGES) { stage = 0; phase ^= 1u; }
        }
      }
    } else if (warp == 1) {
      if (lane == 0) {
        constexpr uint32_t idesc = make_idesc_bf16(128, 128, 1, 1);  // both operands MN-major
        int stage = 0;
        uint32_t phase = 0;
        for (int kbi = kb_beg; kbi < kb_end; ++kbi) {
          mbar_wait(full_bar(stage), phase, 42);
          tc_fence_after();
          const uint32_t s0 = smem_base + stage * GCW_STAGE_BYTES;
          const uint64_t adesc = make_smem_desc_sw128(s0, 8192, 1024);
          for (int j = 0; j < ntap; ++j) {
            const uint64_t bdesc = make_smem_desc_sw128(s0 + (1 + j) * GCW_TILE_BYTES, 8192, 1024);
#pragma unroll
            for (int k = 0; k < 4; ++k)  // 16 K rows = 2048 bytes (>> 4)
              umma_bf16_ss(tmem_base + uint32_t(j * 128), adesc + 128u * uint64_t(k), bdesc + 128u * uint64_t(k), idesc,
                           (kbi != kb_beg || k != 0) ? 1u : 0u);
          }
          umma_commit(empty_bar(stage));
          if (kbi == kb_end - 1) umma_commit(tfull_bar);
          if (++stage == GCW_STAGES) { stage = 0; phase ^= 1u; }
        }
      }
    } else {
      // epilogue: thread = output channel co (TMEM lane), its 64 same-block input channels of every tap
      const int q = warp & 3;                 // TMEM lane quarter this warp may access
      const int co = q * 32 + lane;
      const int blk = co >> 6;                // which of the two 64-channel blocks
      const int o = c0 + co;
      mbar_wait(tfull_bar, 0, 43);
      tc_fence_after();
      {  // the host guarantees C % 128 == 0: every lane owns a real channel (tcgen05.ld is warp-collective)
        const int g0 = ((co & 63) / a.cg) * a.cg;  // first in-block channel of this output channel's group
        float* dst = a.dw + (long long)o * a.cg * a.taps;
        for (int j = 0; j < ntap; ++j) {
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            uint32_t r[32];
            tmem_ld_32x32(tmem_base + (uint32_t(q * 32) << 16) + uint32_t(j * 128 + blk * 64 + h * 32), r);
            tmem_wait_ld();
#pragma unroll
            for (int i = 0; i < 32; ++i) {
              const int ci = h * 32 + i - g0;  // index inside the group
              if (ci >= 0 && ci < a.cg) atomicAdd(dst + (long long)ci * a.taps + tap0 + j, __uint_as_float(r[i]));
            }
          }
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

}  // namespace oron
