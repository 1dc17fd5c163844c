// Grouped 1-D convolution (ConvPositionEmbedding, modules.py:120-141: k = 31, 64-channel groups) as an implicit GEMM whose
// activation window stays RESIDENT in shared memory.
//
// The generic 1-SM kernel (gemm_tcgen05.cuh) fetches the 128 x 64 A tile again for every tap, shifted by one row: 24 KB of
// TMA traffic per tap against 128 cycles of tensor work, i.e. three times the per-SM TMA ingest (64 B/clk). Here
// the (128 + taps - 1)-row window of the tile's 64 input channels is loaded ONCE per tile, and tap j reads it through a
// shared-memory descriptor whose start address is advanced by j rows (j * 128 bytes). A K-major SWIZZLE_128B operand may
// start at any 128-byte row of a tile that TMA wrote: the swizzle is a function of the absolute shared-memory address, so
// the rows a shifted descriptor walks are exactly where TMA put them (tools/desc_shift_test.cu: all 32 shifts bit-exact
// on B200 with the descriptor's base-offset field left 0). Only the weights stream: 8 KB per tap.
//
//   warp 0: TMA producer (window double-buffered across tiles; weight ring of kStages x kTapsPerStage taps)
//   warp 1: MMA issuer (M = 128, N = 64, K = 16; accumulator double-buffered in TMEM)
//   warps 2..9: epilogue (the generic kernel's: bias / Mish + mask / residual ...)
#pragma once
#include "gemm_tcgen05.cuh"

namespace oron {

struct GconvResCfg {
  static constexpr int BN = 64;
  static constexpr int kMaxTaps = 33;
  static constexpr int kWinRows = GEMM_BM + kMaxTaps - 1;      // 160: rows of the TMA box (a multiple of 8 -> 1024-byte atoms)
  static constexpr int kWinBytes = kWinRows * 128;             // 20480
  static constexpr int kTapBytes = BN * GEMM_BK * 2;           // 8192: one tap's 64 x 64 weight block
  static constexpr int kTapsPerStage = 4;
  static constexpr int kStages = 4;
  static constexpr int kStageBytes = kTapsPerStage * kTapBytes;
  static constexpr int kTmemCols = 2 * BN;
  static constexpr int kBarBytes = 256;
  static constexpr int kSmemBytes = 2 * kWinBytes + kStages * kStageBytes + 1024 /*align*/ + kBarBytes + EPI_STAGE_BYTES;
};

template <int EPI>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
gconv_res_tcgen05_kernel(const __grid_constant__ CUtensorMap tmA /* box 64 x kWinRows */,
                         const __grid_constant__ CUtensorMap tmB /* box 64 x 64 */, const GemmArgs args) {
  using Cfg = GconvResCfg;
  constexpr int BN = Cfg::BN, kStages = Cfg::kStages, TPS = Cfg::kTapsPerStage;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t win_base = smem_base;                          // 2 windows
  const uint32_t ring_base = smem_base + 2 * Cfg::kWinBytes;    // weight ring
  const uint32_t bar_base = ring_base + kStages * Cfg::kStageBytes;
  auto bfull = [&](int s) { return bar_base + 8u * s; };
  auto bempty = [&](int s) { return bar_base + 8u * (kStages + s); };
  auto afull = [&](int s) { return bar_base + 8u * (2 * kStages + s); };
  auto aempty = [&](int s) { return bar_base + 8u * (2 * kStages + 2 + s); };
  auto tfull = [&](int s) { return bar_base + 8u * (2 * kStages + 4 + s); };
  auto tempty = [&](int s) { return bar_base + 8u * (2 * kStages + 6 + s); };
  const uint32_t tmem_slot = bar_base + 8u * (2 * kStages + 8);
  const uint32_t epi_stage_base = bar_base + Cfg::kBarBytes;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  const int tiles_m_pb = (args.rows_per_batch + GEMM_BM - 1) / GEMM_BM;
  const int tiles_m = tiles_m_pb * args.nbatch;
  const int tiles_n = args.N / BN;
  const int num_tiles = tiles_m * tiles_n;
  const int taps = args.num_kb;  // cpb == 1: one k-block per tap
  const int nst = (taps + TPS - 1) / TPS;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    for (int s = 0; s < kStages; ++s) {
      mbar_init(bfull(s), 1);
      mbar_init(bempty(s), 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(afull(s), 1);
      mbar_init(aempty(s), 1);
      mbar_init(tfull(s), 1);
      mbar_init(tempty(s), GEMM_EPI_WARPS);
    }
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, Cfg::kTmemCols);
    tmem_relinquish();
  }
  pdl_launch_dependents();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));
  pdl_wait();

  if (warp == 0) {
    if (lane == 0) {
      int stage = 0, it = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
        const int m_tile = tile % tiles_m;
        const int n0 = (tile / tiles_m) * BN;
        const int b = m_tile / tiles_m_pb;
        const int t0 = (m_tile % tiles_m_pb) * GEMM_BM;
        const int ab = it & 1;
        mbar_wait(aempty(ab), ((it >> 1) & 1u) ^ 1u, 1);
        mbar_arrive_expect_tx(afull(ab), Cfg::kWinBytes);  // rows outside [0, rows_per_batch) arrive as zeros: the conv's padding
        tma_load_3d(win_base + ab * Cfg::kWinBytes, &tmA, afull(ab), n0 /* the group's 64 input channels */, t0 - args.pad, b);
        for (int s = 0; s < nst; ++s) {
          const int nt = min(TPS, taps - s * TPS);
          mbar_wait(bempty(stage), phase ^ 1u, 2);
          mbar_arrive_expect_tx(bfull(stage), uint32_t(nt) * Cfg::kTapBytes);
          for (int j = 0; j < nt; ++j)
            tma_load_2d(ring_base + stage * Cfg::kStageBytes + j * Cfg::kTapBytes, &tmB, bfull(stage), (s * TPS + j) * GEMM_BK, n0);
          if (++stage == kStages) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc_bf16(GEMM_BM, BN);
      int stage = 0, it = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
        const int as = it & 1;
        const uint32_t aphase = (it >> 1) & 1u;
        mbar_wait(tempty(as), aphase ^ 1u, 3);
        mbar_wait(afull(as), aphase, 4);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + uint32_t(as * BN);
        const uint32_t win = win_base + as * Cfg::kWinBytes;
        for (int s = 0; s < nst; ++s) {
          const int nt = min(TPS, taps - s * TPS);
          mbar_wait(bfull(stage), phase, 5);
          tc_fence_after();
          for (int j = 0; j < nt; ++j) {
            const int tap = s * TPS + j;
            // tap `tap` of output row r reads window row r + tap: the A tile is the window advanced by `tap` rows
            const uint64_t adesc = make_smem_desc_sw128(win + uint32_t(tap) * 128u, 16, 1024);
            const uint64_t bdesc = make_smem_desc_sw128(ring_base + stage * Cfg::kStageBytes + j * Cfg::kTapBytes, 16, 1024);
#pragma unroll
            for (int k = 0; k < GEMM_BK / 16; ++k)
              umma_bf16_ss(tmem_d, adesc + uint64_t(2 * k), bdesc + uint64_t(2 * k), idesc, (tap | k) != 0 ? 1u : 0u);
          }
          umma_commit(bempty(stage));
          if (++stage == kStages) { stage = 0; phase ^= 1u; }
        }
        umma_commit(aempty(as));  // window free once the tile's MMAs have retired
        umma_commit(tfull(as));
      }
    }
  } else {
    const int q = warp & 3;
    const int chalf = (warp - 2) >> 2;
    constexpr int HN = BN / 2;
    const int cbeg = chalf * HN;
    int it = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
      const int m_tile = tile % tiles_m;
      const int n0 = (tile / tiles_m) * BN;
      const int b = m_tile / tiles_m_pb;
      const int t_base = (m_tile % tiles_m_pb) * GEMM_BM + q * 32;
      const int as = it & 1;
      const uint32_t aphase = (it >> 1) & 1u;
      EpiCols<HN> pc;
      gemm_epilogue_prefetch<BN, EPI, HN>(args, b, n0, cbeg, lane, pc);
      mbar_wait(tfull(as), aphase, 6);
      tc_fence_after();
      const uint32_t trow = tmem_base + (uint32_t(q * 32) << 16) + uint32_t(as * BN);
      gemm_epilogue_tile<BN, EPI, HN>(args, trow, b, t_base, n0, cbeg, epi_stage_base + uint32_t(warp - 2) * EPI_STAGE_BYTES_PER_WARP,
                                      lane, pc);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tempty(as));
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, Cfg::kTmemCols);
  }
}

}  // namespace oron
