// Non-causal multi-head attention with per-sequence key lengths (replaces the reference's
// F.scaled_dot_product_attention + key-padding mask, src/models/modules.py:271-278).
// One CTA = one (batch element, head, 128-query tile). head_dim = 64.
//   S = Q K^T      : tcgen05.mma M=128 N=128 K=64, accumulator in TMEM cols [0,128)
//   P = softmax    : 128 softmax threads, one query row each (TMEM lane == row => no shuffles)
//   O_j = P V_j    : tcgen05.mma M=128 N=64 K=128 (V tile as MN-major B operand), TMEM cols [128,192)
//   O += O_j       : running output kept in registers (64 fp32 per row), rescaled online.
// q/k/v are read straight out of the fused QKV activation [rows, 3*H*64] with one 3-D TMA map.
#pragma once
#include "ptx.cuh"

namespace oron {

struct AttnArgs {
  int rows_per_batch;   // Tpad: rows per batch element in qkv / out
  int nbatch;
  int heads;
  const int* seq_lens;  // [nbatch] valid keys (= valid queries) per batch element, or nullptr
  __nv_bfloat16* out;   // [nbatch*rows_per_batch, ldo], head h at columns [h*64, h*64+64)
  long long ldo;
  float scale_log2;     // softmax scale * log2(e)
};

constexpr int ATT_THREADS = 192;  // warp0 TMA, warp1 MMA + TMEM alloc, warps 2..5 softmax
constexpr int ATT_TILE = 128;
constexpr int ATT_D = 64;
constexpr int ATT_TILE_BYTES = ATT_TILE * ATT_D * 2;  // 16 KB
// smem: Q | K0 K1 | V0 V1 | P(2 slabs) | barriers
constexpr int ATT_SMEM_BYTES = 7 * ATT_TILE_BYTES + 128;
constexpr int ATT_TMEM_COLS = 256;

__global__ void __launch_bounds__(ATT_THREADS, 2)
attn_fwd_tcgen05_kernel(const __grid_constant__ CUtensorMap tmQKV, const AttnArgs args) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t smem_base = smem_u32(smem_raw);
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  const int q_tile = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
  const int q0 = q_tile * ATT_TILE;
  const int len = args.seq_lens ? min(args.seq_lens[b], args.rows_per_batch) : args.rows_per_batch;
  if (q0 >= len) return;  // whole tile is padding: the out-projection masks these rows
  if ((smem_base & 1023u) != 0) {
    if (threadIdx.x == 0) printf("[oron] attention: dynamic smem not 1024-byte aligned\n");
    __trap();
  }
  const int n_kv = (len + ATT_TILE - 1) / ATT_TILE;

  const uint32_t sQ = smem_base;
  auto sK = [&](int s) { return smem_base + (1 + s) * ATT_TILE_BYTES; };
  auto sV = [&](int s) { return smem_base + (3 + s) * ATT_TILE_BYTES; };
  const uint32_t sP = smem_base + 5 * ATT_TILE_BYTES;
  const uint32_t bar_base = smem_base + 7 * ATT_TILE_BYTES;
  const uint32_t q_full = bar_base;
  auto kv_full = [&](int s) { return bar_base + 8u * (1 + s); };
  auto kv_empty = [&](int s) { return bar_base + 8u * (3 + s); };
  const uint32_t s_full = bar_base + 8u * 5;
  const uint32_t p_full = bar_base + 8u * 6;
  const uint32_t o_full = bar_base + 8u * 7;
  const uint32_t tmem_slot = bar_base + 8u * 8;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmQKV);
    mbar_init(q_full, 1);
    for (int s = 0; s < 2; ++s) { mbar_init(kv_full(s), 1); mbar_init(kv_empty(s), 1); }
    mbar_init(s_full, 1);
    mbar_init(p_full, 128);
    mbar_init(o_full, 1);
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, ATT_TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));
  const uint32_t tmem_S = tmem_base;
  const uint32_t tmem_O = tmem_base + 128;

  const int HD = args.heads * ATT_D;
  if (warp == 0) {
    if (lane == 0) {
      mbar_arrive_expect_tx(q_full, ATT_TILE_BYTES);
      tma_load_3d(sQ, &tmQKV, q_full, h * ATT_D, q0, b);
      for (int j = 0; j < n_kv; ++j) {
        const int s = j & 1;
        mbar_wait(kv_empty(s), ((j >> 1) & 1u) ^ 1u, 11);
        mbar_arrive_expect_tx(kv_full(s), 2 * ATT_TILE_BYTES);
        tma_load_3d(sK(s), &tmQKV, kv_full(s), HD + h * ATT_D, j * ATT_TILE, b);
        tma_load_3d(sV(s), &tmQKV, kv_full(s), 2 * HD + h * ATT_D, j * ATT_TILE, b);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc_s = make_idesc_bf16(128, 128, 0, 0);
      constexpr uint32_t idesc_o = make_idesc_bf16(128, 64, 0, 1);  // B = V is MN-major
      const uint64_t qdesc = make_smem_desc_sw128(sQ, 16, 1024);
      auto issue_S = [&](int j) {
        const uint64_t kdesc = make_smem_desc_sw128(sK(j & 1), 16, 1024);
#pragma unroll
        for (int k = 0; k < 4; ++k)
          umma_bf16_ss(tmem_S, qdesc + uint64_t(2 * k), kdesc + uint64_t(2 * k), idesc_s, k != 0);
        umma_commit(s_full);
      };
      mbar_wait(q_full, 0, 12);
      mbar_wait(kv_full(0), 0, 13);
      tc_fence_after();
      issue_S(0);
      for (int j = 0; j < n_kv; ++j) {
        const int s = j & 1;
        mbar_wait(p_full, j & 1u, 14);  // P(j) in smem, S(j) and O(j-1) drained by the softmax threads
        tc_fence_after();
        if (j + 1 < n_kv) {
          mbar_wait(kv_full((j + 1) & 1), ((j + 1) >> 1) & 1u, 15);
          tc_fence_after();
          issue_S(j + 1);
        }
#pragma unroll
        for (int kk = 0; kk < 8; ++kk) {
          const uint64_t pdesc =
              make_smem_desc_sw128(sP + (kk >> 2) * ATT_TILE_BYTES + (kk & 3) * 32, 16, 1024);
          const uint64_t vdesc = make_smem_desc_sw128(sV(s) + kk * 2048, 1024, 1024);
          umma_bf16_ss(tmem_O, pdesc, vdesc, idesc_o, kk != 0);
        }
        umma_commit(o_full);
        umma_commit(kv_empty(s));
      }
    }
  } else {
    // ===================== softmax / output threads =====================
    const int q = warp & 3;
    const int r = q * 32 + lane;  // query row inside the tile == TMEM lane
    const uint32_t lane_off = uint32_t(q * 32) << 16;
    const float c = args.scale_log2;
    float m_run = -INFINITY, l_run = 0.f;
    float o_acc[ATT_D];
#pragma unroll
    for (int i = 0; i < ATT_D; ++i) o_acc[i] = 0.f;
    const uint32_t prow = sP + r * 128;
    const uint32_t sw = uint32_t(r & 7);

    for (int j = 0; j < n_kv; ++j) {
      const int n_valid = min(ATT_TILE, len - j * ATT_TILE);
      mbar_wait(s_full, j & 1u, 16);
      tc_fence_after();
      // pass 1: row maximum over the valid keys of this tile
      float mx = -INFINITY;
#pragma unroll 1
      for (int c0 = 0; c0 < ATT_TILE; c0 += 32) {
        uint32_t v[32];
        tmem_ld_32x32(tmem_S + lane_off + c0, v);
        tmem_wait_ld();
        if (c0 + 32 <= n_valid) {
#pragma unroll
          for (int i = 0; i < 32; ++i) mx = fmaxf(mx, __uint_as_float(v[i]));
        } else {
#pragma unroll
          for (int i = 0; i < 32; ++i)
            if (c0 + i < n_valid) mx = fmaxf(mx, __uint_as_float(v[i]));
        }
      }
      const float m_new = fmaxf(m_run, mx);
      const float alpha = ex2_approx((m_run - m_new) * c);  // m_run = -inf on the first tile -> 0
      if (j > 0) {
        // fold in O(j-1) = P(j-1) V(j-1), which was computed against m_run
        mbar_wait(o_full, (j - 1) & 1u, 17);
        tc_fence_after();
#pragma unroll
        for (int c0 = 0; c0 < ATT_D; c0 += 32) {
          uint32_t v[32];
          tmem_ld_32x32(tmem_O + lane_off + c0, v);
          tmem_wait_ld();
#pragma unroll
          for (int i = 0; i < 32; ++i) o_acc[c0 + i] = (o_acc[c0 + i] + __uint_as_float(v[i])) * alpha;
        }
      }
      l_run *= alpha;
      m_run = m_new;
      const float mc = m_new * c;
      // pass 2: P = exp2(S*c - m*c) -> bf16 -> smem (SW128 K-major, two 64-key slabs)
      float lsum = 0.f;
#pragma unroll 1
      for (int c0 = 0; c0 < ATT_TILE; c0 += 32) {
        uint32_t v[32];
        tmem_ld_32x32(tmem_S + lane_off + c0, v);
        tmem_wait_ld();
        uint32_t pk[16];
#pragma unroll
        for (int i = 0; i < 32; i += 2) {
          float p0 = ex2_approx(__uint_as_float(v[i]) * c - mc);
          float p1 = ex2_approx(__uint_as_float(v[i + 1]) * c - mc);
          if (c0 + i >= n_valid) p0 = 0.f;
          if (c0 + i + 1 >= n_valid) p1 = 0.f;
          lsum += p0 + p1;
          pk[i / 2] = pack_bf16x2(p0, p1);
        }
        const uint32_t slab = prow + (c0 >> 6) * ATT_TILE_BYTES;
        const uint32_t chunk0 = uint32_t(c0 & 63) >> 3;  // first 16-byte chunk of this 32-key group
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          const uint32_t addr = slab + (((chunk0 + g) ^ sw) << 4);
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(pk[4 * g]),
                       "r"(pk[4 * g + 1]), "r"(pk[4 * g + 2]), "r"(pk[4 * g + 3])
                       : "memory");
        }
      }
      l_run += lsum;
      fence_proxy_async_smem();
      tc_fence_before();
      mbar_arrive(p_full);
    }
    // last tile's PV
    mbar_wait(o_full, (n_kv - 1) & 1u, 18);
    tc_fence_after();
    const float inv_l = 1.0f / l_run;
    const int t = q0 + r;
    __nv_bfloat16* orow = args.out + ((long long)b * args.rows_per_batch + t) * args.ldo + h * ATT_D;
#pragma unroll
    for (int c0 = 0; c0 < ATT_D; c0 += 32) {
      uint32_t v[32];
      tmem_ld_32x32(tmem_O + lane_off + c0, v);
      tmem_wait_ld();
      if (t < args.rows_per_batch) {
        uint32_t pk[16];
#pragma unroll
        for (int i = 0; i < 32; i += 2)
          pk[i / 2] = pack_bf16x2((o_acc[c0 + i] + __uint_as_float(v[i])) * inv_l,
                                  (o_acc[c0 + i + 1] + __uint_as_float(v[i + 1])) * inv_l);
        uint4* o4 = reinterpret_cast<uint4*>(orow + c0);
#pragma unroll
        for (int g = 0; g < 4; ++g) o4[g] = make_uint4(pk[4 * g], pk[4 * g + 1], pk[4 * g + 2], pk[4 * g + 3]);
      }
    }
    tc_fence_before();
  }

  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, ATT_TMEM_COLS);
  }
}

}  // namespace oron
