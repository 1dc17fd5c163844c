// Backward of the varlen non-causal attention (head_dim 64) on tcgen05 / TMEM: one CTA per (batch, head, 128-row owner
// tile); the operand tiles of the next iteration arrive by TMA while the current one computes, the accumulating MMAs
// of an iteration stay in flight behind the S / dP MMAs of the next; the S / dP -> CUDA cores -> dS hand-off itself is
// not yet overlapped across iterations (TMEM holds one S and one dP).
//
//   MODE 0 (dQ):    owner = 128 queries. Pre-pass over the key tiles: S = Q K^T -> log2-domain log-sum-exp per row
//                   (written to `lse` together with delta = rowsum(dO * O)). Main pass per key tile j:
//                   S = Q K_j^T, dP = dO V_j^T (TMEM) -> dS = P * (dP - delta) * scale -> bf16 smem -> dQ += dS K_j.
//   MODE 1 (dK,dV): owner = 128 keys. Per query tile i: S^T = K Q_i^T, dP^T = V dO_i^T (TMEM) ->
//                   P^T = exp2(S^T c - lse[q]), dS^T = P^T * (dP^T - delta[q]) * scale -> bf16 smem ->
//                   dV += P^T dO_i, dK += dS^T Q_i. Runs after MODE 0 (needs lse / delta).
// The accumulators are read once at the end; dq / dk are rotated back through RoPE (modules.py:96-104 transposed)
// so that the gradients are w.r.t. the pre-RoPE projections, rows beyond the sequence are written as zeros.
// Operand tiles are 128 x 64 bf16, SW128 K-major (TMA); the second-stage B operands (K_j, dO_i, Q_i) are the same
// tiles read MN-major, exactly like V in the forward kernel (attn_tcgen05.cuh).
#pragma once
#include "ptx.cuh"

namespace oron {

struct AttnBwdArgs {
  int rows_per_batch, nbatch, heads, tiles;
  const int* seq_lens;
  float scale, scale_log2;
  const __nv_bfloat16* o;
  long long ld_o;
  const __nv_bfloat16* d_o;
  long long ld_do;
  __nv_bfloat16* dqkv;
  long long ld_dqkv;
  const float* rope_cos;  // [rows_per_batch, 32]
  const float* rope_sin;
  float* lse;    // [nbatch * heads * rows_per_batch]
  float* delta;
  int have_lse;  // 1: `lse` was written by the forward kernel (oron_attention_fwd_lse): MODE 0 skips its pre-pass
};

constexpr int AB_THREADS = 160;  // warp 0: TMA + MMA issue (+ TMEM alloc); warps 1..4: one thread per owner row
constexpr int AB_TILE = 128;
constexpr int AB_D = 64;
constexpr int AB_TILE_BYTES = AB_TILE * AB_D * 2;  // 16 KB
constexpr int AB_TMEM_COLS = 512;
// smem: X1 | X2 | Y1[0] Y2[0] | Y1[1] Y2[1] | stageA (2 slabs) | stageB (2 slabs) | barriers + lse/delta staging
// (the Y tiles of iteration it + 1 are fetched by TMA while iteration it computes)
constexpr int AB_SMEM_BYTES = 10 * AB_TILE_BYTES + 64 + 4 * 128 * 4 + 1024;

__device__ __forceinline__ void ab_tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) { tmem_ld_32x32(taddr, r); }

// 32 consecutive columns of one staging row (bf16, SW128 K-major tile of 128 rows x 128 columns in two 64-column slabs)
__device__ __forceinline__ void ab_stage_store(uint32_t stage, int r, int c0, const uint32_t (&pk)[16]) {
  const uint32_t slab = stage + uint32_t(c0 >> 6) * AB_TILE_BYTES + uint32_t(r) * 128u;
  const uint32_t chunk0 = uint32_t(c0 & 63) >> 3;
  const uint32_t sw = uint32_t(r & 7);
#pragma unroll
  for (int g = 0; g < 4; ++g) {
    const uint32_t addr = slab + (((chunk0 + g) ^ sw) << 4);
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(pk[4 * g]), "r"(pk[4 * g + 1]),
                 "r"(pk[4 * g + 2]), "r"(pk[4 * g + 3])
                 : "memory");
  }
}

template <int MODE>
__global__ void __launch_bounds__(AB_THREADS, 1)
attn_bwd_tcgen05_kernel(const __grid_constant__ CUtensorMap tmQK, const __grid_constant__ CUtensorMap tmV,
                        const __grid_constant__ CUtensorMap tmDO, const AttnBwdArgs args) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  const int tile = blockIdx.x % args.tiles;
  const int h = (blockIdx.x / args.tiles) % args.heads;
  const int b = blockIdx.x / (args.tiles * args.heads);
  const int HD = args.heads * AB_D;
  const int len = args.seq_lens ? min(args.seq_lens[b], args.rows_per_batch) : args.rows_per_batch;
  const int nt = (len + AB_TILE - 1) / AB_TILE;
  const long long row_base = (long long)b * args.rows_per_batch;

  if (tile >= nt) {  // owner tile entirely beyond the sequence: its gradients are zero
    const int r = int(threadIdx.x);
    const int t = tile * AB_TILE + r;
    if (r < AB_TILE && t < args.rows_per_batch) {
      __nv_bfloat16* p = args.dqkv + (row_base + t) * args.ld_dqkv + h * AB_D;
      const uint4 z = make_uint4(0u, 0u, 0u, 0u);
      if (MODE == 0) {
#pragma unroll
        for (int i = 0; i < 8; ++i) reinterpret_cast<uint4*>(p)[i] = z;
      } else {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          reinterpret_cast<uint4*>(p + HD)[i] = z;
          reinterpret_cast<uint4*>(p + 2 * HD)[i] = z;
        }
      }
    }
    return;
  }

  const uint32_t sX1 = smem_base, sX2 = smem_base + AB_TILE_BYTES;
  auto sY1 = [&](int st) { return smem_base + (2 + 2 * st) * AB_TILE_BYTES; };
  auto sY2 = [&](int st) { return smem_base + (3 + 2 * st) * AB_TILE_BYTES; };
  const uint32_t sA = smem_base + 6 * AB_TILE_BYTES, sB = smem_base + 8 * AB_TILE_BYTES;
  const uint32_t bar_base = smem_base + 10 * AB_TILE_BYTES;
  const uint32_t bar_x = bar_base, bar_s = bar_base + 16, bar_p = bar_base + 24, bar_acc = bar_base + 32,
                 tmem_slot = bar_base + 40;
  auto bar_y = [&](int st) { return bar_base + 8u + 40u * uint32_t(st); };  // +8 and +48
  float* s_stat = reinterpret_cast<float*>(smem_gen + 10 * AB_TILE_BYTES + 64);  // [2][2][128]: buffer, {lse, delta}

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmQK);
    tma_prefetch_desc(&tmV);
    tma_prefetch_desc(&tmDO);
    mbar_init(bar_x, 1);
    mbar_init(bar_y(0), 1);
    mbar_init(bar_y(1), 1);
    mbar_init(bar_s, 1);
    mbar_init(bar_p, 128);
    mbar_init(bar_acc, 1);
    fence_mbar_init();
  }
  if (warp == 0) {
    tmem_alloc(tmem_slot, AB_TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));
  const uint32_t tmem_S = tmem_base, tmem_dP = tmem_base + 128, tmem_acc1 = tmem_base + 256, tmem_acc2 = tmem_base + 320;

  const int n_pre = (MODE == 0 && !args.have_lse) ? nt : 0;  // LSE pre-pass iterations
  const int n_it = n_pre + nt;

  if (warp == 0) {
    if (lane == 0) {
      constexpr uint32_t idesc_s = make_idesc_bf16(128, 128, 0, 0);
      constexpr uint32_t idesc_acc = make_idesc_bf16(128, 64, 0, 1);
      const uint64_t x1desc = make_smem_desc_sw128(sX1, 16, 1024), x2desc = make_smem_desc_sw128(sX2, 16, 1024);
      uint64_t y1desc[2], y2desc[2], y1mn[2], y2mn[2];
#pragma unroll
      for (int st = 0; st < 2; ++st) {
        y1desc[st] = make_smem_desc_sw128(sY1(st), 16, 1024);
        y2desc[st] = make_smem_desc_sw128(sY2(st), 16, 1024);
        y1mn[st] = make_smem_desc_sw128(sY1(st), 1024, 1024);
        y2mn[st] = make_smem_desc_sw128(sY2(st), 1024, 1024);
      }
      const uint64_t a0 = make_smem_desc_sw128(sA, 16, 1024), a1 = make_smem_desc_sw128(sA + AB_TILE_BYTES, 16, 1024);
      const uint64_t b0 = make_smem_desc_sw128(sB, 16, 1024), b1 = make_smem_desc_sw128(sB + AB_TILE_BYTES, 16, 1024);
      // owner tiles
      mbar_arrive_expect_tx(bar_x, 2 * AB_TILE_BYTES);
      if (MODE == 0) {
        tma_load_3d(sX1, &tmQK, bar_x, h * AB_D, tile * AB_TILE, b);   // Q
        tma_load_3d(sX2, &tmDO, bar_x, h * AB_D, tile * AB_TILE, b);   // dO
      } else {
        tma_load_3d(sX1, &tmQK, bar_x, HD + h * AB_D, tile * AB_TILE, b);  // K
        tma_load_3d(sX2, &tmV, bar_x, h * AB_D, tile * AB_TILE, b);        // V
      }
      auto fetch = [&](int it) {  // the Y tiles of iteration `it` into buffer it & 1
        const bool pre = it < n_pre;
        const int j = pre ? it : it - n_pre;
        const int st = it & 1;
        mbar_arrive_expect_tx(bar_y(st), (pre ? 1 : 2) * AB_TILE_BYTES);
        if (MODE == 0) {
          tma_load_3d(sY1(st), &tmQK, bar_y(st), HD + h * AB_D, j * AB_TILE, b);  // K_j
          if (!pre) tma_load_3d(sY2(st), &tmV, bar_y(st), h * AB_D, j * AB_TILE, b);  // V_j
        } else {
          tma_load_3d(sY1(st), &tmQK, bar_y(st), h * AB_D, j * AB_TILE, b);   // Q_i
          tma_load_3d(sY2(st), &tmDO, bar_y(st), h * AB_D, j * AB_TILE, b);   // dO_i
        }
      };
      fetch(0);
      mbar_wait(bar_x, 0, 1);
      int n_acc = 0;
      for (int it = 0; it < n_it; ++it) {
        const bool pre = it < n_pre;
        const int j = pre ? it : it - n_pre;
        const int st = it & 1;
        mbar_wait(bar_y(st), (it >> 1) & 1u, 2);
        tc_fence_after();
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_bf16_ss(tmem_S, x1desc + uint64_t(2 * k), y1desc[st] + uint64_t(2 * k), idesc_s, k != 0);
        if (!pre) {
#pragma unroll
          for (int k = 0; k < 4; ++k) umma_bf16_ss(tmem_dP, x2desc + uint64_t(2 * k), y2desc[st] + uint64_t(2 * k), idesc_s, k != 0);
        }
        umma_commit(bar_s);
        // the accumulating MMAs of the previous iteration were left in flight behind this iteration's S / dP; they
        // must have retired before their Y buffer is refilled (the staging tiles are protected by bar_s, whose commit
        // covers every earlier MMA of this thread)
        if (it > n_pre) mbar_wait(bar_acc, (n_acc - 1) & 1u, 4);
        if (it + 1 < n_it) fetch(it + 1);  // buffer (it + 1) & 1 was last read by the MMAs of iteration it - 1
        mbar_wait(bar_p, it & 1u, 3);  // S (and dP) consumed; staging written
        if (!pre) {
          tc_fence_after();
          const uint32_t accf = j == 0 ? 0u : 1u;
          if (MODE == 0) {
#pragma unroll
            for (int kk = 0; kk < 8; ++kk)  // dQ += dS K_j
              umma_bf16_ss(tmem_acc2, (kk < 4 ? a0 : a1) + uint64_t(2 * (kk & 3)), y1mn[st] + uint64_t(128 * kk), idesc_acc,
                           kk != 0 ? 1u : accf);
          } else {
#pragma unroll
            for (int kk = 0; kk < 8; ++kk)  // dV += P^T dO_i
              umma_bf16_ss(tmem_acc1, (kk < 4 ? a0 : a1) + uint64_t(2 * (kk & 3)), y2mn[st] + uint64_t(128 * kk), idesc_acc,
                           kk != 0 ? 1u : accf);
#pragma unroll
            for (int kk = 0; kk < 8; ++kk)  // dK += dS^T Q_i
              umma_bf16_ss(tmem_acc2, (kk < 4 ? b0 : b1) + uint64_t(2 * (kk & 3)), y1mn[st] + uint64_t(128 * kk), idesc_acc,
                           kk != 0 ? 1u : accf);
          }
          umma_commit(bar_acc);
          ++n_acc;
        }
      }
    }
    __syncwarp();
  } else {
    // ===================== one thread per owner row =====================
    const int q4 = warp & 3;
    const int r = q4 * 32 + lane;
    const uint32_t lane_off = uint32_t(q4 * 32) << 16;
    const int t_own = tile * AB_TILE + r;
    const bool own_valid = t_own < len;
    const float c = args.scale_log2;
    const long long stat_base = ((long long)b * args.heads + h) * args.rows_per_batch;
    float lse2 = 0.f, delta = 0.f;
    if (MODE == 0) {
      if (own_valid) {
        const uint4* po = reinterpret_cast<const uint4*>(args.o + (row_base + t_own) * args.ld_o + h * AB_D);
        const uint4* pd = reinterpret_cast<const uint4*>(args.d_o + (row_base + t_own) * args.ld_do + h * AB_D);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const uint4 a = po[i], d = pd[i];
          const uint32_t aw[4] = {a.x, a.y, a.z, a.w}, dw[4] = {d.x, d.y, d.z, d.w};
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const float2 fa = make_float2(__uint_as_float(aw[k] << 16), __uint_as_float(aw[k] & 0xffff0000u));
            const float2 fd = make_float2(__uint_as_float(dw[k] << 16), __uint_as_float(dw[k] & 0xffff0000u));
            delta = fmaf(fa.x, fd.x, delta);
            delta = fmaf(fa.y, fd.y, delta);
          }
        }
      }
      // ---- pre-pass: log-sum-exp of the row (log2 domain) ----
      float m = -INFINITY, l = 0.f;
      for (int it = 0; it < n_pre; ++it) {
        const int nv = min(AB_TILE, len - it * AB_TILE);
        mbar_wait(bar_s, it & 1u, 5);
        tc_fence_after();
#pragma unroll 1
        for (int c0 = 0; c0 < AB_TILE; c0 += 32) {
          if (c0 >= nv) break;
          uint32_t v[32];
          ab_tmem_ld32(tmem_S + lane_off + c0, v);
          tmem_wait_ld();
          float cm = -INFINITY;
#pragma unroll
          for (int k = 0; k < 32; ++k)
            if (c0 + k < nv) cm = fmaxf(cm, __uint_as_float(v[k]) * c);
          const float mn = fmaxf(m, cm);
          float s = 0.f;
#pragma unroll
          for (int k = 0; k < 32; ++k)
            if (c0 + k < nv) s += ex2_approx(fmaf(__uint_as_float(v[k]), c, -mn));
          l = l * ex2_approx(m - mn) + s;
          m = mn;
        }
        tc_fence_before();
        mbar_arrive(bar_p);
      }
      if (args.have_lse) {
        lse2 = t_own < args.rows_per_batch ? args.lse[stat_base + t_own] : 0.f;
      } else {
        lse2 = m + log2f(l);
      }
      if (t_own < args.rows_per_batch) {
        if (!args.have_lse) args.lse[stat_base + t_own] = lse2;
        args.delta[stat_base + t_own] = delta;
      }
    }
    // ---- main pass ----
    for (int it = n_pre; it < n_it; ++it) {
      const int j = it - n_pre;
      const int nv = min(AB_TILE, len - j * AB_TILE);  // valid columns of this tile (keys in MODE 0, queries in MODE 1)
      const float* st = s_stat + (j & 1) * 256;
      if (MODE == 1) {
        const int tq = j * AB_TILE + r;
        float* sw_ = s_stat + (j & 1) * 256;
        sw_[r] = tq < args.rows_per_batch ? args.lse[stat_base + tq] : 0.f;
        sw_[128 + r] = tq < args.rows_per_batch ? args.delta[stat_base + tq] : 0.f;
        asm volatile("bar.sync 1, 128;" ::: "memory");
      }
      mbar_wait(bar_s, it & 1u, 6);
      tc_fence_after();
#pragma unroll 1
      for (int c0 = 0; c0 < AB_TILE; c0 += 32) {
        uint32_t vs[32], vd[32];
        ab_tmem_ld32(tmem_S + lane_off + c0, vs);
        ab_tmem_ld32(tmem_dP + lane_off + c0, vd);
        tmem_wait_ld();
        uint32_t pp[16], pd[16];
#pragma unroll
        for (int k = 0; k < 32; k += 2) {
          float p0 = 0.f, p1 = 0.f, d0 = 0.f, d1 = 0.f;
          const float l0 = MODE == 0 ? lse2 : st[c0 + k], l1 = MODE == 0 ? lse2 : st[c0 + k + 1];
          const float e0 = MODE == 0 ? delta : st[128 + c0 + k], e1 = MODE == 0 ? delta : st[128 + c0 + k + 1];
          if (own_valid && c0 + k < nv) {
            p0 = ex2_approx(fmaf(__uint_as_float(vs[k]), c, -l0));
            d0 = p0 * (__uint_as_float(vd[k]) - e0) * args.scale;
          }
          if (own_valid && c0 + k + 1 < nv) {
            p1 = ex2_approx(fmaf(__uint_as_float(vs[k + 1]), c, -l1));
            d1 = p1 * (__uint_as_float(vd[k + 1]) - e1) * args.scale;
          }
          pp[k >> 1] = pack_bf16x2(p0, p1);
          pd[k >> 1] = pack_bf16x2(d0, d1);
        }
        if (MODE == 0) {
          ab_stage_store(sA, r, c0, pd);
        } else {
          ab_stage_store(sA, r, c0, pp);
          ab_stage_store(sB, r, c0, pd);
        }
      }
      fence_proxy_async_smem();
      tc_fence_before();
      mbar_arrive(bar_p);
    }
    // ---- read the accumulators ----
    mbar_wait(bar_acc, (nt - 1) & 1u, 7);
    tc_fence_after();
    const bool in_range = t_own < args.rows_per_batch;
    float cs[32], sn[32];
    if (own_valid) {
#pragma unroll
      for (int i = 0; i < 32; ++i) {
        cs[i] = args.rope_cos[(long long)t_own * 32 + i];
        sn[i] = args.rope_sin[(long long)t_own * 32 + i];
      }
    }
    auto emit = [&](uint32_t tm, int col0, bool rope) {
      uint32_t lo[32], hi[32];
      ab_tmem_ld32(tm + lane_off, lo);
      ab_tmem_ld32(tm + lane_off + 32, hi);
      tmem_wait_ld();
      uint32_t out[32];
      if (own_valid) {
#pragma unroll
        for (int i = 0; i < 32; i += 2) {
          float a0 = __uint_as_float(lo[i]), a1 = __uint_as_float(lo[i + 1]);
          float b0 = __uint_as_float(hi[i]), b1 = __uint_as_float(hi[i + 1]);
          if (rope) {  // transpose of q' = q cos + rotate_half(q) sin
            const float x0 = a0 * cs[i] + b0 * sn[i], y0 = b0 * cs[i] - a0 * sn[i];
            const float x1 = a1 * cs[i + 1] + b1 * sn[i + 1], y1 = b1 * cs[i + 1] - a1 * sn[i + 1];
            a0 = x0; b0 = y0; a1 = x1; b1 = y1;
          }
          out[i >> 1] = pack_bf16x2(a0, a1);
          out[16 + (i >> 1)] = pack_bf16x2(b0, b1);
        }
      } else {
#pragma unroll
        for (int i = 0; i < 32; ++i) out[i] = 0u;
      }
      if (in_range) {
        uint4* p = reinterpret_cast<uint4*>(args.dqkv + (row_base + t_own) * args.ld_dqkv + col0 + h * AB_D);
#pragma unroll
        for (int i = 0; i < 8; ++i) p[i] = make_uint4(out[4 * i], out[4 * i + 1], out[4 * i + 2], out[4 * i + 3]);
      }
    };
    if (MODE == 0) {
      emit(tmem_acc2, 0, true);
    } else {
      emit(tmem_acc1, 2 * HD, false);
      emit(tmem_acc2, HD, true);
    }
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem_base, AB_TMEM_COLS);
  }
}

}  // namespace oron
