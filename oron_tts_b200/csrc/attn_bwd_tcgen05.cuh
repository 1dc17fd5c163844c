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
  long long* dbg;  // optional [grid, 8] clock64 stamps of the first compute thread (tools/attn_bwd_trace.py)
  int have_lse;  // 1: `lse` was written by the forward kernel (oron_attention_fwd_lse): MODE 0 skips its pre-pass
};

constexpr int AB_THREADS = 288;  // warp 0: TMA + MMA issue (+ TMEM alloc); warps 1..8: two threads per owner row
                                 // (warp & 3 = TMEM lane quarter, (warp - 1) >> 2 = which 64 of the tile's 128 columns)
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
    mbar_init(bar_p, 256);
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
    // ===================== two threads per owner row =====================
    const int q4 = warp & 3;
    const int half = (warp - 1) >> 2;  // columns [64 * half, 64 * half + 64) of every S / dP tile
    const int r = q4 * 32 + lane;
    const uint32_t lane_off = uint32_t(q4 * 32) << 16;
    const int t_own = tile * AB_TILE + r;
    const bool own_valid = t_own < len;
    const float c = args.scale_log2;
    const long long stat_base = ((long long)b * args.heads + h) * args.rows_per_batch;
    const bool tr = args.dbg != nullptr && threadIdx.x == 32;
#define AB_STAMP(slot) do { if (tr) args.dbg[(long long)blockIdx.x * 8 + (slot)] = clock64(); } while (0)
    AB_STAMP(0);
    float lse2 = 0.f, delta = 0.f;
    if (MODE == 0) {
      if (args.have_lse) lse2 = t_own < args.rows_per_batch ? __ldg(args.lse + stat_base + t_own) : 0.f;  // in flight under the delta loads
      // delta[row] = sum_d dO * O, cooperatively: 8 lanes per row read one 16-byte chunk each (coalesced), shuffle-reduce
      {
        const int tid = int(threadIdx.x) - 32;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int q = i * 256 + tid;
          const int row = q >> 3, k = q & 7;
          const int t = tile * AB_TILE + row;
          float part = 0.f;
          if (t < len) {
            const uint4 a = *reinterpret_cast<const uint4*>(args.o + (row_base + t) * args.ld_o + h * AB_D + 8 * k);
            const uint4 d = *reinterpret_cast<const uint4*>(args.d_o + (row_base + t) * args.ld_do + h * AB_D + 8 * k);
            const uint32_t aw[4] = {a.x, a.y, a.z, a.w}, dw[4] = {d.x, d.y, d.z, d.w};
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              part = fmaf(__uint_as_float(aw[e] << 16), __uint_as_float(dw[e] << 16), part);
              part = fmaf(__uint_as_float(aw[e] & 0xffff0000u), __uint_as_float(dw[e] & 0xffff0000u), part);
            }
          }
          part += __shfl_xor_sync(0xffffffffu, part, 1);
          part += __shfl_xor_sync(0xffffffffu, part, 2);
          part += __shfl_xor_sync(0xffffffffu, part, 4);
          if (k == 0) s_stat[128 + row] = part;
        }
        asm volatile("bar.sync 1, 256;" ::: "memory");
        delta = s_stat[128 + r];
      }
      // ---- pre-pass: log-sum-exp of the row (log2 domain) ----
      float m = -INFINITY, l = 0.f;
      for (int it = 0; it < n_pre; ++it) {  // (fallback path, have_lse == 0: the first thread of each row does all 128 columns)
        const int nv = min(AB_TILE, len - it * AB_TILE);
        mbar_wait(bar_s, it & 1u, 5);
        tc_fence_after();
#pragma unroll 1
        for (int c0 = 0; c0 < AB_TILE; c0 += 32) {
          if (c0 >= nv || half != 0) break;
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
      if (!args.have_lse) {
        if (half == 0) s_stat[r] = m + log2f(l);
        asm volatile("bar.sync 1, 256;" ::: "memory");
        lse2 = s_stat[r];
        asm volatile("bar.sync 1, 256;" ::: "memory");  // s_stat is reused by nobody in MODE 0, kept for symmetry
      }
      if (half == 0 && t_own < args.rows_per_batch) {
        if (!args.have_lse) args.lse[stat_base + t_own] = lse2;
        args.delta[stat_base + t_own] = delta;
      }
    }
    AB_STAMP(1);
    // ---- main pass ----
    for (int it = n_pre; it < n_it; ++it) {
      const int j = it - n_pre;
      const int nv = min(AB_TILE, len - j * AB_TILE);  // valid columns of this tile (keys in MODE 0, queries in MODE 1)
      const float* st = s_stat + (j & 1) * 256;
      if (MODE == 1) {
        const int tq = j * AB_TILE + r;
        float* sw_ = s_stat + (j & 1) * 256;
        const float* src = half == 0 ? args.lse : args.delta;
        sw_[128 * half + r] = tq < args.rows_per_batch ? src[stat_base + tq] : 0.f;
        asm volatile("bar.sync 1, 256;" ::: "memory");
      }
      mbar_wait(bar_s, it & 1u, 6);
      tc_fence_after();
      if (it == n_pre) AB_STAMP(2);
      if (it == n_pre + 1) AB_STAMP(3);
#pragma unroll 1
      for (int c0 = 64 * half; c0 < 64 * half + 64; c0 += 32) {
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
    AB_STAMP(4);
    // ---- read the accumulators ----
    mbar_wait(bar_acc, (nt - 1) & 1u, 7);
    tc_fence_after();
    AB_STAMP(5);
    // Phase A: the accumulator rows go to shared memory as f32 (row = 256 B, 16-byte chunks XOR-swizzled by row & 15)
    // in the idle staging tiles; phase B: all threads store 16-byte bf16 chunks, 8 lanes per row (coalesced), with the
    // RoPE rotation of dq / dk applied on the way (cos / sin read coalesced too).
    auto to_smem = [&](uint32_t tm, uint32_t tile_s) {
      uint32_t lo[32], hi[32];
      ab_tmem_ld32(tm + lane_off, lo);
      ab_tmem_ld32(tm + lane_off + 32, hi);
      tmem_wait_ld();
      const uint32_t rowa = tile_s + uint32_t(r) * 256u;
      const uint32_t sw = uint32_t(r & 15);
#pragma unroll
      for (int cidx = 0; cidx < 8; ++cidx) {
        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(rowa + ((uint32_t(cidx) ^ sw) << 4)), "r"(lo[4 * cidx]),
                     "r"(lo[4 * cidx + 1]), "r"(lo[4 * cidx + 2]), "r"(lo[4 * cidx + 3]) : "memory");
        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(rowa + ((uint32_t(cidx + 8) ^ sw) << 4)), "r"(hi[4 * cidx]),
                     "r"(hi[4 * cidx + 1]), "r"(hi[4 * cidx + 2]), "r"(hi[4 * cidx + 3]) : "memory");
      }
    };
    auto ld_chunk = [&](uint32_t tile_s, int row, int cidx, float (&f)[4]) {
      uint32_t a0, a1, a2, a3;
      asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(a0), "=r"(a1), "=r"(a2), "=r"(a3)
                   : "r"(tile_s + uint32_t(row) * 256u + ((uint32_t(cidx) ^ uint32_t(row & 15)) << 4)) : "memory");
      f[0] = __uint_as_float(a0); f[1] = __uint_as_float(a1); f[2] = __uint_as_float(a2); f[3] = __uint_as_float(a3);
    };
    // q: task index (row = q >> 3, 16-byte output chunk k = q & 7)
    auto store_task = [&](uint32_t tile_s, int q, int col0, bool rope) {
      const int row = q >> 3, k = q & 7;
      const int t = tile * AB_TILE + row;
      if (t >= args.rows_per_batch) return;
      uint4 outv = make_uint4(0u, 0u, 0u, 0u);
      if (t < len) {
        float v[8];
        if (!rope) {
          float x[4], y[4];
          ld_chunk(tile_s, row, 2 * k, x);
          ld_chunk(tile_s, row, 2 * k + 1, y);
#pragma unroll
          for (int e = 0; e < 4; ++e) { v[e] = x[e]; v[4 + e] = y[e]; }
        } else {  // transpose of q' = q cos + rotate_half(q) sin: columns c (< 32) and c + 32 mix
          const int kk = k & 3;
          float a[8], bb[8];
          { float x[4], y[4]; ld_chunk(tile_s, row, 2 * kk, x); ld_chunk(tile_s, row, 2 * kk + 1, y);
#pragma unroll
            for (int e = 0; e < 4; ++e) { a[e] = x[e]; a[4 + e] = y[e]; } }
          { float x[4], y[4]; ld_chunk(tile_s, row, 8 + 2 * kk, x); ld_chunk(tile_s, row, 8 + 2 * kk + 1, y);
#pragma unroll
            for (int e = 0; e < 4; ++e) { bb[e] = x[e]; bb[4 + e] = y[e]; } }
          const float4* pc = reinterpret_cast<const float4*>(args.rope_cos + (long long)t * 32 + 8 * kk);
          const float4* ps = reinterpret_cast<const float4*>(args.rope_sin + (long long)t * 32 + 8 * kk);
          const float4 c0 = __ldg(pc), c1 = __ldg(pc + 1), s0 = __ldg(ps), s1 = __ldg(ps + 1);
          const float cs[8] = {c0.x, c0.y, c0.z, c0.w, c1.x, c1.y, c1.z, c1.w};
          const float sn[8] = {s0.x, s0.y, s0.z, s0.w, s1.x, s1.y, s1.z, s1.w};
#pragma unroll
          for (int e = 0; e < 8; ++e) v[e] = k < 4 ? a[e] * cs[e] + bb[e] * sn[e] : bb[e] * cs[e] - a[e] * sn[e];
        }
        outv = make_uint4(pack_bf16x2(v[0], v[1]), pack_bf16x2(v[2], v[3]), pack_bf16x2(v[4], v[5]), pack_bf16x2(v[6], v[7]));
      }
      *reinterpret_cast<uint4*>(args.dqkv + (row_base + t) * args.ld_dqkv + col0 + h * AB_D + 8 * k) = outv;
    };
    const int tid = int(threadIdx.x) - 32;
    if (MODE == 0) {
      if (half == 0) to_smem(tmem_acc2, sA);
      asm volatile("bar.sync 1, 256;" ::: "memory");
#pragma unroll 1
      for (int i = 0; i < 4; ++i) store_task(sA, i * 256 + tid, 0, true);
    } else {
      if (half == 0) to_smem(tmem_acc1, sA); else to_smem(tmem_acc2, sB);
      asm volatile("bar.sync 1, 256;" ::: "memory");
#pragma unroll 1
      for (int i = 0; i < 8; ++i) {
        if (half == 0) store_task(sA, i * 128 + (tid & 127), 2 * HD, false);
        else store_task(sB, i * 128 + (tid & 127), HD, true);
      }
    }
    AB_STAMP(6);
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem_base, AB_TMEM_COLS);
  }
}

}  // namespace oron
