// Non-causal multi-head attention with per-sequence key lengths (replaces the reference's
// F.scaled_dot_product_attention + key-padding mask, src/models/modules.py:271-278).
// One CTA = one (batch element, head, 128-query tile); two CTAs per SM. head_dim = 64.
//
//   S = Q K^T   : tcgen05.mma M=128 N=128 K=64 -> TMEM cols [0,128)
//   P = softmax : 128 softmax threads, one query row each (TMEM lane == row => no shuffles), two passes over
//                 TMEM (row max, then exp2 -> f16 into a SW128 K-major smem tile)
//   O += P V    : tcgen05.mma (f16 x f16) M=128 N=64 K=128 accumulating IN TMEM (cols [128,192)), V tile as
//                 MN-major B; V is written as f16 by the QKV GEMM epilogue.
//   l           : row sums of the (fp32) probabilities on the CUDA cores. A tensor-core version (P times an
//                 all-ones tile, M=128 N=16) was measured slower: every tcgen05.mma of the dependent accumulate
//                 chain costs ~65 cycles to issue and ~110 to retire whatever its N, and that chain
//                 (p_full -> P V -> o_full) is the per-tile critical loop.
//
// Measured (per-CTA clock64 traces, tools/attn_trace.py, profiles/r01_attn_trace_*.txt):
//   * the exp2 unit is the bound: ~12.5 cycles per MUFU warp instruction, i.e. ~1.6-1.7k cycles per 128x128 tile
//     per scheduler. With two CTAs per SM the first wave (296 of the 352 CTAs at config 2) keeps it ~100 % busy;
//     what is left is the second, 19 %-full wave.
//   * everything else is therefore kept off the softmax threads: O stays in TMEM for the whole KV loop and the
//     running maximum is only raised when a tile exceeds it by more than 2^8 ("lazy rescale": then O is read
//     back, scaled and stored, a warp-uniform rare branch); probabilities carry a 2^7 bias so they use the f16
//     range (cancels in O / l); S is released to the MMA thread after the last TMEM read of pass 2 (the last
//     32-key chunk stays in registers from pass 1) so Q K^T of the next tile runs under the exp phase.
//   * every tcgen05.mma of a dependent accumulate chain costs ~65-80 cycles to issue and ~110 to retire whatever
//     its N, so p_full -> P V -> o_full is a per-tile critical loop: descriptors are loop invariant, the first P
//     slab is handed back after four k-steps (p0_free), and a tensor-core row sum (P x ones, N=16) was dropped.
//   * tried and measured slower (kept in git history): packed f16x2 ex2 (ptxas splits it into two MUFU ops),
//     half of the exponentials as an FMA-pipe polynomial (pass 2 1.7k -> 2.4k cycles), two softmax threads per
//     row with a per-tile row-max exchange (barrier: +1-2k cycles per tile) or with independent key halves and
//     two accumulators (register cap 96 at 2 x 320 threads, same MUFU bound: 64.5 us vs 51 us per call).
// q/k/v are read straight out of the fused QKV activation [rows, 3*H*64] with one 3-D TMA map.
#pragma once
#include <cuda_fp16.h>

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
  long long* dbg;       // optional [grid, 16] clock64 stamps (tools/attn_trace.py); nullptr in production
  // work decomposition (see attn_plan in api.cu): items = (batch, head, 128-query tile) in linear order. The first
  // n_full items run whole, one CTA each; every later ("tail") item is split along the keys into `parts` CTAs whose
  // partial (O, max, sum) results are merged by whichever of them finishes last.
  int q_tiles;          // 128-row query tiles per batch element
  int n_full;
  int parts;            // >= 1
  float* ws_o;          // [n_tail_items * parts][128][64] f32 partial numerators
  float* ws_ml;         // [n_tail_items * parts][128][2] f32 (running max * c, denominator)
  int* ws_cnt;          // [n_tail_items] arrival counters, zero before the first launch (the kernel re-zeroes them)
};

constexpr int ATT_THREADS = 192;  // warp0 TMA, warp1 MMA + TMEM alloc, warps 2..5 softmax
constexpr int ATT_TILE = 128;
constexpr int ATT_D = 64;
constexpr int ATT_TILE_BYTES = ATT_TILE * ATT_D * 2;  // 16 KB
// smem: Q | K0 K1 | V0 V1 | P(2 slabs) | barriers
constexpr int ATT_ONES_BYTES = 512;  // [16 x 16] f16 ones, no-swizzle K-major core matrices
constexpr int ATT_SMEM_BYTES = 7 * ATT_TILE_BYTES + 128 + ATT_ONES_BYTES;
constexpr int ATT_TMEM_COLS = 256;
#define ATT_STAMP(slot) do { if (args.dbg) args.dbg[(long long)blockIdx.x * 16 + (slot)] = clock64(); } while (0)
constexpr float ATT_RESCALE_LOG2 = 8.0f;  // raise the running max only when exceeded by > 2^8
constexpr float ATT_P_EXP_BIAS = 7.0f;    // probabilities are scaled by 2^7 (<= 2^15 in f16); cancels in O / l

__device__ __forceinline__ void tmem_st_32x32(uint32_t taddr, const uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
      :
      : "r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]),
        "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]),
        "r"(r[17]), "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]),
        "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
      : "memory");
}
__device__ __forceinline__ void tmem_wait_st() {
  asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_ld_32x16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_st_32x16(uint32_t taddr, const uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
      :
      : "r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]),
        "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}

// 2^x on the FMA/ALU pipes (no MUFU): round-to-nearest split x = n + f with the 1.5*2^23 magic constant,
// degree-3 minimax polynomial for 2^f on [-0.5, 0.5] (max rel. error 7.5e-5 — f16 resolution is 4.9e-4), and
// n added straight into the exponent field. Valid for -125 <= x < 2^22.
__device__ __forceinline__ float exp2_poly(float x) {
  x = fmaxf(x, -125.0f);
  const float t = x + 12582912.0f;
  const float f = x - (t - 12582912.0f);
  float p = fmaf(0.0551716573536396f, f, 0.2426111251115799f);
  p = fmaf(p, f, 0.6932609677314758f);
  p = fmaf(p, f, 0.9999280571937561f);
  return __int_as_float(__float_as_int(p) + (__float_as_int(t) << 23));
}

// exp2 + swizzled store of one 32-key chunk of a P row. x = s*c - (m*c - 7); results are packed to the f16 P tile.
// MASKED: keys >= n_valid get probability 0.
template <bool MASKED>
__device__ __forceinline__ float softmax_chunk(const uint32_t (&v)[32], const float c, const float mcb, const int c0,
                                               const int n_valid, const uint32_t prow, const uint32_t sw) {
  uint32_t pk[16];
  float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
#pragma unroll
  for (int i = 0; i < 32; i += 4) {
    const float x0 = fmaf(__uint_as_float(v[i]), c, -mcb);
    const float x1 = fmaf(__uint_as_float(v[i + 1]), c, -mcb);
    const float x2 = fmaf(__uint_as_float(v[i + 2]), c, -mcb);
    const float x3 = fmaf(__uint_as_float(v[i + 3]), c, -mcb);
    // all four through the exp2 unit: replacing half of them by exp2_poly (FMA pipe) was measured SLOWER
    // (pass 2: 1.7k -> 2.4k cycles per tile): with one softmax warp per scheduler the loop is bound by the
    // instruction count, not by MUFU throughput.
    float p0 = ex2_approx(x0), p1 = ex2_approx(x1);
    float p2 = ex2_approx(x2), p3 = ex2_approx(x3);
    if (MASKED) {
      if (c0 + i >= n_valid) p0 = 0.f;
      if (c0 + i + 1 >= n_valid) p1 = 0.f;
      if (c0 + i + 2 >= n_valid) p2 = 0.f;
      if (c0 + i + 3 >= n_valid) p3 = 0.f;
    }
    s0 += p0; s1 += p1; s2 += p2; s3 += p3;
    __half2 a = __floats2half2_rn(p0, p1), b = __floats2half2_rn(p2, p3);
    pk[i / 2] = *reinterpret_cast<uint32_t*>(&a);
    pk[i / 2 + 1] = *reinterpret_cast<uint32_t*>(&b);
  }
  const uint32_t slab = prow + (c0 >> 6) * ATT_TILE_BYTES;
  const uint32_t chunk0 = uint32_t(c0 & 63) >> 3;  // first 16-byte chunk of this 32-key group inside its slab row
#pragma unroll
  for (int g = 0; g < 4; ++g) {
    const uint32_t addr = slab + (((chunk0 + g) ^ sw) << 4);
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(pk[4 * g]), "r"(pk[4 * g + 1]),
                 "r"(pk[4 * g + 2]), "r"(pk[4 * g + 3])
                 : "memory");
  }
  return (s0 + s1) + (s2 + s3);
}

__device__ __forceinline__ float max32(const uint32_t (&v)[32]) {
  float m0 = __uint_as_float(v[0]), m1 = __uint_as_float(v[1]), m2 = __uint_as_float(v[2]), m3 = __uint_as_float(v[3]);
#pragma unroll
  for (int i = 4; i < 32; i += 4) {
    m0 = fmaxf(m0, __uint_as_float(v[i]));
    m1 = fmaxf(m1, __uint_as_float(v[i + 1]));
    m2 = fmaxf(m2, __uint_as_float(v[i + 2]));
    m3 = fmaxf(m3, __uint_as_float(v[i + 3]));
  }
  return fmaxf(fmaxf(m0, m1), fmaxf(m2, m3));
}

__global__ void __launch_bounds__(ATT_THREADS, 2)
attn_fwd_tcgen05_kernel(const __grid_constant__ CUtensorMap tmQKV, const AttnArgs args) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t smem_base = smem_u32(smem_raw);
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  pdl_launch_dependents();
  int item = blockIdx.x, part = 0, nparts = 1;
  if (item >= args.n_full) {
    const int u = item - args.n_full;
    item = args.n_full + u / args.parts;
    part = u % args.parts;
    nparts = args.parts;
  }
  const int q_tile = item % args.q_tiles;
  const int h = (item / args.q_tiles) % args.heads;
  const int b = item / (args.q_tiles * args.heads);
  const int q0 = q_tile * ATT_TILE;
  const int len = args.seq_lens ? min(args.seq_lens[b], args.rows_per_batch) : args.rows_per_batch;
  if (q0 >= len) return;  // whole tile is padding (all parts of the item agree): the out-projection masks these rows
  if ((smem_base & 1023u) != 0) {
    if (threadIdx.x == 0) printf("[oron] attention: dynamic smem not 1024-byte aligned\n");
    __trap();
  }
  const int n_kv = (len + ATT_TILE - 1) / ATT_TILE;
  const int j_begin = (part * n_kv) / nparts;                 // this CTA's key tiles: [j_begin, j_begin + n_loc)
  const int n_loc = ((part + 1) * n_kv) / nparts - j_begin;   // may be 0 when the sequence has fewer tiles than parts

  const uint32_t sQ = smem_base;
  auto sK = [&](int s) { return smem_base + (1 + s) * ATT_TILE_BYTES; };
  auto sV = [&](int s) { return smem_base + (3 + s) * ATT_TILE_BYTES; };
  const uint32_t sP = smem_base + 5 * ATT_TILE_BYTES;
  const uint32_t bar_base = smem_base + 7 * ATT_TILE_BYTES;
  const uint32_t q_full = bar_base;
  auto kv_full = [&](int s) { return bar_base + 8u * (1 + s); };
  auto kv_empty = [&](int s) { return bar_base + 8u * (3 + s); };
  const uint32_t s_full = bar_base + 8u * 5;   // MMA -> softmax: S(j) is in TMEM
  const uint32_t s_free = bar_base + 8u * 6;   // softmax -> MMA: S(j) has been read (128 arrivals)
  const uint32_t p_full = bar_base + 8u * 7;   // softmax -> MMA: P(j) is in smem (128 arrivals)
  const uint32_t o_full = bar_base + 8u * 8;   // MMA -> softmax: O includes P(j) V(j)
  const uint32_t p0_free = bar_base + 8u * 9;  // MMA -> softmax: P V has consumed the first 64-key slab of P(j)
  const uint32_t tmem_slot = bar_base + 8u * 10;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmQKV);
    mbar_init(q_full, 1);
    for (int s = 0; s < 2; ++s) { mbar_init(kv_full(s), 1); mbar_init(kv_empty(s), 1); }
    mbar_init(s_full, 1);
    mbar_init(s_free, 128);
    mbar_init(p_full, 128);
    mbar_init(o_full, 1);
    mbar_init(p0_free, 1);
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
  pdl_wait();  // the QKV activations of the previous kernel are visible from here on
  if (threadIdx.x == 0) {
    ATT_STAMP(0);
    if (args.dbg) { unsigned long long gt; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(gt)); args.dbg[(long long)blockIdx.x * 16 + 8] = (long long)gt; }
  }

  const int HD = args.heads * ATT_D;
  if (warp == 0) {
    if (lane == 0) {
      mbar_arrive_expect_tx(q_full, ATT_TILE_BYTES);
      tma_load_3d(sQ, &tmQKV, q_full, h * ATT_D, q0, b);
      for (int j = 0; j < n_loc; ++j) {
        const int s = j & 1;
        mbar_wait(kv_empty(s), ((j >> 1) & 1u) ^ 1u, 11);
        mbar_arrive_expect_tx(kv_full(s), 2 * ATT_TILE_BYTES);
        tma_load_3d(sK(s), &tmQKV, kv_full(s), HD + h * ATT_D, (j_begin + j) * ATT_TILE, b);
        tma_load_3d(sV(s), &tmQKV, kv_full(s), 2 * HD + h * ATT_D, (j_begin + j) * ATT_TILE, b);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc_s = make_idesc_bf16(128, 128, 0, 0);
      constexpr uint32_t idesc_o = make_idesc_f16(128, 64, 0, 1);  // f16 P x f16 V, B = V is MN-major
      // all shared-memory descriptors are loop invariant (two K/V stages): build them once so that the single
      // issuing thread spends its time on tcgen05.mma, not on address arithmetic
      const uint64_t qdesc = make_smem_desc_sw128(sQ, 16, 1024);
      const uint64_t kdesc0 = make_smem_desc_sw128(sK(0), 16, 1024);
      const uint64_t kdesc1 = make_smem_desc_sw128(sK(1), 16, 1024);
      const uint64_t vdesc0 = make_smem_desc_sw128(sV(0), 1024, 1024);
      const uint64_t vdesc1 = make_smem_desc_sw128(sV(1), 1024, 1024);
      const uint64_t pdesc0 = make_smem_desc_sw128(sP, 16, 1024);
      const uint64_t pdesc1 = make_smem_desc_sw128(sP + ATT_TILE_BYTES, 16, 1024);
      auto issue_S = [&](int j) {
        const uint64_t kdesc = (j & 1) ? kdesc1 : kdesc0;
#pragma unroll
        for (int k = 0; k < 4; ++k)
          umma_bf16_ss(tmem_S, qdesc + uint64_t(2 * k), kdesc + uint64_t(2 * k), idesc_s, k != 0);
        umma_commit(s_full);
      };
      if (n_loc > 0) {
        mbar_wait(q_full, 0, 12);
        mbar_wait(kv_full(0), 0, 13);
        tc_fence_after();
        issue_S(0);
      }
      for (int j = 0; j < n_loc; ++j) {
        const int s = j & 1;
        if (j + 1 < n_loc) {
          mbar_wait(s_free, j & 1u, 14);  // S(j) fully read: the S columns may be overwritten
          mbar_wait(kv_full((j + 1) & 1), ((j + 1) >> 1) & 1u, 15);
          tc_fence_after();
          issue_S(j + 1);
        }
        mbar_wait(p_full, j & 1u, 16);  // P(j) in smem (and O rescaled if the running max moved)
        tc_fence_after();
        if (j == 2) ATT_STAMP(12);
        const uint64_t vdesc = s ? vdesc1 : vdesc0;
        const uint32_t acc0 = j != 0 ? 1u : 0u;
#pragma unroll
        for (int kk = 0; kk < 8; ++kk) {
          // P: 16 keys = 32 bytes inside the 128 B swizzle span (>>4 = 2); V: 16 key rows = 2048 bytes (>>4 = 128)
          const uint64_t pdesc = (kk < 4 ? pdesc0 : pdesc1) + uint64_t(2 * (kk & 3));
          umma_bf16_ss(tmem_O, pdesc, vdesc + uint64_t(128 * kk), idesc_o, kk != 0 ? 1u : acc0);
          // the accumulating MMAs form a latency-bound dependent chain (~130 cycles each): hand the first P slab
          // back to the softmax threads as soon as its four k-steps have retired
          if (kk == 3) umma_commit(p0_free);
        }
        umma_commit(o_full);
        umma_commit(kv_empty(s));
        if (j == 2) ATT_STAMP(13);
      }
    }
  } else {
    // ===================== softmax threads =====================
    const int q = warp & 3;
    const int r = q * 32 + lane;  // query row inside the tile == TMEM lane
    const uint32_t lane_off = uint32_t(q * 32) << 16;
    const float c = args.scale_log2;
    float mc = -INFINITY;  // running max (already multiplied by c), possibly stale by < 2^8
    float l_run = 0.f;     // softmax denominator in the same (stale-max, 2^7-biased) scale as O
    const uint32_t prow = sP + r * 128;
    const uint32_t sw = uint32_t(r & 7);
    for (int j = 0; j < n_loc; ++j) {
      const int n_valid = min(ATT_TILE, len - (j_begin + j) * ATT_TILE);
      const bool full_tile = n_valid == ATT_TILE;  // CTA-uniform
      mbar_wait(s_full, j & 1u, 17);
      tc_fence_after();
      const bool tr = threadIdx.x == 64 && j == (nparts > 1 ? 1 : 2);
      if (tr) ATT_STAMP(1);
      // ---- pass 1: row maximum over the valid keys (the last 32-key chunk stays in registers for pass 2) ----
      float mx = -INFINITY;
      uint32_t vlast[32];
#pragma unroll
      for (int c0 = 0; c0 < ATT_TILE; c0 += 32) {
        uint32_t v[32];
        if (c0 < ATT_TILE - 32) {
          tmem_ld_32x32(tmem_S + lane_off + c0, v);
        } else {
          tmem_ld_32x32(tmem_S + lane_off + c0, vlast);
        }
        tmem_wait_ld();
        const uint32_t(&u)[32] = (c0 < ATT_TILE - 32) ? v : vlast;
        if (full_tile || c0 + 32 <= n_valid) {
          mx = fmaxf(mx, max32(u));
        } else {
#pragma unroll
          for (int i = 0; i < 32; ++i)
            if (c0 + i < n_valid) mx = fmaxf(mx, __uint_as_float(u[i]));
        }
      }
      // ---- lazy rescale: only when this tile's max exceeds the running one by more than 2^8 ----
      const float mxc = mx * c;
      const bool need = mxc > mc + ATT_RESCALE_LOG2;
      if (tr) ATT_STAMP(2);
      bool o_done = (j == 0);
      if (j > 0 && __any_sync(0xffffffffu, need)) {
        // rare: P(j-1) V(j-1) must be folded into O before O and l are rescaled
        mbar_wait(o_full, (j - 1) & 1u, 18);
        tc_fence_after();
        o_done = true;
        const float f = need ? ex2_approx(mc - mxc) : 1.0f;
#pragma unroll
        for (int c0 = 0; c0 < ATT_D; c0 += 32) {
          uint32_t v[32];
          tmem_ld_32x32(tmem_O + lane_off + c0, v);
          tmem_wait_ld();
#pragma unroll
          for (int i = 0; i < 32; ++i) v[i] = __float_as_uint(__uint_as_float(v[i]) * f);
          tmem_st_32x32(tmem_O + lane_off + c0, v);
        }
        l_run *= f;
        tmem_wait_st();
      }
      if (j > 0 && !o_done) mbar_wait(p0_free, (j - 1) & 1u, 20);  // first P slab may be overwritten
      if (need) mc = mxc;
      if (tr) ATT_STAMP(3);
      // ---- pass 2: P = 2^7 * exp2(S*c - m) -> f16 -> smem (SW128 K-major, two 64-key slabs) ----
      const float mcb = mc - ATT_P_EXP_BIAS;
#pragma unroll
      for (int c0 = 0; c0 < ATT_TILE - 32; c0 += 32) {
        uint32_t v[32];
        tmem_ld_32x32(tmem_S + lane_off + c0, v);
        tmem_wait_ld();
        if (c0 == ATT_TILE - 64) {
          // last TMEM read of S(j) (the final chunk is still in registers from pass 1): let the MMA thread start
          // Q K^T of the next tile under the remaining half of this pass
          tc_fence_before();
          mbar_arrive(s_free);
        }
        if (c0 == 64 && !o_done) {
          mbar_wait(o_full, (j - 1) & 1u, 18);  // second P slab: all of P(j-1) V(j-1) has retired
          o_done = true;
        }
        l_run += full_tile ? softmax_chunk<false>(v, c, mcb, c0, n_valid, prow, sw)
                           : softmax_chunk<true>(v, c, mcb, c0, n_valid, prow, sw);
      }
      l_run += full_tile ? softmax_chunk<false>(vlast, c, mcb, ATT_TILE - 32, n_valid, prow, sw)
                         : softmax_chunk<true>(vlast, c, mcb, ATT_TILE - 32, n_valid, prow, sw);
      if (tr) ATT_STAMP(4);
      fence_proxy_async_smem();
      tc_fence_before();
      mbar_arrive(p_full);
      if (tr) ATT_STAMP(5);
    }
    // ---- epilogue ----
    if (n_loc > 0) {
      mbar_wait(o_full, (n_loc - 1) & 1u, 19);
      tc_fence_after();
    }
    const int t = q0 + r;
    __nv_bfloat16* orow = args.out + ((long long)b * args.rows_per_batch + t) * args.ldo + h * ATT_D;
    if (nparts == 1) {
      // whole item: O / l
      const float inv_l = 1.0f / l_run;
#pragma unroll
      for (int c0 = 0; c0 < ATT_D; c0 += 32) {
        uint32_t v[32];
        tmem_ld_32x32(tmem_O + lane_off + c0, v);
        tmem_wait_ld();
        if (t < args.rows_per_batch) {
          uint32_t pk[16];
#pragma unroll
          for (int i = 0; i < 32; i += 2)
            pk[i / 2] = pack_bf16x2(__uint_as_float(v[i]) * inv_l, __uint_as_float(v[i + 1]) * inv_l);
          uint4* o4 = reinterpret_cast<uint4*>(orow + c0);
#pragma unroll
          for (int g = 0; g < 4; ++g) o4[g] = make_uint4(pk[4 * g], pk[4 * g + 1], pk[4 * g + 2], pk[4 * g + 3]);
        }
      }
    } else {
      // key-split tail item: publish this part's (O, max, sum); the part that arrives last merges all of them.
      // Nobody waits for anybody (no co-scheduling assumption): ordering is "write, fence, count".
      if (threadIdx.x == 64) ATT_STAMP(6);
      const int tail = item - args.n_full;
      const long long unit = (long long)tail * nparts + part;
      float* wo = args.ws_o + (unit * ATT_TILE + r) * ATT_D;
      if (n_loc > 0) {
#pragma unroll
        for (int c0 = 0; c0 < ATT_D; c0 += 32) {
          uint32_t v[32];
          tmem_ld_32x32(tmem_O + lane_off + c0, v);
          tmem_wait_ld();
#pragma unroll
          for (int i = 0; i < 32; i += 4)
            *reinterpret_cast<uint4*>(wo + c0 + i) = make_uint4(v[i], v[i + 1], v[i + 2], v[i + 3]);
        }
      }
      *reinterpret_cast<float2*>(args.ws_ml + (unit * ATT_TILE + r) * 2) = make_float2(mc, l_run);  // empty part: (-inf, 0)
      if (threadIdx.x == 64) ATT_STAMP(7);
      __threadfence();
      asm volatile("bar.sync 1, 128;" ::: "memory");
      if (threadIdx.x == 64) ATT_STAMP(10);
      const uint32_t flag = sP;  // the P tile is dead by now: reuse its first word as the "I am last" flag
      if (threadIdx.x == 64) {
        const int old = atomicAdd(args.ws_cnt + tail, 1);
        const int last = (old == nparts - 1) ? 1 : 0;
        if (last) args.ws_cnt[tail] = 0;  // ready for the next launch (CUDA-graph replays included)
        asm volatile("st.shared.u32 [%0], %1;" ::"r"(flag), "r"(last) : "memory");
      }
      asm volatile("bar.sync 1, 128;" ::: "memory");
      uint32_t last;
      asm volatile("ld.shared.u32 %0, [%1];" : "=r"(last) : "r"(flag) : "memory");
      if (threadIdx.x == 64) ATT_STAMP(11);
      if (last) {
        __threadfence();
        const long long u0 = (long long)tail * nparts;
        // (1) per row: common maximum, per-part factors 2^(m_p - m) / l  -> smem (the P tile is dead)
        float mp[8], lp[8];
        float m_all = -INFINITY;
#pragma unroll
        for (int p = 0; p < 8; ++p) {
          mp[p] = -INFINITY; lp[p] = 0.f;
          if (p < nparts) {
            const float2 ml = __ldcg(reinterpret_cast<const float2*>(args.ws_ml + ((u0 + p) * ATT_TILE + r) * 2));
            mp[p] = ml.x; lp[p] = ml.y;
          }
          m_all = fmaxf(m_all, mp[p]);
        }
        float l_all = 0.f;
#pragma unroll
        for (int p = 0; p < 8; ++p) {
          mp[p] = lp[p] > 0.f ? ex2_approx(mp[p] - m_all) : 0.f;  // now the factor; empty parts contribute nothing
          l_all = fmaf(lp[p], mp[p], l_all);
        }
        const float inv_l = 1.0f / l_all;
        const uint32_t sF = sP + 16;  // [8 parts][128 rows] f32
#pragma unroll
        for (int p = 0; p < 8; ++p)
          asm volatile("st.shared.f32 [%0], %1;" ::"r"(sF + uint32_t(p * ATT_TILE + r) * 4u), "f"(mp[p] * inv_l) : "memory");
        asm volatile("bar.sync 1, 128;" ::: "memory");
        // (2) cooperative, coalesced merge: 16 threads per output row (4 columns each), 8 rows per step
        const int tid = threadIdx.x - 64;
        const int c4 = (tid & 15) * 4;
        const float* fs = reinterpret_cast<const float*>(__cvta_shared_to_generic(sF));
#pragma unroll 2
        for (int rg = 0; rg < ATT_TILE / 8; ++rg) {
          const int row = rg * 8 + (tid >> 4);
          // all loads of a row group are issued before any of them is consumed (one L2 round trip per group, not
          // one per part). Empty parts are read too: the scratch only ever holds finite values and their factor is 0.
          float f[8];
          float4 o[8];
#pragma unroll
          for (int p = 0; p < 8; ++p) {
            f[p] = 0.f;
            o[p] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (p < nparts) {
              f[p] = fs[p * ATT_TILE + row];
              o[p] = __ldcg(reinterpret_cast<const float4*>(args.ws_o + ((u0 + p) * ATT_TILE + row) * ATT_D + c4));
            }
          }
          float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
          for (int p = 0; p < 8; ++p) {
            acc.x = fmaf(o[p].x, f[p], acc.x); acc.y = fmaf(o[p].y, f[p], acc.y);
            acc.z = fmaf(o[p].z, f[p], acc.z); acc.w = fmaf(o[p].w, f[p], acc.w);
          }
          const int tt = q0 + row;
          if (tt < args.rows_per_batch) {
            __nv_bfloat16* orow2 = args.out + ((long long)b * args.rows_per_batch + tt) * args.ldo + h * ATT_D + c4;
            *reinterpret_cast<uint2*>(orow2) = make_uint2(pack_bf16x2(acc.x, acc.y), pack_bf16x2(acc.z, acc.w));
          }
        }
      }
    }
    tc_fence_before();
    if (threadIdx.x == 64) ATT_STAMP(14);
  }

  __syncthreads();
  if (threadIdx.x == 0) {
    ATT_STAMP(15);
    if (args.dbg) { unsigned long long gt; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(gt)); args.dbg[(long long)blockIdx.x * 16 + 9] = (long long)gt; }
  }
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, ATT_TMEM_COLS);
  }
}

}  // namespace oron
