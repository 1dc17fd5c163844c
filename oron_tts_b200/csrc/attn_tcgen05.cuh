// Non-causal multi-head attention with per-sequence key lengths (replaces the reference's
// F.scaled_dot_product_attention + key-padding mask, src/models/modules.py:271-278).
// One CTA = one (batch element, head, 128-query tile); two CTAs per SM. head_dim = 64.
//
//   S = Q K^T   : tcgen05.mma M=128 N=128 K=64 -> TMEM cols [0,128)
//   P = softmax : 256 softmax threads = two per query row (TMEM lane == row => no shuffles); warps 2..5 ("half A")
//                 own the first 64 keys of every tile, warps 6..9 ("half B") the last 64. The halves are fully
//                 independent online softmaxes — own running maximum, own denominator, own 64-key SW128 slab
//                 of P, own accumulator — and are merged once, after the KV loop (flash-decoding style split,
//                 but inside the CTA, so no extra memory traffic). No per-tile exchange between the halves: a
//                 per-tile row-max exchange through a named barrier was measured at 1-2.4k cycles per tile.
//   O_h += P_h V_h : tcgen05.mma (f16 x f16) M=128 N=64 K=64 accumulating IN TMEM (A: cols [128,192),
//                 B: [192,256)); V tile as MN-major B operand; V is written as f16 by the QKV GEMM epilogue.
//                 Two accumulators also halve the latency-bound dependent MMA chain.
//
// What the per-CTA clock64 traces (tools/attn_trace.py, profiles/r01_attn_trace_*.txt) showed, in order:
//   * O and its rescaling belong on the tensor core / in TMEM: the running maximum is only raised when a tile
//     exceeds it by more than 2^8 ("lazy rescale"); only then O is read back, scaled and stored (warp-uniform,
//     rare). Probabilities carry a 2^7 bias so they use the f16 range; it cancels in O / l.
//   * every tcgen05.mma of a dependent accumulate chain costs ~65-80 cycles to issue and ~110 to retire
//     whatever its N, so the chain p_full -> P V -> o_full is a per-tile critical loop: the first P slab is
//     handed back after four k-steps (p0_free), and a tensor-core row sum (P x ones) was dropped again.
//   * one softmax warp per scheduler is instruction-latency bound (IPC ~0.3); replacing exp2 by a polynomial
//     on the FMA pipe made it slower. Hence two warps per row quarter, a single pass over TMEM, and S released
//     to the MMA thread right after that read so Q K^T of the next tile runs under the whole exp phase.
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
};

constexpr int ATT_THREADS = 320;  // warp0 TMA, warp1 MMA + TMEM alloc, warps 2..9 softmax
constexpr int ATT_SOFTMAX_THREADS = 256;
constexpr int ATT_TILE = 128;
constexpr int ATT_D = 64;
constexpr int ATT_TILE_BYTES = ATT_TILE * ATT_D * 2;  // 16 KB
constexpr int ATT_XCH_BYTES = 512;                    // (max, sum) exchange between the two key halves, once per CTA
// smem: Q | K0 K1 | V0 V1 | P(2 slabs) | barriers | exchange
constexpr int ATT_SMEM_BYTES = 7 * ATT_TILE_BYTES + 128 + ATT_XCH_BYTES;
constexpr int ATT_TMEM_COLS = 256;
#define ATT_STAMP(slot) do { if (args.dbg) args.dbg[(long long)((blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x) * 16 + (slot)] = clock64(); } while (0)
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

// 2^x on the FMA/ALU pipes (kept for reference; measured slower than MUFU here, see header).
__device__ __forceinline__ float exp2_poly(float x) {
  x = fmaxf(x, -125.0f);
  const float t = x + 12582912.0f;
  const float f = x - (t - 12582912.0f);
  float p = fmaf(0.0551716573536396f, f, 0.2426111251115799f);
  p = fmaf(p, f, 0.6932609677314758f);
  p = fmaf(p, f, 0.9999280571937561f);
  return __int_as_float(__float_as_int(p) + (__float_as_int(t) << 23));
}

// exp2 + swizzled store of one 32-key chunk of a P row; returns the chunk's row sum (fp32).
// x = s*c - (m*c - 7); results are packed to the f16 P tile. `cslab`: key offset inside the 64-key slab (0 / 32),
// `kglob`: key offset inside the tile (for masking). MASKED: keys >= n_valid get probability 0.
template <bool MASKED>
__device__ __forceinline__ float softmax_chunk(const uint32_t (&v)[32], const float c, const float mcb, const int cslab,
                                               const int kglob, const int n_valid, const uint32_t prow_slab,
                                               const uint32_t sw) {
  uint32_t pk[16];
  float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
#pragma unroll
  for (int i = 0; i < 32; i += 4) {
    float p0 = ex2_approx(fmaf(__uint_as_float(v[i]), c, -mcb));
    float p1 = ex2_approx(fmaf(__uint_as_float(v[i + 1]), c, -mcb));
    float p2 = ex2_approx(fmaf(__uint_as_float(v[i + 2]), c, -mcb));
    float p3 = ex2_approx(fmaf(__uint_as_float(v[i + 3]), c, -mcb));
    if (MASKED) {
      if (kglob + i >= n_valid) p0 = 0.f;
      if (kglob + i + 1 >= n_valid) p1 = 0.f;
      if (kglob + i + 2 >= n_valid) p2 = 0.f;
      if (kglob + i + 3 >= n_valid) p3 = 0.f;
    }
    s0 += p0; s1 += p1; s2 += p2; s3 += p3;
    __half2 a = __floats2half2_rn(p0, p1), b = __floats2half2_rn(p2, p3);
    pk[i / 2] = *reinterpret_cast<uint32_t*>(&a);
    pk[i / 2 + 1] = *reinterpret_cast<uint32_t*>(&b);
  }
  const uint32_t chunk0 = uint32_t(cslab) >> 3;  // first 16-byte chunk of this 32-key group inside its slab row
#pragma unroll
  for (int g = 0; g < 4; ++g) {
    const uint32_t addr = prow_slab + (((chunk0 + g) ^ sw) << 4);
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
__device__ __forceinline__ float max32_masked(const uint32_t (&v)[32], int k0, int n_valid) {
  float m = -INFINITY;
#pragma unroll
  for (int i = 0; i < 32; ++i)
    if (k0 + i < n_valid) m = fmaxf(m, __uint_as_float(v[i]));
  return m;
}
__device__ __forceinline__ void softmax_bar_sync() {  // named barrier 1: the 256 softmax threads only
  asm volatile("bar.sync 1, 256;" ::: "memory");
}

__global__ void __launch_bounds__(ATT_THREADS, 2)
attn_fwd_tcgen05_kernel(const __grid_constant__ CUtensorMap tmQKV, const AttnArgs args) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t smem_base = smem_u32(smem_raw);
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  pdl_launch_dependents();
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
  const uint32_t s_full = bar_base + 8u * 5;   // MMA -> softmax: S(j) is in TMEM
  const uint32_t s_free = bar_base + 8u * 6;   // softmax -> MMA: S(j) has been read (256 arrivals)
  auto p_full = [&](int hf) { return bar_base + 8u * (7 + hf); };   // softmax half -> MMA: its P slab is in smem (128)
  auto o_full = [&](int hf) { return bar_base + 8u * (9 + hf); };   // MMA -> softmax half: O_h includes P_h(j) V_h(j)
  const uint32_t tmem_slot = bar_base + 8u * 11;
  const uint32_t sXch = bar_base + 128;        // half B -> half A: (running max, denominator) of 64 rows at a time

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmQKV);
    mbar_init(q_full, 1);
    for (int s = 0; s < 2; ++s) { mbar_init(kv_full(s), 1); mbar_init(kv_empty(s), 1); }
    mbar_init(s_full, 1);
    mbar_init(s_free, ATT_SOFTMAX_THREADS);
    for (int hf = 0; hf < 2; ++hf) { mbar_init(p_full(hf), ATT_SOFTMAX_THREADS / 2); mbar_init(o_full(hf), 1); }
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
  const uint32_t tmem_O = tmem_base + 128;  // half A accumulator; half B at +64
  pdl_wait();  // the QKV activations of the previous kernel are visible from here on
  if (threadIdx.x == 0) ATT_STAMP(0);

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
      mbar_wait(q_full, 0, 12);
      mbar_wait(kv_full(0), 0, 13);
      tc_fence_after();
      issue_S(0);
      for (int j = 0; j < n_kv; ++j) {
        const int s = j & 1;
        if (j + 1 < n_kv) {
          mbar_wait(s_free, j & 1u, 14);  // S(j) is in registers: the S columns may be overwritten
          mbar_wait(kv_full((j + 1) & 1), ((j + 1) >> 1) & 1u, 15);
          tc_fence_after();
          issue_S(j + 1);
        }
        const uint64_t vdesc = s ? vdesc1 : vdesc0;
        const uint32_t acc0 = j != 0 ? 1u : 0u;
#pragma unroll
        for (int hf = 0; hf < 2; ++hf) {
          mbar_wait(p_full(hf), j & 1u, 16);  // this half's P slab is in smem (and its O rescaled if its max moved)
          tc_fence_after();
          if (j == 2 && hf == 0) ATT_STAMP(12);
#pragma unroll
          for (int kq = 0; kq < 4; ++kq) {
            const int kk = 4 * hf + kq;
            // P: 16 keys = 32 bytes inside the 128 B swizzle span (>>4 = 2); V: 16 key rows = 2048 bytes (>>4 = 128)
            umma_bf16_ss(tmem_O + 64 * hf, (hf ? pdesc1 : pdesc0) + uint64_t(2 * kq), vdesc + uint64_t(128 * kk), idesc_o,
                         kq != 0 ? 1u : acc0);
          }
          umma_commit(o_full(hf));
        }
        umma_commit(kv_empty(s));
        if (j == 2) ATT_STAMP(13);
      }
    }
  } else {
    // ===================== softmax threads: two per query row =====================
    const int q = warp & 3;                 // TMEM lane quarter
    const int half = (warp - 2) >> 2;       // 0: keys [0,64) of every tile, 1: keys [64,128)
    const int r = q * 32 + lane;            // query row inside the tile == TMEM lane
    const uint32_t lane_off = uint32_t(q * 32) << 16;
    const float c = args.scale_log2;
    float mc = -INFINITY;  // this half's running max (already multiplied by c), possibly stale by < 2^8
    float l_run = 0.f;     // this half's denominator (same stale-max, 2^7-biased scale as its O accumulator)
    const uint32_t prow_slab = sP + half * ATT_TILE_BYTES + r * 128;
    const uint32_t sw = uint32_t(r & 7);
    const int k0 = half * 64;               // first key of this half inside a tile
    const uint32_t tmem_Oh = tmem_O + 64 * half;
    const uint32_t my_p_full = p_full(half), my_o_full = o_full(half);

    for (int j = 0; j < n_kv; ++j) {
      const int n_valid = min(ATT_TILE, len - j * ATT_TILE);
      const bool full_tile = n_valid == ATT_TILE;  // CTA-uniform
      const bool any_key = k0 < n_valid;           // half B of a short last tile may own no valid key at all
      mbar_wait(s_full, j & 1u, 17);
      tc_fence_after();
      const bool tr = threadIdx.x == 64 && (j == 2 || j == 3);
      if (tr) ATT_STAMP(1 + 6 * (j - 2));
      // ---- row maximum over this half's 64 scores (second chunk stays in registers) ----
      uint32_t vb[32];
      float mx;
      {
        uint32_t va[32];
        tmem_ld_32x32(tmem_S + lane_off + k0, va);
        tmem_wait_ld();
        mx = full_tile ? max32(va) : max32_masked(va, k0, n_valid);
      }
      tmem_ld_32x32(tmem_S + lane_off + k0 + 32, vb);
      tmem_wait_ld();
      mx = fmaxf(mx, full_tile ? max32(vb) : max32_masked(vb, k0 + 32, n_valid));
      if (tr) ATT_STAMP(2 + 6 * (j - 2));
      // ---- lazy rescale: only when this tile's max exceeds the running one by more than 2^8 ----
      const float mxc = mx * c;
      const bool need = any_key && (mxc > mc + ATT_RESCALE_LOG2);
      bool o_done = (j == 0);
      if (j > 0 && __any_sync(0xffffffffu, need)) {
        // rare: P_h(j-1) V_h(j-1) must be folded into O_h before it is rescaled
        mbar_wait(my_o_full, (j - 1) & 1u, 18);
        tc_fence_after();
        o_done = true;
        const float f = need ? ex2_approx(mc - mxc) : 1.0f;
#pragma unroll
        for (int c0 = 0; c0 < ATT_D; c0 += 32) {
          uint32_t v[32];
          tmem_ld_32x32(tmem_Oh + lane_off + c0, v);
          tmem_wait_ld();
#pragma unroll
          for (int i = 0; i < 32; ++i) v[i] = __float_as_uint(__uint_as_float(v[i]) * f);
          tmem_st_32x32(tmem_Oh + lane_off + c0, v);
        }
        tmem_wait_st();
        l_run *= f;
      }
      if (need) mc = mxc;
      // this half's P slab must have been consumed by P_h(j-1) V_h(j-1)
      if (!o_done) mbar_wait(my_o_full, (j - 1) & 1u, 18);
      if (tr) ATT_STAMP(3 + 6 * (j - 2));
      // ---- P = 2^7 * exp2(S*c - m) -> f16 -> this half's 64-key slab (SW128 K-major) ----
      // (an all-masked half keeps mc = -inf on its first tile: use 0 so that exp2 sees finite arguments; the
      //  MASKED path zeroes every probability anyway)
      const float mcb = (mc == -INFINITY ? 0.f : mc) - ATT_P_EXP_BIAS;
      if (full_tile) l_run += softmax_chunk<false>(vb, c, mcb, 32, k0 + 32, n_valid, prow_slab, sw);
      else l_run += softmax_chunk<true>(vb, c, mcb, 32, k0 + 32, n_valid, prow_slab, sw);
      {
        uint32_t va[32];
        tmem_ld_32x32(tmem_S + lane_off + k0, va);  // re-read the first chunk (keeps the live registers at 32 + 16)
        tmem_wait_ld();
        tc_fence_before();
        mbar_arrive(s_free);                        // last TMEM read of S(j): Q K^T of the next tile may start
        if (full_tile) l_run += softmax_chunk<false>(va, c, mcb, 0, k0, n_valid, prow_slab, sw);
        else l_run += softmax_chunk<true>(va, c, mcb, 0, k0, n_valid, prow_slab, sw);
      }
      if (tr) ATT_STAMP(4 + 6 * (j - 2));
      fence_proxy_async_smem();
      tc_fence_before();
      mbar_arrive(my_p_full);
      if (tr) ATT_STAMP(5 + 6 * (j - 2));
    }
    // ---- merge the two halves and write O / l: half B publishes (max, denominator), half A owns output columns
    //      [0,32), half B [32,64); both read BOTH accumulators for their columns ----
    mbar_wait(o_full(0), (n_kv - 1) & 1u, 19);
    mbar_wait(o_full(1), (n_kv - 1) & 1u, 19);
    tc_fence_after();
    float m_o = 0.f, l_o = 0.f;
    {
      // 128 rows x 2 halves x (m, l) f32 = 2 KB does not fit the 512 B buffer: four rounds of 32 rows
#pragma unroll 1
      for (int round = 0; round < 4; ++round) {
        const bool mine_now = q == round;
        const uint32_t slot = sXch + uint32_t(half * 32 + lane) * 8u;
        const uint32_t slot_o = sXch + uint32_t((half ^ 1) * 32 + lane) * 8u;
        softmax_bar_sync();
        if (mine_now) asm volatile("st.shared.v2.f32 [%0], {%1, %2};" ::"r"(slot), "f"(mc), "f"(l_run) : "memory");
        softmax_bar_sync();
        if (mine_now) asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(m_o), "=f"(l_o) : "r"(slot_o) : "memory");
      }
    }
    // common scale: m = max(m_A, m_B); a half without any valid key has m = -inf, l = 0 and contributes nothing
    const float m_all = fmaxf(mc, m_o);
    const float f_me = (mc == -INFINITY) ? 0.f : ex2_approx(mc - m_all);
    const float f_ot = (m_o == -INFINITY) ? 0.f : ex2_approx(m_o - m_all);
    const float inv_l = 1.0f / (l_run * f_me + l_o * f_ot);
    const float fA = (half == 0 ? f_me : f_ot) * inv_l;
    const float fB = (half == 0 ? f_ot : f_me) * inv_l;
    const int t = q0 + r;
    __nv_bfloat16* orow = args.out + ((long long)b * args.rows_per_batch + t) * args.ldo + h * ATT_D + half * 32;
#pragma unroll 1
    for (int cc = 0; cc < 32; cc += 16) {
      uint32_t va[16], vb2[16];
      tmem_ld_32x16(tmem_O + lane_off + half * 32 + cc, va);        // O_A, 16 of my 32 output columns
      tmem_ld_32x16(tmem_O + 64 + lane_off + half * 32 + cc, vb2);  // O_B, same columns
      tmem_wait_ld();
      if (t < args.rows_per_batch) {
        uint32_t pk[8];
#pragma unroll
        for (int i = 0; i < 16; i += 2) {
          // an untouched accumulator (half with no valid key) may hold stale TMEM contents: its factor is 0, so
          // guard against 0 * inf/nan by selecting instead of multiplying
          const float a0 = fA != 0.f ? __uint_as_float(va[i]) * fA : 0.f;
          const float a1 = fA != 0.f ? __uint_as_float(va[i + 1]) * fA : 0.f;
          const float b0 = fB != 0.f ? __uint_as_float(vb2[i]) * fB : 0.f;
          const float b1 = fB != 0.f ? __uint_as_float(vb2[i + 1]) * fB : 0.f;
          pk[i / 2] = pack_bf16x2(a0 + b0, a1 + b1);
        }
        uint4* o4 = reinterpret_cast<uint4*>(orow + cc);
        o4[0] = make_uint4(pk[0], pk[1], pk[2], pk[3]);
        o4[1] = make_uint4(pk[4], pk[5], pk[6], pk[7]);
      }
    }
    tc_fence_before();
    if (threadIdx.x == 64) ATT_STAMP(14);
  }

  __syncthreads();
  if (threadIdx.x == 0) ATT_STAMP(15);
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, ATT_TMEM_COLS);
  }
}

}  // namespace oron
