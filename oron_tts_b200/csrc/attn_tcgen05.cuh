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

// One contiguous run of key tiles [j0, j1) of one (batch, head, 128-query tile) item.
struct AttnSeg {
  int b, h, qt;
  int j0, j1;
  int slot;    // -1: the whole item (O / l goes to `out`); else the partial-result slot this segment writes
  int owner;   // partial segment: the CTA that merges the item (the one holding its first key tiles)
  int pad;
};
// The split item a CTA merges at the end of its run: parts live in slots 2*(cta + p) + (p == 0 ? 1 : 0), p = 0..nparts-1.
struct AttnMergeEnt {
  int b, h, qt, nparts;   // nparts == 0: nothing to merge
};
// Workspace of the balanced schedule:
//   AttnPlanHeader | int nseg[grid] | AttnSeg segs[grid][seg_stride] | AttnMergeEnt merge[grid] | int cnt[grid] | ws_ml | ws_o
// Everything up to `merge` is written by attn_plan_kernel; cnt are the arrival counters of the split items (zeroed
// by the plan, left zero by every attention call).
struct AttnPlanHeader {
  unsigned magic;       // ATT_PLAN_MAGIC once attn_plan_kernel has run
  int nbatch, rows, heads;
  int grid;             // CTAs of the balanced launch
  int seg_stride;       // AttnSeg entries reserved per CTA
  int pad[10];
};
constexpr unsigned ATT_PLAN_MAGIC = 0x0A77B201u;

struct AttnArgs {
  int rows_per_batch;   // Tpad: rows per batch element in qkv / out
  int nbatch;
  int heads;
  const int* seq_lens;  // [nbatch] valid keys (= valid queries) per batch element, or nullptr
  __nv_bfloat16* out;   // [nbatch*rows_per_batch, ldo], head h at columns [h*64, h*64+64)
  long long ldo;
  float scale_log2;     // softmax scale * log2(e)
  long long* dbg;       // optional [grid, 16] clock64 stamps (tools/attn_trace.py); nullptr in production
  int q_tiles;          // 128-row query tiles per batch element (ceil(rows_per_batch / 128))
  // balanced schedule (plan != nullptr): the grid is the number of resident CTA slots and CTA c owns an equal share of
  // the flat (item, key tile) list; items that straddle two shares are split along the keys, their normalised partial
  // results go to the workspace and the CTA holding an item's first key tiles combines them at the end of its run. plan == nullptr: one CTA per item.
  const AttnPlanHeader* plan_hdr;
  const int* plan_nseg;
  const AttnSeg* plan_segs;
  const AttnMergeEnt* plan_merge;
  int* ws_cnt;          // [grid] arrival counters of the split items
  __half* ws_o;         // [2 * grid][128][64] f16: O_p / l_p of a partial segment
  float* ws_ml;         // [2 * grid][128][2] f32: (running max * c, l_p)
  float* lse;           // optional [nbatch * heads * rows_per_batch] f32: log2-domain log-sum-exp of every query row
                        // (whole items only: the training forward, which saves it for the backward kernels)
};

constexpr int ATT_THREADS = 192;  // warp0 TMA, warp1 MMA + TMEM alloc, warps 2..5 softmax
constexpr int ATT_TILE = 128;
constexpr int ATT_D = 64;
constexpr int ATT_TILE_BYTES = ATT_TILE * ATT_D * 2;  // 16 KB
// smem: Q | K0 K1 | V0 V1 | P(2 slabs) | barriers
constexpr int ATT_SMEM_BYTES = 7 * ATT_TILE_BYTES + 256;
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

// Cursor over a CTA's segments; every role (TMA, MMA, softmax) runs its own copy and sees the same tile sequence.
struct SegIt {
  const AttnSeg* segs;
  int nseg, s, j;
  AttnSeg cur;
  __device__ __forceinline__ void init(const AttnSeg* p, int n) { segs = p; nseg = n; s = 0; cur = p[0]; j = cur.j0; }
  __device__ __forceinline__ bool valid() const { return s < nseg; }
  __device__ __forceinline__ bool first() const { return j == cur.j0; }
  __device__ __forceinline__ bool last() const { return j == cur.j1 - 1; }
  __device__ __forceinline__ void next() {
    if (++j == cur.j1 && ++s < nseg) { cur = segs[s]; j = cur.j0; }
  }
};

__global__ void __launch_bounds__(ATT_THREADS, 2)
attn_fwd_tcgen05_kernel(const __grid_constant__ CUtensorMap tmQKV, const AttnArgs args) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t smem_base = smem_u32(smem_raw);
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  // one-CTA-per-item mode: the item as a single whole segment, kept after the barriers in dynamic shared memory
  AttnSeg& own_seg = *reinterpret_cast<AttnSeg*>(smem_raw + 7 * ATT_TILE_BYTES + 128);

  pdl_launch_dependents();
  const AttnSeg* segs;
  int nseg;
  if (args.plan_hdr != nullptr) {
    nseg = args.plan_nseg[blockIdx.x];
    segs = args.plan_segs + (long long)blockIdx.x * args.plan_hdr->seg_stride;
    const AttnPlanHeader& ph = *args.plan_hdr;
    if (ph.magic != ATT_PLAN_MAGIC || ph.nbatch != args.nbatch || ph.rows != args.rows_per_batch || ph.heads != args.heads ||
        ph.grid != int(gridDim.x)) {
      if (threadIdx.x == 0 && blockIdx.x == 0)
        printf("[oron] attention: the workspace holds no plan for this shape (call oron_attention_plan first)\n");
      __trap();
    }
  } else {
    const int qt = blockIdx.x % args.q_tiles;
    const int h = (blockIdx.x / args.q_tiles) % args.heads;
    const int b = blockIdx.x / (args.q_tiles * args.heads);
    const int len = args.seq_lens ? min(args.seq_lens[b], args.rows_per_batch) : args.rows_per_batch;
    const int nt = (len + ATT_TILE - 1) / ATT_TILE;
    nseg = qt < nt ? 1 : 0;  // a query tile entirely beyond the sequence: the out-projection masks these rows
    if (threadIdx.x == 0) { own_seg.b = b; own_seg.h = h; own_seg.qt = qt; own_seg.j0 = 0; own_seg.j1 = nt; own_seg.slot = -1; own_seg.owner = -1; own_seg.pad = 0; }
    segs = &own_seg;
  }
  if (nseg == 0) return;  // CTA-uniform, nothing allocated yet
  if ((smem_base & 1023u) != 0) {
    if (threadIdx.x == 0) printf("[oron] attention: dynamic smem not 1024-byte aligned\n");
    __trap();
  }

  const uint32_t sQ = smem_base;
  auto sK = [&](int st) { return smem_base + (1 + st) * ATT_TILE_BYTES; };
  auto sV = [&](int st) { return smem_base + (3 + st) * ATT_TILE_BYTES; };
  const uint32_t sP = smem_base + 5 * ATT_TILE_BYTES;
  const uint32_t bar_base = smem_base + 7 * ATT_TILE_BYTES;
  const uint32_t q_full = bar_base;             // per segment
  auto kv_full = [&](int st) { return bar_base + 8u * (1 + st); };
  auto kv_empty = [&](int st) { return bar_base + 8u * (3 + st); };  // MMA -> TMA: P(i) V(i) (and S(i+1)) have retired
  const uint32_t s_full = bar_base + 8u * 5;    // MMA -> softmax: S(i) is in TMEM
  const uint32_t s_free = bar_base + 8u * 6;    // softmax -> MMA: S(i) has been read (128 arrivals)
  const uint32_t p_full = bar_base + 8u * 7;    // softmax -> MMA: P(i) is in smem (128 arrivals)
  const uint32_t o_full = bar_base + 8u * 8;    // MMA -> softmax: O includes P(i) V(i)
  const uint32_t p0_free = bar_base + 8u * 9;   // MMA -> softmax: P V has consumed the first 64-key slab of P(i)
  const uint32_t tmem_slot = bar_base + 8u * 10;
  const uint32_t o_free = bar_base + 8u * 11;   // softmax -> MMA, per segment: O has been read out (128 arrivals)
  const uint32_t q_empty = bar_base + 8u * 12;  // MMA -> TMA, per segment: the last S of the segment has retired

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmQKV);
    mbar_init(q_full, 1);
    mbar_init(q_empty, 1);
    for (int st = 0; st < 2; ++st) { mbar_init(kv_full(st), 1); mbar_init(kv_empty(st), 1); }
    mbar_init(s_full, 1);
    mbar_init(s_free, 128);
    mbar_init(p_full, 128);
    mbar_init(o_full, 1);
    mbar_init(p0_free, 1);
    mbar_init(o_free, 128);
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
      SegIt w;
      w.init(segs, nseg);
      int seg = 0;
      for (int i = 0; w.valid(); w.next(), ++i) {
        const int st = i & 1;
        if (w.first()) {
          if (seg > 0) mbar_wait(q_empty, (seg - 1) & 1u, 11);  // every S of the previous segment has retired
          ++seg;
          mbar_arrive_expect_tx(q_full, ATT_TILE_BYTES);
          tma_load_3d(sQ, &tmQKV, q_full, w.cur.h * ATT_D, w.cur.qt * ATT_TILE, w.cur.b);
        }
        mbar_wait(kv_empty(st), ((i >> 1) & 1u) ^ 1u, 11);
        mbar_arrive_expect_tx(kv_full(st), 2 * ATT_TILE_BYTES);
        tma_load_3d(sK(st), &tmQKV, kv_full(st), HD + w.cur.h * ATT_D, w.j * ATT_TILE, w.cur.b);
        tma_load_3d(sV(st), &tmQKV, kv_full(st), 2 * HD + w.cur.h * ATT_D, w.j * ATT_TILE, w.cur.b);
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
      int seg_q = 0;  // segments whose Q has been consumed by an S issue
      auto issue_S = [&](bool first_of_seg, bool last_of_seg, int i) {
        if (first_of_seg) { mbar_wait(q_full, seg_q & 1u, 12); ++seg_q; }
        mbar_wait(kv_full(i & 1), (i >> 1) & 1u, 13);
        tc_fence_after();
        const uint64_t kdesc = (i & 1) ? kdesc1 : kdesc0;
#pragma unroll
        for (int k = 0; k < 4; ++k)
          umma_bf16_ss(tmem_S, qdesc + uint64_t(2 * k), kdesc + uint64_t(2 * k), idesc_s, k != 0);
        umma_commit(s_full);
        if (last_of_seg) umma_commit(q_empty);
      };
      SegIt w, wn;            // w: tile whose P V is issued in this iteration; wn: one tile ahead (S(i+1) goes first)
      w.init(segs, nseg);
      wn.init(segs, nseg);
      issue_S(true, wn.last(), 0);
      wn.next();
      int seg_o = 0;          // segments whose first P V has been issued
      for (int i = 0; w.valid(); w.next(), ++i) {
        if (wn.valid()) {
          mbar_wait(s_free, i & 1u, 14);  // S(i) fully read: the S columns may be overwritten
          issue_S(wn.first(), wn.last(), i + 1);
          wn.next();
        }
        mbar_wait(p_full, i & 1u, 16);  // P(i) in smem (and O rescaled if the running max moved)
        const bool first = w.first();
        if (first) {
          if (seg_o > 0) mbar_wait(o_free, (seg_o - 1) & 1u, 17);  // the previous segment's O has been read out
          ++seg_o;
        }
        tc_fence_after();
        if (i == 2) ATT_STAMP(12);
        const uint64_t vdesc = (i & 1) ? vdesc1 : vdesc0;
        const uint32_t acc0 = first ? 0u : 1u;
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
        umma_commit(kv_empty(i & 1));
        if (i == 2) ATT_STAMP(13);
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
    SegIt w;
    w.init(segs, nseg);
    int len = 0;
    for (int i = 0; w.valid(); w.next(), ++i) {
      const bool first = w.first();
      if (first) {
        mc = -INFINITY;
        l_run = 0.f;
        len = args.seq_lens ? min(args.seq_lens[w.cur.b], args.rows_per_batch) : args.rows_per_batch;
      }
      const int n_valid = min(ATT_TILE, len - w.j * ATT_TILE);
      const bool full_tile = n_valid == ATT_TILE;  // CTA-uniform
      mbar_wait(s_full, i & 1u, 17);
      tc_fence_after();
      const bool tr = threadIdx.x == 64 && i == 6;
      if (tr) ATT_STAMP(1);
      if (threadIdx.x == 64 && i == 7) ATT_STAMP(6);
      if (threadIdx.x == 64 && i == 8) ATT_STAMP(7);
      // ---- pass 1: row maximum over the valid keys (the last 32-key chunk stays in registers for pass 2).
      // TMEM loads are software-pipelined: chunk c+1 is in flight while chunk c is reduced (tcgen05.wait::ld waits
      // for everything issued so far, so each load is issued right after the wait for its predecessor).
      float mx = -INFINITY;
      uint32_t vlast[32];
      auto chunk_max = [&](const uint32_t (&u)[32], const int c0) {
        if (full_tile || c0 + 32 <= n_valid) {
          mx = fmaxf(mx, max32(u));
        } else {
#pragma unroll
          for (int k = 0; k < 32; ++k)
            if (c0 + k < n_valid) mx = fmaxf(mx, __uint_as_float(u[k]));
        }
      };
      {
        uint32_t va[32], vb[32];
        tmem_ld_32x32(tmem_S + lane_off + 0, va);
        tmem_wait_ld();
        tmem_ld_32x32(tmem_S + lane_off + 32, vb);
        chunk_max(va, 0);
        tmem_wait_ld();
        tmem_ld_32x32(tmem_S + lane_off + 64, va);
        chunk_max(vb, 32);
        tmem_wait_ld();
        tmem_ld_32x32(tmem_S + lane_off + 96, vlast);
        chunk_max(va, 64);
        tmem_wait_ld();
        chunk_max(vlast, 96);
      }
      // ---- lazy rescale: only when this tile's max exceeds the running one by more than 2^8 ----
      const float mxc = mx * c;
      const bool need = mxc > mc + ATT_RESCALE_LOG2;
      if (tr) ATT_STAMP(2);
      bool o_done = (i == 0);
      if (!first && __any_sync(0xffffffffu, need)) {
        // rare: P(i-1) V(i-1) must be folded into O before O and l are rescaled
        mbar_wait(o_full, (i - 1) & 1u, 18);
        tc_fence_after();
        o_done = true;
        const float f = need ? ex2_approx(mc - mxc) : 1.0f;
#pragma unroll
        for (int c0 = 0; c0 < ATT_D; c0 += 32) {
          uint32_t v[32];
          tmem_ld_32x32(tmem_O + lane_off + c0, v);
          tmem_wait_ld();
#pragma unroll
          for (int k = 0; k < 32; ++k) v[k] = __float_as_uint(__uint_as_float(v[k]) * f);
          tmem_st_32x32(tmem_O + lane_off + c0, v);
        }
        l_run *= f;
        tmem_wait_st();
      }
      if (!o_done) mbar_wait(p0_free, (i - 1) & 1u, 20);  // first P slab may be overwritten
      if (need) mc = mxc;
      if (tr) ATT_STAMP(3);
      // ---- pass 2: P = 2^7 * exp2(S*c - m) -> f16 -> smem (SW128 K-major, two 64-key slabs) ----
      const float mcb = mc - ATT_P_EXP_BIAS;
      auto chunk_exp = [&](const uint32_t (&u)[32], const int c0) {
        l_run += full_tile ? softmax_chunk<false>(u, c, mcb, c0, n_valid, prow, sw)
                           : softmax_chunk<true>(u, c, mcb, c0, n_valid, prow, sw);
      };
      {
        uint32_t va[32], vb[32];
        tmem_ld_32x32(tmem_S + lane_off + 0, va);
        tmem_wait_ld();
        tmem_ld_32x32(tmem_S + lane_off + 32, vb);
        chunk_exp(va, 0);
        tmem_wait_ld();
        tmem_ld_32x32(tmem_S + lane_off + 64, va);
        chunk_exp(vb, 32);
        tmem_wait_ld();
        // last TMEM read of S(i) (the final chunk is still in registers from pass 1): let the MMA thread start
        // Q K^T of the next tile under the remaining half of this pass
        tc_fence_before();
        mbar_arrive(s_free);
        if (!o_done) {
          mbar_wait(o_full, (i - 1) & 1u, 18);  // second P slab: all of P(i-1) V(i-1) has retired
          o_done = true;
        }
        chunk_exp(va, 64);
        chunk_exp(vlast, 96);
      }
      if (tr) ATT_STAMP(4);
      fence_proxy_async_smem();
      tc_fence_before();
      mbar_arrive(p_full);
      if (tr) ATT_STAMP(5);
      if (!w.last()) continue;

      // ---- end of a segment: O / l to the output (whole item) or to the workspace (partial item, f16 + (m, l)) ----
      mbar_wait(o_full, i & 1u, 19);
      tc_fence_after();
      const float inv_l = 1.0f / l_run;
      const int slot = w.cur.slot;
      const int t = w.cur.qt * ATT_TILE + r;
      const bool store = slot >= 0 || t < args.rows_per_batch;
      if (args.lse != nullptr && slot < 0 && t < args.rows_per_batch)  // sum_k 2^(s_k c) = l_run 2^(mc - 7)
        args.lse[((long long)w.cur.b * args.heads + w.cur.h) * args.rows_per_batch + t] = log2f(l_run) + mc - ATT_P_EXP_BIAS;
      // destination row: 64 values, 16 bits each, for both kinds
      uint4* dst = slot >= 0
          ? reinterpret_cast<uint4*>(args.ws_o + ((long long)slot * ATT_TILE + r) * ATT_D)
          : reinterpret_cast<uint4*>(args.out + ((long long)w.cur.b * args.rows_per_batch + t) * args.ldo + w.cur.h * ATT_D);
#pragma unroll 1
      for (int c0 = 0; c0 < ATT_D; c0 += 32) {
        uint32_t v[32];
        tmem_ld_32x32(tmem_O + lane_off + c0, v);
        tmem_wait_ld();
        if (c0 == ATT_D - 32) { tc_fence_before(); mbar_arrive(o_free); }  // O is in registers: the next segment may start
        uint32_t pk[16];
        if (slot >= 0) {
#pragma unroll
          for (int k = 0; k < 32; k += 2) {
            const __half2 hh = __floats2half2_rn(__uint_as_float(v[k]) * inv_l, __uint_as_float(v[k + 1]) * inv_l);
            pk[k / 2] = *reinterpret_cast<const uint32_t*>(&hh);
          }
        } else {
#pragma unroll
          for (int k = 0; k < 32; k += 2)
            pk[k / 2] = pack_bf16x2(__uint_as_float(v[k]) * inv_l, __uint_as_float(v[k + 1]) * inv_l);
        }
        if (store) {
#pragma unroll
          for (int g = 0; g < 4; ++g) dst[c0 / 8 + g] = make_uint4(pk[4 * g], pk[4 * g + 1], pk[4 * g + 2], pk[4 * g + 3]);
        }
      }
      if (slot >= 0) {
        *reinterpret_cast<float2*>(args.ws_ml + ((long long)slot * ATT_TILE + r) * 2) = make_float2(mc, l_run);
        if (w.cur.owner != int(blockIdx.x)) {
          // a part merged by another CTA: publish it ("write, fence, count"; nobody waits here)
          __threadfence();
          asm volatile("bar.sync 1, 128;" ::: "memory");
          if (threadIdx.x == 64) atomicAdd(args.ws_cnt + w.cur.owner, 1);
        }
      }
    }
    // ---- the split item whose first key tiles this CTA computed: combine its parts --------------------------------
    // Partial items come FIRST in every CTA's run, so by now (>= one whole item later) the other parts have long been
    // published by CTAs of this same launch; the wait below is a formality with a watchdog, not a scheduling assumption
    // about other launches.
    if (args.plan_hdr != nullptr) {
      const AttnMergeEnt me = args.plan_merge[blockIdx.x];
      if (me.nparts > 0) {
        const int tid = threadIdx.x - 64;
        if (tid == 0) {
          const long long t_start = clock64();
          volatile int* cnt = args.ws_cnt + blockIdx.x;
          while (*cnt < me.nparts - 1) {
            __nanosleep(100);
            if (clock64() - t_start > 4000000000ll) {
              printf("[oron] attention: merge wait timed out (cta %d)\n", int(blockIdx.x));
              __trap();
            }
          }
          *cnt = 0;  // ready for the next call (CUDA-graph replays included)
        }
        asm volatile("bar.sync 1, 128;" ::: "memory");
        __threadfence();
        auto slot_of = [&](int p) { return (long long)(2 * (int(blockIdx.x) + p) + (p == 0 ? 1 : 0)); };
        // per-row weights w_p = l_p 2^(m_p - m) / sum_q l_q 2^(m_q - m)
        float m_all = -INFINITY;
        for (int p = 0; p < me.nparts; ++p)
          m_all = fmaxf(m_all, __ldcg(args.ws_ml + (slot_of(p) * ATT_TILE + r) * 2));
        float l_all = 0.f;
        for (int p = 0; p < me.nparts; ++p) {
          const float2 ml = __ldcg(reinterpret_cast<const float2*>(args.ws_ml + (slot_of(p) * ATT_TILE + r) * 2));
          l_all = fmaf(ml.y, ex2_approx(ml.x - m_all), l_all);
        }
        const float inv_l = 1.0f / l_all;
        // cooperative, coalesced combine: 8 threads per row (8 columns = 16 bytes of f16 each), 16 rows per step; the
        // weights of up to 8 parts at a time travel through the (idle) P tile
        float* fs = reinterpret_cast<float*>(smem_raw + 5 * ATT_TILE_BYTES);  // [8][128]
        const int c8 = (tid & 7) * 8, rsub = tid >> 3;
        float acc[8][8];
#pragma unroll
        for (int st = 0; st < 8; ++st)
#pragma unroll
          for (int k = 0; k < 8; ++k) acc[st][k] = 0.f;
        for (int p0 = 0; p0 < me.nparts; p0 += 8) {
          const int pn = min(8, me.nparts - p0);
          asm volatile("bar.sync 1, 128;" ::: "memory");  // the previous chunk's weights are consumed
          for (int pp = 0; pp < pn; ++pp) {
            const float2 ml = __ldcg(reinterpret_cast<const float2*>(args.ws_ml + (slot_of(p0 + pp) * ATT_TILE + r) * 2));
            fs[pp * ATT_TILE + r] = ml.y * ex2_approx(ml.x - m_all) * inv_l;
          }
          asm volatile("bar.sync 1, 128;" ::: "memory");
          for (int pp = 0; pp < pn; ++pp) {
            const __half* po = args.ws_o + slot_of(p0 + pp) * (ATT_TILE * ATT_D) + c8;
            uint4 v[8];
#pragma unroll
            for (int st = 0; st < 8; ++st) v[st] = __ldcg(reinterpret_cast<const uint4*>(po + (st * 16 + rsub) * ATT_D));
#pragma unroll
            for (int st = 0; st < 8; ++st) {
              const float f = fs[pp * ATT_TILE + st * 16 + rsub];
              const uint32_t wds[4] = {v[st].x, v[st].y, v[st].z, v[st].w};
#pragma unroll
              for (int k = 0; k < 4; ++k) {
                const float2 fv = __half22float2(*reinterpret_cast<const __half2*>(&wds[k]));
                acc[st][2 * k] = fmaf(fv.x, f, acc[st][2 * k]);
                acc[st][2 * k + 1] = fmaf(fv.y, f, acc[st][2 * k + 1]);
              }
            }
          }
        }
#pragma unroll
        for (int st = 0; st < 8; ++st) {
          const int tt = me.qt * ATT_TILE + st * 16 + rsub;
          if (tt < args.rows_per_batch)
            *reinterpret_cast<uint4*>(args.out + ((long long)me.b * args.rows_per_batch + tt) * args.ldo + me.h * ATT_D + c8) =
                make_uint4(pack_bf16x2(acc[st][0], acc[st][1]), pack_bf16x2(acc[st][2], acc[st][3]),
                           pack_bf16x2(acc[st][4], acc[st][5]), pack_bf16x2(acc[st][6], acc[st][7]));
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

// ------------------------------------------------------------------------------------------------------------
// Balanced schedule, part 1: the plan. One thread per CTA of the balanced launch cuts its equal share [u0, u1) of
// the flat (item, key tile) list into segments: the partial item at either end FIRST, whole items after them (so
// every partial result exists early), and notes for each CTA the split item it has to combine at the end of its run. Runs once per set of sequence lengths (oron_attention_plan), not per call.
// ------------------------------------------------------------------------------------------------------------
struct PlanWalk {
  int heads, nbatch, rows;
  const int* lens;
  int b, item, tile, nt;
  long long item_u0;
  __device__ int len_of(int bb) const { return lens ? min(lens[bb], rows) : rows; }
  __device__ static int tiles_of(int l) { return (l + ATT_TILE - 1) / ATT_TILE; }
  __device__ long long total() const {
    long long t = 0;
    for (int bb = 0; bb < nbatch; ++bb) { const long long n = tiles_of(len_of(bb)); t += n * n * heads; }
    return t;
  }
  __device__ void seek(long long u) {
    long long base = 0;
    for (b = 0; b < nbatch; ++b) {
      nt = tiles_of(len_of(b));
      const long long n = (long long)nt * nt * heads;
      if (u < base + n) break;
      base += n;
    }
    const long long r = u - base;
    item = int(r / nt);
    tile = int(r - (long long)item * nt);
    item_u0 = u - tile;
  }
  __device__ AttnSeg seg(int j0, int j1, int slot) const {
    AttnSeg s;
    s.b = b; s.h = item / nt; s.qt = item % nt; s.j0 = j0; s.j1 = j1; s.slot = slot; s.owner = -1; s.pad = 0;
    return s;
  }
};

__global__ void attn_plan_kernel(AttnPlanHeader* hdr, int* nseg_out, AttnSeg* segs_out, AttnMergeEnt* merge, int* cnt,
                                 const int* seq_lens, int nbatch, int rows, int heads, int grid, int seg_stride) {
  if (threadIdx.x == 0) {
    hdr->nbatch = nbatch; hdr->rows = rows; hdr->heads = heads; hdr->grid = grid; hdr->seg_stride = seg_stride;
  }
  PlanWalk w;
  w.heads = heads; w.nbatch = nbatch; w.rows = rows; w.lens = seq_lens;
  const long long total = w.total();
  const long long G = min((long long)grid, total);
  auto start_of = [&](long long cta) { return (total * cta) / G; };
  auto cta_of = [&](long long u) {
    long long cc = (u * G) / total;
    while (cc + 1 < G && start_of(cc + 1) <= u) ++cc;
    while (cc > 0 && start_of(cc) > u) --cc;
    return cc;
  };
  for (int c = threadIdx.x; c < grid; c += blockDim.x) {
    int n = 0;
    AttnSeg* out = segs_out + (long long)c * seg_stride;
    AttnMergeEnt me;
    me.b = me.h = me.qt = me.nparts = 0;
    if (c < G) {
      const long long u0 = start_of(c), u1 = start_of(c + 1);
      // partial tail of the item the share starts in: merged by the CTA that holds the item's first tile
      w.seek(u0);
      long long a_end = u0;
      if (w.tile != 0) {
        a_end = min(u1, w.item_u0 + w.nt);
        AttnSeg sg = w.seg(w.tile, w.tile + int(a_end - u0), 2 * c);
        sg.owner = int(cta_of(w.item_u0));
        out[n++] = sg;
      }
      // partial head of the item the share ends in: this CTA merges it
      long long b_start = u1;
      if (a_end < u1) {
        w.seek(u1 - 1);
        if (w.item_u0 + w.nt > u1) {
          b_start = w.item_u0;
          AttnSeg sg = w.seg(0, int(u1 - b_start), 2 * c + 1);
          sg.owner = c;
          out[n++] = sg;
          me.b = w.b; me.h = w.item / w.nt; me.qt = w.item % w.nt;
          me.nparts = int(cta_of(w.item_u0 + w.nt - 1) - c + 1);
        }
      }
      // whole items in between
      for (long long u = a_end; u < b_start;) {
        w.seek(u);
        if (n < seg_stride) out[n++] = w.seg(0, w.nt, -1);
        u += w.nt;
      }
    }
    nseg_out[c] = n;
    merge[c] = me;
    cnt[c] = 0;
  }
  __syncthreads();
  if (threadIdx.x == 0) { __threadfence(); hdr->magic = ATT_PLAN_MAGIC; }
}

}  // namespace oron
