// Non-causal multi-head attention with per-sequence key lengths, head_dim 64 (replaces the reference's
// F.scaled_dot_product_attention + key-padding mask, src/models/modules.py:271-278). Fourth generation of the kernel;
// what changed against attn_tcgen05.cuh (round 1) and why is in DESIGN.md section 5.
//
// Work = the flat list of (item, key tile) units, item = (batch element, head, 128-query tile). A launch has at most
// 2 x #SM CTAs (two fit an SM); each takes an equal contiguous share of the list (attn4_plan_kernel, once per set of
// sequence lengths), i.e. a few "segments" = runs of key tiles of one item. An item whose tiles fall into several
// shares is combined by whichever of its CTAs finishes last (partial results in a workspace, an arrival counter per
// item; nobody waits for anybody).
//
// Per CTA: warp 0 = TMA producer (Q double buffered, K two stages, V three stages, separate barriers so that the K of
// tile i+1 never queues behind the V of tile i), warp 1 = tcgen05.mma issuer, warps 2..5 = softmax, one thread per
// query row (TMEM lane == row => no shuffles).
//   S = Q K^T : tcgen05.mma M128 N128 K16 x4, SS operands                      -> TMEM columns [0, 128)
//   softmax   : the whole fp32 row (128 values) is read ONCE into registers and the S columns are handed back to the
//               tensor core at once (S of tile i+1 is computed under the softmax of tile i); row max with 3-input
//               FMNMX, scale/shift with FFMA2, exp2 on the MUFU unit, row sum with FADD2, f16 pack
//   P         : f16, written with tcgen05.st to TMEM columns [128, 192) (2 keys per 32-bit column) -- no shared
//               memory, no swizzle arithmetic, no proxy fence
//   O += P V  : tcgen05.mma M128 N64 K16 x8 with A = P read from TMEM (TS form), B = V (f16, MN-major SW128 tile)
//               accumulating in TMEM columns [192, 256) over the whole segment
//   running max: raised only when a tile exceeds it by more than 2^8 ("lazy rescale": O is then read back, scaled and
//               stored by the softmax threads, a warp-uniform rare branch); probabilities carry a 2^7 bias so they
//               use the f16 range (cancels in O / l).
// q/k/v are read straight out of the fused QKV activation [rows, 3*H*64] with one 3-D TMA map (V third is IEEE f16).
#pragma once
#include <cuda_fp16.h>

#include "ptx.cuh"

namespace oron {

struct Attn4Seg {   // 32 bytes: key tiles [j0, j1) of item (b, h, qt)
  int b, h, qt;
  int j0, j1;
  int slot;    // -1: the whole item (result goes to `out`); else the partial-result slot this segment writes
  int owner;   // partial segment: the CTA holding the item's first key tile (indexes the arrival counter)
  int nparts;  // partial segment: number of consecutive CTAs (owner, owner + 1, ...) that share the item
};
struct Attn4Merge {   // one per CTA of the launch: the split item whose first key tile the CTA holds (nparts == 0: none)
  int b, h, qt, nparts;
};
struct Attn4PlanHeader {
  unsigned magic;       // ATT4_PLAN_MAGIC once attn4_plan_kernel has run
  int nbatch, rows, heads;
  int grid;             // CTAs of the launch
  int seg_stride;       // Attn4Seg entries reserved per CTA
  int pad[10];
};
constexpr unsigned ATT4_PLAN_MAGIC = 0x0A77B204u;

struct Attn4Args {
  int rows_per_batch;   // Tpad: rows per batch element in qkv / out
  int nbatch;
  int heads;
  const int* seq_lens;  // [nbatch] valid keys (= valid queries) per batch element, or nullptr
  __nv_bfloat16* out;   // [nbatch*rows_per_batch, ldo], head h at columns [h*64, h*64+64)
  long long ldo;
  float scale_log2;     // softmax scale * log2(e)
  int q_tiles;          // 128-row query tiles per batch element
  // planned launch (plan_hdr != nullptr): CTA c runs plan_segs[c * seg_stride .. + plan_nseg[c]); else one CTA per item
  const Attn4PlanHeader* plan_hdr;
  const int* plan_nseg;
  const Attn4Seg* plan_segs;
  const Attn4Merge* plan_merge;  // [grid]
  int* ws_cnt;          // [2 grid] arrival / departure counters of the split items (zero between launches), or nullptr: attn4_combine_kernel merges
  __half* ws_o;         // [2 * grid][128][64] f16: O_p / l_p of a partial segment
  float* ws_ml;         // [2 * grid][128][2] f32: (running max * c, l_p)
  float* lse;           // optional [nbatch * heads * rows_per_batch] f32: log2-domain log-sum-exp of every query row
  long long* dbg;       // optional [grid, 16] clock64 stamps (tools/attn_trace.py); nullptr in production
};

constexpr int ATT4_THREADS = 256;  // warp 0 TMA, warp 1 MMA, warps 2-3 combine the split items, warps 4-7 softmax
constexpr int ATT4_REGS_AUX = 72, ATT4_REGS_SOFTMAX = 184;  // setmaxnreg: 128 * (72 + 184) = the CTA's 32768 registers
constexpr int ATT4_TILE = 128;
constexpr int ATT4_D = 64;
constexpr int ATT4_TILE_BYTES = ATT4_TILE * ATT4_D * 2;  // 16 KB
constexpr int ATT4_QS = 2, ATT4_KS = 2, ATT4_VS = 3;
constexpr int ATT4_BAR_OFF = (ATT4_QS + ATT4_KS + ATT4_VS) * ATT4_TILE_BYTES;
constexpr int ATT4_SMEM_BYTES = ATT4_BAR_OFF + 256;
constexpr int ATT4_TMEM_COLS = 256;
#ifndef ATT4_EX2_POLY
#define ATT4_EX2_POLY 1  // 1: part of the exponentials on the FMA pipe (ex2_poly2)
#endif
constexpr float ATT4_RESCALE_LOG2 = 8.0f;  // raise the running max only when exceeded by > 2^8
constexpr float ATT4_P_EXP_BIAS = 7.0f;    // probabilities are scaled by 2^7 (<= 2^15 in f16); cancels in O / l
#define ATT4_STAMP(slot) do { if (DBG) args.dbg[(long long)blockIdx.x * 16 + (slot)] = clock64(); } while (0)
#define ATT4_PUT(slot, val) do { if (DBG) args.dbg[(long long)blockIdx.x * 16 + (slot)] = (val); } while (0)

__device__ __forceinline__ void tmem4_st16(uint32_t taddr, const uint32_t* r) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
      :
      : "r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]),
        "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}
__device__ __forceinline__ void tmem4_st32(uint32_t taddr, const uint32_t* r) {
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
__device__ __forceinline__ void tmem4_wait_st() {
  asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_ld_row32(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]),
        "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]),
        "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]),
        "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
        "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
// D[tmem] (+)= A[tmem] * B[smem desc]: A = 128 lanes x 16 K-elements (f16, two per 32-bit column)
__device__ __forceinline__ void umma_f16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc,
                                            uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t"
      "}\n"
      :
      : "r"(tmem_d), "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ float fmax2(float a, float b) {  // opaque to the compiler, which would fuse pairs into FMNMX3
  float d;
  asm("max.f32 %0, %1, %2;" : "=f"(d) : "f"(a), "f"(b));
  return d;
}
__device__ __forceinline__ float fmax3(float a, float b, float c) {
  float d;
  asm("max.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c));
  return d;
}
// (x0, x1) = (x0, x1) * a + b on the packed-f32 FMA pipe
__device__ __forceinline__ void ffma2(float& x0, float& x1, float a, float b) {
  uint64_t xv, av, bv, r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(xv) : "f"(x0), "f"(x1));
  asm("mov.b64 %0, {%1, %2};" : "=l"(av) : "f"(a), "f"(a));
  asm("mov.b64 %0, {%1, %2};" : "=l"(bv) : "f"(b), "f"(b));
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(xv), "l"(av), "l"(bv));
  asm("mov.b64 {%0, %1}, %2;" : "=f"(x0), "=f"(x1) : "l"(r));
}
__device__ __forceinline__ void fadd2(float& s0, float& s1, float x0, float x1) {
  uint64_t sv, xv, r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(sv) : "f"(s0), "f"(s1));
  asm("mov.b64 %0, {%1, %2};" : "=l"(xv) : "f"(x0), "f"(x1));
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(sv), "l"(xv));
  asm("mov.b64 {%0, %1}, %2;" : "=f"(s0), "=f"(s1) : "l"(r));
}

// Bounded wait without printf (its argument buffer lives in local memory: see below). A stuck pipeline records the tag of
// the barrier in g_att4_fault (read back by the host wrapper after a failed launch) and traps.
__device__ int g_att4_fault = 0;
__device__ __forceinline__ uint32_t mbar_test_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.b32 %0, 1, 0, p;\n\t"
      "}\n"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok;
}
#ifndef ATT4_SPIN
#define ATT4_SPIN 0   // 1: poll with the non-blocking test_wait, 0: suspending try_wait (measured: no difference for the main roles)
#endif
__device__ __forceinline__ void mbar_wait4(uint32_t bar, uint32_t parity, int tag) {
#if ATT4_SPIN
  if (mbar_test_wait(bar, parity)) return;
  const long long t0 = clock64();
  while (!mbar_test_wait(bar, parity)) {
#else
  if (mbar_try_wait(bar, parity)) return;
  const long long t0 = clock64();
  while (!mbar_try_wait(bar, parity)) {
#endif
    if (clock64() - t0 > ORON_WATCHDOG_CYCLES) {
      g_att4_fault = tag * 1000000 + int(blockIdx.x) * 1000 + int(threadIdx.x);
      __trap();
    }
  }
}

// Segment fields are read with scalar loads into registers wherever a role needs them. (A struct cursor kept in local
// memory costs an L2 round trip per access here: with 2 x 112 KB of shared memory per SM the L1 has no capacity left, so
// every local-memory access misses -- that, not the softmax arithmetic, bounded the first versions of this kernel.)
__device__ __forceinline__ int seg_field(const Attn4Seg* segs, int s, int f) {
  return reinterpret_cast<const int*>(segs + s)[f];  // 0 b, 1 h, 2 qt, 3 j0, 4 j1, 5 slot, 6 owner, 7 nparts
}

// Combine the parts of one split item (parts live in slots 2 (owner + p) + (p == 0), p = 0 .. nparts - 1): NT threads,
// 8 threads per row (16 bytes of f16 each), RPT rows per thread per round, four parts per round of loads, all of a round's
// loads in flight at once; online (running-max) combination.
template <int NT, int RPT, int PPR = 4>
__device__ __forceinline__ void att4_combine_item(int tid, int owner, int nparts, int b, int h, int qt, const __half* ws_o,
                                                  const float* ws_ml, __nv_bfloat16* out, long long ldo, float* lse,
                                                  int rows_per_batch, int heads, int row_begin = 0, int row_end = ATT4_TILE) {
  constexpr int ROWS = NT / 8;  // rows per sub-pass
  const int rsub = tid >> 3, c8 = (tid & 7) * 8;
  auto slot_of = [&](int p) { return (long long)(2 * (owner + p) + (p == 0 ? 1 : 0)); };
#pragma unroll 1
  for (int r0 = row_begin; r0 < row_end; r0 += ROWS * RPT) {
    float acc[RPT][8], m_run[RPT], l_run[RPT];
#pragma unroll
    for (int j = 0; j < RPT; ++j) {
      m_run[j] = -INFINITY;
      l_run[j] = 0.f;
#pragma unroll
      for (int k = 0; k < 8; ++k) acc[j][k] = 0.f;
    }
#pragma unroll 1
    for (int p0 = 0; p0 < nparts; p0 += PPR) {
      float2 ml[RPT][PPR];
      uint4 u[RPT][PPR];
#pragma unroll
      for (int i = 0; i < PPR; ++i) {
        const long long sl = slot_of(min(p0 + i, nparts - 1)) * ATT4_TILE;
#pragma unroll
        for (int j = 0; j < RPT; ++j) {
          const long long rr = sl + r0 + j * ROWS + rsub;
          ml[j][i] = __ldcg(reinterpret_cast<const float2*>(ws_ml + rr * 2));
          u[j][i] = __ldcg(reinterpret_cast<const uint4*>(ws_o + rr * ATT4_D + c8));
        }
      }
#pragma unroll
      for (int j = 0; j < RPT; ++j) {
        float m_new = m_run[j];
#pragma unroll
        for (int i = 0; i < PPR; ++i) if (p0 + i < nparts) m_new = fmaxf(m_new, ml[j][i].x);
        const float f_old = ex2_approx(m_run[j] - m_new);  // 0 on the first round (m_run = -inf)
        l_run[j] *= f_old;
#pragma unroll
        for (int k = 0; k < 8; ++k) acc[j][k] *= f_old;
#pragma unroll
        for (int i = 0; i < PPR; ++i) {
          const float wgt = p0 + i < nparts ? ml[j][i].y * ex2_approx(ml[j][i].x - m_new) : 0.f;
          l_run[j] += wgt;
          const uint32_t wd[4] = {u[j][i].x, u[j][i].y, u[j][i].z, u[j][i].w};
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const float2 fv = __half22float2(*reinterpret_cast<const __half2*>(&wd[k]));
            acc[j][2 * k] = fmaf(fv.x, wgt, acc[j][2 * k]);
            acc[j][2 * k + 1] = fmaf(fv.y, wgt, acc[j][2 * k + 1]);
          }
        }
        m_run[j] = m_new;
      }
    }
#pragma unroll
    for (int j = 0; j < RPT; ++j) {
      const int t = qt * ATT4_TILE + r0 + j * ROWS + rsub;
      if (t >= rows_per_batch) continue;
      const float inv_l = 1.0f / l_run[j];
      if (lse != nullptr && c8 == 0)
        lse[((long long)b * heads + h) * rows_per_batch + t] = log2f(l_run[j]) + m_run[j] - ATT4_P_EXP_BIAS;
      *reinterpret_cast<uint4*>(out + ((long long)b * rows_per_batch + t) * ldo + h * ATT4_D + c8) =
          make_uint4(pack_bf16x2(acc[j][0] * inv_l, acc[j][1] * inv_l), pack_bf16x2(acc[j][2] * inv_l, acc[j][3] * inv_l),
                     pack_bf16x2(acc[j][4] * inv_l, acc[j][5] * inv_l), pack_bf16x2(acc[j][6] * inv_l, acc[j][7] * inv_l));
    }
  }
}

// The CTA's segment list: from the plan, or (one CTA per item) the single segment the first thread left in shared memory.
// Called again inside every role after its setmaxnreg: ptxas keeps values that live across that instruction in local
// memory and reloads them at every use.
__device__ __forceinline__ void att4_segments(const Attn4Args& args, const uint8_t* smem_raw, const Attn4Seg*& segs, int& nseg) {
  if (args.plan_hdr != nullptr) {
    nseg = args.plan_nseg[blockIdx.x];
    segs = args.plan_segs + (long long)blockIdx.x * args.plan_hdr->seg_stride;
  } else {
    segs = reinterpret_cast<const Attn4Seg*>(smem_raw + ATT4_BAR_OFF + 192);
    nseg = 1;
  }
}

// DBG: per-phase time accounting of the softmax / MMA threads into args.dbg (tools/attn4_trace.py); production: false.
// ABL: ablation switches for tools/kernel_bench.py (bit 0: no exp2, bit 1: no row max, bit 2: no scale / sum / pack).
template <bool DBG, int ABL = 0>
__global__ void __launch_bounds__(ATT4_THREADS, 2)
attn_fwd4_kernel(const __grid_constant__ CUtensorMap tmQKV, const Attn4Args args) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t smem_base = smem_u32(smem_raw);
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  // one-CTA-per-item launch: the item as a single whole segment, kept behind the barriers in dynamic shared memory
  Attn4Seg& own_seg = *reinterpret_cast<Attn4Seg*>(smem_raw + ATT4_BAR_OFF + 192);

  pdl_launch_dependents();
  const Attn4Seg* segs;
  int nseg;
  if (args.plan_hdr != nullptr) {
    nseg = args.plan_nseg[blockIdx.x];
    segs = args.plan_segs + (long long)blockIdx.x * args.plan_hdr->seg_stride;
    const Attn4PlanHeader& ph = *args.plan_hdr;
    if (ph.magic != ATT4_PLAN_MAGIC || ph.nbatch != args.nbatch || ph.rows != args.rows_per_batch || ph.heads != args.heads ||
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
    const int nt = (len + ATT4_TILE - 1) / ATT4_TILE;
    nseg = qt < nt ? 1 : 0;  // a query tile entirely beyond the sequence: the out-projection masks these rows
    if (threadIdx.x == 0) { own_seg.b = b; own_seg.h = h; own_seg.qt = qt; own_seg.j0 = 0; own_seg.j1 = nt; own_seg.slot = -1; own_seg.owner = -1; own_seg.nparts = 0; }
    segs = &own_seg;
  }
  if (nseg == 0) return;  // CTA-uniform, nothing allocated yet
  if ((smem_base & 1023u) != 0) {
    if (threadIdx.x == 0) printf("[oron] attention: dynamic smem not 1024-byte aligned\n");
    __trap();
  }

  const uint32_t bar_base = smem_base + ATT4_BAR_OFF;
  // barriers (8 bytes each from bar_base; macros so that every role addresses them from its own bar_base register)
#define q_full(st) (bar_base + 8u * (st))          /* TMA -> MMA */
#define q_empty(st) (bar_base + 8u * (2 + (st)))   /* MMA -> TMA: every S of the segment has retired */
#define k_full(st) (bar_base + 8u * (4 + (st)))
#define k_empty(st) (bar_base + 8u * (6 + (st)))   /* MMA -> TMA: S(i) has retired */
#define v_full(st) (bar_base + 8u * (8 + (st)))
#define v_empty(st) (bar_base + 8u * (11 + (st)))  /* MMA -> TMA: P(i) V(i) has retired */
#define s_full (bar_base + 8u * 14)     /* MMA -> softmax: S(i) is in TMEM */
#define s_free (bar_base + 8u * 15)     /* softmax -> MMA: S(i) is in registers (4 warp arrivals) */
#define p_full (bar_base + 8u * 16)     /* softmax -> MMA: P(i) is in TMEM (4 warp arrivals) */
#define o_full (bar_base + 8u * 17)     /* MMA -> softmax: O includes P(i) V(i); also: the P columns are free */
#define tmem_slot (bar_base + 8u * 18)
#define part_ready (bar_base + 8u * 19) /* softmax -> combine warps: a partial result is in the workspace (4 warp arrivals) */

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmQKV);
    for (int st = 0; st < ATT4_QS; ++st) { mbar_init(q_full(st), 1); mbar_init(q_empty(st), 1); }
    for (int st = 0; st < ATT4_KS; ++st) { mbar_init(k_full(st), 1); mbar_init(k_empty(st), 1); }
    for (int st = 0; st < ATT4_VS; ++st) { mbar_init(v_full(st), 1); mbar_init(v_empty(st), 1); }
    mbar_init(s_full, 1);
    mbar_init(s_free, 4);
    mbar_init(p_full, 4);
    mbar_init(o_full, 1);
    mbar_init(part_ready, 4);
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, ATT4_TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  pdl_wait();  // the QKV activations of the previous kernel are visible from here on
  if (threadIdx.x == 0) {
    ATT4_STAMP(0);
    if (args.dbg) { unsigned long long gt; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(gt)); args.dbg[(long long)blockIdx.x * 16 + 8] = (long long)gt; }
  }

  if (warp == 0) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(ATT4_REGS_AUX));
    if ((threadIdx.x & 31) == 0) {
      // ===================== TMA producer =====================
      const Attn4Seg* segs;
      int nseg;
      att4_segments(args, smem_raw, segs, nseg);
      const uint32_t smem_base = smem_u32(smem_raw);
      const uint32_t sQ0 = smem_base, sK0 = smem_base + ATT4_QS * ATT4_TILE_BYTES, sV0 = smem_base + (ATT4_QS + ATT4_KS) * ATT4_TILE_BYTES;
      const uint32_t bar_base = smem_base + ATT4_BAR_OFF;
      const int HD = args.heads * ATT4_D;
      int it = 0, vs = 0, vph = 1;
      for (int s = 0; s < nseg; ++s) {
        const int b = seg_field(segs, s, 0), h = seg_field(segs, s, 1), qt = seg_field(segs, s, 2);
        const int j0 = seg_field(segs, s, 3), j1 = seg_field(segs, s, 4);
        const int qs = s & 1;
        mbar_wait4(q_empty(qs), ((s >> 1) & 1u) ^ 1u, 10);
        mbar_arrive_expect_tx(q_full(qs), ATT4_TILE_BYTES);
        tma_load_3d(sQ0 + qs * ATT4_TILE_BYTES, &tmQKV, q_full(qs), h * ATT4_D, qt * ATT4_TILE, b);
        for (int j = j0; j < j1; ++j, ++it) {
          const int ks = it & 1;
          mbar_wait4(k_empty(ks), ((it >> 1) & 1u) ^ 1u, 11);
          if ((ABL & 8) && it >= 2) mbar_arrive(k_full(ks));
          else {
          mbar_arrive_expect_tx(k_full(ks), ATT4_TILE_BYTES);
          tma_load_3d(sK0 + ks * ATT4_TILE_BYTES, &tmQKV, k_full(ks), HD + h * ATT4_D, j * ATT4_TILE, b);
          }
          mbar_wait4(v_empty(vs), uint32_t(vph), 12);
          if ((ABL & 8) && it >= 3) mbar_arrive(v_full(vs));
          else {
          mbar_arrive_expect_tx(v_full(vs), ATT4_TILE_BYTES);
          tma_load_3d(sV0 + vs * ATT4_TILE_BYTES, &tmQKV, v_full(vs), 2 * HD + h * ATT4_D, j * ATT4_TILE, b);
          }
          if (++vs == ATT4_VS) { vs = 0; vph ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(ATT4_REGS_AUX));
    if ((threadIdx.x & 31) == 0) {
      // ===================== MMA issuer =====================
      const Attn4Seg* segs;
      int nseg;
      att4_segments(args, smem_raw, segs, nseg);
      const uint32_t smem_base = smem_u32(smem_raw);
      const uint32_t sQ0 = smem_base, sK0 = smem_base + ATT4_QS * ATT4_TILE_BYTES, sV0 = smem_base + (ATT4_QS + ATT4_KS) * ATT4_TILE_BYTES;
      const uint32_t bar_base = smem_base + ATT4_BAR_OFF;
      uint32_t tmem_base;
      asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(bar_base + 8u * 18));
      const uint32_t tmem_S = tmem_base, tmem_P = tmem_base + 128, tmem_O = tmem_base + 192;
      constexpr uint32_t idesc_s = make_idesc_bf16(128, 128, 0, 0);
      constexpr uint32_t idesc_o = make_idesc_f16(128, 64, 0, 1);  // f16 P (TMEM) x f16 V, B = V is MN-major
      const uint64_t qdesc0 = make_smem_desc_sw128(sQ0, 16, 1024);
      const uint64_t kdesc0 = make_smem_desc_sw128(sK0, 16, 1024);
      const uint64_t vdesc0 = make_smem_desc_sw128(sV0, 1024, 1024);
      constexpr uint64_t kStageDesc = uint64_t(ATT4_TILE_BYTES >> 4);  // descriptor address field counts 16-byte units
      long long wsf = 0, wkf = 0, wpf = 0, wvf = 0, m_s = 0, m_sc = 0, m_pv = 0, m_pc = 0, m_misc = 0, mt = DBG ? clock64() : 0;
#define ATT4_MK(acc) do { if (DBG) { const long long n_ = clock64(); acc += n_ - mt; mt = n_; } } while (0)
      // two counters over the same tile sequence: "n" = the tile whose S is issued next (one tile ahead), "w" = the tile
      // whose P V is issued in this iteration. Only the tile counts of the segments matter here.
      auto count_of = [&](int s) { return s < nseg ? seg_field(segs, s, 4) - seg_field(segs, s, 3) : 0; };
      int s_n = 0, left_n = count_of(0), next_n = count_of(1);
      int s_w = 0, left_w = left_n, next_w = next_n;
      bool first_n = true, first_w = true;
      auto issue_S = [&](int itn) {
        ATT4_MK(m_misc);
        if (first_n) mbar_wait4(q_full(s_n & 1), (s_n >> 1) & 1u, 13);
        mbar_wait4(k_full(itn & 1), (itn >> 1) & 1u, 14);
        ATT4_MK(wkf);
        tc_fence_after();
        const uint64_t qd = qdesc0 + kStageDesc * uint64_t(s_n & 1);
        const uint64_t kd = kdesc0 + kStageDesc * uint64_t(itn & 1);
#pragma unroll
        for (int k = 0; k < 4; ++k) if (!(ABL & 32) || k == 0) umma_bf16_ss(tmem_S, qd + uint64_t(2 * k), kd + uint64_t(2 * k), idesc_s, k != 0);
        ATT4_MK(m_s);
        umma_commit(s_full);
        umma_commit(k_empty(itn & 1));
        if (left_n == 1) umma_commit(q_empty(s_n & 1));  // last S of the segment: Q may be replaced
        ATT4_MK(m_sc);
        // advance "n"
        first_n = false;
        if (--left_n == 0) { ++s_n; left_n = next_n; next_n = count_of(s_n + 1); first_n = true; }
      };
      issue_S(0);
      int vs = 0, vph = 0;
      for (int it = 0; s_w < nseg; ++it) {
        if (s_n < nseg) {
          ATT4_MK(m_misc);
          mbar_wait4(s_free, it & 1u, 15);  // S(it) is in registers: the S columns may be overwritten
          ATT4_MK(wsf);
          issue_S(it + 1);
        }
        ATT4_MK(m_misc);
        mbar_wait4(p_full, it & 1u, 16);    // P(it) in TMEM (and O rescaled if the running max moved)
        ATT4_MK(wpf);
        mbar_wait4(v_full(vs), uint32_t(vph), 17);
        ATT4_MK(wvf);
        tc_fence_after();
        const uint64_t vd = vdesc0 + kStageDesc * uint64_t(vs);
        const uint32_t acc0 = first_w ? 0u : 1u;
#pragma unroll
        for (int kk = 0; kk < 8; ++kk)  // 16 keys per step: 8 TMEM columns of P, 16 rows (2048 bytes, >> 4 = 128) of V
          if (!(ABL & 16) || kk == 0) umma_f16_ts(tmem_O, tmem_P + 8 * kk, vd + uint64_t(128 * kk), idesc_o, kk != 0 ? 1u : acc0);
        ATT4_MK(m_pv);
        umma_commit(o_full);
        umma_commit(v_empty(vs));
        ATT4_MK(m_pc);
        if (++vs == ATT4_VS) { vs = 0; vph ^= 1; }
        first_w = false;
        if (--left_w == 0) { ++s_w; left_w = next_w; next_w = count_of(s_w + 1); first_w = true; }
      }
#ifdef ATT4_TRACE_MMA
      ATT4_PUT(1, wsf); ATT4_PUT(2, wkf); ATT4_PUT(3, wpf); ATT4_PUT(4, wvf); ATT4_PUT(5, m_s); ATT4_PUT(6, m_sc); ATT4_PUT(7, m_pv); ATT4_PUT(10, m_pc); ATT4_PUT(11, m_misc);
#endif
    }
  } else if (warp < 4) {
    // ===================== combine warps (64 threads): the split items =====================
    // args.ws_cnt != nullptr: the softmax threads only publish a partial result and move on. These two warps count the
    // CTA's parts as they are published (arrival counter of the item, cnt[2 owner]); then, for every split item the CTA
    // has a part of, they wait until all of its parts have arrived and combine THEIR SHARE of the item's 128 rows (part p
    // of n takes row groups [16 p / n, 16 (p + 1) / n) of eight rows): the combination of an item is spread over the CTAs
    // that hold its parts, eight rows per round of loads, and nobody merges a whole item alone. With the two-phase plan
    // (a CFG pair) the parts are the FIRST thing every CTA runs, so all of this happens under the softmax of the whole
    // item that follows. Arrivals are counted for all parts before the first wait: a wait never delays another CTA.
    // The last part to leave (cnt[2 owner + 1]) zeroes both counters for the next launch (graph replays included).
    // args.ws_cnt == nullptr: attn4_combine_kernel does it after the launch.
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(ATT4_REGS_AUX));
    if (args.ws_cnt != nullptr) {
      const Attn4Seg* segs;
      int nseg;
      att4_segments(args, smem_raw, segs, nseg);
      const uint32_t bar_base = smem_u32(smem_raw) + ATT4_BAR_OFF;
      const int hid = threadIdx.x - 64;
      int np = 0;
      for (int sidx = 0; sidx < nseg; ++sidx) {
        if (seg_field(segs, sidx, 5) < 0) continue;
        // a long wait: sleep in the suspending try_wait instead of polling (the softmax warps share these schedulers)
        { const long long t0w = clock64(); while (!mbar_try_wait(part_ready, np & 1u)) { if (clock64() - t0w > ORON_WATCHDOG_CYCLES) { g_att4_fault = 22000000 + int(blockIdx.x) * 1000; __trap(); } } }
        ++np;
        if (hid == 0) {
          __threadfence();  // cumulative: the softmax threads' stores (observed through the barrier) before the count
          atomicAdd(args.ws_cnt + 2 * seg_field(segs, sidx, 6), 1);
        }
      }
      for (int sidx = 0; sidx < nseg; ++sidx) {
        if (seg_field(segs, sidx, 5) < 0) continue;
        const int owner = seg_field(segs, sidx, 6), nparts = seg_field(segs, sidx, 7);
        if (hid == 0) {
          const long long t0w = clock64();
          int seen;
          do {
            asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(seen) : "l"(args.ws_cnt + 2 * owner) : "memory");
            if (seen >= nparts) break;
            __nanosleep(64);
            if (clock64() - t0w > ORON_WATCHDOG_CYCLES) { g_att4_fault = 23000000 + int(blockIdx.x) * 1000; __trap(); }
          } while (true);
        }
        asm volatile("bar.sync 1, 64;" ::: "memory");
        __threadfence();
        const int p = int(blockIdx.x) - owner;
        att4_combine_item<64, 1, 4>(hid, owner, nparts, seg_field(segs, sidx, 0), seg_field(segs, sidx, 1), seg_field(segs, sidx, 2),
                                    args.ws_o, args.ws_ml, args.out, args.ldo, args.lse, args.rows_per_batch, args.heads,
                                    8 * ((16 * p) / nparts), 8 * ((16 * (p + 1)) / nparts));
        asm volatile("bar.sync 1, 64;" ::: "memory");  // every read of the parts is done
        if (hid == 0) {
          const int old = atomicAdd(args.ws_cnt + 2 * owner + 1, 1);
          if (old == nparts - 1) { args.ws_cnt[2 * owner] = 0; args.ws_cnt[2 * owner + 1] = 0; }
        }
      }
    }
  } else {
    // ===================== softmax threads =====================
    asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(ATT4_REGS_SOFTMAX));
    const Attn4Seg* segs;
    int nseg;
    att4_segments(args, smem_raw, segs, nseg);
    const uint32_t bar_base = smem_u32(smem_raw) + ATT4_BAR_OFF;
    uint32_t tmem_base;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(bar_base + 8u * 18));
    const uint32_t tmem_S = tmem_base, tmem_P = tmem_base + 128, tmem_O = tmem_base + 192;
    const int lane = threadIdx.x & 31;
    const int q = (threadIdx.x >> 5) & 3;
    const int r = q * 32 + lane;  // query row inside the tile == TMEM lane
    const uint32_t lane_off = uint32_t(q * 32) << 16;
    const float c = args.scale_log2;
    float mc = -INFINITY;  // running max (already multiplied by c), possibly stale by < 2^8
    float l_run = 0.f;     // softmax denominator in the same (stale-max, 2^7-biased) scale as O
    long long ws_ = 0, wo_ = 0, wep = 0, a_ld = 0, a_max = 0, a_exp = 0, a_st = 0, a_misc = 0, tt = 0;
#define ATT4_TK(acc) do { if (DBG) { const long long n_ = clock64(); acc += n_ - tt; tt = n_; } } while (0)
    int ntiles = 0;
    const long long tloop0 = DBG ? clock64() : 0;
    tt = tloop0;
    // per-segment state in registers: tiles left, valid keys left from the current tile on, and the same for the next
    // segment (prefetched one segment ahead so that a segment change costs no memory round trip)
    auto keys_of = [&](int s) {  // valid keys from the segment's first tile on
      const int b = seg_field(segs, s, 0);
      const int len = args.seq_lens ? min(args.seq_lens[b], args.rows_per_batch) : args.rows_per_batch;
      return len - seg_field(segs, s, 3) * ATT4_TILE;
    };
    auto count_of = [&](int s) { return seg_field(segs, s, 4) - seg_field(segs, s, 3); };
    int s = 0, left = count_of(0), rem = keys_of(0);
    int next_left = nseg > 1 ? count_of(1) : 0, next_rem = nseg > 1 ? keys_of(1) : 0;
    bool first = true;
    for (int it = 0; s < nseg; ++it) {
      ++ntiles;
      if (first) { mc = -INFINITY; l_run = 0.f; }
      const int n_valid = min(ATT4_TILE, rem);  // CTA-uniform
      const bool last = left == 1;
      ATT4_TK(a_misc);
      mbar_wait4(s_full, it & 1u, 18);
      ATT4_TK(ws_);
      tc_fence_after();
      // ---- the whole row into registers, S columns back to the tensor core ----
      uint32_t v[128];
      tmem_ld_row32(tmem_S + lane_off + 0, v);
      if (ABL & 64) {
#pragma unroll
        for (int k = 32; k < 128; ++k) v[k] = v[k & 31] + k;
      } else {
      tmem_ld_row32(tmem_S + lane_off + 32, v + 32);
      tmem_ld_row32(tmem_S + lane_off + 64, v + 64);
      tmem_ld_row32(tmem_S + lane_off + 96, v + 96);
      }
      tmem_wait_ld();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(s_free);
      ATT4_TK(a_ld);
      if (n_valid < ATT4_TILE) {
#pragma unroll
        for (int k = 0; k < 128; ++k)
          if (k >= n_valid) v[k] = 0xff800000u;  // -inf: exp2 gives probability 0
      }
      // ---- row maximum ----
      float mm[8];  // plain 2-input FMNMX: the 3-input form issues at a quarter of the rate on this part
#pragma unroll
      for (int k = 0; k < 8; ++k) mm[k] = fmax2(__uint_as_float(v[k]), __uint_as_float(v[k + 8]));
#pragma unroll
      for (int k0 = 16; k0 < 128; k0 += 8) {
#pragma unroll
        for (int k = 0; k < 8; ++k) mm[k] = fmax2(mm[k], __uint_as_float(v[k0 + k]));
      }
      float mxc = fmaxf(fmaxf(fmaxf(mm[0], mm[1]), fmaxf(mm[2], mm[3])), fmaxf(fmaxf(mm[4], mm[5]), fmaxf(mm[6], mm[7]))) * c;
      if (ABL & 2) mxc = __uint_as_float(v[0]) * c;
      // ---- lazy rescale: only when this tile's max exceeds the running one by more than 2^8 ----
      const bool need = mxc > mc + ATT4_RESCALE_LOG2;
      bool o_done = first;  // first tile of a segment: the P columns are free (the previous segment's epilogue waited)
      if (!first && __any_sync(0xffffffffu, need)) {
        // rare: P(it-1) V(it-1) must be folded into O before O and l are rescaled
        mbar_wait4(o_full, (it - 1) & 1u, 19);
        tc_fence_after();
        o_done = true;
        const float f = need ? ex2_approx(mc - mxc) : 1.0f;
#pragma unroll 1
        for (int c0 = 0; c0 < ATT4_D; c0 += 32) {
          uint32_t o[32];
          tmem_ld_row32(tmem_O + lane_off + c0, o);
          tmem_wait_ld();
#pragma unroll
          for (int k = 0; k < 32; ++k) o[k] = __float_as_uint(__uint_as_float(o[k]) * f);
          tmem4_st32(tmem_O + lane_off + c0, o);
        }
        l_run *= f;
        tmem4_wait_st();
      }
      if (need) mc = mxc;
      ATT4_TK(a_max);
      // ---- P = 2^7 * exp2(S*c - m) -> f16, two keys per 32-bit word ----
      const float nmcb = ATT4_P_EXP_BIAS - mc;
      uint32_t pk[64];
      float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
#pragma unroll
      for (int k = 0; k < 128; k += 4) {
        float x0 = __uint_as_float(v[k]), x1 = __uint_as_float(v[k + 1]);
        float x2 = __uint_as_float(v[k + 2]), x3 = __uint_as_float(v[k + 3]);
        if (ABL & 4) {
          pk[k / 2] = __float_as_uint(x0) ^ __float_as_uint(x1);
          pk[k / 2 + 1] = __float_as_uint(x2) ^ __float_as_uint(x3);
          s0 += x0;
          continue;
        }
        ffma2(x0, x1, c, nmcb);
        ffma2(x2, x3, c, nmcb);
        if (!(ABL & 1)) {
          x0 = ex2_approx(x0); x1 = ex2_approx(x1);
          constexpr int POLYSEL = (ABL >> 7) & 3;  // tuning variants: 0 = 1 in 8 (default), 1 = none, 2 = 1 in 4, 3 = 3 in 8
          const bool poly = POLYSEL == 1 ? false
                          : POLYSEL == 0 ? ((k & 4) && k >= 32 && k < 96)
                          : POLYSEL == 2 ? ((k & 4) && k >= 0 && k < 128)
                                         : (((k & 4) && k >= 16) || ((k & 8) && !(k & 4) && k >= 32 && k < 96));
          if (ATT4_EX2_POLY && poly) {
            ex2_poly2(x2, x3);
          } else {
            x2 = ex2_approx(x2); x3 = ex2_approx(x3);
          }
        }
        fadd2(s0, s1, x0, x1);
        fadd2(s2, s3, x2, x3);
        const __half2 a = __floats2half2_rn(x0, x1), b = __floats2half2_rn(x2, x3);
        pk[k / 2] = *reinterpret_cast<const uint32_t*>(&a);
        pk[k / 2 + 1] = *reinterpret_cast<const uint32_t*>(&b);
      }
      l_run += (s0 + s1) + (s2 + s3);
      ATT4_TK(a_exp);
      if (!o_done) {
        mbar_wait4(o_full, (it - 1) & 1u, 20);  // P(it-1) V(it-1) has retired: the P columns may be overwritten
        tc_fence_after();
      }
      ATT4_TK(wo_);
      tmem4_st16(tmem_P + lane_off + 0, pk);
      if (!(ABL & 64)) {
      tmem4_st16(tmem_P + lane_off + 16, pk + 16);
      tmem4_st16(tmem_P + lane_off + 32, pk + 32);
      tmem4_st16(tmem_P + lane_off + 48, pk + 48);
      }
      tmem4_wait_st();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(p_full);
      ATT4_TK(a_st);
      rem -= ATT4_TILE;
      --left;
      first = false;
      if (!last) continue;

      // ---- end of a segment: O / l to the output (whole item) or to the workspace (partial item, f16 + (m, l)) ----
      const int sg_b = seg_field(segs, s, 0), sg_h = seg_field(segs, s, 1), sg_qt = seg_field(segs, s, 2);
      const int slot = seg_field(segs, s, 5);
      mbar_wait4(o_full, it & 1u, 21);
      tc_fence_after();
      float o[64];
      tmem_ld_row32(tmem_O + lane_off + 0, reinterpret_cast<uint32_t*>(o));
      tmem_ld_row32(tmem_O + lane_off + 32, reinterpret_cast<uint32_t*>(o) + 32);
      tmem_wait_ld();
      const int t = sg_qt * ATT4_TILE + r;
      const float inv_l = 1.0f / l_run;
      if (slot >= 0) {
        // partial item: O_p / l_p (f16) and (m_p, l_p) to the workspace; attn4_combine_kernel merges the parts
        uint4* dst = reinterpret_cast<uint4*>(args.ws_o + ((long long)slot * ATT4_TILE + r) * ATT4_D);
#pragma unroll
        for (int g = 0; g < 8; ++g) {
          uint32_t hw[4];
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const __half2 hh = __floats2half2_rn(o[8 * g + 2 * k] * inv_l, o[8 * g + 2 * k + 1] * inv_l);
            hw[k] = *reinterpret_cast<const uint32_t*>(&hh);
          }
          __stcg(dst + g, make_uint4(hw[0], hw[1], hw[2], hw[3]));
        }
        __stcg(reinterpret_cast<float2*>(args.ws_ml + ((long long)slot * ATT4_TILE + r) * 2), make_float2(mc, l_run));
        if (args.ws_cnt != nullptr) {
          __syncwarp();
          if (lane == 0) mbar_arrive(part_ready);
        }
      } else if (t < args.rows_per_batch) {
        if (args.lse != nullptr)  // log2 sum_k 2^(s_k c)
          args.lse[((long long)sg_b * args.heads + sg_h) * args.rows_per_batch + t] = log2f(l_run) + mc - ATT4_P_EXP_BIAS;
        uint4* dst = reinterpret_cast<uint4*>(args.out + ((long long)sg_b * args.rows_per_batch + t) * args.ldo + sg_h * ATT4_D);
#pragma unroll
        for (int g = 0; g < 8; ++g)
          dst[g] = make_uint4(pack_bf16x2(o[8 * g] * inv_l, o[8 * g + 1] * inv_l), pack_bf16x2(o[8 * g + 2] * inv_l, o[8 * g + 3] * inv_l),
                              pack_bf16x2(o[8 * g + 4] * inv_l, o[8 * g + 5] * inv_l), pack_bf16x2(o[8 * g + 6] * inv_l, o[8 * g + 7] * inv_l));
      }
      // next segment (its counts were fetched a segment ago)
      ++s;
      left = next_left;
      rem = next_rem;
      first = true;
      if (s + 1 < nseg) { next_left = count_of(s + 1); next_rem = keys_of(s + 1); }
      ATT4_TK(wep);
    }
    tc_fence_before();
#if !defined(ATT4_TRACE_MMA) && !defined(ATT4_TRACE_HELPER)
    if (threadIdx.x == 128) { ATT4_STAMP(14); ATT4_PUT(1, ws_); ATT4_PUT(2, wo_); ATT4_PUT(3, wep); ATT4_PUT(13, (long long)ntiles); ATT4_PUT(4, a_ld); ATT4_PUT(5, a_max); ATT4_PUT(6, a_exp); ATT4_PUT(7, a_st); ATT4_PUT(12, clock64() - tloop0); ATT4_PUT(10, a_misc); }
#endif
  }

  __syncthreads();
  if (threadIdx.x == 0) {
    ATT4_STAMP(15);
    if (args.dbg) { unsigned long long gt; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(gt)); args.dbg[(long long)blockIdx.x * 16 + 9] = (long long)gt; }
  }
  if ((threadIdx.x >> 5) == 1) {
    tc_fence_after();
    const uint32_t bar_base = smem_u32(smem_raw) + ATT4_BAR_OFF;
    uint32_t tmem_base;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));
    tmem_dealloc(tmem_base, ATT4_TMEM_COLS);
  }
#undef q_full
#undef q_empty
#undef k_full
#undef k_empty
#undef v_full
#undef v_empty
#undef s_full
#undef s_free
#undef p_full
#undef o_full
#undef tmem_slot
#undef part_ready
}

// ------------------------------------------------------------------------------------------------------------
// The plan: one thread per CTA of the launch cuts its share [u0, u1) of the flat (item, key tile) list into segments.
// When there are no more items than CTAs every CTA takes one whole item (no splits); otherwise the shares are equal
// (total * c / G) and an item that straddles shares is split along its keys. Runs once per set of sequence lengths
// (oron_attention_plan), not per call.
// ------------------------------------------------------------------------------------------------------------
struct Plan4Walk {
  int heads, nbatch, rows;
  const int* lens;
  int b, item, tile, nt;   // item = h * nt + qt inside batch element b
  long long item_u0;
  __device__ int len_of(int bb) const { return lens ? min(lens[bb], rows) : rows; }
  __device__ static int tiles_of(int l) { return (l + ATT4_TILE - 1) / ATT4_TILE; }
  __device__ void totals(long long& units, long long& items) const {
    units = 0; items = 0;
    for (int bb = 0; bb < nbatch; ++bb) { const long long n = tiles_of(len_of(bb)); units += n * n * heads; items += n * heads; }
  }
  __device__ void seek(long long u) {
    long long base = 0;
    for (b = 0; b < nbatch; ++b) {
      nt = tiles_of(len_of(b));
      const long long n = (long long)nt * nt * heads;
      if (u < base + n) break;
      base += n;
    }
    const long long rr = u - base;
    item = int(rr / nt);
    tile = int(rr - (long long)item * nt);
    item_u0 = u - tile;
  }
  // first unit of the `i`-th item of the whole problem
  __device__ long long item_start(long long i) const {
    long long base = 0;
    for (int bb = 0; bb < nbatch; ++bb) {
      const long long n = tiles_of(len_of(bb));
      if (i < n * heads) return base + i * n;
      i -= n * heads;
      base += n * n * heads;
    }
    return base;
  }
};

__global__ void attn4_plan_kernel(Attn4PlanHeader* hdr, int* nseg_out, Attn4Seg* segs_out, Attn4Merge* merge, int* cnt, const int* seq_lens,
                                  int nbatch, int rows, int heads, int grid, int seg_stride, int force_flat, float skew) {
  if (threadIdx.x == 0) {
    hdr->nbatch = nbatch; hdr->rows = rows; hdr->heads = heads; hdr->grid = grid; hdr->seg_stride = seg_stride;
  }
  Plan4Walk w;
  w.heads = heads; w.nbatch = nbatch; w.rows = rows; w.lens = seq_lens;
  long long total, items;
  w.totals(total, items);
  bool uniform = true;  // every batch element has the same number of key tiles
  const int nt0 = Plan4Walk::tiles_of(w.len_of(0));
  for (int bb = 1; bb < nbatch; ++bb) uniform = uniform && Plan4Walk::tiles_of(w.len_of(bb)) == nt0;
  // Three shapes of plan:
  //  per_item : no more items than CTAs -> one whole item per CTA, nothing is split.
  //  two_phase: more items than CTAs, all of one length nt0 (a CFG pair, a training batch): every CTA takes W = items / grid
  //             whole items; the R = items - W * grid left-over items are cut into equal pieces, one per CTA, which the
  //             CTA runs FIRST -- every split item is complete (and combined by the combine warps) early in the launch,
  //             and the launch ends with plain whole-item epilogues.
  //  flat     : mixed lengths (or force_flat, a test aid: many parts per item): equal contiguous shares of the flat
  //             (item, key tile) list; an item that straddles shares is split along its keys.
  const bool per_item = items <= grid && !force_flat;
  const bool two_phase = !per_item && uniform && !force_flat;
  const long long G = per_item ? items : min((long long)grid, total);  // every share holds >= 1 unit: the parts of an item sit in consecutive CTAs
  auto start_of = [&](long long cta) { return per_item ? w.item_start(cta) : (total * cta) / G; };
  auto cta_of = [&](long long u) {   // flat shares only
    long long cc = (u * G) / total;
    while (cc + 1 < G && start_of(cc + 1) <= u) ++cc;
    while (cc > 0 && start_of(cc) > u) --cc;
    return cc;
  };
  // two_phase: pieces of the left-over items' flat list [0, RU)
  const long long W = items / grid, R = items - W * grid, RU = R * nt0, Gp = min((long long)grid, RU);
  // Skew: the first CTA placed on an SM runs measurably faster than the second one (17.9 vs 21.5 us at config 2 with equal
  // work: the older CTA's warps win the issue / MUFU arbitration), and an SM whose CTAs finish apart spends the difference
  // with one CTA alone. Blocks 0 .. grid/2 - 1 (placed first, one per SM) therefore take pieces that are `skew` * share
  // tiles longer, the others as much shorter, so that the two CTAs of an SM end together.
  const double share = double(total) / double(grid);
  const long long half = Gp / 2;
  double d_units = (Gp == grid && R > 0) ? double(skew) * share : 0.0;
  const double mean_piece = Gp > 0 ? double(RU) / double(Gp) : 0.0;
  if (d_units > mean_piece - 0.5) d_units = mean_piece > 0.5 ? mean_piece - 0.5 : 0.0;  // every piece keeps >= half a tile
  auto pstart_of = [&](long long cta) {
    if (cta >= Gp) return RU;
    const double x = cta <= half ? (mean_piece + d_units) * double(cta)
                                 : (mean_piece + d_units) * double(half) + (mean_piece - d_units) * double(cta - half);
    long long v = (long long)(x + 1e-6);
    return v < 0 ? 0LL : (v > RU ? RU : v);
  };
  auto pcta_of = [&](long long u) {
    long long cc = (u * Gp) / RU;
    if (cc >= Gp) cc = Gp - 1;
    while (cc + 1 < Gp && pstart_of(cc + 1) <= u) ++cc;
    while (cc > 0 && pstart_of(cc) > u) --cc;
    return cc;
  };
  auto item_seg = [&](long long id) {  // uniform lengths: item id -> (b, h, qt)
    Attn4Seg sg;
    sg.b = int(id / ((long long)heads * nt0)); sg.h = int((id / nt0) % heads); sg.qt = int(id % nt0);
    sg.j0 = 0; sg.j1 = nt0; sg.slot = -1; sg.owner = -1; sg.nparts = 0;
    return sg;
  };
  for (int c = threadIdx.x; c < grid; c += blockDim.x) {
    int n = 0;
    Attn4Seg* out = segs_out + (long long)c * seg_stride;
    Attn4Merge me;
    me.b = me.h = me.qt = me.nparts = 0;
    if (two_phase) {
      if (c < Gp) {
        const long long u0 = pstart_of(c), u1 = pstart_of(c + 1);
        long long u = u0;
        while (u < u1 && n < seg_stride) {
          const long long li = u / nt0;
          const int j0 = int(u - li * nt0);
          const int j1 = int(min((long long)nt0, (long long)j0 + (u1 - u)));
          Attn4Seg sg = item_seg(W * grid + li);
          sg.j0 = j0; sg.j1 = j1;
          if (!(j0 == 0 && j1 == nt0)) {
            const long long own = pcta_of(li * nt0);
            sg.owner = int(own);
            sg.nparts = int(pcta_of(li * nt0 + nt0 - 1) - own + 1);
            sg.slot = 2 * c + (own == c ? 1 : 0);
            if (own == c) { me.b = sg.b; me.h = sg.h; me.qt = sg.qt; me.nparts = sg.nparts; }
          }
          out[n++] = sg;
          u += j1 - j0;
        }
      }
      for (long long k = 0; k < W && n < seg_stride; ++k) out[n++] = item_seg((long long)c * W + k);
    } else if (c < G) {
      const long long u0 = start_of(c), u1 = start_of(c + 1);
      long long u = u0;
      while (u < u1 && n < seg_stride) {
        w.seek(u);
        const int j0 = w.tile;
        const int j1 = int(min((long long)w.nt, (long long)w.tile + (u1 - u)));
        Attn4Seg sg;
        sg.b = w.b; sg.h = w.item / w.nt; sg.qt = w.item % w.nt; sg.j0 = j0; sg.j1 = j1;
        sg.slot = -1; sg.owner = -1; sg.nparts = 0;
        if (!(j0 == 0 && j1 == w.nt)) {
          const long long own = cta_of(w.item_u0);
          sg.owner = int(own);
          sg.nparts = int(cta_of(w.item_u0 + w.nt - 1) - own + 1);
          sg.slot = 2 * c + (own == c ? 1 : 0);
          if (own == c) { me.b = sg.b; me.h = sg.h; me.qt = sg.qt; me.nparts = sg.nparts; }
        }
        out[n++] = sg;
        u += j1 - j0;
      }
    }
    nseg_out[c] = n;
    merge[c] = me;
    cnt[2 * c] = 0;
    cnt[2 * c + 1] = 0;
  }
  __syncthreads();
  if (threadIdx.x == 0) { __threadfence(); hdr->magic = ATT4_PLAN_MAGIC; }
}

// ------------------------------------------------------------------------------------------------------------
// The split items: CTAs 4c .. 4c+3 combine 32 rows each of the item whose first key tile CTA c of the attention launch
// held (parts live in slots 2 (c + p) + (p == 0), p = 0 .. nparts - 1). 8 threads per row (16 bytes of f16 each); eight
// parts per round of loads, all of them in flight at once (one L2 round trip for the usual 5-7 parts); online
// (running-max) combination. Launched right behind the attention kernel (programmatic dependent launch): no atomics, no
// inter-CTA waiting. It sits on the critical path between the attention kernel and the out-projection, so its own
// latency matters: the first version (one CTA per item, four passes of two rounds) cost ~6 us per call.
// ------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
attn4_combine_kernel(const Attn4Merge* __restrict__ merge, const __half* __restrict__ ws_o, const float* __restrict__ ws_ml,
                     __nv_bfloat16* __restrict__ out, long long ldo, float* __restrict__ lse, int rows_per_batch, int heads) {
  pdl_launch_dependents();
  const Attn4Merge me = merge[blockIdx.x >> 2];  // plan data, written long before the attention launch
  // every CTA waits for the attention kernel, also those with nothing to combine: a grid whose CTAs all left without
  // waiting would count as complete and release the NEXT kernel of the stream before the attention results exist
  pdl_wait();
  if (me.nparts == 0) return;
  const int r0 = int(blockIdx.x & 3) * 32;
  att4_combine_item<256, 1, 8>(threadIdx.x, int(blockIdx.x >> 2), me.nparts, me.b, me.h, me.qt, ws_o, ws_ml, out, ldo, lse, rows_per_batch, heads, r0, r0 + 32);
}

}  // namespace oron
