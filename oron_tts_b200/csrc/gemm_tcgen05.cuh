// Persistent, warp-specialised bf16 GEMM for sm_100a:
//   D[rows, N] = A[rows, K] * W[N, K]^T   (nn.Linear layout, both operands K-major)
// TMA (SWIZZLE_128B) -> smem ring -> tcgen05.mma (M=128, N=BN, K=16) -> TMEM (double
// buffered) -> tcgen05.ld epilogue fused with the op that follows the linear layer in the
// reference (bias, GELU, RoPE, gated residual, Mish+mask ...).
//
// The same kernel runs the 1-D convolutions of the path as implicit GEMMs: the A tile for
// k-block `kb` is fetched at row offset (kb / cpb - pad) and, for grouped convs, at the
// column block of the output group; TMA zero-fills rows outside [0, rows_per_batch).
#pragma once
#include <cuda_fp16.h>

#include "ptx.cuh"

namespace oron {

enum GemmEpilogue : int {
  EPI_BF16 = 0,            // out_bf16 = act(acc + bias)              (act: see GemmArgs::act)
  EPI_F32 = 1,             // out_f32  = acc + bias (+ addend[row, col])
  EPI_QKV_ROPE = 2,        // out_bf16 = rope(acc + bias) for cols < rope_cols, else acc + bias
  EPI_GATE_RESID = 3,      // resid[row, col] += gate[b, col] * (acc + bias); masked rows untouched
  EPI_EMBED_DUAL = 4,      // v = valid ? acc + addend[row, col] : 0 ; out_f32 = v ; out_bf16 = v
  EPI_MISH_MASK_BF16 = 5,  // out_bf16 = valid ? mish(acc + bias) : 0
  EPI_MISH_MASK_RESID = 6, // out_f32 = (valid ? mish(acc + bias) : 0) + addend[row, col]
  EPI_SCALE_RESID = 7,     // out_f32 = valid ? addend[row,col] + colscale[col]*(acc+bias) : 0 ; opt. out_bf16
  // training-step fusions (oron_tts_b200/train.py): FeedForward up-projection forward and its data gradient
  EPI_GELU_DROP_DUAL = 8,  // out2_bf16 = pre = acc + bias ; out_bf16 = dropout(gelu_tanh(pre))   (modules.py:294-297)
  EPI_GELU_DROP_BWD = 9,   // out_bf16 = acc * gelu_tanh'(aux_bf16[row, col]) * dropout mask; aux = out2 (read only)
  EPI_GATE_RESID_DUAL = 10,  // y = acc + bias -> out2_bf16 ; out_f32 = addend + gate[b] * dropout(valid ? y : 0)  (modules.py:338, 343)
};

enum GemmAct : int { ACT_NONE = 0, ACT_GELU_TANH = 1, ACT_GELU_ERF = 2, ACT_SILU = 3 };

// LayerNorm + modulation of the UPDATED residual rows as the tail of a 2-SM EPI_GATE_RESID launch (gemm2 kernel, LNT = 1):
// replaces the ln_modulate launch that follows the attention out-projection / FeedForward down-projection of a DiTBlock
// (modules.py:338-343). counters: int32 [(ceil(m_tiles / 2) + 1) * 32] (one 128-byte line each), zero on entry; the launch leaves them zero.
struct LnTail {
  int* counters;
  const float* scale;        // (1 + scale) and shift rows of the modulation table: ptr + step * step_stride + (b % mod_nb) * mod_ld
  const float* shift;
  long long mod_ld;
  int mod_nb;
  long long step_stride;
  float eps;
  int add_one;
  __nv_bfloat16* out;        // [rows, ldo]
  long long ldo;
};

struct GemmArgs {
  // problem
  int rows_per_batch;  // rows of A / D per batch element (M when nbatch == 1)
  int nbatch;
  int N;               // valid output columns
  int num_kb;          // K / 64 (k-blocks of BLOCK_K)
  // implicit-GEMM addressing of A (plain GEMM: cpb = num_kb, pad = 0, grouped = 0)
  int cpb;             // k-blocks per tap
  int pad;             // rows of left padding (taps centred)
  int grouped;         // 0: dense. else group size in channels (multiple of 64): A column origin =
                       //    (n0 / grouped) * grouped, so an output tile only reads its own group's inputs
  // epilogue operands
  int act;
  const float* bias;           // [N] or nullptr
  void* out;                   // bf16 or f32 [rows, ldo]
  long long ldo;
  void* out2;                  // secondary bf16 output (EPI_EMBED_DUAL / EPI_SCALE_RESID) or nullptr
  long long ldo2;
  const float* addend;         // f32 [rows, ld_add]
  long long ld_add;
  const float* gate;           // f32 modulation vectors: gate + step*gate_step_stride + (b % gate_nb)*gate_ld
  long long gate_ld;           //   (EPI_SCALE_RESID: plain per-column scale [N] or nullptr)
  int gate_nb;                 // number of distinct modulation rows (1 = shared by all batch elements)
  long long gate_step_stride;  // elements between consecutive ODE steps of the table
  const int* step_ptr;         // device-side ODE step counter (nullptr -> step 0); keeps CUDA graphs replayable
  const float* rope_cos;       // f32 [rows_per_batch, 32]
  const float* rope_sin;
  int rope_cols;               // columns [0, rope_cols) are rotated (q and k)
  int f16_from_col;            // EPI_QKV_ROPE: columns >= this are stored as IEEE f16 instead of bf16 (V for attention)
  const int* seq_lens;         // [nbatch] valid rows per batch element or nullptr (all valid)
  const unsigned char* row_valid;  // [rows] explicit per-row validity (overrides seq_lens) or nullptr
  int mask_rows;               // EPI_GATE_RESID: skip rows t >= seq_len
  int stream_k;                // 2-SM EPI_GATE_RESID / EPI_F32: split the K loops evenly over the SM pairs, partial sums land with f32
                               // vector reductions (EPI_F32: out += A W^T, the caller zeroes or accumulates into out)
  long long* dbg;              // optional [grid, 16] clock64 stamps (tools/kernel_bench.py --trace); nullptr in production
  // 2-SM kernel only: operands whose K dimension runs along the ROWS of the source (the backward-pass GEMMs):
  //   a_mn: A source is [K, M] row-major (D = Asrc^T ...), b_mn: B source is [K, N] row-major (D = ... Bsrc).
  // The TMA box is 64 (MN, contiguous) x 64 (K rows); a 128-wide operand tile is two such 8 KB boxes, consumed through
  // MN-major SW128 descriptors (SBO = 1024: next 8 K rows, LBO = 8192: next 64 MN elements).
  int a_mn, b_mn;
  DropCfg drop;  // EPI_GELU_DROP_*: stateless dropout mask over element index row * N + col (thresh 0: off)
  LnTail ln;     // gemm2 kernel with LNT = 1 only
};

#define ORON_STAMP(slot) do { if (args.dbg) args.dbg[(long long)blockIdx.x * 16 + (slot)] = clock64(); } while (0)
__device__ __forceinline__ long long global_ns() {
  long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}
#define ORON_STAMP_NS(slot) do { if (args.dbg) args.dbg[(long long)blockIdx.x * 16 + (slot)] = global_ns(); } while (0)

constexpr int GEMM_BM = 128;
constexpr int GEMM_BK = 64;
constexpr int GEMM_THREADS = 320;  // warp0 TMA, warp1 MMA + TMEM alloc, warps 2..9 epilogue
constexpr int GEMM_EPI_WARPS = 8;   // two warps per TMEM lane quarter, each owning half of the BN columns

constexpr int EPI_STAGE_LD = 36;                                  // floats per staged row (pad -> conflict-free v4)
constexpr int EPI_STAGE_BYTES_PER_WARP = 32 * EPI_STAGE_LD * 4;   // 4608
constexpr int EPI_STAGE_BYTES = GEMM_EPI_WARPS * EPI_STAGE_BYTES_PER_WARP;

template <int BN>
struct GemmCfg {
  static constexpr int kStages = (BN == 256) ? 3 : (BN == 128 ? 5 : 7);
  static constexpr int kABytes = GEMM_BM * GEMM_BK * 2;
  static constexpr int kBBytes = BN * GEMM_BK * 2;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kTmemCols = (2 * BN < 32) ? 32 : 2 * BN;  // double-buffered accumulator
  static constexpr int kSmemBytes = kStages * kStageBytes + 1024 /*align*/ + 256 /*barriers*/ + EPI_STAGE_BYTES;
};

// ---------------------------------------------------------------------------------------------------
// Epilogue. tcgen05.ld hands every thread one accumulator ROW (TMEM lane == row), so writing straight from
// those registers makes each warp-wide global access touch 32 different cache lines (measured: 7.7k-21k
// cycles per 128x256 tile, the dominant cost of the v1 kernel — profiles/r01_gemm_trace_v1.txt). Each
// epilogue warp therefore transposes 32x32 fp32 blocks through a private, padded smem staging buffer and
// runs the fused op in the "coalesced domain": lane l owns columns 4*(l&7)..+3 of row 4*it + (l>>3), so every
// global load/store of the warp covers 4 rows x 128 contiguous bytes.
// ---------------------------------------------------------------------------------------------------

__device__ __forceinline__ void stage_store_row(uint32_t stage, int lane, const uint32_t (&r)[32]) {
  const uint32_t base = stage + uint32_t(lane) * (EPI_STAGE_LD * 4);
#pragma unroll
  for (int j = 0; j < 8; ++j)
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(base + 16u * j), "r"(r[4 * j]), "r"(r[4 * j + 1]),
                 "r"(r[4 * j + 2]), "r"(r[4 * j + 3])
                 : "memory");
}
__device__ __forceinline__ float4 stage_load4(uint32_t stage, int row, int c4) {
  float4 v;
#ifdef ORON_EXP_NOLDS
  return make_float4(float(row), float(c4), 1.f, 2.f);
#endif
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
               : "r"(stage + uint32_t(row * EPI_STAGE_LD + c4) * 4u)
               : "memory");
  return v;
}
__device__ __forceinline__ float4 ldg4_guard(const float* p, int col, int n) {
  // p points at column `col`; columns >= n read as 0
  if (col + 4 <= n) return __ldg(reinterpret_cast<const float4*>(p));
  float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
  if (col < n) v.x = __ldg(p);
  if (col + 1 < n) v.y = __ldg(p + 1);
  if (col + 2 < n) v.z = __ldg(p + 2);
  return v;
}
__device__ __forceinline__ void st_bf16x4(__nv_bfloat16* p, float4 v, int col, int n) {
#ifdef ORON_EXP_NOSTORE
  if (v.x != 123456.789f) return;
#endif
  if (col + 4 <= n) {
    *reinterpret_cast<uint2*>(p) = make_uint2(pack_bf16x2(v.x, v.y), pack_bf16x2(v.z, v.w));
  } else {
    if (col < n) p[0] = __float2bfloat16(v.x);
    if (col + 1 < n) p[1] = __float2bfloat16(v.y);
    if (col + 2 < n) p[2] = __float2bfloat16(v.z);
  }
}
__device__ __forceinline__ void st_f32x4(float* p, float4 v, int col, int n) {
  if (col + 4 <= n) {
    *reinterpret_cast<float4*>(p) = v;
  } else {
    if (col < n) p[0] = v.x;
    if (col + 1 < n) p[1] = v.y;
    if (col + 2 < n) p[2] = v.z;
  }
}
template <typename F>
__device__ __forceinline__ float4 map4(float4 v, F f) {
  return make_float4(f(v.x), f(v.y), f(v.z), f(v.w));
}

template <int ACT>
__device__ __forceinline__ float4 act4(float4 v) {
  if constexpr (ACT == ACT_GELU_TANH) return map4(v, gelu_tanh_f);
  if constexpr (ACT == ACT_GELU_ERF) return map4(v, gelu_erf_f);
  if constexpr (ACT == ACT_SILU) return map4(v, silu_f);
  return v;
}
__device__ __forceinline__ float4 add4(float4 a, float4 b) { return make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w); }
__device__ __forceinline__ float4 fma4(float4 g, float4 v, float4 a) {
  return make_float4(fmaf(g.x, v.x, a.x), fmaf(g.y, v.y, a.y), fmaf(g.z, v.z, a.z), fmaf(g.w, v.w, a.w));
}
__device__ __forceinline__ uint2 pack4(float4 v) { return make_uint2(pack_bf16x2(v.x, v.y), pack_bf16x2(v.z, v.w)); }
__device__ __forceinline__ uint2 pack4_f16(float4 v) {
  __half2 a = __floats2half2_rn(v.x, v.y), b = __floats2half2_rn(v.z, v.w);
  return make_uint2(*reinterpret_cast<uint32_t*>(&a), *reinterpret_cast<uint32_t*>(&b));
}

__device__ __forceinline__ void red_add4(float* p, float4 v) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__device__ __forceinline__ float4 mul4(float4 a, float4 b) { return make_float4(a.x * b.x, a.y * b.y, a.z * b.z, a.w * b.w); }

// Fast path of one staged 32x32 block: every row and column of the block is in range, so the 8 row-groups are
// loaded up front (ILP) and written with unguarded vector accesses. The epilogue is instruction-latency bound
// (2 warps per scheduler), so straight-line code with independent chains matters more than instruction count.
template <int EPI, int ACT>
__device__ __forceinline__ void epi_block_fast(const GemmArgs& args, const float* sp /*stage + rsub*LD + c4*/,
                                               const long long grow0 /*row0 + rsub*/, const int col, const float4 b4,
                                               const float4 g4, const int t0 /*t_base + rsub*/, const int seq_len) {
  float4 v[8];
#pragma unroll
  for (int it = 0; it < 8; ++it) v[it] = add4(*reinterpret_cast<const float4*>(sp + it * 4 * EPI_STAGE_LD), b4);
  if constexpr (EPI == EPI_BF16) {
    __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(args.out) + grow0 * args.ldo + col;
#pragma unroll
    for (int it = 0; it < 8; ++it) *reinterpret_cast<uint2*>(o + (long long)it * 4 * args.ldo) = pack4(act4<ACT>(v[it]));
  } else if constexpr (EPI == EPI_GELU_DROP_DUAL) {
    __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(args.out) + grow0 * args.ldo + col;
    __nv_bfloat16* o2 = reinterpret_cast<__nv_bfloat16*>(args.out2) + grow0 * args.ldo2 + col;
#pragma unroll
    for (int it = 0; it < 8; ++it) {
      const unsigned long long e = (unsigned long long)(grow0 + 4 * it) * args.N + col;
      const float2 k01 = drop_scale2(args.drop, e), k23 = drop_scale2(args.drop, e + 2);
      *reinterpret_cast<uint2*>(o2 + (long long)it * 4 * args.ldo2) = pack4(v[it]);
      const float4 h = map4(v[it], gelu_tanh_f);
      *reinterpret_cast<uint2*>(o + (long long)it * 4 * args.ldo) =
          pack4(make_float4(h.x * k01.x, h.y * k01.y, h.z * k23.x, h.w * k23.y));
    }
  } else if constexpr (EPI == EPI_GATE_RESID_DUAL) {
    float* o = reinterpret_cast<float*>(args.out) + grow0 * args.ldo + col;
    __nv_bfloat16* o2 = reinterpret_cast<__nv_bfloat16*>(args.out2) + grow0 * args.ldo2 + col;
    const float* ad = args.addend + grow0 * args.ld_add + col;
    float4 x[8];
#pragma unroll
    for (int it = 0; it < 8; ++it) x[it] = __ldg(reinterpret_cast<const float4*>(ad + (long long)it * 4 * args.ld_add));
#pragma unroll
    for (int it = 0; it < 8; ++it) {
      *reinterpret_cast<uint2*>(o2 + (long long)it * 4 * args.ldo2) = pack4(v[it]);
      float4 r = x[it];
      if (!(args.mask_rows && (t0 + 4 * it >= seq_len))) {
        const unsigned long long e = (unsigned long long)(grow0 + 4 * it) * args.N + col;
        const float2 k01 = drop_scale2(args.drop, e), k23 = drop_scale2(args.drop, e + 2);
        r = fma4(g4, make_float4(v[it].x * k01.x, v[it].y * k01.y, v[it].z * k23.x, v[it].w * k23.y), x[it]);
      }
      *reinterpret_cast<float4*>(o + (long long)it * 4 * args.ldo) = r;
    }
  } else if constexpr (EPI == EPI_GELU_DROP_BWD) {
    __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(args.out) + grow0 * args.ldo + col;
    const __nv_bfloat16* ax = reinterpret_cast<const __nv_bfloat16*>(args.out2) + grow0 * args.ldo2 + col;
    uint2 pre[8];
#pragma unroll
    for (int it = 0; it < 8; ++it) pre[it] = __ldg(reinterpret_cast<const uint2*>(ax + (long long)it * 4 * args.ldo2));
#pragma unroll
    for (int it = 0; it < 8; ++it) {
      const unsigned long long e = (unsigned long long)(grow0 + 4 * it) * args.N + col;
      const float2 k01 = drop_scale2(args.drop, e), k23 = drop_scale2(args.drop, e + 2);
      const float p0 = __uint_as_float(pre[it].x << 16), p1 = __uint_as_float(pre[it].x & 0xffff0000u);
      const float p2 = __uint_as_float(pre[it].y << 16), p3 = __uint_as_float(pre[it].y & 0xffff0000u);
      *reinterpret_cast<uint2*>(o + (long long)it * 4 * args.ldo) =
          pack4(make_float4(v[it].x * gelu_tanh_grad(p0) * k01.x, v[it].y * gelu_tanh_grad(p1) * k01.y,
                            v[it].z * gelu_tanh_grad(p2) * k23.x, v[it].w * gelu_tanh_grad(p3) * k23.y));
    }
  } else if constexpr (EPI == EPI_F32) {
    float* o = reinterpret_cast<float*>(args.out) + grow0 * args.ldo + col;
    if (args.addend != nullptr) {
      const float* a = args.addend + grow0 * args.ld_add + col;
      float4 x[8];
#pragma unroll
      for (int it = 0; it < 8; ++it) x[it] = __ldg(reinterpret_cast<const float4*>(a + (long long)it * 4 * args.ld_add));
#pragma unroll
      for (int it = 0; it < 8; ++it) v[it] = add4(v[it], x[it]);
    }
    if (args.stream_k) {  // this CTA holds part of the K sum: out += partial (the caller zeroed or is accumulating into out)
#pragma unroll
      for (int it = 0; it < 8; ++it) red_add4(o + (long long)it * 4 * args.ldo, v[it]);
    } else {
#pragma unroll
      for (int it = 0; it < 8; ++it) *reinterpret_cast<float4*>(o + (long long)it * 4 * args.ldo) = v[it];
    }
  } else if constexpr (EPI == EPI_GATE_RESID) {
    float* o = reinterpret_cast<float*>(args.out) + grow0 * args.ldo + col;
    if (args.stream_k) {
      // this CTA holds only part of the K sum: x += gate * partial, fire-and-forget vector reductions
#pragma unroll
      for (int it = 0; it < 8; ++it) {
        if (args.mask_rows && (t0 + 4 * it >= seq_len)) continue;
        red_add4(o + (long long)it * 4 * args.ldo, mul4(g4, v[it]));
      }
    } else {
      float4 x[8];
#pragma unroll
      for (int it = 0; it < 8; ++it) x[it] = *reinterpret_cast<const float4*>(o + (long long)it * 4 * args.ldo);
#pragma unroll
      for (int it = 0; it < 8; ++it) {
        if (args.mask_rows && (t0 + 4 * it >= seq_len)) continue;
        *reinterpret_cast<float4*>(o + (long long)it * 4 * args.ldo) = fma4(g4, v[it], x[it]);
      }
    }
  } else {
    // epilogues with an addend and a validity mask
    const float* a = args.addend + grow0 * args.ld_add + col;
    float4 x[8];
#pragma unroll
    for (int it = 0; it < 8; ++it) x[it] = *reinterpret_cast<const float4*>(a + (long long)it * 4 * args.ld_add);
    const float4 zero = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int it = 0; it < 8; ++it) {
      const long long grow = grow0 + 4 * it;
      bool valid = (t0 + 4 * it) < seq_len;
      if (args.row_valid != nullptr) valid = args.row_valid[grow] != 0;
      float4 w;
      if constexpr (EPI == EPI_EMBED_DUAL) {
        w = valid ? add4(v[it], x[it]) : zero;
        *reinterpret_cast<float4*>(reinterpret_cast<float*>(args.out) + grow * args.ldo + col) = w;
        *reinterpret_cast<uint2*>(reinterpret_cast<__nv_bfloat16*>(args.out2) + grow * args.ldo2 + col) = pack4(w);
      } else if constexpr (EPI == EPI_MISH_MASK_BF16) {
        w = valid ? map4(v[it], mish_f) : zero;
        *reinterpret_cast<uint2*>(reinterpret_cast<__nv_bfloat16*>(args.out) + grow * args.ldo + col) = pack4(w);
      } else if constexpr (EPI == EPI_MISH_MASK_RESID) {
        w = add4(valid ? map4(v[it], mish_f) : zero, x[it]);
        *reinterpret_cast<float4*>(reinterpret_cast<float*>(args.out) + grow * args.ldo + col) = w;
      } else if constexpr (EPI == EPI_SCALE_RESID) {
        w = valid ? fma4(g4, v[it], x[it]) : zero;
        *reinterpret_cast<float4*>(reinterpret_cast<float*>(args.out) + grow * args.ldo + col) = w;
        if (args.out2 != nullptr)
          *reinterpret_cast<uint2*>(reinterpret_cast<__nv_bfloat16*>(args.out2) + grow * args.ldo2 + col) = pack4(w);
      }
    }
  }
}

// Guarded path (ragged rows / columns): one row-group at a time with per-element bounds checks.
template <int EPI>
__device__ __forceinline__ void epi_block_slow(const GemmArgs& args, const uint32_t stage, const int rsub, const int c4,
                                               const long long row0, const int t_base, const int col, const float4 b4,
                                               const float4 g4, const int seq_len) {
  const int N = args.N;
#pragma unroll 1
  for (int it = 0; it < 8; ++it) {
    const int rr = 4 * it + rsub;
    const int t = t_base + rr;
    if (t >= args.rows_per_batch) continue;
    const long long grow = row0 + rr;
    bool valid = t < seq_len;
    if (args.row_valid != nullptr) valid = args.row_valid[grow] != 0;
    float4 v = add4(stage_load4(stage, rr, c4), b4);
    const float4 zero = make_float4(0.f, 0.f, 0.f, 0.f);
    if constexpr (EPI == EPI_BF16) {
      if (args.act == ACT_GELU_TANH) v = map4(v, gelu_tanh_f);
      else if (args.act == ACT_GELU_ERF) v = map4(v, gelu_erf_f);
      else if (args.act == ACT_SILU) v = map4(v, silu_f);
      st_bf16x4(reinterpret_cast<__nv_bfloat16*>(args.out) + grow * args.ldo + col, v, col, N);
    } else if constexpr (EPI == EPI_GATE_RESID_DUAL) {
      *reinterpret_cast<uint2*>(reinterpret_cast<__nv_bfloat16*>(args.out2) + grow * args.ldo2 + col) = pack4(v);
      float4 r = *reinterpret_cast<const float4*>(args.addend + grow * args.ld_add + col);
      if (!(args.mask_rows && !valid)) {
        const unsigned long long e = (unsigned long long)grow * N + col;
        const float2 k01 = drop_scale2(args.drop, e), k23 = drop_scale2(args.drop, e + 2);
        r = fma4(g4, make_float4(v.x * k01.x, v.y * k01.y, v.z * k23.x, v.w * k23.y), r);
      }
      *reinterpret_cast<float4*>(reinterpret_cast<float*>(args.out) + grow * args.ldo + col) = r;
    } else if constexpr (EPI == EPI_GELU_DROP_DUAL || EPI == EPI_GELU_DROP_BWD) {
      // the host guarantees N % 32 == 0 for these: only ragged ROWS reach this path
      const unsigned long long e = (unsigned long long)grow * N + col;
      const float2 k01 = drop_scale2(args.drop, e), k23 = drop_scale2(args.drop, e + 2);
      __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(args.out) + grow * args.ldo + col;
      if constexpr (EPI == EPI_GELU_DROP_DUAL) {
        *reinterpret_cast<uint2*>(reinterpret_cast<__nv_bfloat16*>(args.out2) + grow * args.ldo2 + col) = pack4(v);
        const float4 h = map4(v, gelu_tanh_f);
        *reinterpret_cast<uint2*>(o) = pack4(make_float4(h.x * k01.x, h.y * k01.y, h.z * k23.x, h.w * k23.y));
      } else {
        const uint2 pre = *reinterpret_cast<const uint2*>(reinterpret_cast<const __nv_bfloat16*>(args.out2) + grow * args.ldo2 + col);
        const float p0 = __uint_as_float(pre.x << 16), p1 = __uint_as_float(pre.x & 0xffff0000u);
        const float p2 = __uint_as_float(pre.y << 16), p3 = __uint_as_float(pre.y & 0xffff0000u);
        *reinterpret_cast<uint2*>(o) = pack4(make_float4(v.x * gelu_tanh_grad(p0) * k01.x, v.y * gelu_tanh_grad(p1) * k01.y,
                                                         v.z * gelu_tanh_grad(p2) * k23.x, v.w * gelu_tanh_grad(p3) * k23.y));
      }
    } else if constexpr (EPI == EPI_F32) {
      if (args.addend != nullptr) v = add4(v, ldg4_guard(args.addend + grow * args.ld_add + col, col, N));
      if (args.stream_k) {
        if (col < N) red_add4(reinterpret_cast<float*>(args.out) + grow * args.ldo + col, v);  // host guarantees N % 4 == 0
      } else {
        st_f32x4(reinterpret_cast<float*>(args.out) + grow * args.ldo + col, v, col, N);
      }
    } else if constexpr (EPI == EPI_GATE_RESID) {
      if (args.mask_rows && !valid) continue;
      float4* o = reinterpret_cast<float4*>(reinterpret_cast<float*>(args.out) + grow * args.ldo + col);
      if (args.stream_k) red_add4(reinterpret_cast<float*>(o), mul4(g4, v));  // host guarantees N % 4 == 0 here
      else *o = fma4(g4, v, *o);
    } else if constexpr (EPI == EPI_EMBED_DUAL) {
      const float4 a = *reinterpret_cast<const float4*>(args.addend + grow * args.ld_add + col);
      v = valid ? add4(v, a) : zero;
      *reinterpret_cast<float4*>(reinterpret_cast<float*>(args.out) + grow * args.ldo + col) = v;
      st_bf16x4(reinterpret_cast<__nv_bfloat16*>(args.out2) + grow * args.ldo2 + col, v, col, N);
    } else if constexpr (EPI == EPI_MISH_MASK_BF16) {
      v = valid ? map4(v, mish_f) : zero;
      st_bf16x4(reinterpret_cast<__nv_bfloat16*>(args.out) + grow * args.ldo + col, v, col, N);
    } else if constexpr (EPI == EPI_MISH_MASK_RESID) {
      const float4 a = *reinterpret_cast<const float4*>(args.addend + grow * args.ld_add + col);
      v = valid ? map4(v, mish_f) : zero;
      *reinterpret_cast<float4*>(reinterpret_cast<float*>(args.out) + grow * args.ldo + col) = add4(v, a);
    } else if constexpr (EPI == EPI_SCALE_RESID) {
      const float4 a = *reinterpret_cast<const float4*>(args.addend + grow * args.ld_add + col);
      v = valid ? fma4(g4, v, a) : zero;
      *reinterpret_cast<float4*>(reinterpret_cast<float*>(args.out) + grow * args.ldo + col) = v;
      if (args.out2 != nullptr) st_bf16x4(reinterpret_cast<__nv_bfloat16*>(args.out2) + grow * args.ldo2 + col, v, col, N);
    }
  }
}

// Drains columns [cbeg, cbeg + HN) of one 128-row accumulator tile. `trow`: TMEM address of this warp's lane
// quarter at column 0 of the tile; `t_base`: row (inside batch element b) of the warp's first lane, or
// >= rows_per_batch for a phantom tile; `stage`: this warp's staging buffer (shared-window address).
// Per-column operands (bias, gate / column scale) are fetched for the whole column range BEFORE the caller's
// accumulator-ready wait would expose their latency: see gemm_epilogue_prefetch.
template <int HN>
struct EpiCols {
  float4 b4[HN / 32];
  float4 g4[HN / 32];
};

template <int BN, int EPI, int HN>
__device__ __forceinline__ void gemm_epilogue_prefetch(const GemmArgs& args, const int b, const int n0, const int cbeg,
                                                       const int lane, EpiCols<HN>& pc, const bool with_bias = true) {
  const int c4 = (lane & 7) * 4;
#pragma unroll
  for (int j = 0; j < HN / 32; ++j) {
    const int col = n0 + cbeg + 32 * j + c4;
    pc.b4[j] = make_float4(0.f, 0.f, 0.f, 0.f);
    pc.g4[j] = make_float4(1.f, 1.f, 1.f, 1.f);
    if (col < args.N) {
      if (with_bias && args.bias != nullptr) pc.b4[j] = ldg4_guard(args.bias + col, col, args.N);
      if constexpr (EPI == EPI_GATE_RESID || EPI == EPI_GATE_RESID_DUAL) {
        const long long step = args.step_ptr ? (long long)__ldg(args.step_ptr) : 0ll;
        pc.g4[j] = __ldg(reinterpret_cast<const float4*>(args.gate + step * args.gate_step_stride +
                                                        (long long)(b % args.gate_nb) * args.gate_ld + col));
      }
      if constexpr (EPI == EPI_SCALE_RESID) {
        if (args.gate != nullptr) pc.g4[j] = __ldg(reinterpret_cast<const float4*>(args.gate + col));
      }
    }
  }
}

// FACT >= 0 (EPI_BF16 only): the activation is a compile-time choice, so only that body is instantiated (ffn_tcgen05.cuh)
// nch: 32-column chunks this warp drains (HN / 32, fewer for the upper half of an uneven split: BN = 224 is 128 + 96)
template <int BN, int EPI, int HN, int FACT = -1>
__device__ __forceinline__ void gemm_epilogue_tile(const GemmArgs& args, const uint32_t trow, const int b,
                                                   const int t_base, const int n0, const int cbeg,
                                                   const uint32_t stage, const int lane, const EpiCols<HN>& pc,
                                                   const int nch = HN / 32) {
  const int rsub = lane >> 3;        // row inside a 4-row group
  const int c4 = (lane & 7) * 4;     // first of the 4 columns this lane owns inside a 32-column chunk
  const int seq_len = args.seq_lens ? args.seq_lens[b] : args.rows_per_batch;
  const long long row0 = (long long)b * args.rows_per_batch + t_base;
  const int N = args.N;
  const bool rows_full = t_base + 32 <= args.rows_per_batch;  // warp-uniform
  const float* sp = reinterpret_cast<const float*>(__cvta_shared_to_generic(stage)) + rsub * EPI_STAGE_LD + c4;

  if constexpr (EPI == EPI_QKV_ROPE) {
    __nv_bfloat16* out = reinterpret_cast<__nv_bfloat16*>(args.out);
#pragma unroll
    for (int j = 0; j < HN / 64; ++j) {   // one 64-wide head at a time
      const int c0 = cbeg + 64 * j;
      uint32_t r[32];
      const int col = n0 + c0 + c4;                     // this lane's columns in the first half of the head
      const bool live = (n0 + c0) < N;                  // warp-uniform
      tmem_ld_32x32(trow + c0, r);
      tmem_wait_ld();
      float4 xa[8];
      if (live) {
        stage_store_row(stage, lane, r);
        __syncwarp();
#pragma unroll
        for (int it = 0; it < 8; ++it) xa[it] = add4(*reinterpret_cast<const float4*>(sp + it * 4 * EPI_STAGE_LD), pc.b4[2 * j]);
        __syncwarp();
      }
      tmem_ld_32x32(trow + c0 + 32, r);
      tmem_wait_ld();
      if (live) {
        stage_store_row(stage, lane, r);
        __syncwarp();
        const bool rot = (n0 + c0) < args.rope_cols;
        const bool as_f16 = (n0 + c0) >= args.f16_from_col;  // warp-uniform
        float4 xb[8];
#pragma unroll
        for (int it = 0; it < 8; ++it) xb[it] = add4(*reinterpret_cast<const float4*>(sp + it * 4 * EPI_STAGE_LD), pc.b4[2 * j + 1]);
        if (rot) {
          // rotate_half: out[i] = x[i] cos - x[i+32] sin ; out[i+32] = x[i+32] cos + x[i] sin
#pragma unroll
          for (int h = 0; h < 2; ++h) {   // two batches of 4 row-groups keep the cos/sin registers bounded
            float4 cs[4], sn[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const int t = min(t_base + 4 * (4 * h + i) + rsub, args.rows_per_batch - 1);
              cs[i] = __ldg(reinterpret_cast<const float4*>(args.rope_cos + (long long)t * 32 + c4));
              sn[i] = __ldg(reinterpret_cast<const float4*>(args.rope_sin + (long long)t * 32 + c4));
            }
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const int it = 4 * h + i;
              const float4 x1 = xa[it], x2 = xb[it], c = cs[i], s_ = sn[i];
              xa[it] = make_float4(x1.x * c.x - x2.x * s_.x, x1.y * c.y - x2.y * s_.y, x1.z * c.z - x2.z * s_.z, x1.w * c.w - x2.w * s_.w);
              xb[it] = make_float4(x2.x * c.x + x1.x * s_.x, x2.y * c.y + x1.y * s_.y, x2.z * c.z + x1.z * s_.z, x2.w * c.w + x1.w * s_.w);
            }
          }
        }
        __nv_bfloat16* o = out + (row0 + rsub) * args.ldo + col;
#pragma unroll
        for (int it = 0; it < 8; ++it) {
          if (rows_full || (t_base + 4 * it + rsub) < args.rows_per_batch) {
            *reinterpret_cast<uint2*>(o + (long long)it * 4 * args.ldo) = as_f16 ? pack4_f16(xa[it]) : pack4(xa[it]);
            *reinterpret_cast<uint2*>(o + (long long)it * 4 * args.ldo + 32) = as_f16 ? pack4_f16(xb[it]) : pack4(xb[it]);
          }
        }
        __syncwarp();
      }
    }
  } else {
    // NOT unrolled: one 32-column block body is ~0.5 k instructions per activation; unrolled over the 4 blocks of
    // a 128-column half tile the straight-line epilogue outgrew the instruction cache (ncu on the FFN-up GEMM:
    // 74 % icc hit rate, stall_no_instruction second only to the accumulator wait). The prefetched per-column
    // operands are picked out of their registers with selects.
#pragma unroll 1
    for (int j = 0; j < nch; ++j) {
      const int c0 = cbeg + 32 * j;
      float4 b4 = pc.b4[0], g4 = pc.g4[0];
#pragma unroll
      for (int q = 1; q < HN / 32; ++q)
        if (j == q) { b4 = pc.b4[q]; g4 = pc.g4[q]; }
      uint32_t r[32];
      tmem_ld_32x32(trow + c0, r);
      tmem_wait_ld();
      if (n0 + c0 >= N) continue;  // warp-uniform
      stage_store_row(stage, lane, r);
      __syncwarp();
      const int col = n0 + c0 + c4;
      if (rows_full && (n0 + c0 + 32 <= N)) {
        if constexpr (EPI == EPI_BF16 && FACT >= 0) {
          epi_block_fast<EPI, FACT>(args, sp, row0 + rsub, col, b4, g4, t_base + rsub, seq_len);
        } else if constexpr (EPI == EPI_BF16) {
          switch (args.act) {  // hoisted out of the row loop: one straight-line body per activation
            case ACT_GELU_TANH: epi_block_fast<EPI, ACT_GELU_TANH>(args, sp, row0 + rsub, col, b4, g4, t_base + rsub, seq_len); break;
            case ACT_GELU_ERF: epi_block_fast<EPI, ACT_GELU_ERF>(args, sp, row0 + rsub, col, b4, g4, t_base + rsub, seq_len); break;
            case ACT_SILU: epi_block_fast<EPI, ACT_SILU>(args, sp, row0 + rsub, col, b4, g4, t_base + rsub, seq_len); break;
            default: epi_block_fast<EPI, ACT_NONE>(args, sp, row0 + rsub, col, b4, g4, t_base + rsub, seq_len); break;
          }
        } else {
          epi_block_fast<EPI, ACT_NONE>(args, sp, row0 + rsub, col, b4, g4, t_base + rsub, seq_len);
        }
      } else {
        epi_block_slow<EPI>(args, stage, rsub, c4, row0, t_base, col, b4, g4, seq_len);
      }
      __syncwarp();
    }
  }
}

template <int BN, int EPI>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_bf16_tcgen05_kernel(const __grid_constant__ CUtensorMap tmA,
                         const __grid_constant__ CUtensorMap tmB, const GemmArgs args) {
  using Cfg = GemmCfg<BN>;
  constexpr int kStages = Cfg::kStages;
  extern __shared__ uint8_t smem_raw[];
  // SWIZZLE_128B tiles need 1024-byte alignment
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t bar_base = smem_base + kStages * Cfg::kStageBytes;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (kStages + s); };
  auto tfull_bar = [&](int s) { return bar_base + 8u * (2 * kStages + s); };
  auto tempty_bar = [&](int s) { return bar_base + 8u * (2 * kStages + 2 + s); };
  const uint32_t tmem_slot = bar_base + 8u * (2 * kStages + 4);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  const int tiles_m_pb = (args.rows_per_batch + GEMM_BM - 1) / GEMM_BM;
  const int tiles_m = tiles_m_pb * args.nbatch;
  const int tiles_n = (args.N + BN - 1) / BN;
  const int num_tiles = tiles_m * tiles_n;
  const int num_kb = args.num_kb;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    for (int s = 0; s < kStages; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(tfull_bar(s), 1);
      mbar_init(tempty_bar(s), GEMM_EPI_WARPS);  // one arrive per epilogue warp
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
  pdl_wait();  // everything above overlapped the previous kernel's tail; its outputs are visible from here on

  if (warp == 0) {
    // ===================== TMA producer (one thread) =====================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int m_tile = tile % tiles_m;
        const int n_tile = tile / tiles_m;
        const int b = m_tile / tiles_m_pb;
        const int t0 = (m_tile % tiles_m_pb) * GEMM_BM;
        const int n0 = n_tile * BN;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(empty_bar(stage), phase ^ 1u, 1);
          mbar_arrive_expect_tx(full_bar(stage), Cfg::kStageBytes);
          const uint32_t sa = smem_base + stage * Cfg::kStageBytes;
          const uint32_t sb = sa + Cfg::kABytes;
          const int a_col = (args.grouped ? (n0 / args.grouped) * args.grouped : 0) + (kb % args.cpb) * GEMM_BK;
          const int a_row = t0 + kb / args.cpb - args.pad;
          tma_load_3d(sa, &tmA, full_bar(stage), a_col, a_row, b);
          tma_load_2d(sb, &tmB, full_bar(stage), kb * GEMM_BK, n0);
          if (++stage == kStages) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (one thread) =====================
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc_bf16(GEMM_BM, BN);
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
        const int as = it & 1;
        const uint32_t aphase = (it >> 1) & 1u;
        mbar_wait(tempty_bar(as), aphase ^ 1u, 2);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + uint32_t(as * BN);
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(full_bar(stage), phase, 3);
          tc_fence_after();
          const uint32_t sa = smem_base + stage * Cfg::kStageBytes;
          const uint32_t sb = sa + Cfg::kABytes;
          const uint64_t adesc = make_smem_desc_sw128(sa, 16, 1024);
          const uint64_t bdesc = make_smem_desc_sw128(sb, 16, 1024);
#pragma unroll
          for (int k = 0; k < GEMM_BK / 16; ++k) {
            // advance 16 K-elements = 32 bytes inside the 128 B swizzle span (encoded >> 4)
            umma_bf16_ss(tmem_d, adesc + uint64_t(2 * k), bdesc + uint64_t(2 * k), idesc,
                         (kb | k) != 0 ? 1u : 0u);
          }
          umma_commit(empty_bar(stage));  // frees the smem slot once these MMAs retire
          if (kb == num_kb - 1) umma_commit(tfull_bar(as));
          if (++stage == kStages) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else {
    // ===================== epilogue warps (2..9) =====================
    const int q = warp & 3;               // TMEM lane quarter this warp may access
    const int chalf = (warp - 2) >> 2;    // which half of the tile's columns this warp drains
    constexpr int HN = BN / 2;
    const int cbeg = chalf * HN;
    int it = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
      const int m_tile = tile % tiles_m;
      const int n_tile = tile / tiles_m;
      const int b = m_tile / tiles_m_pb;
      const int t_base = (m_tile % tiles_m_pb) * GEMM_BM + q * 32;  // first row of this warp inside the batch element
      const int n0 = n_tile * BN;
      const int as = it & 1;
      const uint32_t aphase = (it >> 1) & 1u;
      EpiCols<HN> pc;
      gemm_epilogue_prefetch<BN, EPI, HN>(args, b, n0, cbeg, lane, pc);
      mbar_wait(tfull_bar(as), aphase, 4);
      tc_fence_after();

      const uint32_t trow = tmem_base + (uint32_t(q * 32) << 16) + uint32_t(as * BN);
      gemm_epilogue_tile<BN, EPI, HN>(args, trow, b, t_base, n0, cbeg,
                                      bar_base + 256u + uint32_t(warp - 2) * EPI_STAGE_BYTES_PER_WARP, lane, pc);
      // accumulator drained -> hand the TMEM buffer back to the MMA warp
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tempty_bar(as));
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, Cfg::kTmemCols);
  }
}

// =====================================================================================================
// 2-SM variant: a cluster of two CTAs (one SM pair) computes a 256 x BN tile with tcgen05.mma.cta_group::2.
// Each CTA stages its own 128 A rows and HALF of the B tile (BN/2 weight rows), so per MMA the tensor core
// reads 2/3 (BN=256) of the shared-memory bytes of the 1-SM kernel and TMA writes 2/3 as many: the 1-SM
// kernel is shared-memory-bandwidth bound (profiles/r01_ncu_full_gemm_v1.txt: tensor pipe 33-40 % active
// with L2 and DRAM far from their limits). Only the leader CTA issues MMAs; barriers:
//   full[s]   (leader)  <- TMA bytes of both CTAs             empty[s] (both) <- multicast tcgen05.commit
//   tfull[a]  (both)    <- multicast commit of the last k-block   tempty[a] (leader) <- 16 epilogue warps
// The two CTAs' row blocks are consecutive 128-row m-tiles (2*pm, 2*pm+1); they need not be adjacent in
// memory (each CTA addresses its own (batch, t0) through the 3-D A map).
// =====================================================================================================
// Work walker of the 2-SM kernel. Plain: whole tiles, strided over the SM pairs. stream_k: the flat list of
// (tile, k-block) units is cut into equal contiguous shares, so a pair runs up to two partial tiles ("segments")
// and every pair is busy for the same time whatever the tile count (44 tiles on 74 pairs for the N=1024 GEMMs).
struct SegWalk {
  long long u, u_end;
  int nkb, step;
  bool sk;
  int tile, kb0, kb1;
  __device__ SegWalk(bool stream_k, int pair_id, int num_pairs, int num_tiles, int num_kb) {
    sk = stream_k; nkb = num_kb; step = num_pairs;
    if (sk) {
      const long long U = (long long)num_tiles * num_kb;
      u = (U * pair_id) / num_pairs;
      u_end = (U * (pair_id + 1)) / num_pairs;
    } else {
      u = pair_id;
      u_end = num_tiles;
    }
    tile = 0; kb0 = 0; kb1 = 0;
  }
  __device__ bool next() {
    if (u >= u_end) return false;
    if (sk) {
      tile = int(u / nkb);
      kb0 = int(u - (long long)tile * nkb);
      const long long n = min((long long)(nkb - kb0), u_end - u);
      kb1 = kb0 + int(n);
      u += n;
    } else {
      tile = int(u); kb0 = 0; kb1 = nkb;
      u += step;
    }
    return true;
  }
};

// ---- LayerNorm tail of the 2-SM kernel (LNT = 1) ----
// Every epilogue warp publishes the k-blocks it has folded into the residual rows of its m-tile pair (fence, then one
// relaxed add per warp); once all n-tiles x k-blocks x 16 warps of a pair have arrived its 256 rows are final. The rows
// are dealt round-robin to all epilogue warps of the grid (every CTA is resident: grid <= SM count, one CTA per SM, and
// nobody waits before its own GEMM work is done, so the waits cannot deadlock; they are bounded all the same).
__device__ __forceinline__ int ld_acquire_gpu(const int* p) {
  int v;
  asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
// counters sit LN_CNT_STRIDE ints (one 128-byte line) apart; ONE lane per warp polls (37 k threads polling a dozen words would
// queue up in front of the L2 slices that also have to take the arrivals), the warp barrier hands the acquire to the rest
constexpr int LN_CNT_STRIDE = 32;
__device__ __forceinline__ void ln_tail_wait(const int* cnt, int target, int lane) {
  if (lane == 0 && ld_acquire_gpu(cnt) < target) {
    const long long t0 = clock64();
    while (ld_acquire_gpu(cnt) < target) {
      __nanosleep(100);
      if (clock64() - t0 > ORON_WATCHDOG_CYCLES) __trap();
    }
  }
  __syncwarp();
}
// One warp, up to LN_TAIL_NR rows at a time: with eight warps per SM the tail is bound by memory LATENCY, so the rows of a batch
// are fetched together -- with 16-byte cp.async.cg copies (L2, past L1: other SMs wrote x during this launch) into the warp's
// slice of the operand ring, which is idle once the CTA's last accumulator has been drained (the register file has no room
// for a second row: 168 registers per thread with ten warps per CTA). Every lane copies exactly the 16-byte pieces it later
// reads, so its own wait_group is all the synchronisation needed. Per row the arithmetic of ln_modulate_kernel<C>
// (rowwise.cuh), statement for statement: the tail and the separate launch agree bit for bit. C = 128 * V4 <= 1024.
constexpr int LN_TAIL_NR = 4;
constexpr int LN_TAIL_WARP_BYTES = LN_TAIL_NR * 4096;
template <int V4>
__device__ __forceinline__ void ln_tail_rows_t(const GemmArgs& a, const long long (&gr)[LN_TAIL_NR], const long long (&moff)[LN_TAIL_NR],
                                               const int nr, const int lane, const uint32_t buf) {
  constexpr int C = 128 * V4, NR = LN_TAIL_NR;
#pragma unroll
  for (int r = 0; r < NR; ++r)
    if (r < nr) {
      const float4* xr = reinterpret_cast<const float4*>(reinterpret_cast<const float*>(a.out) + gr[r] * a.ldo);
#pragma unroll
      for (int i = 0; i < V4; ++i)
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(buf + uint32_t(r * 4096 + (lane + 32 * i) * 16)), "l"(xr + lane + 32 * i)
                     : "memory");
    }
  asm volatile("cp.async.commit_group;" ::: "memory");
  asm volatile("cp.async.wait_group 0;" ::: "memory");
  const float one = a.ln.add_one ? 1.f : 0.f;
#pragma unroll
  for (int r = 0; r < NR; ++r) {
    if (r >= nr) break;
    float4 v[V4];
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < V4; ++i) {
      asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];"
                   : "=f"(v[i].x), "=f"(v[i].y), "=f"(v[i].z), "=f"(v[i].w)
                   : "r"(buf + uint32_t(r * 4096 + (lane + 32 * i) * 16))
                   : "memory");
      s += v[i].x + v[i].y + v[i].z + v[i].w;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    const float mean = s * (1.0f / C);
    float ss = 0.f;
#pragma unroll
    for (int i = 0; i < V4; ++i) {
      v[i].x -= mean; v[i].y -= mean; v[i].z -= mean; v[i].w -= mean;
      ss += v[i].x * v[i].x + v[i].y * v[i].y + v[i].z * v[i].z + v[i].w * v[i].w;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
    const float rstd = rsqrtf(ss * (1.0f / C) + a.ln.eps);
    const float4* sc = reinterpret_cast<const float4*>(a.ln.scale + moff[r]);
    const float4* sh = a.ln.shift ? reinterpret_cast<const float4*>(a.ln.shift + moff[r]) : nullptr;
    uint2* orow = reinterpret_cast<uint2*>(a.ln.out + gr[r] * a.ln.ldo);
#pragma unroll
    for (int i = 0; i < V4; ++i) {
      const float4 g = __ldg(sc + lane + 32 * i);
      float4 h = make_float4(0.f, 0.f, 0.f, 0.f);
      if (sh) h = __ldg(sh + lane + 32 * i);
      float4 y;
      y.x = v[i].x * rstd * (one + g.x) + h.x;
      y.y = v[i].y * rstd * (one + g.y) + h.y;
      y.z = v[i].z * rstd * (one + g.z) + h.z;
      y.w = v[i].w * rstd * (one + g.w) + h.w;
      orow[lane + 32 * i] = make_uint2(pack_bf16x2(y.x, y.y), pack_bf16x2(y.z, y.w));
    }
  }
}
__device__ __forceinline__ void ln_tail_rows(const GemmArgs& a, const long long (&gr)[LN_TAIL_NR], const long long (&moff)[LN_TAIL_NR],
                                             const int nr, const int lane, const uint32_t buf) {
  switch (a.N >> 7) {  // warp-uniform
    case 8: ln_tail_rows_t<8>(a, gr, moff, nr, lane, buf); break;
    case 6: ln_tail_rows_t<6>(a, gr, moff, nr, lane, buf); break;
    case 4: ln_tail_rows_t<4>(a, gr, moff, nr, lane, buf); break;
    case 2: ln_tail_rows_t<2>(a, gr, moff, nr, lane, buf); break;
    case 1: ln_tail_rows_t<1>(a, gr, moff, nr, lane, buf); break;
    default: break;  // the host admits only these widths (the ones ln_modulate has kernels for)
  }
}

template <int BN>
struct Gemm2Cfg {
  static constexpr int kABytes = GEMM_BM * GEMM_BK * 2;
  static constexpr int kBBytes = (BN / 2) * GEMM_BK * 2;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kStages = (BN == 256 || BN == 224) ? 5 : (BN == 192 ? 6 : 7);
  static constexpr int kTmemCols = (2 * BN <= 256) ? 256 : 512;  // double-buffered accumulator; allocations are powers of two
  static constexpr int kSmemBytes = kStages * kStageBytes + 1024 + 256 + EPI_STAGE_BYTES;
};

// MNM: bit 0 = A is MN-major, bit 1 = B is MN-major (compile-time: the K-major instantiations used by the inference path keep
// constant descriptors / instruction descriptor in the single-thread MMA issue loop)
template <int BN, int EPI, int MNM = 0, int LNT = 0>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(GEMM_THREADS, 1)
gemm2_bf16_tcgen05_kernel(const __grid_constant__ CUtensorMap tmA,
                          const __grid_constant__ CUtensorMap tmB, const GemmArgs args) {
  using Cfg = Gemm2Cfg<BN>;
  constexpr int kStages = Cfg::kStages;
  static_assert(LNT == 0 || kStages * Cfg::kStageBytes >= GEMM_EPI_WARPS * LN_TAIL_WARP_BYTES, "LN tail row buffers live in the operand ring");
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t bar_base = smem_base + kStages * Cfg::kStageBytes;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (kStages + s); };
  auto tfull_bar = [&](int s) { return bar_base + 8u * (2 * kStages + s); };
  auto tempty_bar = [&](int s) { return bar_base + 8u * (2 * kStages + 2 + s); };
  const uint32_t tmem_slot = bar_base + 8u * (2 * kStages + 4);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int rank = int(cluster_ctarank());
  const bool leader = rank == 0;
  const int pair_id = blockIdx.x >> 1;
  const int num_pairs = gridDim.x >> 1;

  const int tiles_m_pb = (args.rows_per_batch + GEMM_BM - 1) / GEMM_BM;
  const int tiles_m = tiles_m_pb * args.nbatch;
  const int tiles_mp = (tiles_m + 1) / 2;
  const int tiles_n = (args.N + BN - 1) / BN;
  const int num_tiles = tiles_mp * tiles_n;
  const int num_kb = args.num_kb;
  const bool sk = (EPI == EPI_GATE_RESID || EPI == EPI_F32) && args.stream_k != 0;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    for (int s = 0; s < kStages; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(tfull_bar(s), 1);
      mbar_init(tempty_bar(s), 2 * GEMM_EPI_WARPS);
    }
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc_2sm(tmem_slot, Cfg::kTmemCols);
    tmem_relinquish_2sm();
  }
  pdl_launch_dependents();
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));
  pdl_wait();  // everything above overlapped the previous kernel's tail; its outputs are visible from here on
  if (threadIdx.x == 0) ORON_STAMP(0);

  if (warp == 0) {
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      bool first_seg = true;
      for (SegWalk w(sk, pair_id, num_pairs, num_tiles, num_kb); w.next(); first_seg = false) {
        const int tile = w.tile;
        const int m_tile = 2 * (tile % tiles_mp) + rank;
        const int n_tile = tile / tiles_mp;
        const int b = m_tile / tiles_m_pb;  // a phantom m-tile (odd tile count) lands past the last batch element: zero fill
        const int t0 = (m_tile % tiles_m_pb) * GEMM_BM;
        const int n0 = n_tile * BN;
        for (int kb = w.kb0; kb < w.kb1; ++kb) {
          mbar_wait(empty_bar(stage), phase ^ 1u, 21);
          if (leader) mbar_arrive_expect_tx(full_bar(stage), 2 * Cfg::kStageBytes);
          const uint32_t sa = smem_base + stage * Cfg::kStageBytes;
          const uint32_t sb = sa + Cfg::kABytes;
          const int a_col = (args.grouped ? (n0 / args.grouped) * args.grouped : 0) + (kb % args.cpb) * GEMM_BK;
          const int a_row = t0 + kb / args.cpb - args.pad;
          if constexpr ((MNM & 1) != 0) {
#pragma unroll
            for (int i = 0; i < GEMM_BM / 64; ++i)
              tma_load_2d_2sm(sa + i * 8192, &tmA, full_bar(stage), m_tile * GEMM_BM + 64 * i, kb * GEMM_BK);
          } else {
            tma_load_3d_2sm(sa, &tmA, full_bar(stage), a_col, a_row, b);
          }
          if constexpr ((MNM & 2) != 0) {
#pragma unroll
            for (int i = 0; i < BN / 128; ++i)
              tma_load_2d_2sm(sb + i * 8192, &tmB, full_bar(stage), n0 + rank * (BN / 2) + 64 * i, kb * GEMM_BK);
          } else {
            tma_load_2d_2sm(sb, &tmB, full_bar(stage), kb * GEMM_BK, n0 + rank * (BN / 2));
          }
          if (kb == w.kb0 && first_seg) ORON_STAMP(1);
          if (kb == w.kb1 - 1 && first_seg) ORON_STAMP(2);
          if (++stage == kStages) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    if (leader && lane == 0) {
      constexpr uint32_t idesc = make_idesc_bf16(2 * GEMM_BM, BN, MNM & 1, (MNM >> 1) & 1);
      // K-major: 16 K elements = 32 bytes inside the 128-byte rows; MN-major: 16 K rows = 2048 bytes (>> 4)
      constexpr uint64_t a_kstep = (MNM & 1) ? 128u : 2u, b_kstep = (MNM & 2) ? 128u : 2u;
      constexpr uint32_t a_lbo = (MNM & 1) ? 8192u : 16u, b_lbo = (MNM & 2) ? 8192u : 16u;
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      for (SegWalk w(sk, pair_id, num_pairs, num_tiles, num_kb); w.next(); ++it) {
        const int as = it & 1;
        const uint32_t aphase = (it >> 1) & 1u;
        mbar_wait(tempty_bar(as), aphase ^ 1u, 22);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + uint32_t(as * BN);
        for (int kb = w.kb0; kb < w.kb1; ++kb) {
          mbar_wait(full_bar(stage), phase, 23);
          tc_fence_after();
          if (kb == w.kb0 && it == 0) ORON_STAMP(3);
          const uint32_t sa = smem_base + stage * Cfg::kStageBytes;
          const uint32_t sb = sa + Cfg::kABytes;
          const uint64_t adesc = make_smem_desc_sw128(sa, a_lbo, 1024);
          const uint64_t bdesc = make_smem_desc_sw128(sb, b_lbo, 1024);
#pragma unroll
          for (int k = 0; k < GEMM_BK / 16; ++k)
            umma_bf16_ss_2sm(tmem_d, adesc + a_kstep * uint64_t(k), bdesc + b_kstep * uint64_t(k), idesc, (kb != w.kb0 || k != 0) ? 1u : 0u);
          umma_commit_2sm(empty_bar(stage), 3);
          if (kb == w.kb1 - 1) { umma_commit_2sm(tfull_bar(as), 3); if (it < 2) ORON_STAMP(4 + it); }
          if (++stage == kStages) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else {
    const int q = warp & 3;
    const int chalf = (warp - 2) >> 2;
    // the tile's columns are split between the two warps of a lane quarter in 32-column chunks: evenly, except BN = 224 = 128 + 96
    constexpr int HN = (BN == 224) ? 128 : BN / 2;
    static_assert(EPI != EPI_QKV_ROPE || BN != 224, "the RoPE epilogue walks whole 64-wide heads per half");
    const int cbeg = chalf * HN;
    const int nch = (BN == 224 && chalf == 1) ? 3 : HN / 32;
    int it = 0;
    for (SegWalk w(sk, pair_id, num_pairs, num_tiles, num_kb); w.next(); ++it) {
      const int tile = w.tile;
      const int m_tile = 2 * (tile % tiles_mp) + rank;
      const int n_tile = tile / tiles_mp;
      const int as = it & 1;
      const uint32_t aphase = (it >> 1) & 1u;
      const int b = m_tile < tiles_m ? m_tile / tiles_m_pb : 0;
      EpiCols<HN> pc;
      gemm_epilogue_prefetch<BN, EPI, HN>(args, b, n_tile * BN, cbeg, lane, pc, w.kb0 == 0);  // the bias rides with the first k-block
      mbar_wait(tfull_bar(as), aphase, 24);
      tc_fence_after();
      if (threadIdx.x == 64 && it < 2) ORON_STAMP(6 + 2 * it);
      // phantom tile: push the row index out of range so nothing is stored
      const int t_base = m_tile < tiles_m ? (m_tile % tiles_m_pb) * GEMM_BM + q * 32 : args.rows_per_batch;
      const uint32_t trow = tmem_base + (uint32_t(q * 32) << 16) + uint32_t(as * BN);
      gemm_epilogue_tile<BN, EPI, HN>(args, trow, b, t_base, n_tile * BN, cbeg,
                                      bar_base + 256u + uint32_t(warp - 2) * EPI_STAGE_BYTES_PER_WARP, lane, pc, nch);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_leader(tempty_bar(as));
      if (threadIdx.x == 64 && it < 2) ORON_STAMP(7 + 2 * it);
      if constexpr (LNT != 0) {  // this warp's share of k-blocks [kb0, kb1) of the tile is in the residual rows
        __threadfence();
        __syncwarp();
        if (lane == 0) atomicAdd(args.ln.counters + (tile % tiles_mp) * LN_CNT_STRIDE, w.kb1 - w.kb0);
      }
    }
    if constexpr (LNT != 0) {
      const long long rows = (long long)args.rows_per_batch * args.nbatch;
      const int total_warps = int(gridDim.x) * GEMM_EPI_WARPS;
      const int target = tiles_n * num_kb * 2 * GEMM_EPI_WARPS;
      const long long step = args.step_ptr ? (long long)__ldg(args.step_ptr) : 0ll;
      for (long long g0 = (long long)blockIdx.x * GEMM_EPI_WARPS + (warp - 2); g0 < rows; g0 += (long long)LN_TAIL_NR * total_warps) {
        long long gr[LN_TAIL_NR], moff[LN_TAIL_NR];
        int nr = 0;
#pragma unroll
        for (int r = 0; r < LN_TAIL_NR; ++r) {
          gr[r] = g0 + (long long)r * total_warps;
          moff[r] = 0;
          if (gr[r] < rows) {
            nr = r + 1;
            const int b = int(gr[r] / args.rows_per_batch);
            const int t = int(gr[r] - (long long)b * args.rows_per_batch);
            moff[r] = step * args.ln.step_stride + (long long)(b % args.ln.mod_nb) * args.ln.mod_ld;
            ln_tail_wait(args.ln.counters + ((b * tiles_m_pb + t / GEMM_BM) >> 1) * LN_CNT_STRIDE, target, lane);
          }
        }
        ln_tail_rows(args, gr, moff, nr, lane, smem_base + uint32_t(warp - 2) * LN_TAIL_WARP_BYTES);
      }
      __syncwarp();
      if (lane == 0) {  // the last warp out of the whole grid re-arms the counters for the next launch
        if (atomicAdd(args.ln.counters + tiles_mp * LN_CNT_STRIDE, 1) == total_warps - 1) {
          for (int j = 0; j <= tiles_mp; ++j) args.ln.counters[j * LN_CNT_STRIDE] = 0;
        }
      }
    }
  }

  tc_fence_before();
  cluster_sync_all();
  if (threadIdx.x == 0) ORON_STAMP(10);
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc_2sm(tmem_base, Cfg::kTmemCols);
  }
}

}  // namespace oron
