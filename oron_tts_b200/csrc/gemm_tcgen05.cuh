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
};

enum GemmAct : int { ACT_NONE = 0, ACT_GELU_TANH = 1, ACT_GELU_ERF = 2, ACT_SILU = 3 };

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
  const int* seq_lens;         // [nbatch] valid rows per batch element or nullptr (all valid)
  const unsigned char* row_valid;  // [rows] explicit per-row validity (overrides seq_lens) or nullptr
  int mask_rows;               // EPI_GATE_RESID: skip rows t >= seq_len
};

constexpr int GEMM_BM = 128;
constexpr int GEMM_BK = 64;
constexpr int GEMM_THREADS = 320;  // warp0 TMA, warp1 MMA + TMEM alloc, warps 2..9 epilogue
constexpr int GEMM_EPI_WARPS = 8;   // two warps per TMEM lane quarter, each owning half of the BN columns

template <int BN>
struct GemmCfg {
  static constexpr int kStages = (BN == 256) ? 4 : (BN == 128 ? 6 : 8);
  static constexpr int kABytes = GEMM_BM * GEMM_BK * 2;
  static constexpr int kBBytes = BN * GEMM_BK * 2;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kTmemCols = (2 * BN < 32) ? 32 : 2 * BN;  // double-buffered accumulator
  static constexpr int kSmemBytes = kStages * kStageBytes + 1024 /*align*/ + 256 /*barriers*/;
};

// Drains columns [cbeg, cbeg + HN) of one 128-row accumulator tile (TMEM address `trow` = this warp's lane
// quarter, column 0 of the tile) through the fused epilogue. Thread = one output row t of batch element b.
template <int BN, int EPI, int HN>
__device__ __forceinline__ void gemm_epilogue_tile(const GemmArgs& args, const uint32_t trow, const int b,
                                                   const int t, const int n0, const int cbeg) {
  const bool in_range = t < args.rows_per_batch;
  const int seq_len = args.seq_lens ? args.seq_lens[b] : args.rows_per_batch;
  const long long grow = (long long)b * args.rows_per_batch + t;
  bool valid = in_range && (t < seq_len);
  if (args.row_valid != nullptr) valid = in_range && (args.row_valid[in_range ? grow : 0] != 0);

  if constexpr (EPI == EPI_QKV_ROPE) {
    // BN is a multiple of 64: each 64-column group is one head.
    float cs[32], sn[32];
    if (in_range) {
      const float4* c4 = reinterpret_cast<const float4*>(args.rope_cos + (long long)t * 32);
      const float4* s4 = reinterpret_cast<const float4*>(args.rope_sin + (long long)t * 32);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float4 c = __ldg(c4 + i), s = __ldg(s4 + i);
        cs[4 * i] = c.x; cs[4 * i + 1] = c.y; cs[4 * i + 2] = c.z; cs[4 * i + 3] = c.w;
        sn[4 * i] = s.x; sn[4 * i + 1] = s.y; sn[4 * i + 2] = s.z; sn[4 * i + 3] = s.w;
      }
    } else {
#pragma unroll
      for (int i = 0; i < 32; ++i) { cs[i] = 1.f; sn[i] = 0.f; }
    }
    __nv_bfloat16* out = reinterpret_cast<__nv_bfloat16*>(args.out) + grow * args.ldo;
#pragma unroll 1
    for (int c0 = cbeg; c0 < cbeg + HN; c0 += 64) {
      uint32_t ra[32], rb[32];
      tmem_ld_32x32(trow + c0, ra);
      tmem_ld_32x32(trow + c0 + 32, rb);
      tmem_wait_ld();
      const int col = n0 + c0;
      if (in_range && col < args.N) {
        const bool rot = col < args.rope_cols;
        uint32_t pa[16], pb[16];
#pragma unroll
        for (int i = 0; i < 32; i += 2) {
          float x1a = __uint_as_float(ra[i]) + __ldg(args.bias + col + i);
          float x1b = __uint_as_float(ra[i + 1]) + __ldg(args.bias + col + i + 1);
          float x2a = __uint_as_float(rb[i]) + __ldg(args.bias + col + 32 + i);
          float x2b = __uint_as_float(rb[i + 1]) + __ldg(args.bias + col + 32 + i + 1);
          if (rot) {
            // rotate_half: out[i] = x[i] cos - x[i+32] sin ; out[i+32] = x[i+32] cos + x[i] sin
            const float o1a = x1a * cs[i] - x2a * sn[i];
            const float o2a = x2a * cs[i] + x1a * sn[i];
            const float o1b = x1b * cs[i + 1] - x2b * sn[i + 1];
            const float o2b = x2b * cs[i + 1] + x1b * sn[i + 1];
            x1a = o1a; x2a = o2a; x1b = o1b; x2b = o2b;
          }
          pa[i / 2] = pack_bf16x2(x1a, x1b);
          pb[i / 2] = pack_bf16x2(x2a, x2b);
        }
        uint4* o4 = reinterpret_cast<uint4*>(out + col);
#pragma unroll
        for (int i = 0; i < 4; ++i)
          o4[i] = make_uint4(pa[4 * i], pa[4 * i + 1], pa[4 * i + 2], pa[4 * i + 3]);
#pragma unroll
        for (int i = 0; i < 4; ++i)
          o4[4 + i] = make_uint4(pb[4 * i], pb[4 * i + 1], pb[4 * i + 2], pb[4 * i + 3]);
      }
    }
  } else {
#pragma unroll 1
    for (int c0 = cbeg; c0 < cbeg + HN; c0 += 32) {
      uint32_t r[32];
      tmem_ld_32x32(trow + c0, r);
      tmem_wait_ld();
      const int col = n0 + c0;
      if (!in_range || col >= args.N) continue;
      const bool full = (col + 32 <= args.N);
      float v[32];
#pragma unroll
      for (int i = 0; i < 32; ++i) {
        float bv = 0.f;
        if (args.bias != nullptr && (full || col + i < args.N)) bv = __ldg(args.bias + col + i);
        v[i] = __uint_as_float(r[i]) + bv;
      }

      if constexpr (EPI == EPI_BF16) {
        if (args.act == ACT_GELU_TANH) {
#pragma unroll
          for (int i = 0; i < 32; ++i) v[i] = gelu_tanh_f(v[i]);
        } else if (args.act == ACT_GELU_ERF) {
#pragma unroll
          for (int i = 0; i < 32; ++i) v[i] = gelu_erf_f(v[i]);
        } else if (args.act == ACT_SILU) {
#pragma unroll
          for (int i = 0; i < 32; ++i) v[i] = silu_f(v[i]);
        }
        __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(args.out) + grow * args.ldo + col;
        if (full) {
          uint4* o4 = reinterpret_cast<uint4*>(o);
#pragma unroll
          for (int i = 0; i < 4; ++i)
            o4[i] = make_uint4(pack_bf16x2(v[8 * i], v[8 * i + 1]), pack_bf16x2(v[8 * i + 2], v[8 * i + 3]),
                               pack_bf16x2(v[8 * i + 4], v[8 * i + 5]), pack_bf16x2(v[8 * i + 6], v[8 * i + 7]));
        } else {
#pragma unroll
          for (int i = 0; i < 32; ++i)
            if (col + i < args.N) o[i] = __float2bfloat16(v[i]);
        }
      } else if constexpr (EPI == EPI_F32) {
        float* o = reinterpret_cast<float*>(args.out) + grow * args.ldo + col;
        const float* ad = args.addend ? args.addend + grow * args.ld_add + col : nullptr;
        if (full) {
          float4* o4 = reinterpret_cast<float4*>(o);
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            float4 w = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
            if (ad) {
              const float4 a = *reinterpret_cast<const float4*>(ad + 4 * i);
              w.x += a.x; w.y += a.y; w.z += a.z; w.w += a.w;
            }
            o4[i] = w;
          }
        } else {
#pragma unroll
          for (int i = 0; i < 32; ++i)
            if (col + i < args.N) o[i] = v[i] + (ad ? ad[i] : 0.f);
        }
      } else if constexpr (EPI == EPI_GATE_RESID) {
        if (args.mask_rows && !valid) continue;
        float* o = reinterpret_cast<float*>(args.out) + grow * args.ldo + col;
        const long long step = args.step_ptr ? (long long)__ldg(args.step_ptr) : 0ll;
        const float* g = args.gate + step * args.gate_step_stride +
                         (long long)(b % args.gate_nb) * args.gate_ld + col;
        float4* o4 = reinterpret_cast<float4*>(o);
        const float4* g4 = reinterpret_cast<const float4*>(g);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          float4 x = o4[i];
          const float4 gg = __ldg(g4 + i);
          x.x += gg.x * v[4 * i]; x.y += gg.y * v[4 * i + 1];
          x.z += gg.z * v[4 * i + 2]; x.w += gg.w * v[4 * i + 3];
          o4[i] = x;
        }
      } else if constexpr (EPI == EPI_EMBED_DUAL) {
        float* o = reinterpret_cast<float*>(args.out) + grow * args.ldo + col;
        __nv_bfloat16* o2 = reinterpret_cast<__nv_bfloat16*>(args.out2) + grow * args.ldo2 + col;
        const float* ad = args.addend + grow * args.ld_add + col;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float4 a = *reinterpret_cast<const float4*>(ad + 4 * i);
          v[4 * i] = valid ? v[4 * i] + a.x : 0.f;
          v[4 * i + 1] = valid ? v[4 * i + 1] + a.y : 0.f;
          v[4 * i + 2] = valid ? v[4 * i + 2] + a.z : 0.f;
          v[4 * i + 3] = valid ? v[4 * i + 3] + a.w : 0.f;
          reinterpret_cast<float4*>(o)[i] = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
        }
        uint4* o4 = reinterpret_cast<uint4*>(o2);
#pragma unroll
        for (int i = 0; i < 4; ++i)
          o4[i] = make_uint4(pack_bf16x2(v[8 * i], v[8 * i + 1]), pack_bf16x2(v[8 * i + 2], v[8 * i + 3]),
                             pack_bf16x2(v[8 * i + 4], v[8 * i + 5]), pack_bf16x2(v[8 * i + 6], v[8 * i + 7]));
      } else if constexpr (EPI == EPI_MISH_MASK_BF16) {
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] = valid ? mish_f(v[i]) : 0.f;
        uint4* o4 = reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(args.out) + grow * args.ldo + col);
#pragma unroll
        for (int i = 0; i < 4; ++i)
          o4[i] = make_uint4(pack_bf16x2(v[8 * i], v[8 * i + 1]), pack_bf16x2(v[8 * i + 2], v[8 * i + 3]),
                             pack_bf16x2(v[8 * i + 4], v[8 * i + 5]), pack_bf16x2(v[8 * i + 6], v[8 * i + 7]));
      } else if constexpr (EPI == EPI_MISH_MASK_RESID) {
        float* o = reinterpret_cast<float*>(args.out) + grow * args.ldo + col;
        const float* ad = args.addend + grow * args.ld_add + col;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float4 a = *reinterpret_cast<const float4*>(ad + 4 * i);
          float4 w;
          w.x = (valid ? mish_f(v[4 * i]) : 0.f) + a.x;
          w.y = (valid ? mish_f(v[4 * i + 1]) : 0.f) + a.y;
          w.z = (valid ? mish_f(v[4 * i + 2]) : 0.f) + a.z;
          w.w = (valid ? mish_f(v[4 * i + 3]) : 0.f) + a.w;
          reinterpret_cast<float4*>(o)[i] = w;
        }
      } else if constexpr (EPI == EPI_SCALE_RESID) {
        float* o = reinterpret_cast<float*>(args.out) + grow * args.ldo + col;
        const float* ad = args.addend + grow * args.ld_add + col;
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          const float sc = args.gate ? __ldg(args.gate + col + i) : 1.f;
          v[i] = valid ? ad[i] + sc * v[i] : 0.f;
        }
#pragma unroll
        for (int i = 0; i < 8; ++i)
          reinterpret_cast<float4*>(o)[i] = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
        if (args.out2 != nullptr) {
          uint4* o4 = reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(args.out2) + grow * args.ldo2 + col);
#pragma unroll
          for (int i = 0; i < 4; ++i)
            o4[i] = make_uint4(pack_bf16x2(v[8 * i], v[8 * i + 1]), pack_bf16x2(v[8 * i + 2], v[8 * i + 3]),
                               pack_bf16x2(v[8 * i + 4], v[8 * i + 5]), pack_bf16x2(v[8 * i + 6], v[8 * i + 7]));
        }
      }
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
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));

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
      const int t = (m_tile % tiles_m_pb) * GEMM_BM + q * 32 + lane;  // row inside the batch element
      const int n0 = n_tile * BN;
      const int as = it & 1;
      const uint32_t aphase = (it >> 1) & 1u;
      mbar_wait(tfull_bar(as), aphase, 4);
      tc_fence_after();

      const uint32_t trow = tmem_base + (uint32_t(q * 32) << 16) + uint32_t(as * BN);
      gemm_epilogue_tile<BN, EPI, HN>(args, trow, b, t, n0, cbeg);
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
template <int BN>
struct Gemm2Cfg {
  static constexpr int kABytes = GEMM_BM * GEMM_BK * 2;
  static constexpr int kBBytes = (BN / 2) * GEMM_BK * 2;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kStages = (BN == 256) ? 6 : 8;
  static constexpr int kTmemCols = 2 * BN;
  static constexpr int kSmemBytes = kStages * kStageBytes + 1024 + 256;
};

template <int BN, int EPI>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(GEMM_THREADS, 1)
gemm2_bf16_tcgen05_kernel(const __grid_constant__ CUtensorMap tmA,
                          const __grid_constant__ CUtensorMap tmB, const GemmArgs args) {
  using Cfg = Gemm2Cfg<BN>;
  constexpr int kStages = Cfg::kStages;
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
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));

  if (warp == 0) {
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = pair_id; tile < num_tiles; tile += num_pairs) {
        const int m_tile = 2 * (tile % tiles_mp) + rank;
        const int n_tile = tile / tiles_mp;
        const int b = m_tile / tiles_m_pb;  // a phantom m-tile (odd tile count) lands past the last batch element: zero fill
        const int t0 = (m_tile % tiles_m_pb) * GEMM_BM;
        const int n0 = n_tile * BN;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(empty_bar(stage), phase ^ 1u, 21);
          if (leader) mbar_arrive_expect_tx(full_bar(stage), 2 * Cfg::kStageBytes);
          const uint32_t sa = smem_base + stage * Cfg::kStageBytes;
          const uint32_t sb = sa + Cfg::kABytes;
          const int a_col = (args.grouped ? (n0 / args.grouped) * args.grouped : 0) + (kb % args.cpb) * GEMM_BK;
          const int a_row = t0 + kb / args.cpb - args.pad;
          tma_load_3d_2sm(sa, &tmA, full_bar(stage), a_col, a_row, b);
          tma_load_2d_2sm(sb, &tmB, full_bar(stage), kb * GEMM_BK, n0 + rank * (BN / 2));
          if (++stage == kStages) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    if (leader && lane == 0) {
      constexpr uint32_t idesc = make_idesc_bf16(2 * GEMM_BM, BN);
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      for (int tile = pair_id; tile < num_tiles; tile += num_pairs, ++it) {
        const int as = it & 1;
        const uint32_t aphase = (it >> 1) & 1u;
        mbar_wait(tempty_bar(as), aphase ^ 1u, 22);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + uint32_t(as * BN);
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(full_bar(stage), phase, 23);
          tc_fence_after();
          const uint32_t sa = smem_base + stage * Cfg::kStageBytes;
          const uint32_t sb = sa + Cfg::kABytes;
          const uint64_t adesc = make_smem_desc_sw128(sa, 16, 1024);
          const uint64_t bdesc = make_smem_desc_sw128(sb, 16, 1024);
#pragma unroll
          for (int k = 0; k < GEMM_BK / 16; ++k)
            umma_bf16_ss_2sm(tmem_d, adesc + uint64_t(2 * k), bdesc + uint64_t(2 * k), idesc, (kb | k) != 0 ? 1u : 0u);
          umma_commit_2sm(empty_bar(stage), 3);
          if (kb == num_kb - 1) umma_commit_2sm(tfull_bar(as), 3);
          if (++stage == kStages) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else {
    const int q = warp & 3;
    const int chalf = (warp - 2) >> 2;
    constexpr int HN = BN / 2;
    const int cbeg = chalf * HN;
    int it = 0;
    for (int tile = pair_id; tile < num_tiles; tile += num_pairs, ++it) {
      const int m_tile = 2 * (tile % tiles_mp) + rank;
      const int n_tile = tile / tiles_mp;
      const int as = it & 1;
      const uint32_t aphase = (it >> 1) & 1u;
      mbar_wait(tfull_bar(as), aphase, 24);
      tc_fence_after();
      const int b = m_tile < tiles_m ? m_tile / tiles_m_pb : 0;
      // phantom tile: push the row index out of range so nothing is stored
      const int t = m_tile < tiles_m ? (m_tile % tiles_m_pb) * GEMM_BM + q * 32 + lane : args.rows_per_batch;
      const uint32_t trow = tmem_base + (uint32_t(q * 32) << 16) + uint32_t(as * BN);
      gemm_epilogue_tile<BN, EPI, HN>(args, trow, b, t, n_tile * BN, cbeg);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_leader(tempty_bar(as));
    }
  }

  tc_fence_before();
  cluster_sync_all();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc_2sm(tmem_base, Cfg::kTmemCols);
  }
}

}  // namespace oron
