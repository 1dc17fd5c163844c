// Backward of the varlen non-causal attention (head_dim 64) on tcgen05 / TMEM: one CTA per (batch, head, 128-row owner
// tile). The 128-row operand tiles arrive by TMA (double buffered); S and dP are 128 x 128 accumulators in TMEM that the
// CUDA cores copy to registers and release at once, so the tensor core computes S / dP of tile j + 1 (N = 128 MMAs: the
// single issuing thread is the bound, wide instructions are cheapest per FLOP) and the accumulating MMAs of tile j while
// the CUDA cores turn tile j into dS (and P^T).
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
  long long* dbg;  // optional [grid, 16] clock64 stamps: slots 0-7 first compute thread, 8-15 MMA thread (tools/attn_bwd_trace.py)
  int have_lse;  // 1: `lse` was written by the forward kernel (oron_attention_fwd_lse): MODE 0 skips its pre-pass
};

constexpr int AB_THREADS = 544;  // warp 0: TMA + MMA issue (+ TMEM alloc); warps 1..16: four threads per owner row
                                 // (warp & 3 = TMEM lane quarter, (warp - 1) >> 2 = which 32 of the tile's 128 columns)
constexpr int AB_TILE = 128;
constexpr int AB_D = 64;
constexpr int AB_TILE_BYTES = AB_TILE * AB_D * 2;  // 16 KB
constexpr int AB_TMEM_COLS = 512;
// smem: X1 | X2 | Y1[0] Y2[0] | Y1[1] Y2[1] | Y1[2] Y2[2] | stageA (2 slabs) | stageB (2 slabs) | barriers + lse/delta staging
// Three Y buffers: the tiles of iteration it + 2 are fetched at the top of iteration it. With two buffers the fetch of
// it + 1 could only be issued once the accumulating MMAs of it - 1 had retired, i.e. right before S / dP of it + 1 needed
// it: the MMA thread sat ~1.2 k cycles per tile in the TMA latency (trace: "issue S/dP(n+1)" 1766 cycles for 8 MMAs).
constexpr int AB_YBUF = 3;
#ifndef AB_TS
#define AB_TS 1  // 1: dS / P^T reach the accumulating MMAs through TMEM (TS-form tcgen05.mma, A operand = 64 TMEM columns of packed
                 // bf16 per 128 x 128 tile) instead of a swizzled shared-memory staging tile: no staging stores, no cross-proxy
                 // fence, and a third less shared-memory operand traffic for the tensor pipe
#endif
#ifndef AB_POLY
#define AB_POLY 4  // of every 16 pairs of exponentials, this many are evaluated on the FMA pipe (ex2_poly2) instead of the MUFU unit
#endif
constexpr int AB_SMEM_BYTES = (6 + 2 * AB_YBUF) * AB_TILE_BYTES + 128 + 4 * 128 * 4 + 8 * 128 * 4 + 1024;

__device__ __forceinline__ void ab_tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) { tmem_ld_32x32(taddr, r); }
// 16 packed bf16x2 words of one row -> 16 TMEM columns (the A operand layout of a TS-form MMA: two K elements per column)
__device__ __forceinline__ void ab_tmem_st16(uint32_t taddr, const uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
      :
      : "r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]),
        "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}
__device__ __forceinline__ void ab_tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
// D[tmem] (+)= A[tmem] * B[smem desc]: A = 128 lanes x 16 K elements (bf16, two per 32-bit column)
__device__ __forceinline__ void ab_umma_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
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

// 17 warps: one SM sub-partition holds five of them, so the register budget is 16384 / (5 * 32) -> 96 per thread (ptxas picks
// it from the launch bounds; __maxnreg__(120), which the whole register file could hold, fails to launch)
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
  const uint32_t sA = smem_base + (2 + 2 * AB_YBUF) * AB_TILE_BYTES, sB = smem_base + (4 + 2 * AB_YBUF) * AB_TILE_BYTES;
  const uint32_t bar_base = smem_base + (6 + 2 * AB_YBUF) * AB_TILE_BYTES;
  // barriers (8 bytes each): x | y0 y1 | s0 s1 | p0 p1 | acc ; then the TMEM slot
  const uint32_t bar_x = bar_base, bar_acc = bar_base + 56, tmem_slot = bar_base + 80;
  auto bar_y = [&](int st) { return st < 2 ? bar_base + 8u + 8u * uint32_t(st) : bar_base + 72u; };
  auto bar_s = [&](int u) { return bar_base + 24u + 8u * uint32_t(u); };   // MMA -> CUDA cores: S_u / dP_u are in TMEM
  auto bar_p = [&](int u) { return bar_base + 40u + 8u * uint32_t(u); };   // CUDA cores -> MMA: S_u / dP_u read, slab u written
  float* s_stat = reinterpret_cast<float*>(smem_gen + (6 + 2 * AB_YBUF) * AB_TILE_BYTES + 128);  // [2][2][128]: buffer, {lse, delta}

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmQK);
    tma_prefetch_desc(&tmV);
    tma_prefetch_desc(&tmDO);
    mbar_init(bar_x, 1);
    mbar_init(bar_y(2), 1);
    for (int u = 0; u < 2; ++u) {
      mbar_init(bar_y(u), 1);
      mbar_init(bar_s(u), 1);
      mbar_init(bar_p(u), 512);
    }
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
  // TMEM columns: S (128) | dP (128) | two 64-column accumulators
  auto tmem_S = [&](int u) { return tmem_base + 128u * uint32_t(u); };  // u = 0: S, u = 1: dP
  const uint32_t tmem_acc1 = tmem_base + 256, tmem_acc2 = tmem_base + 320;
  // TS form: packed bf16 A operands, 64 columns each -- MODE 0: dS | MODE 1: P^T, dS^T
  const uint32_t tmem_a1 = tmem_base + 384, tmem_a2 = tmem_base + 448;

  // Iterations: one per 128-row operand tile (an optional pre-pass over the same tiles computes the log-sum-exp when the
  // forward did not supply it). S and dP are single 128-column TMEM tiles: the CUDA cores copy their 64 columns to
  // registers first and release the tiles at once (bar_free), so the tensor core computes S / dP of tile j + 1 while
  // tile j is turned into dS / P^T and fed to the accumulating MMAs.
  const int n_pre = (MODE == 0 && !args.have_lse) ? nt : 0;
  const int n_it = n_pre + nt;
  const uint32_t bar_free = bar_p(1);  // the second bar_p slot is unused in this scheme

  if (warp == 0) {
    if (lane == 0) {
      constexpr uint32_t idesc_s = make_idesc_bf16(128, 128, 0, 0);
      constexpr uint32_t idesc_acc = make_idesc_bf16(128, 64, 0, 1);
      const uint64_t x1desc = make_smem_desc_sw128(sX1, 16, 1024), x2desc = make_smem_desc_sw128(sX2, 16, 1024);
      // every descriptor is base + a small multiple: no arrays (dynamic indexing would put them in local memory and each
      // tcgen05.mma issue behind a load). Buffer st of the Y tiles is 2 tiles further (>> 4: 2048), staging slab 1 one
      // tile further (1024).
      const uint64_t y1d = make_smem_desc_sw128(sY1(0), 16, 1024), y2d = make_smem_desc_sw128(sY2(0), 16, 1024);
      const uint64_t y1m = make_smem_desc_sw128(sY1(0), 1024, 1024), y2m = make_smem_desc_sw128(sY2(0), 1024, 1024);
      const uint64_t ad = make_smem_desc_sw128(sA, 16, 1024), bd = make_smem_desc_sw128(sB, 16, 1024);
      constexpr uint64_t kBuf = uint64_t(2 * AB_TILE_BYTES) >> 4, kSlab = uint64_t(AB_TILE_BYTES) >> 4;
      // owner tiles
      mbar_arrive_expect_tx(bar_x, 2 * AB_TILE_BYTES);
      if (MODE == 0) {
        tma_load_3d(sX1, &tmQK, bar_x, h * AB_D, tile * AB_TILE, b);   // Q
        tma_load_3d(sX2, &tmDO, bar_x, h * AB_D, tile * AB_TILE, b);   // dO
      } else {
        tma_load_3d(sX1, &tmQK, bar_x, HD + h * AB_D, tile * AB_TILE, b);  // K
        tma_load_3d(sX2, &tmV, bar_x, h * AB_D, tile * AB_TILE, b);        // V
      }
      auto fetch = [&](int it) {  // the Y tiles of iteration `it` into buffer it % 3
        const bool pre = it < n_pre;
        const int j = pre ? it : it - n_pre;
        const int st = it % AB_YBUF;
        mbar_arrive_expect_tx(bar_y(st), (pre ? 1 : 2) * AB_TILE_BYTES);
        if (MODE == 0) {
          tma_load_3d(sY1(st), &tmQK, bar_y(st), HD + h * AB_D, j * AB_TILE, b);  // K_j
          if (!pre) tma_load_3d(sY2(st), &tmV, bar_y(st), h * AB_D, j * AB_TILE, b);  // V_j
        } else {
          tma_load_3d(sY1(st), &tmQK, bar_y(st), h * AB_D, j * AB_TILE, b);   // Q_i
          tma_load_3d(sY2(st), &tmDO, bar_y(st), h * AB_D, j * AB_TILE, b);   // dO_i
        }
      };
      auto issue_sdp = [&](int it) {
        const int st = it % AB_YBUF;
        mbar_wait(bar_y(st), uint32_t(it / AB_YBUF) & 1u, 2);
        tc_fence_after();
        const bool with_dp = it >= n_pre;
#pragma unroll
        for (int k = 0; k < 4; ++k) {  // the S and dP chains interleaved (independent accumulators)
          umma_bf16_ss(tmem_S(0), x1desc + uint64_t(2 * k), (y1d + kBuf * uint64_t(st)) + uint64_t(2 * k), idesc_s, k != 0);
          if (with_dp) umma_bf16_ss(tmem_S(1), x2desc + uint64_t(2 * k), (y2d + kBuf * uint64_t(st)) + uint64_t(2 * k), idesc_s, k != 0);
        }
        umma_commit(bar_s(0));
      };
#define AB_CSTAMP(slot) do { if (args.dbg != nullptr && it == n_pre + 3) args.dbg[(long long)blockIdx.x * 16 + (slot)] = clock64(); } while (0)
      fetch(0);
      if (n_it > 1) fetch(1);
      mbar_wait(bar_x, 0, 1);
      issue_sdp(0);
      int n_acc = 0;  // bar_acc phases committed so far (one per main tile)
      for (int it = 0; it < n_it; ++it) {
        const int st = it % AB_YBUF;
        const bool pre = it < n_pre;
        AB_CSTAMP(8);
        AB_CSTAMP(9);
        mbar_wait(bar_free, it & 1u, 8);  // S / dP of this iteration are in registers: the tiles may be overwritten
        if (it + 1 < n_it) issue_sdp(it + 1);  // its tiles were fetched one iteration ago
        if (it + 2 < n_it) {
          // buffer (it + 2) % 3 was last read by the MMAs of iteration it - 1 (S / dP: retired, their tiles were consumed;
          // accumulating MMAs: committed to bar_acc, and by now they have had the whole S / dP issue to retire)
          if (n_acc > 0 && it - 1 >= n_pre) mbar_wait(bar_acc, (n_acc - 1) & 1u, 4);
          fetch(it + 2);
        }
        AB_CSTAMP(10);
        mbar_wait(bar_p(0), it & 1u, 3);  // staging written
        AB_CSTAMP(11);
        if (!pre) {
          tc_fence_after();
          const int j = it - n_pre;
          const uint32_t accf = j == 0 ? 0u : 1u;
#if AB_TS
          (void)ad; (void)bd; (void)kSlab;
          if (MODE == 0) {
#pragma unroll
            for (int kk = 0; kk < 8; ++kk)  // dQ += dS K_j; even / odd k-steps into two accumulators, summed in the epilogue
              ab_umma_ts((kk & 1) ? tmem_acc1 : tmem_acc2, tmem_a1 + 8u * uint32_t(kk),
                         (y1m + kBuf * uint64_t(st)) + uint64_t(128 * kk), idesc_acc, kk >= 2 ? 1u : accf);
          } else {
#pragma unroll
            for (int kk = 0; kk < 8; ++kk) {  // dV += P^T dO_i and dK += dS^T Q_i, interleaved
              ab_umma_ts(tmem_acc1, tmem_a1 + 8u * uint32_t(kk), (y2m + kBuf * uint64_t(st)) + uint64_t(128 * kk), idesc_acc,
                         kk != 0 ? 1u : accf);
              ab_umma_ts(tmem_acc2, tmem_a2 + 8u * uint32_t(kk), (y1m + kBuf * uint64_t(st)) + uint64_t(128 * kk), idesc_acc,
                         kk != 0 ? 1u : accf);
            }
          }
#else
          if (MODE == 0) {
#pragma unroll
            for (int kk = 0; kk < 8; ++kk)  // dQ += dS K_j; even / odd k-steps into two accumulators, summed in the epilogue
              umma_bf16_ss((kk & 1) ? tmem_acc1 : tmem_acc2, (ad + kSlab * uint64_t(kk >> 2)) + uint64_t(2 * (kk & 3)),
                           (y1m + kBuf * uint64_t(st)) + uint64_t(128 * kk), idesc_acc, kk >= 2 ? 1u : accf);
          } else {
#pragma unroll
            for (int kk = 0; kk < 8; ++kk) {  // dV += P^T dO_i and dK += dS^T Q_i, interleaved
              umma_bf16_ss(tmem_acc1, (ad + kSlab * uint64_t(kk >> 2)) + uint64_t(2 * (kk & 3)),
                           (y2m + kBuf * uint64_t(st)) + uint64_t(128 * kk), idesc_acc, kk != 0 ? 1u : accf);
              umma_bf16_ss(tmem_acc2, (bd + kSlab * uint64_t(kk >> 2)) + uint64_t(2 * (kk & 3)),
                           (y1m + kBuf * uint64_t(st)) + uint64_t(128 * kk), idesc_acc, kk != 0 ? 1u : accf);
            }
          }
#endif
          umma_commit(bar_acc);
          ++n_acc;
        }
        AB_CSTAMP(12);
      }
    }
    __syncwarp();
  } else {
    // ===================== four threads per owner row =====================
    const int q4 = warp & 3;
    const int quad = (warp - 1) >> 2;  // columns [32 * quad, 32 * quad + 32) of every 128-column tile
    const int r = q4 * 32 + lane;
    const uint32_t lane_off = uint32_t(q4 * 32) << 16;
    const int t_own = tile * AB_TILE + r;
    const bool own_valid = t_own < len;
    const float c = args.scale_log2;
    const long long stat_base = ((long long)b * args.heads + h) * args.rows_per_batch;
    const bool tr = args.dbg != nullptr && threadIdx.x == 32;
#define AB_STAMP(slot) do { if (tr) args.dbg[(long long)blockIdx.x * 16 + (slot)] = clock64(); } while (0)
    AB_STAMP(0);
    float lse2 = 0.f, delta = 0.f;
    if (MODE == 0) {
      if (args.have_lse) lse2 = t_own < args.rows_per_batch ? __ldg(args.lse + stat_base + t_own) : 0.f;  // in flight under the delta loads
      // delta[row] = sum_d dO * O, cooperatively: 8 lanes per row read one 16-byte chunk each (coalesced), shuffle-reduce
      {
        const int tid = int(threadIdx.x) - 32;
        // all four 16-byte loads first: the two rounds used to run back to back (load - reduce - store, twice: two exposed
        // memory latencies at the head of every CTA, with S / dP of the first tile already waiting in TMEM)
        uint4 av[2], dv[2];
#pragma unroll
        for (int i = 0; i < 2; ++i) {
          const int q = i * 512 + tid;
          const int row = q >> 3, k = q & 7;
          const int t = tile * AB_TILE + row;
          av[i] = make_uint4(0u, 0u, 0u, 0u);
          dv[i] = make_uint4(0u, 0u, 0u, 0u);
          if (t < len) {
            av[i] = __ldg(reinterpret_cast<const uint4*>(args.o + (row_base + t) * args.ld_o + h * AB_D + 8 * k));
            dv[i] = __ldg(reinterpret_cast<const uint4*>(args.d_o + (row_base + t) * args.ld_do + h * AB_D + 8 * k));
          }
        }
#pragma unroll
        for (int i = 0; i < 2; ++i) {
          const int q = i * 512 + tid;
          const int row = q >> 3, k = q & 7;
          float part = 0.f;
          const uint32_t aw[4] = {av[i].x, av[i].y, av[i].z, av[i].w}, dw[4] = {dv[i].x, dv[i].y, dv[i].z, dv[i].w};
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            part = fmaf(__uint_as_float(aw[e] << 16), __uint_as_float(dw[e] << 16), part);
            part = fmaf(__uint_as_float(aw[e] & 0xffff0000u), __uint_as_float(dw[e] & 0xffff0000u), part);
          }
          part += __shfl_xor_sync(0xffffffffu, part, 1);
          part += __shfl_xor_sync(0xffffffffu, part, 2);
          part += __shfl_xor_sync(0xffffffffu, part, 4);
          if (k == 0) s_stat[128 + row] = part;
        }
        asm volatile("bar.sync 1, 512;" ::: "memory");
        delta = s_stat[128 + r];
      }
      // ---- pre-pass (no lse from the forward): log-sum-exp of the row, log2 domain; each thread covers its 64 columns of
      // every tile, the two partial (max, sum) pairs of a row are combined through shared memory ----
      float m = -INFINITY, l = 0.f;
      for (int it = 0; it < n_pre; ++it) {
        const int nv = min(AB_TILE, len - it * AB_TILE);
        mbar_wait(bar_s(0), it & 1u, 5);
        tc_fence_after();
        for (int c0 = 32 * quad; c0 < 32 * quad + 32; c0 += 32) {
          if (c0 >= nv) break;
          uint32_t v[32];
          ab_tmem_ld32(tmem_S(0) + lane_off + c0, v);
          tmem_wait_ld();
          float cm = -INFINITY;
#pragma unroll
          for (int k = 0; k < 32; ++k)
            if (c0 + k < nv) cm = fmaxf(cm, __uint_as_float(v[k]) * c);
          const float mn = fmaxf(m, cm);
          float sum = 0.f;
#pragma unroll
          for (int k = 0; k < 32; ++k)
            if (c0 + k < nv) sum += ex2_approx(fmaf(__uint_as_float(v[k]), c, -mn));
          l = l * ex2_approx(m - mn) + sum;
          m = mn;
        }
        tc_fence_before();
        mbar_arrive(bar_free);
        mbar_arrive(bar_p(0));
      }
      if (!args.have_lse) {
        s_stat[256 + 2 * (128 * quad + r)] = m;
        s_stat[256 + 2 * (128 * quad + r) + 1] = l;
        asm volatile("bar.sync 1, 512;" ::: "memory");
        float mm = -INFINITY;
#pragma unroll
        for (int qq = 0; qq < 4; ++qq) mm = fmaxf(mm, s_stat[256 + 2 * (128 * qq + r)]);
        float ll = 0.f;
#pragma unroll
        for (int qq = 0; qq < 4; ++qq) {
          const float mq = s_stat[256 + 2 * (128 * qq + r)];
          if (mq > -INFINITY) ll += s_stat[256 + 2 * (128 * qq + r) + 1] * ex2_approx(mq - mm);
        }
        lse2 = mm + log2f(ll);
      }
      if (quad == 0 && t_own < args.rows_per_batch) {
        if (!args.have_lse) args.lse[stat_base + t_own] = lse2;
        args.delta[stat_base + t_own] = delta;
      }
    }
    if (MODE == 0) lse2 -= log2f(args.scale);  // ps = scale * P straight out of the exp2
    AB_STAMP(1);
    // ---- main pass ----
    // MODE 1: the column statistics (lse, delta of the query tile) are read one iteration ahead: the load of tile j + 1 is in
    // flight under the arithmetic of tile j (it used to sit, ~800 cycles of L2 latency, at the top of every iteration)
    float sv_next = 0.f;
    const float* stat_src = quad == 0 ? args.lse : args.delta;
    if (MODE == 1 && quad < 2 && r < args.rows_per_batch) sv_next = stat_src[stat_base + r];
    for (int it = n_pre; it < n_it; ++it) {
      const int j = it - n_pre;
      const int nv = min(AB_TILE, len - j * AB_TILE);  // valid columns of this tile (keys in MODE 0, queries in MODE 1)
      const float* st = s_stat + (j & 1) * 256;
      if (MODE == 1) {
        float* sw_ = s_stat + (j & 1) * 256;
        if (quad < 2) {
          sw_[128 * quad + r] = quad == 0 ? sv_next : -args.scale * sv_next;  // (lse, -scale * delta)
          const int tqn = (j + 1) * AB_TILE + r;
          sv_next = (it + 1 < n_it && tqn < args.rows_per_batch) ? stat_src[stat_base + tqn] : 0.f;
        }
        asm volatile("bar.sync 1, 512;" ::: "memory");
      }
      if (it == n_pre + 3) AB_STAMP(13);
      mbar_wait(bar_s(0), it & 1u, 6);
      tc_fence_after();
      if (it == n_pre) AB_STAMP(2);
      if (it == n_pre + 1) AB_STAMP(3);
      if (it == n_pre + 3) AB_STAMP(14);
      // this thread's 32 columns of S and dP to registers, then the TMEM tiles are free for the next iteration's MMAs
      uint32_t vs[1][32], vd[1][32];
      ab_tmem_ld32(tmem_S(0) + lane_off + 32 * quad, vs[0]);
      ab_tmem_ld32(tmem_S(1) + lane_off + 32 * quad, vd[0]);
      tmem_wait_ld();
      tc_fence_before();
      mbar_arrive(bar_free);
      const bool fast = own_valid && nv == AB_TILE;  // nv is CTA-uniform
#pragma unroll
      for (int cc = 0; cc < 1; ++cc) {
        const int c0 = 32 * quad;  // column inside the 128-wide tile
        uint32_t pp[16], pd[16];
        // MODE 0: the softmax scale is folded into the exponent (lse2 holds lse - log2(scale)): ps = scale * P in one
        // exp2, dS = ps * (dP - delta). MODE 1: the staged column statistics are (lse, -scale * delta), so that
        // dS^T = P^T * fma(dP^T, scale, -scale * delta). Full tiles of valid rows skip the masks.
        if (fast) {
          if (MODE == 0) {
#pragma unroll
            for (int k = 0; k < 32; k += 2) {
              float p0 = fmaf(__uint_as_float(vs[cc][k]), c, -lse2), p1 = fmaf(__uint_as_float(vs[cc][k + 1]), c, -lse2);
              // the MUFU unit (16 exp2 per clock per SM) is the narrowest pipe of this loop (ncu: a quarter of all stall
              // samples sit on MUFU.EX2 with the MIO queue full): AB_POLY pairs of 16 go through the FMA-pipe polynomial
              if ((k >> 1) % 16 < AB_POLY) {
                ex2_poly2(p0, p1);
              } else {
                p0 = ex2_approx(p0);
                p1 = ex2_approx(p1);
              }
              pd[k >> 1] = pack_bf16x2(p0 * (__uint_as_float(vd[cc][k]) - delta), p1 * (__uint_as_float(vd[cc][k + 1]) - delta));
            }
          } else {
            const float4* sl = reinterpret_cast<const float4*>(st + c0);
            const float4* se = reinterpret_cast<const float4*>(st + 128 + c0);
#pragma unroll
            for (int k4 = 0; k4 < 8; ++k4) {
              const float4 l4 = sl[k4], e4 = se[k4];
              const float ll[4] = {l4.x, l4.y, l4.z, l4.w}, ee[4] = {e4.x, e4.y, e4.z, e4.w};
              float pv[4], dv[4];
#pragma unroll
              for (int e = 0; e < 4; ++e) pv[e] = fmaf(__uint_as_float(vs[cc][4 * k4 + e]), c, -ll[e]);
              if ((2 * k4) % 16 < AB_POLY) ex2_poly2(pv[0], pv[1]); else { pv[0] = ex2_approx(pv[0]); pv[1] = ex2_approx(pv[1]); }
              if ((2 * k4 + 1) % 16 < AB_POLY) ex2_poly2(pv[2], pv[3]); else { pv[2] = ex2_approx(pv[2]); pv[3] = ex2_approx(pv[3]); }
#pragma unroll
              for (int e = 0; e < 4; ++e) dv[e] = pv[e] * fmaf(__uint_as_float(vd[cc][4 * k4 + e]), args.scale, ee[e]);
              pp[2 * k4] = pack_bf16x2(pv[0], pv[1]);
              pp[2 * k4 + 1] = pack_bf16x2(pv[2], pv[3]);
              pd[2 * k4] = pack_bf16x2(dv[0], dv[1]);
              pd[2 * k4 + 1] = pack_bf16x2(dv[2], dv[3]);
            }
          }
        } else {
#pragma unroll
          for (int k = 0; k < 32; k += 2) {
            float p0 = 0.f, p1 = 0.f, d0 = 0.f, d1 = 0.f;
            const float l0 = MODE == 0 ? lse2 : st[c0 + k], l1 = MODE == 0 ? lse2 : st[c0 + k + 1];
            const float e0 = MODE == 0 ? delta : st[128 + c0 + k], e1 = MODE == 0 ? delta : st[128 + c0 + k + 1];
            if (own_valid && c0 + k < nv) {
              p0 = ex2_approx(fmaf(__uint_as_float(vs[cc][k]), c, -l0));
              d0 = MODE == 0 ? p0 * (__uint_as_float(vd[cc][k]) - e0) : p0 * fmaf(__uint_as_float(vd[cc][k]), args.scale, e0);
            }
            if (own_valid && c0 + k + 1 < nv) {
              p1 = ex2_approx(fmaf(__uint_as_float(vs[cc][k + 1]), c, -l1));
              d1 = MODE == 0 ? p1 * (__uint_as_float(vd[cc][k + 1]) - e1) : p1 * fmaf(__uint_as_float(vd[cc][k + 1]), args.scale, e1);
            }
            pp[k >> 1] = pack_bf16x2(p0, p1);
            pd[k >> 1] = pack_bf16x2(d0, d1);
          }
        }
        // the staging tiles still feed the previous iteration's accumulating MMAs: wait for them only now, with this
        // iteration's arithmetic already done
        if (j > 0) mbar_wait(bar_acc, (j - 1) & 1u, 9);
#if AB_TS
        // this thread's 32 columns = 16 packed words of its row, at columns [16 quad, 16 quad + 16) of the operand tile
        tc_fence_after();
        if (MODE == 0) {
          ab_tmem_st16(tmem_a1 + lane_off + 16u * uint32_t(quad), pd);
        } else {
          ab_tmem_st16(tmem_a1 + lane_off + 16u * uint32_t(quad), pp);
          ab_tmem_st16(tmem_a2 + lane_off + 16u * uint32_t(quad), pd);
        }
        ab_tmem_wait_st();
#else
        if (MODE == 0) {
          ab_stage_store(sA, r, c0, pd);
        } else {
          ab_stage_store(sA, r, c0, pp);
          ab_stage_store(sB, r, c0, pd);
        }
#endif
      }
#if !AB_TS
      fence_proxy_async_smem();
#endif
      tc_fence_before();
      mbar_arrive(bar_p(0));
      if (it == n_pre + 3) AB_STAMP(15);
    }
    AB_STAMP(4);
    // ---- read the accumulators ----
    mbar_wait(bar_acc, (nt - 1) & 1u, 7);
    tc_fence_after();
    AB_STAMP(5);
    // Phase A: the accumulator rows go to shared memory as f32 (row = 256 B, 16-byte chunks XOR-swizzled by row & 15)
    // in the idle staging tiles; phase B: all threads store 16-byte bf16 chunks, 8 lanes per row (coalesced), with the
    // RoPE rotation of dq / dk applied on the way (cos / sin read coalesced too).
    // columns [32 ch, 32 ch + 32) of an accumulator row (ch = 0 / 1), plus optionally the same columns of a second one
    auto to_smem = [&](uint32_t tm, uint32_t tile_s, uint32_t tm_add, int ch) {
      uint32_t lo[32];
      ab_tmem_ld32(tm + lane_off + 32 * ch, lo);
      tmem_wait_ld();
      if (tm_add != 0u) {
        uint32_t lo2[32];
        ab_tmem_ld32(tm_add + lane_off + 32 * ch, lo2);
        tmem_wait_ld();
#pragma unroll
        for (int i = 0; i < 32; ++i) lo[i] = __float_as_uint(__uint_as_float(lo[i]) + __uint_as_float(lo2[i]));
      }
      const uint32_t rowa = tile_s + uint32_t(r) * 256u;
      const uint32_t sw = uint32_t(r & 15);
#pragma unroll
      for (int cidx = 0; cidx < 8; ++cidx)
        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(rowa + ((uint32_t(cidx + 8 * ch) ^ sw) << 4)), "r"(lo[4 * cidx]),
                     "r"(lo[4 * cidx + 1]), "r"(lo[4 * cidx + 2]), "r"(lo[4 * cidx + 3]) : "memory");
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
      if (quad < 2) to_smem(tmem_acc2, sA, tmem_acc1, quad);
      asm volatile("bar.sync 1, 512;" ::: "memory");
#pragma unroll 1
      for (int i = 0; i < 2; ++i) store_task(sA, i * 512 + tid, 0, true);
    } else {
      if (quad < 2) to_smem(tmem_acc1, sA, 0u, quad); else to_smem(tmem_acc2, sB, 0u, quad - 2);
      asm volatile("bar.sync 1, 512;" ::: "memory");
#pragma unroll 1
      for (int i = 0; i < 4; ++i) {
        if (quad < 2) store_task(sA, i * 256 + (tid & 255), 2 * HD, false);
        else store_task(sB, i * 256 + (tid & 255), HD, true);
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
