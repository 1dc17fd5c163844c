// FeedForward as ONE persistent launch (modules.py:294-299 + the gated residual of DiTBlock, modules.py:343):
//   phase 1   H = gelu_tanh(A W1^T + b1)          whole 256 x 256 tiles per SM pair, bf16 H written to global memory (L2)
//   phase 2   x += gate * (H W2^T + b2)           stream-K: equal shares of the flat (tile, k-block) list, f32 vector reductions
// Two separate launches of gemm2_bf16_tcgen05_kernel lose, at config 2, a kernel boundary (pipeline drain + fill), the
// wave quantisation of 176 phase-1 tiles on 74 SM pairs (2.38 waves) and the tail of the stream-K launch. Here every
// SM pair gets the same number of k-block units over BOTH phases: pair p runs the phase-1 tiles p, p + P, p + 2P ...
// (2 or 3 of them) and a phase-2 share sized so that (phase-1 units + phase-2 units) is U / P for everybody.
//
// Dependencies travel through global flags: flag[m_tile][n1] counts the epilogue warps (8 per CTA) that have stored
// their part of H[m_tile rows, n1 * 256 ..). The TMA producer of a phase-2 segment polls the flag of the k-blocks it is
// about to fetch (4 k-blocks per flag), then crosses to the async proxy. No CTA waits on anything before its own
// phase-1 tiles are issued and all CTAs are co-resident (grid <= SM count, 1 CTA per SM), so the waits cannot cycle.
// H becomes available in waves (first, second, third phase-1 tile of every pair = rising n1 = rising k-blocks of phase 2)
// while the pairs with fewer phase-1 tiles are already in phase 2 (measured: 16-21 k cycles of flag waits per CTA when a
// share was a contiguous k range). A share is therefore a contiguous range of a PERMUTED k index (k-block =
// position * kstride mod num_kb2, kstride ~ num_kb2 / golden ratio): every share samples the whole k range evenly and
// walks its k-blocks in ascending order, i.e. in the order in which H arrives.
// The last CTA to leave zeroes the flags (CUDA-graph replays find them clean).
#pragma once
#include "gemm_tcgen05.cuh"

namespace oron {

struct FfnSync {
  int* done;    // [1] CTAs that have left
  int* flags;   // [tiles_m_padded * tiles_n1]
  int nflags;
  int mode;                  // experiment knob (ORON_FFN_MODE): bit 0 = no fence.proxy.async, bit 1 = no fence.acq_rel.gpu after the polls
  int kstride, kstride_inv;  // phase-2 k interleave: position j of a tile's unit list holds k-block (j * kstride) % num_kb2
};

struct FfnWalk {
  int P, T1, nkb1, nkb2;
  int t1;
  long long lo, hi;
  int phase, tile, kb0, kb1;  // phase 0: k-blocks [kb0, kb1). phase 1: permuted positions [kb0, kb1), see member()
  int ks, ksinv;              // k-block kb of a phase-2 tile sits at position (kb * ksinv) % nkb2 of the flat unit list
  __device__ static long long c2(long long p, int P, int T1, int nkb1, long long U2) {
    const long long U = (long long)T1 * nkb1 + U2;
    const long long cum1 = p * (T1 / P) + min(p, (long long)(T1 % P));
    long long v = (p * U) / P - cum1 * nkb1;
    return v < 0 ? 0 : (v > U2 ? U2 : v);
  }
  __device__ FfnWalk(int pair_id, int num_pairs, int tiles1, int num_kb1, int tiles2, int num_kb2, int kstride, int kstride_inv) {
    P = num_pairs; T1 = tiles1; nkb1 = num_kb1; nkb2 = num_kb2; ks = kstride; ksinv = kstride_inv;
    t1 = pair_id;
    const long long U2 = (long long)tiles2 * num_kb2;
    lo = c2(pair_id, P, T1, nkb1, U2);
    hi = pair_id + 1 == P ? U2 : c2(pair_id + 1, P, T1, nkb1, U2);
    if (hi < lo) hi = lo;
    phase = 0; tile = 0; kb0 = 0; kb1 = 0;
  }
  __device__ bool next() {
    if (t1 < T1) {
      phase = 0; tile = t1; kb0 = 0; kb1 = nkb1; t1 += P;
      return true;
    }
    if (lo < hi) {  // last segment of the remaining share
      phase = 1;
      tile = int((hi - 1) / nkb2);
      const long long tb = (long long)tile * nkb2;
      const long long s = lo > tb ? lo : tb;
      kb0 = int(s - tb);
      kb1 = int(hi - tb);
      hi = s;
      return true;
    }
    return false;
  }
  // Visit order of a segment: k-blocks kb = first() .. last() - 1 with `pos` = position of kb in the tile's unit list, kept
  // incrementally (pos += ksinv mod nkb2); phase 2 skips the k-blocks whose position lies outside [kb0, kb1).
  __device__ int first() const { return phase == 0 ? kb0 : 0; }
  __device__ int last() const { return phase == 0 ? kb1 : nkb2; }
  __device__ bool member(int pos) const { return phase == 0 || (pos >= kb0 && pos < kb1); }
  __device__ int advance(int pos) const { pos += ksinv; return pos >= nkb2 ? pos - nkb2 : pos; }
};

// ld_acquire_gpu: gemm_tcgen05.cuh
__device__ __forceinline__ void red_release_gpu_add(int* p, int v) {
  asm volatile("red.release.gpu.global.add.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ int ld_relaxed_gpu(const int* p) {
  int v;
  asm volatile("ld.relaxed.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void fence_acq_rel_gpu() { asm volatile("fence.acq_rel.gpu;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async_all() { asm volatile("fence.proxy.async;" ::: "memory"); }

__device__ __forceinline__ void flag_wait(const int* p, int target) {
  if (ld_acquire_gpu(p) >= target) return;
  const long long t0 = clock64();
  while (ld_acquire_gpu(p) < target) {
    __nanosleep(32);
    if (clock64() - t0 > ORON_WATCHDOG_CYCLES) {
      printf("[oron] ffn flag watchdog: block %d flag %p target %d\n", blockIdx.x, (const void*)p, target);
      __trap();
    }
  }
}

// a1: phase 1 (EPI_BF16, act = ACT_GELU_TANH or any), a2: phase 2 (EPI_GATE_RESID, stream_k = 1). BN = 256 for both.
template <int ACT1>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(GEMM_THREADS, 1)
ffn2_bf16_tcgen05_kernel(const __grid_constant__ CUtensorMap tmA1, const __grid_constant__ CUtensorMap tmB1,
                         const __grid_constant__ CUtensorMap tmA2, const __grid_constant__ CUtensorMap tmB2,
                         const GemmArgs a1, const GemmArgs a2, const FfnSync sync) {
  constexpr int BN = 256;
  using Cfg = Gemm2Cfg<BN>;
  constexpr int kStages = Cfg::kStages;
  constexpr int KB_PER_FLAG = BN / GEMM_BK;  // k-blocks of phase 2 covered by one phase-1 tile
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
  const GemmArgs& args = a1;  // ORON_STAMP

  const int tiles_m_pb = (a1.rows_per_batch + GEMM_BM - 1) / GEMM_BM;
  const int tiles_m = tiles_m_pb * a1.nbatch;
  const int tiles_mp = (tiles_m + 1) / 2;
  const int tiles_n1 = (a1.N + BN - 1) / BN;
  const int tiles_n2 = (a2.N + BN - 1) / BN;
  const int T1 = tiles_mp * tiles_n1, T2 = tiles_mp * tiles_n2;
  const int nkb1 = a1.num_kb, nkb2 = a2.num_kb;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmA1);
    tma_prefetch_desc(&tmB1);
    tma_prefetch_desc(&tmA2);
    tma_prefetch_desc(&tmB2);
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
  pdl_wait();
  if (threadIdx.x == 0) { ORON_STAMP(0); ORON_STAMP_NS(11); }

  if (warp == 0) {
    // ===================== TMA producer (one thread per CTA) =====================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      bool stamped = false;
      const bool tr = args.dbg != nullptr;
      long long w_flag = 0, w_empty1 = 0, w_empty2 = 0;
      for (FfnWalk w(pair_id, num_pairs, T1, nkb1, T2, nkb2, sync.kstride, sync.kstride_inv); w.next();) {
        const int m_tile = 2 * (w.tile % tiles_mp) + rank;
        const int n_tile = w.tile / tiles_mp;
        const int b = m_tile / tiles_m_pb;  // phantom m-tile (odd tile count): past the last batch element -> zero fill
        const int t0 = (m_tile % tiles_m_pb) * GEMM_BM;
        const int n0 = n_tile * BN;
        const CUtensorMap* ta = w.phase ? &tmA2 : &tmA1;
        const CUtensorMap* tb = w.phase ? &tmB2 : &tmB1;
        // H tiles of this m-tile already known complete (bit f = columns [256 f, 256 f + 256)): one batch of relaxed loads per
        // segment instead of a dependent L2 round trip per flag (measured: ~1.3 k cycles per poll on the producer's critical path)
        uint32_t ready = 0u;
        const int* frow = sync.flags + m_tile * tiles_n1;
        auto snapshot = [&]() {
          uint32_t m = 0u;
          for (int f0 = 0; f0 < tiles_n1 && f0 < 32; f0 += 8) {
            int v[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) v[i] = (f0 + i < tiles_n1) ? ld_relaxed_gpu(frow + f0 + i) : 0;
#pragma unroll
            for (int i = 0; i < 8; ++i) m |= (v[i] >= GEMM_EPI_WARPS ? 1u : 0u) << (f0 + i);
          }
          if (!(sync.mode & 2)) fence_acq_rel_gpu();
          if (!(sync.mode & 1)) fence_proxy_async_all();
          return m;
        };
        if (w.phase) {
          const long long c0 = tr ? clock64() : 0;
          ready = snapshot();
          if (tr) w_flag += clock64() - c0;
          if (!stamped) { ORON_STAMP(2); ORON_STAMP_NS(13); stamped = true; }
        }
        int pos = 0;
        for (int kb = w.first(); kb < w.last(); ++kb, pos = w.advance(pos)) {
          if (!w.member(pos)) continue;
          if (w.phase) {
            const int f = kb / KB_PER_FLAG;
            if (f >= 32 || !((ready >> f) & 1u)) {  // not complete at the last look
              const long long c0 = tr ? clock64() : 0;
              flag_wait(frow + f, GEMM_EPI_WARPS);
              ready = f < 32 ? (snapshot() | (1u << f)) : 0u;  // (the snapshot also crosses to the async proxy)
              if (f >= 32 && !(sync.mode & 1)) fence_proxy_async_all();
              if (tr) w_flag += clock64() - c0;
            }
          }
          {
            const long long c0 = tr ? clock64() : 0;
            mbar_wait(empty_bar(stage), phase ^ 1u, 21);
            if (tr) { if (w.phase) w_empty2 += clock64() - c0; else w_empty1 += clock64() - c0; }
          }
          if (leader) mbar_arrive_expect_tx(full_bar(stage), 2 * Cfg::kStageBytes);
          const uint32_t sa = smem_base + stage * Cfg::kStageBytes;
          const uint32_t sb = sa + Cfg::kABytes;
          tma_load_3d_2sm(sa, ta, full_bar(stage), kb * GEMM_BK, t0, b);
          tma_load_2d_2sm(sb, tb, full_bar(stage), kb * GEMM_BK, n0 + rank * (BN / 2));
          if (++stage == kStages) { stage = 0; phase ^= 1u; }
        }
      }
      ORON_STAMP(3);
      if (tr) { args.dbg[(long long)blockIdx.x * 16 + 1] = w_empty1; args.dbg[(long long)blockIdx.x * 16 + 9] = w_empty2;
                args.dbg[(long long)blockIdx.x * 16 + 8] = w_flag; }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (leader CTA, one thread) =====================
    if (leader && lane == 0) {
      constexpr uint32_t idesc = make_idesc_bf16(2 * GEMM_BM, BN, 0, 0);
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      const bool tr = args.dbg != nullptr;
      long long w_full1 = 0, w_full2 = 0, w_tempty = 0;
      bool stamped = false;
      for (FfnWalk w(pair_id, num_pairs, T1, nkb1, T2, nkb2, sync.kstride, sync.kstride_inv); w.next(); ++it) {
        const int as = it & 1;
        const uint32_t aphase = (it >> 1) & 1u;
        {
          const long long c0 = tr ? clock64() : 0;
          mbar_wait(tempty_bar(as), aphase ^ 1u, 22);
          if (tr) w_tempty += clock64() - c0;
        }
        tc_fence_after();
        if (w.phase && !stamped) { ORON_STAMP(14); stamped = true; }
        const uint32_t tmem_d = tmem_base + uint32_t(as * BN);
        int left = w.kb1 - w.kb0;  // members of this segment (phase 2: positions of the permuted list)
        uint32_t acc = 0u;
        int pos = 0;
        for (int kb = w.first(); kb < w.last(); ++kb, pos = w.advance(pos)) {
          if (!w.member(pos)) continue;
          {
            const long long c0 = tr ? clock64() : 0;
            mbar_wait(full_bar(stage), phase, 23);
            if (tr) { if (w.phase) w_full2 += clock64() - c0; else w_full1 += clock64() - c0; }
          }
          tc_fence_after();
          const uint32_t sa = smem_base + stage * Cfg::kStageBytes;
          const uint32_t sb = sa + Cfg::kABytes;
          const uint64_t adesc = make_smem_desc_sw128(sa, 16, 1024);
          const uint64_t bdesc = make_smem_desc_sw128(sb, 16, 1024);
#pragma unroll
          for (int k = 0; k < GEMM_BK / 16; ++k)
            umma_bf16_ss_2sm(tmem_d, adesc + 2u * uint64_t(k), bdesc + 2u * uint64_t(k), idesc, (acc | uint32_t(k)) != 0 ? 1u : 0u);
          acc = 1u;
          umma_commit_2sm(empty_bar(stage), 3);
          if (--left == 0) umma_commit_2sm(tfull_bar(as), 3);
          if (++stage == kStages) { stage = 0; phase ^= 1u; }
        }
      }
      ORON_STAMP(4);
      if (tr) { args.dbg[(long long)blockIdx.x * 16 + 5] = w_full1; args.dbg[(long long)blockIdx.x * 16 + 6] = w_full2;
                args.dbg[(long long)blockIdx.x * 16 + 7] = w_tempty; }
    }
  } else {
    // ===================== epilogue warps (2..9 of both CTAs) =====================
    const int q = warp & 3;
    const int chalf = (warp - 2) >> 2;
    constexpr int HN = BN / 2;
    const int cbeg = chalf * HN;
    const uint32_t stage_buf = bar_base + 256u + uint32_t(warp - 2) * EPI_STAGE_BYTES_PER_WARP;
    int it = 0;
    for (FfnWalk w(pair_id, num_pairs, T1, nkb1, T2, nkb2, sync.kstride, sync.kstride_inv); w.next(); ++it) {
      const int m_tile = 2 * (w.tile % tiles_mp) + rank;
      const int n_tile = w.tile / tiles_mp;
      const int as = it & 1;
      const uint32_t aphase = (it >> 1) & 1u;
      const int b = m_tile < tiles_m ? m_tile / tiles_m_pb : 0;
      // phantom tile: push the row index out of range so nothing is stored
      const int t_base = m_tile < tiles_m ? (m_tile % tiles_m_pb) * GEMM_BM + q * 32 : a1.rows_per_batch;
      const uint32_t trow = tmem_base + (uint32_t(q * 32) << 16) + uint32_t(as * BN);
      if (w.phase == 0) {
        EpiCols<HN> pc;
        gemm_epilogue_prefetch<BN, EPI_BF16, HN>(a1, b, n_tile * BN, cbeg, lane, pc, true);
        mbar_wait(tfull_bar(as), aphase, 24);
        tc_fence_after();
        gemm_epilogue_tile<BN, EPI_BF16, HN, ACT1>(a1, trow, b, t_base, n_tile * BN, cbeg, stage_buf, lane, pc);
        tc_fence_before();
        __syncwarp();
        if (lane == 0) {
          mbar_arrive_leader(tempty_bar(as));
          // publish this warp's 32 x 128 part of H: the stores of all lanes are ordered before the reduction by the
          // __syncwarp above and the gpu-scope release
          __threadfence();
          red_release_gpu_add(sync.flags + m_tile * tiles_n1 + n_tile, 1);
        }
      } else {
        EpiCols<HN> pc;
        gemm_epilogue_prefetch<BN, EPI_GATE_RESID, HN>(a2, b, n_tile * BN, cbeg, lane, pc, w.kb0 == 0);  // the bias rides with k-block 0
        mbar_wait(tfull_bar(as), aphase, 25);
        tc_fence_after();
        gemm_epilogue_tile<BN, EPI_GATE_RESID, HN>(a2, trow, b, t_base, n_tile * BN, cbeg, stage_buf, lane, pc);
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_leader(tempty_bar(as));
      }
    }
  }

  tc_fence_before();
  cluster_sync_all();
  // last CTA out resets the flags: every other CTA has finished polling (its increment follows its last poll)
  __shared__ int s_last;
  if (threadIdx.x == 0) {
    ORON_STAMP(10);
    ORON_STAMP_NS(12);
    __threadfence();
    s_last = atomicAdd(sync.done, 1) == int(gridDim.x) - 1;
  }
  __syncthreads();
  if (s_last) {
    __threadfence();
    for (int i = threadIdx.x; i < sync.nflags; i += blockDim.x) sync.flags[i] = 0;
    if (threadIdx.x == 0) *sync.done = 0;
  }
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc_2sm(tmem_base, Cfg::kTmemCols);
  }
}

}  // namespace oron
