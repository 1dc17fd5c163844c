// Thin inline-PTX wrappers for sm_100a: mbarrier, TMA (cp.async.bulk.tensor),
// tcgen05 (alloc / mma / commit / ld / fences) and UMMA descriptor builders.
// Everything here is device-only; no CUTLASS/CuTe dependency.
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

namespace oron {

// ---------------------------------------------------------------------------
// misc
// ---------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

__device__ __forceinline__ bool elect_one() {
  uint32_t pred = 0;
  asm volatile(
      "{\n\t"
      ".reg .b32 rx;\n\t"
      ".reg .pred px;\n\t"
      "elect.sync rx|px, %1;\n\t"
      "@px mov.s32 %0, 1;\n\t"
      "}\n"
      : "+r"(pred)
      : "r"(0xFFFFFFFFu));
  return pred != 0;
}

// ---------------------------------------------------------------------------
// programmatic dependent launch: a kernel launched with programmaticStreamSerialization may start while its
// predecessor is still running; it must not touch the predecessor's outputs before pdl_wait().
// Both are no-ops for ordinary launches.
// ---------------------------------------------------------------------------
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

// ---------------------------------------------------------------------------
// mbarrier
// ---------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ uint32_t mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.b32 %0, 1, 0, p;\n\t"
      "}\n"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok;
}

#ifndef ORON_WATCHDOG_CYCLES
#define ORON_WATCHDOG_CYCLES (6000000000LL)  // ~3 s at 2 GHz; a stuck pipeline traps instead of hanging the box
#endif

// Bounded wait: a wrong barrier protocol must fail loudly (trap), never hang the GPU.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity, int tag = 0) {
  if (mbar_try_wait(bar, parity)) return;
  const long long t0 = clock64();
  while (!mbar_try_wait(bar, parity)) {
    if (clock64() - t0 > ORON_WATCHDOG_CYCLES) {
      printf("[oron] mbarrier watchdog: block (%d,%d,%d) thread %d tag %d parity %u\n", blockIdx.x,
             blockIdx.y, blockIdx.z, threadIdx.x, tag, parity);
      __trap();
    }
  }
}

// ---------------------------------------------------------------------------
// TMA
// ---------------------------------------------------------------------------
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* m) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(m)) : "memory");
}
// L2 prefetch of one box (no shared-memory destination, no barrier): warms L2 for a later tma_load_2d of the same box
__device__ __forceinline__ void tma_prefetch_2d(const CUtensorMap* m, int32_t c0, int32_t c1) {
  asm volatile("cp.async.bulk.prefetch.tensor.2d.L2.global.tile [%0, {%1, %2}];" ::"l"(reinterpret_cast<uint64_t>(m)), "r"(c0), "r"(c1)
               : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t smem_dst, const CUtensorMap* m, uint32_t bar,
                                            int32_t c0, int32_t c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4}], [%2];"
      :
      : "r"(smem_dst), "l"(reinterpret_cast<uint64_t>(m)), "r"(bar), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_load_3d(uint32_t smem_dst, const CUtensorMap* m, uint32_t bar,
                                            int32_t c0, int32_t c1, int32_t c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4, %5}], [%2];"
      :
      : "r"(smem_dst), "l"(reinterpret_cast<uint64_t>(m)), "r"(bar), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
// generic-proxy smem writes -> visible to the async proxy (UMMA / TMA reads)
__device__ __forceinline__ void fence_proxy_async_smem() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

// ---------------------------------------------------------------------------
// tcgen05: TMEM management
// ---------------------------------------------------------------------------
// Must be executed by one full warp. Writes the TMEM base address to *smem_dst.
__device__ __forceinline__ void tmem_alloc(uint32_t smem_dst, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_dst),
               "r"(ncols)
               : "memory");
}
__device__ __forceinline__ void tmem_relinquish() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols)
               : "memory");
}
__device__ __forceinline__ void tc_fence_before() {
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_after() {
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tmem_wait_ld() {
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// ---------------------------------------------------------------------------
// tcgen05: MMA + commit
// ---------------------------------------------------------------------------
// D[tmem] (+)= A[smem desc] * B[smem desc]; kind::f16 covers bf16/fp16 inputs, fp32 accumulate.
__device__ __forceinline__ void umma_bf16_ss(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc,
                                             uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}\n"
      :
      : "r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// Arrive on an mbarrier once all previously issued tcgen05.mma of this thread have completed.
// (implies tcgen05.fence::before_thread_sync)
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(
                   bar)
               : "memory");
}

// ---------------------------------------------------------------------------
// 2-SM (cta_group::2) variants: a CTA pair of one cluster issues one M=256 MMA; only the leader
// (even) CTA issues tcgen05.mma / commit, both CTAs load their operand slices with TMA and both
// own 128 accumulator rows in their TMEM. Barrier addresses with bit 24 cleared name the leader
// CTA's barrier from either CTA of the pair (shared::cluster window).
// ---------------------------------------------------------------------------
constexpr uint32_t kPeerBitMask = 0xFEFFFFFFu;

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_alloc_2sm(uint32_t smem_dst, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_dst),
               "r"(ncols)
               : "memory");
}
__device__ __forceinline__ void tmem_relinquish_2sm() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_2sm(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols)
               : "memory");
}
// TMA loads whose completion bytes are credited to the LEADER CTA's mbarrier
__device__ __forceinline__ void tma_load_2d_2sm(uint32_t smem_dst, const CUtensorMap* m, uint32_t bar,
                                                int32_t c0, int32_t c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4}], [%2];"
      :
      : "r"(smem_dst), "l"(reinterpret_cast<uint64_t>(m)), "r"(bar & kPeerBitMask), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_load_3d_2sm(uint32_t smem_dst, const CUtensorMap* m, uint32_t bar,
                                                int32_t c0, int32_t c1, int32_t c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4, %5}], [%2];"
      :
      : "r"(smem_dst), "l"(reinterpret_cast<uint64_t>(m)), "r"(bar & kPeerBitMask), "r"(c0), "r"(c1),
        "r"(c2)
      : "memory");
}
__device__ __forceinline__ void umma_bf16_ss_2sm(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc,
                                                 uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}\n"
      :
      : "r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrive (once the issued MMAs retire) on the barrier at the same smem offset in the CTAs of cta_mask
__device__ __forceinline__ void umma_commit_2sm(uint32_t bar, uint16_t cta_mask) {
  asm volatile(
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
          bar),
      "h"(cta_mask)
      : "memory");
}
// plain arrive on the leader CTA's barrier from either CTA of the pair
__device__ __forceinline__ void mbar_arrive_leader(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(bar & kPeerBitMask) : "memory");
}

// Instruction descriptor (cute::UMMA::InstrDescriptor bit layout): bf16 x bf16 -> fp32.
// bits [4,6) c_format=1 (F32); [7,10) a_format=1 (BF16); [10,13) b_format=1 (BF16);
// bit 15 a_major, bit 16 b_major (0 = K-major, 1 = MN-major); [17,23) N>>3; [24,29) M>>4.
__host__ __device__ constexpr uint32_t make_idesc_bf16(int M, int N, int a_mn_major = 0,
                                                       int b_mn_major = 0) {
  return (1u << 4) | (1u << 7) | (1u << 10) | (uint32_t(a_mn_major) << 15) |
         (uint32_t(b_mn_major) << 16) | (uint32_t(N >> 3) << 17) | (uint32_t(M >> 4) << 24);
}
// same, IEEE f16 x f16 -> fp32 (a_format = b_format = 0)
__host__ __device__ constexpr uint32_t make_idesc_f16(int M, int N, int a_mn_major = 0, int b_mn_major = 0) {
  return (1u << 4) | (uint32_t(a_mn_major) << 15) | (uint32_t(b_mn_major) << 16) | (uint32_t(N >> 3) << 17) |
         (uint32_t(M >> 4) << 24);
}
// no-swizzle K-major descriptor (8-row x 16-byte core matrices): LBO = bytes between K-adjacent core matrices,
// SBO = bytes between 8-row groups
__device__ __forceinline__ uint64_t make_smem_desc_noswz(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= uint64_t((smem_addr >> 4) & 0x3FFF);
  d |= uint64_t((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= uint64_t((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= uint64_t(1) << 46;
  return d;
}
// two exp2 per MUFU op: packed IEEE-f16 in, packed f16 out
__device__ __forceinline__ uint32_t ex2_f16x2(uint32_t x) {
  uint32_t y;
  asm("ex2.approx.f16x2 %0, %1;" : "=r"(y) : "r"(x));
  return y;
}

// Shared-memory matrix descriptor (cute::UMMA::SmemDescriptor bit layout), 128-byte swizzle.
//   [0,14)  start address >> 4        [16,30) leading byte offset >> 4
//   [32,46) stride byte offset >> 4   [46,48) version = 1 (sm_100)
//   [61,64) layout type: 2 = SWIZZLE_128B
// K-major SW128 tile (rows of 128 B, 8-row 1024 B swizzle atoms): SBO = 1024, LBO unused (=1).
// MN-major SW128 tile (64 MN-elements contiguous per 128 B row, 8 K-rows per atom):
//   SBO = 1024 (next 8 K rows), LBO = stride to the next 64 MN elements (unused when MN = 64).
__device__ __forceinline__ uint64_t make_smem_desc_sw128(uint32_t smem_addr, uint32_t lbo_bytes,
                                                         uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= uint64_t((smem_addr >> 4) & 0x3FFF);
  d |= uint64_t((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= uint64_t((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= uint64_t(1) << 46;
  d |= uint64_t(2) << 61;
  return d;
}

// ---------------------------------------------------------------------------
// tcgen05.ld: 32 lanes x 32 columns of fp32; thread i of the warp receives lane (base+i),
// registers j = columns (col+j). A warp may only touch lanes [32*(warp_id%4), +32).
// ---------------------------------------------------------------------------
__device__ __forceinline__ void tmem_ld_32x32(uint32_t taddr, uint32_t (&r)[32]) {
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

// ---------------------------------------------------------------------------
// small math helpers shared by epilogues
// ---------------------------------------------------------------------------
__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}
__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// 2^x for a pair on the FMA / ALU pipes instead of the MUFU unit (one in four exponentials goes this way: the MUFU unit,
// 16 results per clock per SM, is the softmax's narrowest pipe). x = floor(x) + f by adding 1.5 * 2^23 with round-down,
// degree-3 polynomial for 2^f on [0, 1) (max relative error 9e-5, f16 resolution is 4.9e-4), floor(x) added into the
// exponent field. Valid for -127 <= x < 128; smaller x are clamped (result ~0).
__device__ __forceinline__ void ex2_poly2(float& x0, float& x1) {
  uint64_t xv, mg, t, r, f, p, c3, c2, c1, c0;
  const float y0 = fmaxf(x0, -127.0f), y1 = fmaxf(x1, -127.0f);
  asm("mov.b64 %0, {%1, %2};" : "=l"(xv) : "f"(y0), "f"(y1));
  asm("mov.b64 %0, {%1, %1};" : "=l"(mg) : "f"(12582912.0f));
  asm("mov.b64 %0, {%1, %1};" : "=l"(c3) : "f"(0.07711965f));
  asm("mov.b64 %0, {%1, %1};" : "=l"(c2) : "f"(0.22756439f));
  asm("mov.b64 %0, {%1, %1};" : "=l"(c1) : "f"(0.69514614f));
  asm("mov.b64 %0, {%1, %1};" : "=l"(c0) : "f"(1.0f));
  asm("add.rm.ftz.f32x2 %0, %1, %2;" : "=l"(t) : "l"(xv), "l"(mg));
  asm("sub.rn.ftz.f32x2 %0, %1, %2;" : "=l"(r) : "l"(t), "l"(mg));
  asm("sub.rn.ftz.f32x2 %0, %1, %2;" : "=l"(f) : "l"(xv), "l"(r));
  asm("fma.rn.ftz.f32x2 %0, %1, %2, %3;" : "=l"(p) : "l"(f), "l"(c3), "l"(c2));
  asm("fma.rn.ftz.f32x2 %0, %1, %2, %3;" : "=l"(p) : "l"(p), "l"(f), "l"(c1));
  asm("fma.rn.ftz.f32x2 %0, %1, %2, %3;" : "=l"(p) : "l"(p), "l"(f), "l"(c0));
  uint32_t t0, t1, p0, p1;
  asm("mov.b64 {%0, %1}, %2;" : "=r"(t0), "=r"(t1) : "l"(t));
  asm("mov.b64 {%0, %1}, %2;" : "=r"(p0), "=r"(p1) : "l"(p));
  x0 = __uint_as_float(p0 + (t0 << 23));
  x1 = __uint_as_float(p1 + (t1 << 23));
}

__device__ __forceinline__ float gelu_tanh_f(float x) {
  // 0.5 x (1 + tanh(sqrt(2/pi) (x + 0.044715 x^3))) with the hardware tanh (one MUFU op, max rel. error 2^-11:
  // below the bf16 rounding of the result). The epilogue that uses it is instruction-bound.
  const float x2 = x * x;
  const float u = x * fmaf(x2, 0.7978845608028654f * 0.044715f, 0.7978845608028654f);
  float th;
  asm("tanh.approx.f32 %0, %1;" : "=f"(th) : "f"(u));
  const float hx = 0.5f * x;
  return fmaf(hx, th, hx);
}
// Exact-form GELU, 0.5 x (1 + erf(x / sqrt 2)). erf through Abramowitz-Stegun 7.1.26 (|error| <= 1.5e-7, below
// fp32 erff's own error after the bf16 rounding that follows) on the MUFU rcp / ex2 units: ~14 instructions and
// no divergent range branches, against ~30 for erff -- the Vocos pointwise-conv epilogue is bound by this.
__device__ __forceinline__ float gelu_erf_f(float x) {
  const float z = fabsf(x) * 0.7071067811865476f;
  float t;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t) : "f"(fmaf(0.3275911f, z, 1.0f)));
  float p = fmaf(1.061405429f, t, -1.453152027f);
  p = fmaf(p, t, 1.421413741f);
  p = fmaf(p, t, -0.284496736f);
  p = fmaf(p, t, 0.254829592f);
  const float e = ex2_approx(-1.4426950408889634f * z * z);
  const float erf_abs = fmaf(-p * t, e, 1.0f);
  const float hx = 0.5f * x;
  return fmaf(fabsf(hx), erf_abs, hx);  // 0.5 x + 0.5 |x| erf(|x|/sqrt2) == 0.5 x (1 + erf(x/sqrt2))
}
__device__ __forceinline__ float silu_f(float x) { return __fdividef(x, 1.0f + __expf(-x)); }
__device__ __forceinline__ float mish_f(float x) {
  // x * tanh(softplus(x)) with tanh(log(1 + n)) = ((1 + n)^2 - 1) / ((1 + n)^2 + 1) = p / (p + 2), p = n (n + 2), n = e^x:
  // one exponential and one division, no cancellation for x << 0 (rel. error 3e-7 against float64 over [-30, 30]). The
  // clamp plays the part of torch's softplus threshold (20): there p / (p + 2) rounds to 1 and mish(x) = x.
  const float n = __expf(fminf(x, 20.0f));
  const float p = n * (n + 2.0f);
  return x * __fdividef(p, p + 2.0f);
}

// Dropout (modules.py:253, 297) as a stateless mask: element `idx` of a tensor survives iff hash(seed, idx) >= p * 2^32,
// survivors are scaled by 1 / (1 - p). Forward and backward recompute the same mask from (seed, idx): nothing is stored.
struct DropCfg {
  unsigned int key;     // host-mixed (splitmix64) 32-bit key of the call's seed
  unsigned int thresh;  // p * 2^32 (0: dropout off)
  float inv_keep;       // 1 / (1 - p)
};
// one hash decides the element pair (idx, idx + 1), idx even: its two 16-bit halves against p * 2^16
__device__ __forceinline__ float2 drop_scale2(const DropCfg& d, unsigned long long idx) {
  if (d.thresh == 0u) return make_float2(1.0f, 1.0f);
  unsigned int x = (unsigned int)(idx >> 1) * 0x9E3779B1u + d.key;  // lowbias32 finaliser
  x ^= x >> 16;
  x *= 0x7feb352du;
  x ^= x >> 15;
  x *= 0x846ca68bu;
  x ^= x >> 16;
  const unsigned int t16 = d.thresh >> 16;
  return make_float2((x & 0xffffu) >= t16 ? d.inv_keep : 0.0f, (x >> 16) >= t16 ? d.inv_keep : 0.0f);
}

// derivative of GELU(tanh) (forward form: gelu_tanh_f above)
__device__ __forceinline__ float gelu_tanh_grad(float x) {
  const float k = 0.7978845608028654f, a = 0.044715f;
  const float x2 = x * x;
  const float u = k * x * fmaf(a, x2, 1.0f);
  float t;
  asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(u));  // same unit as the forward epilogue (max rel. error 2^-11)
  return 0.5f * (1.0f + t) + 0.5f * x * (1.0f - t * t) * k * fmaf(3.0f * a, x2, 1.0f);
}

}  // namespace oron
