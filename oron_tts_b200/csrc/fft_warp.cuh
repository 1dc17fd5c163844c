// 1024-point complex FFT owned by ONE warp: 32 points per lane, two in-register radix-32 passes and a
// single transpose through a warp-private shared-memory tile (no block barrier anywhere).
//
//   n = 32*n1 + n2, k = k1 + 32*k2:
//   X[k1 + 32 k2] = sum_{n2} W32^{n2 k2} * ( W1024^{n2 k1} * sum_{n1} W32^{n1 k1} x[32 n1 + n2] )
//
//   pass 1: lane n2 holds x[32 n1 + n2] (n1 = register index), DFT-32 over n1 in registers
//   twiddle: * W1024^{n2 k1} from a [k1][n2] table (lanes read consecutive words: conflict-free)
//   transpose: write row k1 / read row n2 of a 32 x 33 float2 tile (both conflict-free)
//   pass 2: lane k1 holds n2 = register index, DFT-32 over n2 -> X[k1 + 32 k2] in register k2
//
// Instruction budget per lane: 2 x ~260 packed-f32x2 butterfly instructions + 31 x 2 (twiddles) + 64 smem accesses (the
// scalar version of the same transform: 2 x ~460 + 31 x 4; the block-wide radix-4 Stockham transform before it: ~850 per
// thread x 8 warps with five barriers).
#pragma once
#include <cuda_runtime.h>

#include <type_traits>

namespace oron {
namespace fw {

constexpr int XB_LD = 33;                       // float2 row stride of the exchange tile
constexpr int XB_ELEMS = 32 * XB_LD;            // float2 per warp (8448 B)

// cos / sin of 2*pi*m/32, m = 0..15
constexpr float kCos[16] = {1.0f, 0.98078528040323f, 0.92387953251129f, 0.83146961230255f, 0.70710678118655f,
                            0.55557023301960f, 0.38268343236509f, 0.19509032201613f, 0.0f, -0.19509032201613f,
                            -0.38268343236509f, -0.55557023301960f, -0.70710678118655f, -0.83146961230255f,
                            -0.92387953251129f, -0.98078528040323f};
constexpr float kSin[16] = {0.0f, 0.19509032201613f, 0.38268343236509f, 0.55557023301960f, 0.70710678118655f,
                            0.83146961230255f, 0.92387953251129f, 0.98078528040323f, 1.0f, 0.98078528040323f,
                            0.92387953251129f, 0.83146961230255f, 0.70710678118655f, 0.55557023301960f,
                            0.38268343236509f, 0.19509032201613f};

__host__ __device__ constexpr int brev5(int r) {
  return ((r & 1) << 4) | ((r & 2) << 2) | (r & 4) | ((r & 8) >> 2) | ((r & 16) >> 4);
}

template <int I, int N, class F>
__device__ __forceinline__ void static_for(F&& f) {
  if constexpr (I < N) {
    f(std::integral_constant<int, I>{});
    static_for<I + 1, N>(f);
  }
}

// ---- packed complex arithmetic: one 64-bit register pair = (re, im), every operation one FADD2 / FMUL2 / FFMA2 (the
// swap of the two halves and the sign pattern of a complex product are operand modifiers of FFMA2 on sm_100: free).
// A radix-2 butterfly with twiddle is 4 instructions instead of 10 scalar ones; these kernels are bound by fp32
// instruction issue ("audio kernels").
typedef unsigned long long c64;
__device__ __forceinline__ c64 cpack(float re, float im) {
  c64 r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(re), "f"(im));
  return r;
}
__device__ __forceinline__ float2 cunpack(c64 v) {
  float2 r;
  asm("mov.b64 {%0, %1}, %2;" : "=f"(r.x), "=f"(r.y) : "l"(v));
  return r;
}
__device__ __forceinline__ c64 cswap(c64 v) {
  const float2 t = cunpack(v);
  return cpack(t.y, t.x);
}
__device__ __forceinline__ c64 cadd(c64 a, c64 b) {
  c64 r;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ c64 csub(c64 a, c64 b) {
  c64 r;
  asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ c64 cmul2(c64 a, c64 b) {  // element-wise
  c64 r;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ c64 cfma2(c64 a, c64 b, c64 c) {  // element-wise a * b + c
  c64 r;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
  return r;
}
// d * (wr + i wi)
__device__ __forceinline__ c64 cmul(c64 d, float wr, float wi) {
  return cfma2(cswap(d), cpack(-wi, wi), cmul2(d, cpack(wr, wr)));
}

// d * exp(-2 pi i M / 32) (forward) or d * exp(+2 pi i M / 32) (inverse)
template <int M, bool INV>
__device__ __forceinline__ c64 mul_w32(c64 d) {
  if constexpr (M == 0) {
    return d;
  } else if constexpr (M == 8) {  // -i d = (d.y, -d.x);  +i d = (-d.y, d.x)
    return cmul2(cswap(d), INV ? cpack(-1.0f, 1.0f) : cpack(1.0f, -1.0f));
  } else {
    constexpr float c = kCos[M], s = kSin[M];
    return INV ? cmul(d, c, s) : cmul(d, c, -s);
  }
}

// In-register 32-point DFT, decimation in frequency: v[brev5(k)] = sum_n v_in[n] W32^{nk} on return.
template <bool INV>
__device__ __forceinline__ void dft32(c64 (&v)[32]) {
  static_for<0, 5>([&](auto S) {
    constexpr int half = 16 >> decltype(S)::value;
    static_for<0, 16>([&](auto I) {
      constexpr int i = decltype(I)::value;
      constexpr int i0 = (i / half) * 2 * half + (i % half), i1 = i0 + half;
      constexpr int m = (i % half) * (16 / half);
      const c64 a = v[i0], c = v[i1];
      v[i0] = cadd(a, c);
      v[i1] = mul_w32<m, INV>(csub(a, c));
    });
  });
}

// tw[k1 * 32 + n2] = exp(-2 pi i k1 n2 / 1024); filled once per CTA by all threads.
__device__ __forceinline__ void fill_twiddle_table(float2* tw) {
  for (int i = threadIdx.x; i < 1024; i += blockDim.x) {
    const int p = ((i >> 5) * (i & 31)) & 1023;
    float s, c;
    sincospif(float(p) * (2.0f / 1024.0f), &s, &c);
    tw[i] = make_float2(c, -s);
  }
}

// in:  v[n1] = x[32 n1 + lane];  out: v[k2] = X[lane + 32 k2].  `xb` is this warp's XB_ELEMS tile.
// Unnormalised in both directions.  Ends with a __syncwarp(): the tile may be reused right away.
template <bool INV>
__device__ __forceinline__ void fft1024_warp(c64 (&v)[32], float2* xb, const float2* tw, int lane) {
  dft32<INV>(v);
  c64* xq = reinterpret_cast<c64*>(xb);
  const c64* twq = reinterpret_cast<const c64*>(tw);
  static_for<0, 32>([&](auto K) {
    constexpr int k1 = decltype(K)::value;
    c64 y = v[brev5(k1)];
    if constexpr (k1 > 0) {
      // y * (wr + i wi) with a run-time twiddle: one FMUL2 (wr broadcast) and two scalar FFMAs on the halves (the packed
      // form needs the pair (-wi, wi) built from a register: two more instructions)
      const float2 w = cunpack(twq[k1 * 32 + lane]);  // exp(-2 pi i k1 lane / 1024); conjugate for the inverse
      const float wi = INV ? -w.y : w.y;
      const float2 d = cunpack(y), t = cunpack(cmul2(y, cpack(w.x, w.x)));
      y = cpack(fmaf(-d.y, wi, t.x), fmaf(d.x, wi, t.y));
    }
    xq[k1 * XB_LD + lane] = y;
  });
  __syncwarp();
#pragma unroll
  for (int n2 = 0; n2 < 32; ++n2) v[n2] = xq[lane * XB_LD + n2];
  __syncwarp();
  dft32<INV>(v);
  c64 t[32];
  static_for<0, 32>([&](auto K) { t[decltype(K)::value] = v[brev5(decltype(K)::value)]; });
#pragma unroll
  for (int k = 0; k < 32; ++k) v[k] = t[k];
}
// float2 interface (same contract)
template <bool INV>
__device__ __forceinline__ void fft1024_warp(float2 (&v)[32], float2* xb, const float2* tw, int lane) {
  c64 q[32];
#pragma unroll
  for (int k = 0; k < 32; ++k) q[k] = cpack(v[k].x, v[k].y);
  fft1024_warp<INV>(q, xb, tw, lane);
#pragma unroll
  for (int k = 0; k < 32; ++k) v[k] = cunpack(q[k]);
}

}  // namespace fw
}  // namespace oron
