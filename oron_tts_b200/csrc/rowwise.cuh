// HBM/L2-bound row-wise kernels of the DiT / TextEmbedding / Vocos path.
// Layout everywhere: activations are [nbatch * rows_per_batch, C] row-major ("frame-major").
#pragma once
#include "ptx.cuh"

namespace oron {

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// ---------------------------------------------------------------------------
// LayerNorm (eps, biased variance, fp32 statistics) fused with either
//   AdaLN modulation  y = LN(x) * (1 + scale[b]) + shift[b]   (modules.py:218, :234, :341)
//   affine            y = LN(x) * weight + bias               (modules.py:169 / Vocos norms)
// One warp per row; C = 32 * VPL * 4 ... handled generically with float4 lanes.
// mod vectors live in a per-step table: ptr + step*step_stride + (b % mod_nb)*mod_ld.
// ---------------------------------------------------------------------------
struct LnArgs {
  const float* x;        // [rows, ldx]
  long long ldx;
  int rows_per_batch, nbatch, C;
  float eps;
  const float* scale;    // modulation: (1+scale) ; affine: weight (add_one = 0)
  const float* shift;    // modulation shift / affine bias (nullptr -> 0)
  long long mod_ld;      // batch stride of scale/shift (0 for affine)
  int mod_nb;
  long long step_stride;
  const int* step_ptr;
  int add_one;
  __nv_bfloat16* out_bf16;  // [rows, ldo] or nullptr
  float* out_f32;           // [rows, ldo] or nullptr
  long long ldo;
};

template <int C>
__global__ void __launch_bounds__(256) ln_modulate_kernel(const LnArgs a) {
  constexpr int V4 = C / 128;  // float4 per lane
  pdl_launch_dependents();
  pdl_wait();
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  const long long rows = (long long)a.rows_per_batch * a.nbatch;
  if (warp >= rows) return;
  const int b = warp / a.rows_per_batch;
  const float4* xr = reinterpret_cast<const float4*>(a.x + (long long)warp * a.ldx);
  float4 v[V4];
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < V4; ++i) {
    v[i] = xr[lane + 32 * i];
    s += v[i].x + v[i].y + v[i].z + v[i].w;
  }
  const float mean = warp_sum(s) * (1.0f / C);
  float ss = 0.f;
#pragma unroll
  for (int i = 0; i < V4; ++i) {
    v[i].x -= mean; v[i].y -= mean; v[i].z -= mean; v[i].w -= mean;
    ss += v[i].x * v[i].x + v[i].y * v[i].y + v[i].z * v[i].z + v[i].w * v[i].w;
  }
  const float rstd = rsqrtf(warp_sum(ss) * (1.0f / C) + a.eps);
  const long long step = a.step_ptr ? (long long)__ldg(a.step_ptr) : 0ll;
  const long long moff = step * a.step_stride + (long long)(b % a.mod_nb) * a.mod_ld;
  const float4* sc = reinterpret_cast<const float4*>(a.scale + moff);
  const float4* sh = a.shift ? reinterpret_cast<const float4*>(a.shift + moff) : nullptr;
  const float one = a.add_one ? 1.f : 0.f;
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
    if (a.out_bf16) {
      uint2 p = make_uint2(pack_bf16x2(y.x, y.y), pack_bf16x2(y.z, y.w));
      reinterpret_cast<uint2*>(a.out_bf16 + (long long)warp * a.ldo)[lane + 32 * i] = p;
    }
    if (a.out_f32) reinterpret_cast<float4*>(a.out_f32 + (long long)warp * a.ldo)[lane + 32 * i] = y;
  }
}

// Narrow rows (C = 64: the text embedding of small models): two columns per lane, same arithmetic.
__global__ void __launch_bounds__(256) ln_modulate64_kernel(const LnArgs a) {
  pdl_launch_dependents();
  pdl_wait();
  constexpr int C = 64;
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  const long long rows = (long long)a.rows_per_batch * a.nbatch;
  if (warp >= rows) return;
  const int b = warp / a.rows_per_batch;
  float2 v = reinterpret_cast<const float2*>(a.x + (long long)warp * a.ldx)[lane];
  const float mean = warp_sum(v.x + v.y) * (1.0f / C);
  v.x -= mean;
  v.y -= mean;
  const float rstd = rsqrtf(warp_sum(v.x * v.x + v.y * v.y) * (1.0f / C) + a.eps);
  const long long step = a.step_ptr ? (long long)__ldg(a.step_ptr) : 0ll;
  const long long moff = step * a.step_stride + (long long)(b % a.mod_nb) * a.mod_ld;
  const float2 g = reinterpret_cast<const float2*>(a.scale + moff)[lane];
  float2 h = make_float2(0.f, 0.f);
  if (a.shift) h = reinterpret_cast<const float2*>(a.shift + moff)[lane];
  const float one = a.add_one ? 1.f : 0.f;
  const float2 y = make_float2(v.x * rstd * (one + g.x) + h.x, v.y * rstd * (one + g.y) + h.y);
  if (a.out_bf16) reinterpret_cast<uint32_t*>(a.out_bf16 + (long long)warp * a.ldo)[lane] = pack_bf16x2(y.x, y.y);
  if (a.out_f32) reinterpret_cast<float2*>(a.out_f32 + (long long)warp * a.ldo)[lane] = y;
}

// ---------------------------------------------------------------------------
// CFG combine + Euler update (flow.py:266-267, 295-299), one launch per ODE step:
//   v = v_c + (v_c - v_u) * cfg ;  x += v * dt[step]
// also refreshes the bf16 GEMM operand of the next step (both CFG halves) and the trajectory
// slot, then (last thread) advances the device-side step counter.
// ---------------------------------------------------------------------------
struct EulerArgs {
  float* x;              // [nb*rows_per_batch, n_mels] fp32 ODE state (in place)
  const float* v;        // [(cfg?2:1)*nb*rows_per_batch, ldv] fp32 velocity (cond rows first)
  long long ldv;
  int nb, rows_per_batch, n_mels;
  int has_uncond;
  float cfg;
  const float* dt;       // [steps]
  int* step_ptr;         // read, then incremented by one
  __nv_bfloat16* xb;     // [(cfg?2:1)*nb*rows_per_batch, ldxb] bf16 copy for the input projection
  long long ldxb;
  float* traj;           // [steps+1, nb*rows_per_batch, n_mels] or nullptr; slot step+1 is written
  float* v_out;          // optional [nb*rows_per_batch, n_mels]: the guided velocity (parity checks)
  int midpoint;          // 1: explicit midpoint rule. The counter counts EVALUATIONS e: interval i = e >> 1; even e:
                         // the operand becomes x + v dt_i / 2 (x itself untouched), odd e: x += v dt_i, trajectory slot i + 1
};

__global__ void __launch_bounds__(256) cfg_euler_kernel(const EulerArgs a) {
  pdl_launch_dependents();
  pdl_wait();
  const long long rows = (long long)a.nb * a.rows_per_batch;
  const long long total = rows * a.n_mels;
  const int e = *a.step_ptr;
  const int step = a.midpoint ? (e >> 1) : e;
  const bool half_step = a.midpoint && (e & 1) == 0;  // first evaluation of a midpoint interval
  const float dt = half_step ? 0.5f * a.dt[step] : a.dt[step];
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const long long row = i / a.n_mels;
    const int c = int(i - row * a.n_mels);
    const float vc = a.v[row * a.ldv + c];
    float vv = vc;
    if (a.has_uncond) {
      const float vu = a.v[(rows + row) * a.ldv + c];
      vv = vc + (vc - vu) * a.cfg;
    }
    const float xn = a.x[i] + vv * dt;
    if (!half_step) {
      a.x[i] = xn;
      if (a.traj) a.traj[(long long)(step + 1) * total + i] = xn;
    }
    if (a.v_out) a.v_out[i] = vv;
    const __nv_bfloat16 xh = __float2bfloat16(xn);
    a.xb[row * a.ldxb + c] = xh;
    if (a.has_uncond) a.xb[(rows + row) * a.ldxb + c] = xh;
  }
}
__global__ void step_advance_kernel(int* step_ptr) {
  pdl_launch_dependents();
  pdl_wait();
  *step_ptr += 1;
}

// fp32 [rows, C] -> bf16 [rows, ldo] (cols >= C untouched), optionally replicated `reps` times
__global__ void __launch_bounds__(256)
cast_rows_bf16_kernel(const float* x, long long ldx, long long rows, int C, __nv_bfloat16* out,
                      long long ldo, int reps) {
  const long long total = rows * C;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const long long row = i / C;
    const int c = int(i - row * C);
    const __nv_bfloat16 h = __float2bfloat16(x[row * ldx + c]);
    for (int r = 0; r < reps; ++r) out[(r * rows + row) * ldo + c] = h;
  }
}

// ---------------------------------------------------------------------------
// Timestep features (modules.py:39-45): [sin(1000 t f_j) | cos(1000 t f_j)], f_j = exp(-j ln(1e4)/127)
// -> bf16 [n, 256]
// ---------------------------------------------------------------------------
__global__ void time_sinusoid_kernel(const float* t, int n, __nv_bfloat16* out, long long ldo) {
  const int i = blockIdx.x;
  const int j = threadIdx.x;  // 0..127
  if (i >= n) return;
  const float emb = expf(float(j) * (-9.210340371976184f / 127.0f));
  const float e = 1000.0f * t[i] * emb;
  out[(long long)i * ldo + j] = __float2bfloat16(sinf(e));
  out[(long long)i * ldo + 128 + j] = __float2bfloat16(cosf(e));
}

// ---------------------------------------------------------------------------
// TextEmbedding front (encoder.py:68-91): ids(+1, 0 = filler) -> emb + sinusoidal abs-pos, fillers zeroed.
//   ids:   [nb, rows_per_batch] int32, already shifted by +1 and cropped/padded with 0
//   drop:  per batch element flag: look up row 0 for every position (text dropped) but keep the
//          filler mask of the *original* ids (encoder.py:77-80)
// outputs fp32 x [rows, C] and row_valid (1 = real token) consumed by the ConvNeXt epilogues.
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(128)
text_embed_front_kernel(const int* ids, const unsigned char* drop, const float* table,
                        const float* pos_table, int rows_per_batch, int nb, int C, float* x,
                        long long ldx, unsigned char* row_valid) {
  const long long row = blockIdx.x;
  const int b = int(row / rows_per_batch);
  const int t = int(row - (long long)b * rows_per_batch);
  const int id = ids[row];
  const bool filler = (id == 0);
  if (threadIdx.x == 0) row_valid[row] = filler ? 0 : 1;
  const int look = drop[b] ? 0 : id;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    // pos_table = precompute_freqs_cis(C, 8192) built on the host exactly as modules.py:191-196
    const float v = filler ? 0.f : table[(long long)look * C + c] + pos_table[(long long)t * C + c];
    x[row * ldx + c] = v;
  }
}

// ---------------------------------------------------------------------------
// Depthwise conv1d k=7 pad=3 over frames (per channel) fused with the LayerNorm that follows it
// in every ConvNeXt block of the path (modules.py:178-180; Vocos ConvNeXtBlock dwconv+norm).
// One warp per output frame; the 7 input rows are L1/L2 hits shared with neighbouring warps.
//   w: [C, 7] (Conv1d weight [C,1,7]), zero padding at sequence ends: rows outside [0,len) read 0.
// ---------------------------------------------------------------------------
struct DwLnArgs {
  const float* x;       // [rows, ldx]
  long long ldx;
  int rows_per_batch, nbatch;
  const int* seq_lens;  // conv sees zeros for t >= len (nullptr -> rows_per_batch)
  const float* w;       // [C,7]
  const float* wb;      // [C]
  const float* ln_w;    // [C]
  const float* ln_b;    // [C]
  float eps;
  __nv_bfloat16* out;   // [rows, ldo]
  long long ldo;
};

template <int C>
__global__ void __launch_bounds__(256) dwconv7_ln_kernel(const DwLnArgs a) {
  constexpr int VPL = C / 32;
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  const long long rows = (long long)a.rows_per_batch * a.nbatch;
  if (warp >= rows) return;
  const int b = warp / a.rows_per_batch;
  const int t = warp - b * a.rows_per_batch;
  const int len = a.seq_lens ? a.seq_lens[b] : a.rows_per_batch;
  float y[VPL];
#pragma unroll
  for (int i = 0; i < VPL; ++i) y[i] = __ldg(a.wb + lane + 32 * i);
#pragma unroll
  for (int k = 0; k < 7; ++k) {
    const int tt = t + k - 3;
    if (tt < 0 || tt >= len) continue;
    const float* xr = a.x + ((long long)b * a.rows_per_batch + tt) * a.ldx;
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
      const int c = lane + 32 * i;
      y[i] += __ldg(a.w + c * 7 + k) * xr[c];
    }
  }
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < VPL; ++i) s += y[i];
  const float mean = warp_sum(s) * (1.0f / C);
  float ss = 0.f;
#pragma unroll
  for (int i = 0; i < VPL; ++i) { y[i] -= mean; ss += y[i] * y[i]; }
  const float rstd = rsqrtf(warp_sum(ss) * (1.0f / C) + a.eps);
  __nv_bfloat16* o = a.out + (long long)warp * a.ldo;
#pragma unroll
  for (int i = 0; i < VPL; ++i) {
    const int c = lane + 32 * i;
    o[c] = __float2bfloat16(y[i] * rstd * __ldg(a.ln_w + c) + __ldg(a.ln_b + c));
  }
}

// Sliding-window variant (the one the launcher picks when rows are 16-byte aligned): a CTA of C/4 threads owns
// 4 channels per thread and walks a run of consecutive frames, keeping the 7-row window and the 7 taps of its
// channels in registers, so every input row is read from HBM once (+6 halo rows per run) and nothing is
// re-fetched through L1. Rows arrive through a shared-memory ring filled by 16-byte cp.async copies issued
// DW_AHEAD rows ahead (each thread copies exactly the 16 bytes it later reads, so the ring needs no barrier):
// ~24 KB in flight per CTA is what covers HBM latency at four or five CTAs per SM.
//
// The kernel is bound by instruction issue, not by HBM (ncu: issue 71 % at 44 % of HBM before this version), so the
// arithmetic is packed: channel pairs live in 64-bit registers and the taps, the statistics and the affine are
// FFMA2 / FMUL2 / FADD2; the 8-row window is a circular buffer addressed statically (the frame loop is unrolled over
// one full rotation: no register moves); the four LayerNorm sums of a frame pair (sum and sum of squares of both
// frames) are reduced together by one 6-shuffle butterfly and cross the C/128 warps through ONE barrier per pair
// (double-buffered slots). Variance is E[y^2] - mean^2 in fp32, clamped at 0 (the two-pass form costs a second
// butterfly and barrier; the difference is below the bf16 rounding of the output: tests/test_kernels_gpu.py).
constexpr int DW_AHEAD = 12;            // rows in flight per CTA (even)
constexpr int DW_RING = DW_AHEAD + 2;   // + the pair consumed in the previous iteration (write-after-read safety)

template <int N>
__device__ __forceinline__ void cp_async_wait_group() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

typedef unsigned long long dw_f2;  // two packed floats
__device__ __forceinline__ dw_f2 dw_pack(float lo, float hi) {
  dw_f2 r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ float2 dw_unpack(dw_f2 v) {
  float2 r;
  asm("mov.b64 {%0, %1}, %2;" : "=f"(r.x), "=f"(r.y) : "l"(v));
  return r;
}
__device__ __forceinline__ dw_f2 dw_fma(dw_f2 a, dw_f2 b, dw_f2 c) {
  dw_f2 r;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
  return r;
}
__device__ __forceinline__ dw_f2 dw_mul(dw_f2 a, dw_f2 b) {
  dw_f2 r;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ dw_f2 dw_add(dw_f2 a, dw_f2 b) {
  dw_f2 r;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}

template <int C>
__global__ void __launch_bounds__(C / 4) dwconv7_ln_run_kernel(const DwLnArgs a, int run) {
  constexpr int NW = C / 128;
  __shared__ __align__(16) float4 red[2][NW];  // (sum0, sum1, sumsq0, sumsq1) of a frame pair per warp
  __shared__ __align__(16) float4 ring[DW_RING][C / 4];
  const int c = 4 * threadIdx.x;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int runs_per_seq = (a.rows_per_batch + run - 1) / run;

  // taps of channels c..c+3: 28 consecutive floats of w[C,7], repacked as (c, c+1) and (c+2, c+3) pairs per tap
  dw_f2 wlo[7], whi[7];
  {
    const float4* wp = reinterpret_cast<const float4*>(a.w + (long long)c * 7);
    float raw[28];
#pragma unroll
    for (int i = 0; i < 7; ++i) {
      const float4 q = __ldg(wp + i);
      raw[4 * i] = q.x; raw[4 * i + 1] = q.y; raw[4 * i + 2] = q.z; raw[4 * i + 3] = q.w;
    }
#pragma unroll
    for (int k = 0; k < 7; ++k) {
      wlo[k] = dw_pack(raw[k], raw[7 + k]);
      whi[k] = dw_pack(raw[14 + k], raw[21 + k]);
    }
  }
  const float4 bias4 = __ldg(reinterpret_cast<const float4*>(a.wb + c));
  const float4 g4 = __ldg(reinterpret_cast<const float4*>(a.ln_w + c));
  const float4 be4 = __ldg(reinterpret_cast<const float4*>(a.ln_b + c));
  const dw_f2 bias_lo = dw_pack(bias4.x, bias4.y), bias_hi = dw_pack(bias4.z, bias4.w);
  const dw_f2 g_lo = dw_pack(g4.x, g4.y), g_hi = dw_pack(g4.z, g4.w);
  const dw_f2 be_lo = dw_pack(be4.x, be4.y), be_hi = dw_pack(be4.z, be4.w);
  const bool up16 = (lane & 16) != 0, up8 = (lane & 8) != 0;

  // persistent over (sequence, run) items: grid is sized to the resident CTA slots, so there is no partial last wave
  for (int item = blockIdx.x; item < runs_per_seq * a.nbatch; item += gridDim.x) {
    const int b = item / runs_per_seq;
    const int t0 = (item - b * runs_per_seq) * run;
    const int t1 = min(t0 + run, a.rows_per_batch);
    const int len = a.seq_lens ? a.seq_lens[b] : a.rows_per_batch;
    const float* xb = a.x + (long long)b * a.rows_per_batch * a.ldx + c;
    __nv_bfloat16* ob = a.out + (long long)b * a.rows_per_batch * a.ldo + c;
    const int t_last = min(len - 1, t1 + 3);  // last row anyone in this run reads
    // next ring pair slot to fill, the first of its two rows and that row's address; rows past t_last are zero-filled
    // by the copy itself (source size 0), so the consumer reads the ring unconditionally
    int issue_sl = 0, issue_t = t0 + 3;
    const float* issue_p = xb + (long long)issue_t * a.ldx;
    const long long ldx2 = 2 * a.ldx;
    auto issue_pair = [&]() {
      const uint32_t dst = (uint32_t)__cvta_generic_to_shared(&ring[issue_sl][threadIdx.x]);
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(issue_p), "r"(issue_t <= t_last ? 16 : 0) : "memory");
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst + (uint32_t)sizeof(float4) * (C / 4)),
                   "l"(issue_p + a.ldx), "r"(issue_t + 1 <= t_last ? 16 : 0) : "memory");
      asm volatile("cp.async.commit_group;" ::: "memory");
      issue_t += 2;
      issue_p += ldx2;
      issue_sl = issue_sl + 2 == DW_RING ? 0 : issue_sl + 2;
    };
#pragma unroll
    for (int q = 0; q < DW_AHEAD / 2; ++q) issue_pair();
    // circular 8-row window: slot (2j + k) % 8 holds row t - 3 + k of the j-th frame pair of a rotation
    dw_f2 wl[8], wh[8];
#pragma unroll
    for (int k = 0; k < 6; ++k) {
      const int tt = t0 - 3 + k;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (tt >= 0 && tt < len) v = *reinterpret_cast<const float4*>(xb + (long long)tt * a.ldx);
      wl[k] = dw_pack(v.x, v.y);
      wh[k] = dw_pack(v.z, v.w);
    }

    int read_sl = 0, par = 0;
    __nv_bfloat16* op = ob + (long long)t0 * a.ldo;
    for (int tb = t0; tb < t1; tb += 8) {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int t = tb + 2 * j;
        if (t >= t1) break;  // uniform over the CTA
        cp_async_wait_group<DW_AHEAD / 2 - 1>();
        {
          const float4 v6 = ring[read_sl][threadIdx.x], v7 = ring[read_sl + 1][threadIdx.x];
          wl[(2 * j + 6) & 7] = dw_pack(v6.x, v6.y); wh[(2 * j + 6) & 7] = dw_pack(v6.z, v6.w);
          wl[(2 * j + 7) & 7] = dw_pack(v7.x, v7.y); wh[(2 * j + 7) & 7] = dw_pack(v7.z, v7.w);
          read_sl = read_sl + 2 == DW_RING ? 0 : read_sl + 2;
        }
        issue_pair();  // overwrites the pair read one iteration ago
        dw_f2 y0l = bias_lo, y0h = bias_hi, y1l = bias_lo, y1h = bias_hi;
#pragma unroll
        for (int k = 0; k < 7; ++k) {
          y0l = dw_fma(wlo[k], wl[(2 * j + k) & 7], y0l);
          y0h = dw_fma(whi[k], wh[(2 * j + k) & 7], y0h);
          y1l = dw_fma(wlo[k], wl[(2 * j + k + 1) & 7], y1l);
          y1h = dw_fma(whi[k], wh[(2 * j + k + 1) & 7], y1h);
        }
        // (sum0, sum1, sumsq0, sumsq1) over the warp: exchange halves at lane bit 4, then bit 3, then a plain butterfly
        const float2 sa = dw_unpack(dw_add(y0l, y0h)), sb = dw_unpack(dw_add(y1l, y1h));
        const float2 qa = dw_unpack(dw_fma(y0l, y0l, dw_mul(y0h, y0h))), qb = dw_unpack(dw_fma(y1l, y1l, dw_mul(y1h, y1h)));
        const float s0 = sa.x + sa.y, s1 = sb.x + sb.y, q0 = qa.x + qa.y, q1 = qb.x + qb.y;
        float k0 = up16 ? q0 : s0, k1 = up16 ? q1 : s1;
        k0 += __shfl_xor_sync(0xffffffffu, up16 ? s0 : q0, 16);
        k1 += __shfl_xor_sync(0xffffffffu, up16 ? s1 : q1, 16);
        float kk = up8 ? k1 : k0;
        kk += __shfl_xor_sync(0xffffffffu, up8 ? k0 : k1, 8);
        kk += __shfl_xor_sync(0xffffffffu, kk, 4);
        kk += __shfl_xor_sync(0xffffffffu, kk, 2);
        kk += __shfl_xor_sync(0xffffffffu, kk, 1);
        if ((lane & 7) == 0) reinterpret_cast<float*>(&red[par][warp])[lane >> 3] = kk;
        __syncthreads();
        const ulonglong2* rp = reinterpret_cast<const ulonglong2*>(&red[par][0]);
        ulonglong2 acc = rp[0];
#pragma unroll
        for (int i = 1; i < NW; ++i) {
          const ulonglong2 r = rp[i];
          acc.x = dw_add(acc.x, r.x);
          acc.y = dw_add(acc.y, r.y);
        }
        const float2 ts = dw_unpack(acc.x), tq = dw_unpack(acc.y);
        const float4 tot = make_float4(ts.x, ts.y, tq.x, tq.y);
        par ^= 1;
        const float m0 = tot.x * (1.0f / C), m1 = tot.y * (1.0f / C);
        const float r0 = rsqrtf(fmaxf(tot.z * (1.0f / C) - m0 * m0, 0.f) + a.eps);
        const float r1 = rsqrtf(fmaxf(tot.w * (1.0f / C) - m1 * m1, 0.f) + a.eps);
        // (y - m) * r * g + be = y * (r g) + (be - m r g)
        const dw_f2 a0l = dw_mul(dw_pack(r0, r0), g_lo), a0h = dw_mul(dw_pack(r0, r0), g_hi);
        const dw_f2 a1l = dw_mul(dw_pack(r1, r1), g_lo), a1h = dw_mul(dw_pack(r1, r1), g_hi);
        const float2 o0l = dw_unpack(dw_fma(y0l, a0l, dw_fma(dw_pack(-m0, -m0), a0l, be_lo)));
        const float2 o0h = dw_unpack(dw_fma(y0h, a0h, dw_fma(dw_pack(-m0, -m0), a0h, be_hi)));
        const float2 o1l = dw_unpack(dw_fma(y1l, a1l, dw_fma(dw_pack(-m1, -m1), a1l, be_lo)));
        const float2 o1h = dw_unpack(dw_fma(y1h, a1h, dw_fma(dw_pack(-m1, -m1), a1h, be_hi)));
        *reinterpret_cast<uint2*>(op) = make_uint2(pack_bf16x2(o0l.x, o0l.y), pack_bf16x2(o0h.x, o0h.y));
        if (t + 1 < t1)
          *reinterpret_cast<uint2*>(op + a.ldo) = make_uint2(pack_bf16x2(o1l.x, o1l.y), pack_bf16x2(o1h.x, o1h.y));
        op += 2 * a.ldo;
      }
    }
    cp_async_wait_group<0>();
  }
}

// ---------------------------------------------------------------------------
// GRN (modules.py:153-156): gx[b,c] = ||h[b,:,c]||_2 over the frames of the sequence,
// nx = gx / (mean_c gx + 1e-6), y = gamma * (h * nx) + beta + h.   Two kernels:
//   grn_sumsq : sums of squares per (b, c) into gx2, fixed summation order
//   grn_apply : every block re-derives mean_c from gx2 (C <= 2048) and rewrites h in place (bf16)
// ---------------------------------------------------------------------------
// One block per (32 channels, batch element): 8 row lanes x 32 channels; every thread sums its rows in order, the eight
// row lanes are added in a fixed order -- no atomics, so the result (and with it a seeded CFM.sample) is bit-reproducible.
__global__ void __launch_bounds__(256)
grn_sumsq_kernel(const __nv_bfloat16* h, long long ldh, int rows_per_batch, int nb, const int* seq_lens,
                 int C, float* gx2) {
  __shared__ float part[8][33];
  const int b = blockIdx.y;
  const int len = seq_lens ? min(seq_lens[b], rows_per_batch) : rows_per_batch;
  const int cx = threadIdx.x & 31, ry = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + cx;
  float s = 0.f;
  if (c < C) {
    for (int t = ry; t < len; t += 8) {
      const float v = __bfloat162float(h[((long long)b * rows_per_batch + t) * ldh + c]);
      s += v * v;
    }
  }
  part[ry][cx] = s;
  __syncthreads();
  if (ry == 0 && c < C) {
    float tot = part[0][cx];
#pragma unroll
    for (int k = 1; k < 8; ++k) tot += part[k][cx];
    gx2[(long long)b * C + c] = tot;
  }
}

__global__ void __launch_bounds__(256)
grn_apply_kernel(__nv_bfloat16* h, long long ldh, int rows_per_batch, int nb, int C, int rows_per_block,
                 const float* gx2, const float* gamma, const float* beta) {
  __shared__ float red[8];
  __shared__ float s_mean;
  const int b = blockIdx.y;
  float s = 0.f;
  for (int c = threadIdx.x; c < C; c += blockDim.x) s += sqrtf(gx2[(long long)b * C + c]);
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    float tot = 0.f;
    for (int i = 0; i < (blockDim.x >> 5); ++i) tot += red[i];
    s_mean = tot / float(C);
  }
  __syncthreads();
  const float inv = 1.0f / (s_mean + 1e-6f);
  const int t0 = blockIdx.x * rows_per_block;
  const int t1 = min(t0 + rows_per_block, rows_per_batch);
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    const float nx = sqrtf(gx2[(long long)b * C + c]) * inv;
    const float g = gamma[c], be = beta[c];
    for (int t = t0; t < t1; ++t) {
      __nv_bfloat16* p = h + ((long long)b * rows_per_batch + t) * ldh + c;
      const float v = __bfloat162float(*p);
      *p = __float2bfloat16(g * (v * nx) + be + v);
    }
  }
}

}  // namespace oron
