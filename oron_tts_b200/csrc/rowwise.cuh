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
// ~24 KB in flight per CTA is what covers HBM latency at four CTAs per SM. Two frames per iteration; the
// LayerNorm statistics cross the C/128 warps through shared memory (mean, then centred variance: the same
// two-pass arithmetic as the warp-per-row kernel).
constexpr int DW_AHEAD = 12;            // rows in flight per CTA (even)
constexpr int DW_RING = DW_AHEAD + 2;   // + the pair consumed in the previous iteration (write-after-read safety)

template <int N>
__device__ __forceinline__ void cp_async_wait_group() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

template <int C>
__global__ void __launch_bounds__(C / 4) dwconv7_ln_run_kernel(const DwLnArgs a, int run) {
  constexpr int NW = C / 128;
  __shared__ float red[2][NW][2];
  __shared__ __align__(16) float4 ring[DW_RING][C / 4];
  const int c = 4 * threadIdx.x;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int runs_per_seq = (a.rows_per_batch + run - 1) / run;

  // taps of channels c..c+3: 28 consecutive floats of w[C,7]
  float wt[4][7];
  {
    const float4* wp = reinterpret_cast<const float4*>(a.w + (long long)c * 7);
    float raw[28];
#pragma unroll
    for (int i = 0; i < 7; ++i) {
      const float4 q = __ldg(wp + i);
      raw[4 * i] = q.x; raw[4 * i + 1] = q.y; raw[4 * i + 2] = q.z; raw[4 * i + 3] = q.w;
    }
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
      for (int k = 0; k < 7; ++k) wt[j][k] = raw[7 * j + k];
  }
  const float4 bias = __ldg(reinterpret_cast<const float4*>(a.wb + c));
  const float4 g = __ldg(reinterpret_cast<const float4*>(a.ln_w + c));
  const float4 be = __ldg(reinterpret_cast<const float4*>(a.ln_b + c));

  const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
  // persistent over (sequence, run) items: grid is sized to the resident CTA slots, so there is no partial last wave
  for (int item = blockIdx.x; item < runs_per_seq * a.nbatch; item += gridDim.x) {
  const int b = item / runs_per_seq;
  const int t0 = (item - b * runs_per_seq) * run;
  const int t1 = min(t0 + run, a.rows_per_batch);
  const int len = a.seq_lens ? a.seq_lens[b] : a.rows_per_batch;
  const float* xb = a.x + (long long)b * a.rows_per_batch * a.ldx + c;
  __nv_bfloat16* ob = a.out + (long long)b * a.rows_per_batch * a.ldo + c;
  const int t_last = min(len - 1, t1 + 3);  // last row anyone in this run reads
  auto issue_pair = [&](int q) {  // rows t0 + 2q + 3, +4 -> ring pair slot q % (DW_RING / 2)
    const int ts = t0 + 2 * q + 3;
    const int sl = 2 * (q % (DW_RING / 2));
    if (ts >= 0 && ts <= t_last) {
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(&ring[sl][threadIdx.x])),
                   "l"(xb + (long long)ts * a.ldx) : "memory");
    }
    if (ts + 1 >= 0 && ts + 1 <= t_last) {
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(&ring[sl + 1][threadIdx.x])),
                   "l"(xb + (long long)(ts + 1) * a.ldx) : "memory");
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
#pragma unroll
  for (int q = 0; q < DW_AHEAD / 2; ++q) issue_pair(q);
  float4 win[8];  // rows t-3 .. t+4 of the current frame pair; the first six come straight from global memory
#pragma unroll
  for (int k = 0; k < 6; ++k) {
    const int tt = t0 - 3 + k;
    win[k] = (tt >= 0 && tt < len) ? *reinterpret_cast<const float4*>(xb + (long long)tt * a.ldx) : zero4;
  }

  int p = 0;
  for (int t = t0; t < t1; t += 2, ++p) {
    cp_async_wait_group<DW_AHEAD / 2 - 1>();
    {
      const int sl = 2 * (p % (DW_RING / 2));
      win[6] = (t + 3 <= t_last) ? ring[sl][threadIdx.x] : zero4;
      win[7] = (t + 4 <= t_last) ? ring[sl + 1][threadIdx.x] : zero4;
    }
    issue_pair(p + DW_AHEAD / 2);  // overwrites the pair read one iteration ago
    float4 y0 = bias, y1 = bias;
#pragma unroll
    for (int k = 0; k < 7; ++k) {
      y0.x += wt[0][k] * win[k].x; y0.y += wt[1][k] * win[k].y; y0.z += wt[2][k] * win[k].z; y0.w += wt[3][k] * win[k].w;
      y1.x += wt[0][k] * win[k + 1].x; y1.y += wt[1][k] * win[k + 1].y; y1.z += wt[2][k] * win[k + 1].z; y1.w += wt[3][k] * win[k + 1].w;
    }
#pragma unroll
    for (int k = 0; k < 6; ++k) win[k] = win[k + 2];

    float s0 = warp_sum(y0.x + y0.y + y0.z + y0.w), s1 = warp_sum(y1.x + y1.y + y1.z + y1.w);
    if (lane == 0) { red[0][warp][0] = s0; red[0][warp][1] = s1; }
    __syncthreads();
    s0 = 0.f; s1 = 0.f;
#pragma unroll
    for (int i = 0; i < NW; ++i) { s0 += red[0][i][0]; s1 += red[0][i][1]; }
    const float m0 = s0 * (1.0f / C), m1 = s1 * (1.0f / C);
    y0.x -= m0; y0.y -= m0; y0.z -= m0; y0.w -= m0;
    y1.x -= m1; y1.y -= m1; y1.z -= m1; y1.w -= m1;
    float q0 = warp_sum(y0.x * y0.x + y0.y * y0.y + y0.z * y0.z + y0.w * y0.w);
    float q1 = warp_sum(y1.x * y1.x + y1.y * y1.y + y1.z * y1.z + y1.w * y1.w);
    if (lane == 0) { red[1][warp][0] = q0; red[1][warp][1] = q1; }
    __syncthreads();
    q0 = 0.f; q1 = 0.f;
#pragma unroll
    for (int i = 0; i < NW; ++i) { q0 += red[1][i][0]; q1 += red[1][i][1]; }
    const float r0 = rsqrtf(q0 * (1.0f / C) + a.eps), r1 = rsqrtf(q1 * (1.0f / C) + a.eps);
    uint2 o0, o1;
    o0.x = pack_bf16x2(y0.x * r0 * g.x + be.x, y0.y * r0 * g.y + be.y);
    o0.y = pack_bf16x2(y0.z * r0 * g.z + be.z, y0.w * r0 * g.w + be.w);
    o1.x = pack_bf16x2(y1.x * r1 * g.x + be.x, y1.y * r1 * g.y + be.y);
    o1.y = pack_bf16x2(y1.z * r1 * g.z + be.z, y1.w * r1 * g.w + be.w);
    *reinterpret_cast<uint2*>(ob + (long long)t * a.ldo) = o0;
    if (t + 1 < t1) *reinterpret_cast<uint2*>(ob + (long long)(t + 1) * a.ldo) = o1;
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
