// fp32 ("1e-4") mode helpers (include/oron_b200_precise.h): everything of the DiT forward that is not a GEMM, in fp32 with
// libm-accurate transcendental functions, plus the 3-way bf16 split that lets the tcgen05 GEMM carry fp32 operands.
#include <cmath>
#include <cstring>

#include "../../include/oron_b200_precise.h"
#include "host_util.h"
#include "ptx.cuh"

using namespace oron;
#define ST(s) reinterpret_cast<cudaStream_t>(s)

static inline int pr_blocks(long long total) {
  long long b = (total + 255) / 256;
  const long long cap = (long long)num_sms() * 16;
  return int(b < 1 ? 1 : (b > cap ? cap : b));
}

namespace {

__global__ void __launch_bounds__(256) split3_kernel(const float* x, long long ldx, long long rows, int C, __nv_bfloat16* hi,
                                                     __nv_bfloat16* mid, __nv_bfloat16* lo, long long ldo) {
  const long long total = rows * C;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / C;
    const int c = int(i - r * C);
    const float v = x[r * ldx + c];
    const __nv_bfloat16 h = __float2bfloat16(v);
    const float r1 = v - __bfloat162float(h);
    const __nv_bfloat16 m = __float2bfloat16(r1);
    const float r2 = r1 - __bfloat162float(m);
    hi[r * ldo + c] = h;
    mid[r * ldo + c] = m;
    lo[r * ldo + c] = __float2bfloat16(r2);
  }
}

__global__ void time_sinusoid_f32_kernel(const float* t, int n, float* out, long long ldo) {
  const int i = blockIdx.x, j = threadIdx.x;  // 0..127
  if (i >= n) return;
  const float emb = expf(float(j) * (-9.210340371976184f / 127.0f));
  const float e = 1000.0f * t[i] * emb;
  out[(long long)i * ldo + j] = sinf(e);
  out[(long long)i * ldo + 128 + j] = cosf(e);
}

__global__ void __launch_bounds__(256) rope_f32_kernel(float* qkv, long long ld, int rows_per_batch, int nbatch, int heads,
                                                       const float* cs, const float* sn) {
  const long long total = (long long)rows_per_batch * nbatch * heads * 2 * 32;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int j = int(i & 31);
    const long long rest = i >> 5;
    const int hh = int(rest % (2 * heads));  // q heads then k heads: column block hh * 64
    const long long row = rest / (2 * heads);
    const int t = int(row % rows_per_batch);
    float* p = qkv + row * ld + hh * 64;
    const float a = p[j], b = p[j + 32];
    const float c = cs[(long long)t * 32 + j], s = sn[(long long)t * 32 + j];
    p[j] = a * c - b * s;
    p[j + 32] = b * c + a * s;
  }
}

// one thread per query row, 128 queries per CTA; K / V tiles of 32 keys staged in shared memory
constexpr int AF_Q = 128, AF_K = 32;
__global__ void __launch_bounds__(AF_Q) attention_f32_kernel(const float* qkv, long long ld, float* out, long long ldo,
                                                             int rows_per_batch, int heads, const int* seq_lens, float scale) {
  __shared__ float sk[AF_K][64];
  __shared__ float sv[AF_K][64];
  const int qt = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
  const int HD = heads * 64;
  const int len = seq_lens ? min(seq_lens[b], rows_per_batch) : rows_per_batch;
  const int t = qt * AF_Q + threadIdx.x;
  const long long base = (long long)b * rows_per_batch;
  const bool in_range = t < rows_per_batch;
  const bool valid = t < len;
  float q[64], o[64];
#pragma unroll
  for (int d = 0; d < 64; ++d) {
    q[d] = valid ? qkv[(base + t) * ld + h * 64 + d] * scale : 0.f;
    o[d] = 0.f;
  }
  float m = -INFINITY, l = 0.f;
  for (int k0 = 0; k0 < len; k0 += AF_K) {
    __syncthreads();
    for (int i = threadIdx.x; i < AF_K * 64; i += AF_Q) {
      const int kk = i >> 6, d = i & 63;
      const bool ok = k0 + kk < len;
      sk[kk][d] = ok ? qkv[(base + k0 + kk) * ld + HD + h * 64 + d] : 0.f;
      sv[kk][d] = ok ? qkv[(base + k0 + kk) * ld + 2 * HD + h * 64 + d] : 0.f;
    }
    __syncthreads();
    const int nk = min(AF_K, len - k0);
    float s[AF_K];
    float tm = -INFINITY;
#pragma unroll
    for (int kk = 0; kk < AF_K; ++kk) {
      float acc = 0.f;
#pragma unroll
      for (int d = 0; d < 64; ++d) acc = fmaf(q[d], sk[kk][d], acc);
      s[kk] = kk < nk ? acc : -INFINITY;
      tm = fmaxf(tm, s[kk]);
    }
    const float mn = fmaxf(m, tm);
    const float f = expf(m - mn);  // exp(-inf) = 0 on the first tile
    l *= f;
#pragma unroll
    for (int d = 0; d < 64; ++d) o[d] *= f;
#pragma unroll
    for (int kk = 0; kk < AF_K; ++kk) {
      const float p = expf(s[kk] - mn);  // masked keys: exp(-inf) = 0
      l += p;
#pragma unroll
      for (int d = 0; d < 64; ++d) o[d] = fmaf(p, sv[kk][d], o[d]);
    }
    m = mn;
  }
  if (in_range) {
    const float inv = valid ? 1.0f / l : 0.f;
#pragma unroll
    for (int d = 0; d < 64; ++d) out[(base + t) * ldo + h * 64 + d] = valid ? o[d] * inv : 0.f;
  }
}

__global__ void __launch_bounds__(256) grn_f32_sumsq_kernel(const float* h, long long ldh, int rows_per_batch, const int* seq_lens,
                                                            int C, int rows_per_block, float* gx2) {
  const int b = blockIdx.y;
  const int len = seq_lens ? min(seq_lens[b], rows_per_batch) : rows_per_batch;
  const int t0 = blockIdx.x * rows_per_block, t1 = min(t0 + rows_per_block, len);
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float s = 0.f;
    for (int t = t0; t < t1; ++t) {
      const float v = h[((long long)b * rows_per_batch + t) * ldh + c];
      s += v * v;
    }
    if (t1 > t0) atomicAdd(gx2 + (long long)b * C + c, s);
  }
}
__global__ void __launch_bounds__(256) grn_f32_apply_kernel(float* h, long long ldh, int rows_per_batch, int C, int rows_per_block,
                                                            const float* gx2, const float* gamma, const float* beta) {
  __shared__ float red[8];
  __shared__ float s_mean;
  const int b = blockIdx.y;
  float s = 0.f;
  for (int c = threadIdx.x; c < C; c += blockDim.x) s += sqrtf(gx2[(long long)b * C + c]);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    float tot = 0.f;
    for (int i = 0; i < 8; ++i) tot += red[i];
    s_mean = tot / float(C);
  }
  __syncthreads();
  const float inv = 1.0f / (s_mean + 1e-6f);
  const int t0 = blockIdx.x * rows_per_block, t1 = min(t0 + rows_per_block, rows_per_batch);
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    const float nx = sqrtf(gx2[(long long)b * C + c]) * inv;
    const float g = gamma[c], be = beta[c];
    for (int t = t0; t < t1; ++t) {
      float* p = h + ((long long)b * rows_per_batch + t) * ldh + c;
      const float v = *p;
      *p = g * (v * nx) + be + v;
    }
  }
}

__device__ __forceinline__ float act_precise(int act, float x) {
  switch (act) {
    case 1: return 0.5f * x * (1.0f + tanhf(0.7978845608028654f * (x + 0.044715f * x * x * x)));
    case 2: return 0.5f * x * (1.0f + erff(x * 0.7071067811865476f));
    case 3: return x / (1.0f + expf(-x));
    case 4: return x * tanhf(x > 20.0f ? x : log1pf(expf(x)));
    default: return x;
  }
}
__global__ void __launch_bounds__(256) act_f32_precise_kernel(const float* in, long long ld_in, long long rows, int C, int act,
                                                              float* out, long long ld_out, int rows_per_batch,
                                                              const int* seq_lens) {
  const long long total = rows * C;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / C;
    const int c = int(i - r * C);
    float v = 0.f;
    if (seq_lens == nullptr || int(r % rows_per_batch) < seq_lens[r / rows_per_batch]) v = act_precise(act, in[r * ld_in + c]);
    out[r * ld_out + c] = v;
  }
}
__global__ void __launch_bounds__(256) add_f32_kernel(const float* a, long long lda, const float* b, long long ldb, long long rows,
                                                      int C, float* out, long long ldo) {
  const long long total = rows * C;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / C;
    const int c = int(i - r * C);
    out[r * ldo + c] = a[r * lda + c] + b[r * ldb + c];
  }
}

}  // namespace

extern "C" int oron_split3_bf16(const float* x, int64_t ldx, int64_t rows, int32_t C, void* hi, void* mid, void* lo,
                                int64_t ldo, oron_stream_t stream) {
  if (!x || !hi || !mid || !lo || rows <= 0 || C <= 0) return fail(ORON_ERR_BAD_ARG, "split3: bad argument");
  split3_kernel<<<pr_blocks(rows * C), 256, 0, ST(stream)>>>(x, ldx, rows, C, reinterpret_cast<__nv_bfloat16*>(hi),
                                                           reinterpret_cast<__nv_bfloat16*>(mid),
                                                           reinterpret_cast<__nv_bfloat16*>(lo), ldo);
  return check_launch("split3");
}
extern "C" int oron_time_sinusoid_f32(const float* t, int32_t n, float* out, int64_t ldo, oron_stream_t stream) {
  if (!t || !out || n <= 0) return fail(ORON_ERR_BAD_ARG, "time_sinusoid_f32: bad argument");
  time_sinusoid_f32_kernel<<<n, 128, 0, ST(stream)>>>(t, n, out, ldo);
  return check_launch("time_sinusoid_f32");
}
extern "C" int oron_rope_f32(float* qkv, int64_t ld, int32_t rows_per_batch, int32_t nbatch, int32_t heads,
                             const float* rope_cos, const float* rope_sin, oron_stream_t stream) {
  if (!qkv || !rope_cos || !rope_sin) return fail(ORON_ERR_BAD_ARG, "rope_f32: null pointer");
  const long long total = (long long)rows_per_batch * nbatch * heads * 2 * 32;
  rope_f32_kernel<<<pr_blocks(total), 256, 0, ST(stream)>>>(qkv, ld, rows_per_batch, nbatch, heads, rope_cos, rope_sin);
  return check_launch("rope_f32");
}
extern "C" int oron_attention_f32(const float* qkv, int64_t ld, float* out, int64_t ldo, int32_t nbatch,
                                  int32_t rows_per_batch, int32_t heads, const int32_t* seq_lens, float scale,
                                  oron_stream_t stream) {
  if (!qkv || !out || nbatch <= 0 || rows_per_batch <= 0 || heads <= 0) return fail(ORON_ERR_BAD_ARG, "attention_f32: bad argument");
  dim3 grid(unsigned((rows_per_batch + AF_Q - 1) / AF_Q), unsigned(heads), unsigned(nbatch));
  attention_f32_kernel<<<grid, AF_Q, 0, ST(stream)>>>(qkv, ld, out, ldo, rows_per_batch, heads, seq_lens, scale);
  return check_launch("attention_f32");
}
extern "C" int oron_grn_f32(float* h, int64_t ldh, int32_t rows_per_batch, int32_t nb, int32_t C, const int32_t* seq_lens,
                            const float* gamma, const float* beta, float* gx2, oron_stream_t stream) {
  if (!h || !gamma || !beta || !gx2) return fail(ORON_ERR_BAD_ARG, "grn_f32: null pointer");
  cudaError_t e = cudaMemsetAsync(gx2, 0, sizeof(float) * size_t(nb) * C, ST(stream));
  if (e != cudaSuccess) return fail(int(e), "grn_f32 memset: %s", cudaGetErrorString(e));
  const int rpb = 32;
  dim3 grid(unsigned((rows_per_batch + rpb - 1) / rpb), unsigned(nb));
  grn_f32_sumsq_kernel<<<grid, 256, 0, ST(stream)>>>(h, ldh, rows_per_batch, seq_lens, C, rpb, gx2);
  int rc = check_launch("grn_f32_sumsq");
  if (rc) return rc;
  grn_f32_apply_kernel<<<grid, 256, 0, ST(stream)>>>(h, ldh, rows_per_batch, C, rpb, gx2, gamma, beta);
  return check_launch("grn_f32_apply");
}
extern "C" int oron_act_f32_precise(const float* in, int64_t ld_in, int64_t rows, int32_t C, int32_t act, float* out,
                                    int64_t ld_out, int32_t rows_per_batch, const int32_t* seq_lens, oron_stream_t stream) {
  if (!in || !out || (seq_lens && rows_per_batch <= 0)) return fail(ORON_ERR_BAD_ARG, "act_f32_precise: bad argument");
  if (rows <= 0) return 0;
  act_f32_precise_kernel<<<pr_blocks(rows * C), 256, 0, ST(stream)>>>(in, ld_in, rows, C, act, out, ld_out, rows_per_batch,
                                                                    seq_lens);
  return check_launch("act_f32_precise");
}
extern "C" int oron_add_f32(const float* a, int64_t lda, const float* b, int64_t ldb, int64_t rows, int32_t C, float* out,
                            int64_t ldo, oron_stream_t stream) {
  if (!a || !b || !out) return fail(ORON_ERR_BAD_ARG, "add_f32: null pointer");
  if (rows <= 0) return 0;
  add_f32_kernel<<<pr_blocks(rows * C), 256, 0, ST(stream)>>>(a, lda, b, ldb, rows, C, out, ldo);
  return check_launch("add_f32");
}
