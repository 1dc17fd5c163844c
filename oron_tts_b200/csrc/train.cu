// C ABI (include/oron_b200_train.h) of the OT-CFM training-step kernels: argument checks and launches.
#include <cudaTypedefs.h>

#include <cstring>

#include "../../include/oron_b200_train.h"
#include "attn_bwd_tcgen05.cuh"
#include "gconv_wgrad_tcgen05.cuh"
#include "host_util.h"
#include "train_rowwise.cuh"

using namespace oron;

#define ST(s) reinterpret_cast<cudaStream_t>(s)

static inline DropCfg drop_cfg(float p, uint64_t seed) {
  DropCfg d;
  uint64_t z = seed + 0x9E3779B97F4A7C15ull;  // splitmix64: nearby seeds (per-layer offsets) give unrelated keys
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  z ^= z >> 31;
  d.key = (unsigned int)(z >> 32) ^ (unsigned int)z;
  if (!(p > 0.f)) { d.thresh = 0u; d.inv_keep = 1.f; return d; }
  if (p > 0.999f) p = 0.999f;
  d.thresh = (unsigned int)((double)p * 4294967296.0);
  if (d.thresh == 0u) d.thresh = 1u;
  d.inv_keep = 1.0f / (1.0f - p);
  return d;
}

static inline int ew_blocks(long long total, int per_block = 256) {
  long long b = (total + per_block - 1) / per_block;
  const long long cap = (long long)num_sms() * 16;
  if (b > cap) b = cap;
  if (b < 1) b = 1;
  return int(b);
}

extern "C" int oron_transpose_bf16(const void* in, int64_t ld_in, int32_t rows_per_batch, int32_t nbatch, int32_t C,
                                   const int32_t* seq_lens, void* out, int64_t ld_out, float* colsum,
                                   oron_stream_t stream) {
  if (!in || !out || rows_per_batch <= 0 || nbatch <= 0 || C <= 0) return fail(ORON_ERR_BAD_ARG, "transpose: bad argument");
  const long long R = (long long)rows_per_batch * nbatch;
  if ((C & 1) || (ld_in & 1) || (ld_out & 1) || (R & 1)) return fail(ORON_ERR_BAD_ARG, "transpose: C, rows and leading dimensions must be even");
  TransposeArgs a{reinterpret_cast<const __nv_bfloat16*>(in), ld_in, rows_per_batch, nbatch, C, seq_lens,
                  reinterpret_cast<__nv_bfloat16*>(out), ld_out, colsum};
  dim3 grid(unsigned((R + 63) / 64), unsigned((C + 63) / 64));
  transpose_bf16_kernel<<<grid, 256, 0, ST(stream)>>>(a);
  return check_launch("transpose_bf16");
}


extern "C" int oron_colsum_bf16(const void* in, int64_t ld_in, int64_t rows, int32_t C, float* out, oron_stream_t stream) {
  if (!in || !out || rows <= 0 || C <= 0 || (C & 1) || (ld_in & 1)) return fail(ORON_ERR_BAD_ARG, "colsum: bad argument");
  if ((C & 7) == 0 && (ld_in & 7) == 0 && (reinterpret_cast<uintptr_t>(in) & 15) == 0) {
    dim3 grid8(unsigned((C + 127) / 128), unsigned((rows + 255) / 256));
    colsum_bf16x8_kernel<<<grid8, 256, 0, ST(stream)>>>(reinterpret_cast<const __nv_bfloat16*>(in), ld_in, rows, C, out);
    return check_launch("colsum_bf16");
  }
  dim3 grid(unsigned((C + 63) / 64), unsigned((rows + 511) / 512));
  colsum_bf16_kernel<<<grid, 256, 0, ST(stream)>>>(reinterpret_cast<const __nv_bfloat16*>(in), ld_in, rows, C, out);
  return check_launch("colsum_bf16");
}

#define DISPATCH_V2(C, CALL)                                                                  \
  switch (C) {                                                                                \
    case 64: { constexpr int V2 = 1; CALL; break; }                                           \
    case 128: { constexpr int V2 = 2; CALL; break; }                                          \
    case 256: { constexpr int V2 = 4; CALL; break; }                                          \
    case 512: { constexpr int V2 = 8; CALL; break; }                                          \
    case 1024: { constexpr int V2 = 16; CALL; break; }                                        \
    default: return fail(ORON_ERR_UNSUPPORTED, "C = %d not in {64, 128, 256, 512, 1024}", C); \
  }

extern "C" int oron_ln_bwd(const float* x, int64_t ldx, const void* dy_bf16, int64_t lddy, int32_t rows_per_batch,
                           int32_t nbatch, int32_t C, float eps, const float* scale, int64_t mod_ld, int32_t add_one,
                           const int32_t* seq_lens, float* dx, int64_t lddx, int32_t accumulate, float* dscale,
                           float* dshift, int64_t dmod_ld, oron_stream_t stream) {
  if (!x || !dy_bf16 || !scale || !dx) return fail(ORON_ERR_BAD_ARG, "ln_bwd: null pointer");
  const int rpc = tr_rows_for((long long)rows_per_batch * nbatch, num_sms());
  LnBwdArgs a{x, ldx, reinterpret_cast<const __nv_bfloat16*>(dy_bf16), lddy, rows_per_batch, nbatch, eps, scale, mod_ld,
              add_one, seq_lens, dx, lddx, accumulate, dscale, dshift, dmod_ld, rpc};
  const bool vec4 = C % 128 == 0 && ldx % 4 == 0 && lddx % 4 == 0 && lddy % 4 == 0 && mod_ld % 4 == 0 && dmod_ld % 4 == 0 &&
                    ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(dx) | reinterpret_cast<uintptr_t>(scale) |
                      reinterpret_cast<uintptr_t>(dscale) | reinterpret_cast<uintptr_t>(dshift)) & 15) == 0 &&
                    (reinterpret_cast<uintptr_t>(dy_bf16) & 7) == 0;
  if (vec4) {
    dim3 g2(unsigned((rows_per_batch + 4 * TR2_ROWS - 1) / (4 * TR2_ROWS)), unsigned(nbatch));
    switch (C) {
      case 128: ln_bwd2_kernel<1><<<g2, 128, 0, ST(stream)>>>(a); break;
      case 256: ln_bwd2_kernel<2><<<g2, 128, 0, ST(stream)>>>(a); break;
      case 512: ln_bwd2_kernel<4><<<g2, 128, 0, ST(stream)>>>(a); break;
      case 1024: ln_bwd2_kernel<8><<<g2, 128, 0, ST(stream)>>>(a); break;
      default: return fail(ORON_ERR_UNSUPPORTED, "ln_bwd: C = %d", C);
    }
    return check_launch("ln_bwd");
  }
  dim3 grid(unsigned((rows_per_batch + rpc - 1) / rpc), unsigned(nbatch));
  DISPATCH_V2(C, (ln_bwd_kernel<V2><<<grid, 256, 0, ST(stream)>>>(a)));
  return check_launch("ln_bwd");
}

template <typename TI, typename TO>
static void launch_act_fwd(const void* in, int64_t ld_in, int64_t rows, int32_t C, int32_t act, void* out, int64_t ld_out,
                           int rpb, const int* sl, const DropCfg& dc, cudaStream_t st) {
  act_fwd_kernel<TI, TO><<<ew_blocks(rows * (C / 2)), 256, 0, st>>>(reinterpret_cast<const TI*>(in), ld_in, rows, C, act,
                                                                    reinterpret_cast<TO*>(out), ld_out, rpb, sl, dc);
}
extern "C" int oron_act_fwd(const void* in, int32_t in_f32, int64_t ld_in, int64_t rows, int32_t C, int32_t act, void* out,
                            int32_t out_f32, int64_t ld_out, int32_t rows_per_batch, const int32_t* seq_lens,
                            float dropout_p, uint64_t dropout_seed, oron_stream_t stream) {
  if (!in || !out || (C & 1) || (seq_lens && rows_per_batch <= 0)) return fail(ORON_ERR_BAD_ARG, "act_fwd: bad argument");
  if (rows <= 0) return 0;
  cudaStream_t st = ST(stream);
  const DropCfg dc = drop_cfg(dropout_p, dropout_seed);
  if (!in_f32 && !out_f32 && C % 8 == 0 && ld_in % 8 == 0 && ld_out % 8 == 0 &&
      ((reinterpret_cast<uintptr_t>(in) | reinterpret_cast<uintptr_t>(out)) & 15) == 0) {
    act_fwd_bf16x8_kernel<<<ew_blocks(rows * (C / 8)), 256, 0, st>>>(reinterpret_cast<const __nv_bfloat16*>(in), ld_in, rows, C, act,
                                                                    reinterpret_cast<__nv_bfloat16*>(out), ld_out, rows_per_batch,
                                                                    seq_lens, dc);
    return check_launch("act_fwd");
  }
  if (in_f32 && out_f32) launch_act_fwd<float, float>(in, ld_in, rows, C, act, out, ld_out, rows_per_batch, seq_lens, dc, st);
  else if (in_f32) launch_act_fwd<float, __nv_bfloat16>(in, ld_in, rows, C, act, out, ld_out, rows_per_batch, seq_lens, dc, st);
  else if (out_f32) launch_act_fwd<__nv_bfloat16, float>(in, ld_in, rows, C, act, out, ld_out, rows_per_batch, seq_lens, dc, st);
  else launch_act_fwd<__nv_bfloat16, __nv_bfloat16>(in, ld_in, rows, C, act, out, ld_out, rows_per_batch, seq_lens, dc, st);
  return check_launch("act_fwd");
}
template <typename TD, typename TP, typename TO>
static void launch_act_bwd(const void* dy, int64_t ld_dy, const void* pre, int64_t ld_pre, int64_t rows, int32_t C,
                           int32_t act, void* out, int64_t ld_out, int rpb, const int* sl, const DropCfg& dc, cudaStream_t st) {
  act_bwd_kernel<TD, TP, TO><<<ew_blocks(rows * (C / 2)), 256, 0, st>>>(
      reinterpret_cast<const TD*>(dy), ld_dy, reinterpret_cast<const TP*>(pre), ld_pre, rows, C, act,
      reinterpret_cast<TO*>(out), ld_out, rpb, sl, dc);
}
extern "C" int oron_act_bwd(const void* dy, int32_t dy_f32, int64_t ld_dy, const void* pre, int32_t pre_f32, int64_t ld_pre,
                            int64_t rows, int32_t C, int32_t act, void* out, int32_t out_f32, int64_t ld_out,
                            int32_t rows_per_batch, const int32_t* seq_lens, float dropout_p, uint64_t dropout_seed,
                            oron_stream_t stream) {
  if (!dy || !pre || !out || (C & 1) || (seq_lens && rows_per_batch <= 0)) return fail(ORON_ERR_BAD_ARG, "act_bwd: bad argument");
  if (rows <= 0) return 0;
  cudaStream_t st = ST(stream);
  const DropCfg dc = drop_cfg(dropout_p, dropout_seed);
  using bf = __nv_bfloat16;
  const int key = (dy_f32 ? 4 : 0) | (pre_f32 ? 2 : 0) | (out_f32 ? 1 : 0);
  if (key == 0 && C % 8 == 0 && ld_dy % 8 == 0 && ld_pre % 8 == 0 && ld_out % 8 == 0 &&
      ((reinterpret_cast<uintptr_t>(dy) | reinterpret_cast<uintptr_t>(pre) | reinterpret_cast<uintptr_t>(out)) & 15) == 0) {
    act_bwd_bf16x8_kernel<<<ew_blocks(rows * (C / 8)), 256, 0, st>>>(
        reinterpret_cast<const bf*>(dy), ld_dy, reinterpret_cast<const bf*>(pre), ld_pre, rows, C, act, reinterpret_cast<bf*>(out),
        ld_out, rows_per_batch, seq_lens, dc);
    return check_launch("act_bwd");
  }
  switch (key) {
    case 0: launch_act_bwd<bf, bf, bf>(dy, ld_dy, pre, ld_pre, rows, C, act, out, ld_out, rows_per_batch, seq_lens, dc, st); break;
    case 7: launch_act_bwd<float, float, float>(dy, ld_dy, pre, ld_pre, rows, C, act, out, ld_out, rows_per_batch, seq_lens, dc, st); break;
    case 1: launch_act_bwd<bf, bf, float>(dy, ld_dy, pre, ld_pre, rows, C, act, out, ld_out, rows_per_batch, seq_lens, dc, st); break;
    case 4: launch_act_bwd<float, bf, bf>(dy, ld_dy, pre, ld_pre, rows, C, act, out, ld_out, rows_per_batch, seq_lens, dc, st); break;
    case 5: launch_act_bwd<float, bf, float>(dy, ld_dy, pre, ld_pre, rows, C, act, out, ld_out, rows_per_batch, seq_lens, dc, st); break;
    default: return fail(ORON_ERR_UNSUPPORTED, "act_bwd: dtype combination %d not instantiated", key);
  }
  return check_launch("act_bwd");
}

extern "C" int oron_gate_resid(const float* x, int64_t ldx, const void* y_bf16, int64_t ldy, int32_t rows_per_batch,
                               int32_t nbatch, int32_t C, const float* gate, int64_t gate_ld, const int32_t* seq_lens,
                               int32_t mask_rows, float dropout_p, uint64_t dropout_seed, float* out, int64_t ldo,
                               oron_stream_t stream) {
  if (!x || !y_bf16 || !gate || !out || (C & 1)) return fail(ORON_ERR_BAD_ARG, "gate_resid: bad argument");
  const long long total = (long long)rows_per_batch * nbatch * (C / 2);
  gate_resid_kernel<<<ew_blocks(total), 256, 0, ST(stream)>>>(x, ldx, reinterpret_cast<const __nv_bfloat16*>(y_bf16), ldy,
                                                              rows_per_batch, nbatch, C, gate, gate_ld, seq_lens, mask_rows,
                                                              drop_cfg(dropout_p, dropout_seed), out, ldo);
  return check_launch("gate_resid");
}
extern "C" int oron_gate_bwd(const float* dx, int64_t lddx, const void* y_bf16, int64_t ldy, int32_t rows_per_batch,
                             int32_t nbatch, int32_t C, const float* gate, int64_t gate_ld, const int32_t* seq_lens,
                             void* dy_bf16, int64_t lddy, float* dgate, int64_t dgate_ld, float* dbias, float dropout_p,
                             uint64_t dropout_seed, oron_stream_t stream) {
  if (!dx || !y_bf16 || !gate || !dy_bf16) return fail(ORON_ERR_BAD_ARG, "gate_bwd: null pointer");
  const int rpc = tr_rows_for((long long)rows_per_batch * nbatch, num_sms());
  GateBwdArgs a{dx, lddx, reinterpret_cast<const __nv_bfloat16*>(y_bf16), ldy, rows_per_batch, nbatch, gate, gate_ld,
                seq_lens, reinterpret_cast<__nv_bfloat16*>(dy_bf16), lddy, dgate, dgate_ld, rpc, drop_cfg(dropout_p, dropout_seed), dbias};
  const bool vec4 = C % 128 == 0 && lddx % 4 == 0 && ldy % 4 == 0 && lddy % 4 == 0 && gate_ld % 4 == 0 && dgate_ld % 4 == 0 &&
                    ((reinterpret_cast<uintptr_t>(dx) | reinterpret_cast<uintptr_t>(gate) | reinterpret_cast<uintptr_t>(dgate) |
                      reinterpret_cast<uintptr_t>(dbias)) & 15) == 0 &&
                    ((reinterpret_cast<uintptr_t>(y_bf16) | reinterpret_cast<uintptr_t>(dy_bf16)) & 7) == 0;
  if (vec4) {
    dim3 g2(unsigned((rows_per_batch + 4 * TR2_ROWS - 1) / (4 * TR2_ROWS)), unsigned(nbatch));
    switch (C) {
      case 128: gate_bwd2_kernel<1><<<g2, 128, 0, ST(stream)>>>(a); break;
      case 256: gate_bwd2_kernel<2><<<g2, 128, 0, ST(stream)>>>(a); break;
      case 512: gate_bwd2_kernel<4><<<g2, 128, 0, ST(stream)>>>(a); break;
      case 1024: gate_bwd2_kernel<8><<<g2, 128, 0, ST(stream)>>>(a); break;
      default: return fail(ORON_ERR_UNSUPPORTED, "gate_bwd: C = %d", C);
    }
    return check_launch("gate_bwd");
  }
  dim3 grid(unsigned((rows_per_batch + rpc - 1) / rpc), unsigned(nbatch));
  DISPATCH_V2(C, (gate_bwd_kernel<V2><<<grid, 256, 0, ST(stream)>>>(a)));
  return check_launch("gate_bwd");
}

extern "C" int oron_dwconv7(const float* x, int64_t ldx, int32_t rows_per_batch, int32_t nbatch, int32_t C,
                            const int32_t* seq_lens, const float* w, const float* bias, int32_t flip, float* out,
                            int64_t ldo, int32_t accumulate, oron_stream_t stream) {
  if (!x || !w || !out) return fail(ORON_ERR_BAD_ARG, "dwconv7: null pointer");
  const long long total = (long long)rows_per_batch * nbatch * C;
  dwconv7_kernel<<<ew_blocks(total), 256, 0, ST(stream)>>>(x, ldx, rows_per_batch, nbatch, C, seq_lens, w, bias, flip, out,
                                                           ldo, accumulate);
  return check_launch("dwconv7");
}
extern "C" int oron_dwconv7_wgrad(const float* x, int64_t ldx, const float* dy, int64_t lddy, int32_t rows_per_batch,
                                  int32_t nbatch, int32_t C, const int32_t* seq_lens, float* dw, float* db,
                                  oron_stream_t stream) {
  if (!x || !dy || !dw) return fail(ORON_ERR_BAD_ARG, "dwconv7_wgrad: null pointer");
  dim3 grid(unsigned((rows_per_batch + 63) / 64), unsigned(nbatch), unsigned((C + 127) / 128));
  dwconv7_wgrad_kernel<<<grid, 128, 0, ST(stream)>>>(x, ldx, dy, lddy, rows_per_batch, nbatch, C, seq_lens, dw, db);
  return check_launch("dwconv7_wgrad");
}

extern "C" int oron_grn_bwd_reduce(const void* dy_bf16, int64_t lddy, const void* pre_bf16, int64_t ldpre,
                                   int32_t rows_per_batch, int32_t nb, int32_t C, const int32_t* seq_lens, float* A,
                                   float* dbeta, oron_stream_t stream) {
  if (!dy_bf16 || !pre_bf16 || !A || !dbeta) return fail(ORON_ERR_BAD_ARG, "grn_bwd_reduce: null pointer");
  GrnBwdArgs a;
  memset(&a, 0, sizeof(a));
  a.dy = reinterpret_cast<const __nv_bfloat16*>(dy_bf16);
  a.lddy = lddy;
  a.pre = reinterpret_cast<const __nv_bfloat16*>(pre_bf16);
  a.ldpre = ldpre;
  a.rows_per_batch = rows_per_batch;
  a.nb = nb;
  a.seq_lens = seq_lens;
  a.A = A;
  a.dbeta = dbeta;
  cudaError_t e = cudaMemsetAsync(A, 0, sizeof(float) * size_t(nb) * C, ST(stream));
  if (e != cudaSuccess) return fail(int(e), "grn_bwd memset: %s", cudaGetErrorString(e));
  a.rows_per_cta = tr_rows_for((long long)rows_per_batch * nb, num_sms());
  dim3 grid(unsigned((rows_per_batch + a.rows_per_cta - 1) / a.rows_per_cta), unsigned(nb));
  DISPATCH_V2(C, (grn_bwd_reduce_kernel<V2><<<grid, 256, 0, ST(stream)>>>(a)));
  return check_launch("grn_bwd_reduce");
}
extern "C" int oron_grn_bwd_coef(const float* A, const float* gx2, int32_t nb, int32_t C, const float* gamma, float* coef,
                                 float* nx, float* dgamma, oron_stream_t stream) {
  if (!A || !gx2 || !gamma || !coef || !nx || !dgamma) return fail(ORON_ERR_BAD_ARG, "grn_bwd_coef: null pointer");
  grn_bwd_coef_kernel<<<nb, 256, 0, ST(stream)>>>(A, gx2, nb, C, gamma, coef, nx, dgamma);
  return check_launch("grn_bwd_coef");
}
extern "C" int oron_grn_bwd_apply(const void* dy_bf16, int64_t lddy, const void* pre_bf16, int64_t ldpre,
                                  int32_t rows_per_batch, int32_t nb, int32_t C, const int32_t* seq_lens,
                                  const float* gamma, const float* nx, const float* coef, void* dpre_bf16, int64_t ldo,
                                  oron_stream_t stream) {
  if (!dy_bf16 || !pre_bf16 || !gamma || !nx || !coef || !dpre_bf16 || (C & 1))
    return fail(ORON_ERR_BAD_ARG, "grn_bwd_apply: bad argument");
  GrnBwdArgs a;
  memset(&a, 0, sizeof(a));
  a.dy = reinterpret_cast<const __nv_bfloat16*>(dy_bf16);
  a.lddy = lddy;
  a.pre = reinterpret_cast<const __nv_bfloat16*>(pre_bf16);
  a.ldpre = ldpre;
  a.rows_per_batch = rows_per_batch;
  a.nb = nb;
  a.seq_lens = seq_lens;
  a.gamma = gamma;
  a.nx = nx;
  a.coef = coef;
  a.dpre = reinterpret_cast<__nv_bfloat16*>(dpre_bf16);
  a.ldo = ldo;
  const long long total = (long long)rows_per_batch * nb * (C / 2);
  grn_bwd_apply_kernel<<<ew_blocks(total), 256, 0, ST(stream)>>>(a, C);
  return check_launch("grn_bwd_apply");
}

extern "C" int oron_text_embed_bwd(const int32_t* ids, const uint8_t* drop, const float* dx, int64_t lddx,
                                   int32_t rows_per_batch, int32_t nb, int32_t C, float* dtable, oron_stream_t stream) {
  if (!ids || !drop || !dx || !dtable) return fail(ORON_ERR_BAD_ARG, "text_embed_bwd: null pointer");
  const long long total = (long long)rows_per_batch * nb * C;
  text_embed_bwd_kernel<<<ew_blocks(total), 256, 0, ST(stream)>>>(ids, drop, dx, lddx, rows_per_batch, nb, C, dtable);
  return check_launch("text_embed_bwd");
}

extern "C" int oron_skinny_dgrad(const float* dY, int64_t lddy, int32_t nb, int32_t N, const void* W_bf16, int64_t ldw,
                                 int32_t K, float* dX, int64_t lddx, oron_stream_t stream) {
  if (!dY || !W_bf16 || !dX || nb <= 0 || N <= 0 || K <= 0 || (K & 1) || (ldw & 1))
    return fail(ORON_ERR_BAD_ARG, "skinny_dgrad: bad argument");
  if (nb <= 8 && (K & 7) == 0 && (ldw & 7) == 0 && (reinterpret_cast<uintptr_t>(W_bf16) & 15) == 0) {
    const int gy = (K + 511) / 512, chunks = (N + 255) / 256;
    int gx = (2 * num_sms() + gy - 1) / gy;
    if (gx > chunks) gx = chunks;
    dim3 grid8((unsigned)gx, (unsigned)gy);
    skinny_dgrad8_kernel<<<grid8, 256, 0, ST(stream)>>>(dY, lddy, nb, N, reinterpret_cast<const __nv_bfloat16*>(W_bf16), ldw, K, dX, lddx);
    return check_launch("skinny_dgrad");
  }
  dim3 grid(unsigned((N + 1023) / 1024), unsigned((K + 511) / 512));
  skinny_dgrad_kernel<<<grid, 256, 0, ST(stream)>>>(dY, lddy, nb, N, reinterpret_cast<const __nv_bfloat16*>(W_bf16), ldw, K,
                                                    dX, lddx);
  return check_launch("skinny_dgrad");
}
extern "C" int oron_skinny_wgrad(const float* dY, int64_t lddy, const float* X, int64_t ldx, int32_t nb, int32_t N,
                                 int32_t K, float* dW, int64_t lddw, float* db, int32_t accumulate, oron_stream_t stream) {
  if (!dY || !X || !dW || nb <= 0 || nb > 64 || N <= 0 || K <= 0) return fail(ORON_ERR_BAD_ARG, "skinny_wgrad: bad argument (nb <= 64)");
  dim3 grid(unsigned((N + 7) / 8), unsigned((K + 255) / 256));
  skinny_wgrad_kernel<<<grid, 256, 0, ST(stream)>>>(dY, lddy, X, ldx, nb, N, K, dW, lddw, db, accumulate);
  return check_launch("skinny_wgrad");
}

extern "C" int oron_gconv_wgrad(const void* x_bf16, int64_t ldx, const void* dy_bf16, int64_t lddy, int32_t rows_per_batch,
                                int32_t nbatch, int32_t C, int32_t cg, int32_t taps, const int32_t* seq_lens, float* dw,
                                float* db, oron_stream_t stream) {
  if (!x_bf16 || !dy_bf16 || !dw) return fail(ORON_ERR_BAD_ARG, "gconv_wgrad: null pointer");
  if (cg <= 0 || 64 % cg != 0 || C % cg != 0 || (C & 1) || taps <= 0)
    return fail(ORON_ERR_UNSUPPORTED, "gconv_wgrad: group width %d must divide 64", cg);
  GconvWgradArgs a{reinterpret_cast<const __nv_bfloat16*>(x_bf16), ldx, reinterpret_cast<const __nv_bfloat16*>(dy_bf16), lddy,
                   rows_per_batch, nbatch, C, cg, taps, seq_lens, dw, db};
  dim3 grid(unsigned(((C + 63) / 64) * taps), unsigned(nbatch));
  gconv_wgrad_kernel<<<grid, 256, 0, ST(stream)>>>(a);
  return check_launch("gconv_wgrad");
}

// tensor-core variant (gconv_wgrad_tcgen05.cuh): dw += ..., no bias gradient (use oron_colsum_bf16), rows beyond each
// sequence's length must already be zero in x and dy
extern "C" int oron_gconv_wgrad_tc(const void* x_bf16, int64_t ldx, const void* dy_bf16, int64_t lddy, int32_t rows_per_batch,
                                   int32_t nbatch, int32_t C, int32_t cg, int32_t taps, float* dw, oron_stream_t stream) {
  if (!x_bf16 || !dy_bf16 || !dw) return fail(ORON_ERR_BAD_ARG, "gconv_wgrad_tc: null pointer");
  if (cg <= 0 || 64 % cg != 0 || C % 128 != 0 || taps <= 0 || rows_per_batch <= 0 || rows_per_batch % 64 != 0 || nbatch <= 0)
    return fail(ORON_ERR_UNSUPPORTED, "gconv_wgrad_tc: needs C %% 128 == 0, a group width dividing 64 and rows_per_batch %% 64 == 0");
  CUtensorMap tx, tdy;
  int rc = make_tmap_bf16(&tx, x_bf16, uint64_t(C), uint64_t(rows_per_batch), uint64_t(nbatch), uint64_t(ldx),
                          uint64_t(ldx) * uint64_t(rows_per_batch), 64, 3);
  if (rc) return rc;
  rc = make_tmap_bf16(&tdy, dy_bf16, uint64_t(C), uint64_t(rows_per_batch), uint64_t(nbatch), uint64_t(lddy),
                      uint64_t(lddy) * uint64_t(rows_per_batch), 64, 3);
  if (rc) return rc;
  GconvTcArgs a;
  a.C = C; a.cg = cg; a.taps = taps; a.pad = taps / 2;
  a.kb_per_batch = rows_per_batch / 64;
  a.kb_total = a.kb_per_batch * nbatch;
  const int gx = C / 128, gy = (taps + GCW_TAPS - 1) / GCW_TAPS;
  int nchunk = (2 * num_sms() + gx * gy - 1) / (gx * gy);  // about two CTAs' worth of work per SM
  if (nchunk > a.kb_total) nchunk = a.kb_total;
  if (nchunk < 1) nchunk = 1;
  a.nchunk = nchunk;
  a.dw = dw;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(gconv_wgrad_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, GCW_SMEM_BYTES);
    if (e != cudaSuccess) return fail(int(e), "gconv_wgrad_tc smem attribute: %s", cudaGetErrorString(e));
    configured = true;
  }
  gconv_wgrad_tc_kernel<<<dim3(gx, gy, nchunk), GCW_THREADS, GCW_SMEM_BYTES, ST(stream)>>>(tx, tdy, a);
  return check_launch("gconv_wgrad_tc");
}

extern "C" int oron_cfm_loss(const float* pred, int64_t ldp, const float* flow, const uint8_t* span, const int32_t* count,
                             int64_t rows, int32_t n_mels, float* loss_sum, void* dpred_bf16, int64_t ldd,
                             oron_stream_t stream) {
  if (!pred || !flow || !span || !count || !loss_sum || !dpred_bf16 || ldd < n_mels)
    return fail(ORON_ERR_BAD_ARG, "cfm_loss: bad argument");
  cfm_loss_kernel<<<ew_blocks(rows * ldd), 256, 0, ST(stream)>>>(pred, ldp, flow, span, count, rows, n_mels, loss_sum,
                                                                 reinterpret_cast<__nv_bfloat16*>(dpred_bf16), ldd);
  return check_launch("cfm_loss");
}

extern "C" int oron_sumsq(const float* g, int64_t n, float* sumsq, oron_stream_t stream) {
  if (!g || !sumsq || n < 0) return fail(ORON_ERR_BAD_ARG, "sumsq: bad argument");
  if ((reinterpret_cast<uintptr_t>(g) & 15) != 0) return fail(ORON_ERR_BAD_ARG, "sumsq: arena must be 16-byte aligned");
  sumsq_kernel<<<ew_blocks(n / 4 + 1), 256, 0, ST(stream)>>>(g, n, sumsq);
  return check_launch("sumsq");
}
extern "C" int oron_adamw_clip(float* p, const float* g, float* m, float* v, void* p_bf16, int64_t n, const float* sumsq,
                               float grad_scale, float max_norm, float lr, float beta1, float beta2, float eps, float wd,
                               float bc1, float bc2, int32_t* skipped, oron_stream_t stream) {
  if (!p || !g || !m || !v || !sumsq || n < 0) return fail(ORON_ERR_BAD_ARG, "adamw_clip: bad argument");
  if (((reinterpret_cast<uintptr_t>(p) | reinterpret_cast<uintptr_t>(g) | reinterpret_cast<uintptr_t>(m) |
        reinterpret_cast<uintptr_t>(v)) & 15) != 0 || (reinterpret_cast<uintptr_t>(p_bf16) & 7) != 0)
    return fail(ORON_ERR_BAD_ARG, "adamw_clip: arenas must be 16-byte aligned (bf16 copy: 8)");
  AdamArgs a{p, g, m, v, reinterpret_cast<__nv_bfloat16*>(p_bf16), n, sumsq, grad_scale, max_norm, lr, beta1, beta2, eps, wd,
             bc1, bc2, skipped};
  adamw_clip_kernel<<<ew_blocks(n / 4 + 1), 256, 0, ST(stream)>>>(a);
  return check_launch("adamw_clip");
}
extern "C" int oron_mask_rows_f32(float* x, int64_t ldx, int64_t rows, int32_t C, const uint8_t* row_valid,
                                  oron_stream_t stream) {
  if (!x || !row_valid) return fail(ORON_ERR_BAD_ARG, "mask_rows: null pointer");
  mask_rows_kernel<<<ew_blocks(rows * C), 256, 0, ST(stream)>>>(x, ldx, rows, C, row_valid);
  return check_launch("mask_rows");
}
extern "C" int oron_f16_to_bf16(const void* in, int64_t ld_in, int64_t rows, int32_t C, void* out, int64_t ld_out,
                                oron_stream_t stream) {
  if (!in || !out || (C & 1) || (ld_in & 1) || (ld_out & 1)) return fail(ORON_ERR_BAD_ARG, "f16_to_bf16: bad argument");
  if ((C & 7) == 0 && (ld_in & 7) == 0 && (ld_out & 7) == 0 && ((reinterpret_cast<uintptr_t>(in) | reinterpret_cast<uintptr_t>(out)) & 15) == 0)
    f16_to_bf16x8_kernel<<<ew_blocks(rows * (C / 8)), 256, 0, ST(stream)>>>(reinterpret_cast<const __half*>(in), ld_in, rows, C,
                                                                            reinterpret_cast<__nv_bfloat16*>(out), ld_out);
  else
    f16_to_bf16_kernel<<<ew_blocks(rows * (C / 2)), 256, 0, ST(stream)>>>(reinterpret_cast<const __half*>(in), ld_in, rows, C,
                                                                          reinterpret_cast<__nv_bfloat16*>(out), ld_out);
  return check_launch("f16_to_bf16");
}

// ---------------------------------------------------------------------------------------------------------
// attention backward
// ---------------------------------------------------------------------------------------------------------
static long long* g_attn_bwd_dbg = nullptr;
// profiling aid (tools/attn_bwd_trace.py): int64 [2 * grid, 16] device buffer (dQ launch first, then dK/dV), or NULL
extern "C" void oron_debug_set_attention_bwd_stamps(void* buf) { g_attn_bwd_dbg = reinterpret_cast<long long*>(buf); }

extern "C" int oron_attention_bwd(const void* qk, int64_t ld_qk, const void* v, int64_t ld_v, const void* o, int64_t ld_o,
                                  const void* d_o, int64_t ld_do, void* dqkv, int64_t ld_dqkv, int32_t nbatch,
                                  int32_t rows_per_batch, int32_t heads, const int32_t* seq_lens, float scale,
                                  const float* rope_cos, const float* rope_sin, float* lse, float* delta, int32_t have_lse,
                                  oron_stream_t stream) {
  if (!qk || !v || !o || !d_o || !dqkv || !rope_cos || !rope_sin || !lse || !delta || nbatch <= 0 || rows_per_batch <= 0 ||
      heads <= 0)
    return fail(ORON_ERR_BAD_ARG, "attention_bwd: bad argument");
  if (ld_o % 8 || ld_do % 8 || ld_dqkv % 8) return fail(ORON_ERR_BAD_ARG, "attention_bwd: leading dimensions must be multiples of 8");
  const int HD = heads * AB_D;
  CUtensorMap tqk, tv, tdo;
  int rc = make_tmap_bf16(&tqk, qk, uint64_t(2 * HD), uint64_t(rows_per_batch), uint64_t(nbatch), uint64_t(ld_qk),
                          uint64_t(ld_qk) * uint64_t(rows_per_batch), AB_TILE, 3);
  if (rc) return rc;
  rc = make_tmap_bf16(&tv, v, uint64_t(HD), uint64_t(rows_per_batch), uint64_t(nbatch), uint64_t(ld_v),
                      uint64_t(ld_v) * uint64_t(rows_per_batch), AB_TILE, 3);
  if (rc) return rc;
  rc = make_tmap_bf16(&tdo, d_o, uint64_t(HD), uint64_t(rows_per_batch), uint64_t(nbatch), uint64_t(ld_do),
                      uint64_t(ld_do) * uint64_t(rows_per_batch), AB_TILE, 3);
  if (rc) return rc;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(attn_bwd_tcgen05_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, AB_SMEM_BYTES);
    if (e == cudaSuccess)
      e = cudaFuncSetAttribute(attn_bwd_tcgen05_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, AB_SMEM_BYTES);
    if (e != cudaSuccess) return fail(int(e), "attention_bwd smem attribute: %s", cudaGetErrorString(e));
    configured = true;
  }
  AttnBwdArgs a;
  memset(&a, 0, sizeof(a));
  a.rows_per_batch = rows_per_batch;
  a.nbatch = nbatch;
  a.heads = heads;
  a.tiles = (rows_per_batch + AB_TILE - 1) / AB_TILE;
  a.seq_lens = seq_lens;
  a.scale = scale;
  a.scale_log2 = scale * 1.4426950408889634f;
  a.o = reinterpret_cast<const __nv_bfloat16*>(o);
  a.ld_o = ld_o;
  a.d_o = reinterpret_cast<const __nv_bfloat16*>(d_o);
  a.ld_do = ld_do;
  a.dqkv = reinterpret_cast<__nv_bfloat16*>(dqkv);
  a.ld_dqkv = ld_dqkv;
  a.rope_cos = rope_cos;
  a.rope_sin = rope_sin;
  a.lse = lse;
  a.delta = delta;
  a.have_lse = have_lse ? 1 : 0;
  const unsigned grid = unsigned(a.tiles) * unsigned(heads) * unsigned(nbatch);
  cudaStream_t st = ST(stream);
  a.dbg = g_attn_bwd_dbg;
  attn_bwd_tcgen05_kernel<0><<<grid, AB_THREADS, AB_SMEM_BYTES, st>>>(tqk, tv, tdo, a);
  rc = check_launch("attn_bwd_dq");
  if (rc) return rc;
  if (a.dbg) a.dbg += (long long)grid * 16;
  attn_bwd_tcgen05_kernel<1><<<grid, AB_THREADS, AB_SMEM_BYTES, st>>>(tqk, tv, tdo, a);
  return check_launch("attn_bwd_dkdv");
}
