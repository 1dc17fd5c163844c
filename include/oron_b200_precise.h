/*
 * oron_b200_precise.h — C ABI of the fp32 ("1e-4") mode helpers of the DiT forward (north_star: per-NFE-step velocity
 * within 1e-4 relative L2 of the fp32 reference). Same conventions as oron_b200.h.
 *
 * In this mode every activation stays fp32. The dense contractions still run on the tcgen05 GEMM (oron_gemm_bf16): an
 * fp32 operand is split into three bf16 terms (x = hi + mid + lo, 24 mantissa bits) and a Linear becomes six accumulating
 * passes (hi*hi, hi*mid, mid*hi, hi*lo, lo*hi, mid*mid) through the ORON_EPI_F32 / GATE_RESID / SCALE_RESID epilogues,
 * whose f32 outputs accumulate. What cannot be expressed that way is below. Slow by design (a parity / debugging mode).
 */
#ifndef ORON_B200_PRECISE_H_
#define ORON_B200_PRECISE_H_

#include "oron_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

/* x f32 [rows, ldx] columns [0, C) -> hi, mid, lo bf16 [rows, ldo]: hi = bf16(x), mid = bf16(x - hi), lo = bf16(x - hi - mid). */
int oron_split3_bf16(const float* x, int64_t ldx, int64_t rows, int32_t C, void* hi, void* mid, void* lo, int64_t ldo,
                     oron_stream_t stream);
/* SinusoidalEmbedding(256) in fp32 (modules.py:39-45) -> f32 [n, ldo]. */
int oron_time_sinusoid_f32(const float* t, int32_t n, float* out, int64_t ldo, oron_stream_t stream);
/* RoPE (modules.py:96-104) in place on the q and k columns [0, 2*heads*64) of f32 qkv [nbatch*rows_per_batch, ld]. */
int oron_rope_f32(float* qkv, int64_t ld, int32_t rows_per_batch, int32_t nbatch, int32_t heads, const float* rope_cos,
                  const float* rope_sin, oron_stream_t stream);
/* softmax(Q K^T scale + key mask) V in fp32 on the CUDA cores (modules.py:271-278); qkv f32 [rows, ld] (q | k | v),
 * out f32 [rows, ldo]; rows t >= seq_lens[b] are written as zeros. head_dim 64. */
int oron_attention_f32(const float* qkv, int64_t ld, float* out, int64_t ldo, int32_t nbatch, int32_t rows_per_batch,
                       int32_t heads, const int32_t* seq_lens, float scale, oron_stream_t stream);
/* GRN (modules.py:153-156) in place on f32 h [rows, ldh]; gx2: f32 [nb, C] workspace. */
int oron_grn_f32(float* h, int64_t ldh, int32_t rows_per_batch, int32_t nb, int32_t C, const int32_t* seq_lens,
                 const float* gamma, const float* beta, float* gx2, oron_stream_t stream);
/* out = act(in) with libm-accurate functions (act: oron_act / 4 = Mish / 0 = copy); rows t >= seq_lens[b] -> 0 when
 * seq_lens != NULL. f32 in / out, [rows, C] with leading dimensions. */
int oron_act_f32_precise(const float* in, int64_t ld_in, int64_t rows, int32_t C, int32_t act, float* out, int64_t ld_out,
                         int32_t rows_per_batch, const int32_t* seq_lens, oron_stream_t stream);
/* out[r, c] = a[r, c] + b[r, c] (f32, [rows, C], one leading dimension each). */
int oron_add_f32(const float* a, int64_t lda, const float* b, int64_t ldb, int64_t rows, int32_t C, float* out,
                 int64_t ldo, oron_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* ORON_B200_PRECISE_H_ */
