/*
 * oron_b200_train.h — C ABI of the B200 (sm_100a) kernels behind the OT-CFM training step
 * (CFM.forward flow.py:69-159, autograd backward of DiT.forward, clip + AdamW of
 * F5Trainer._optimizer_step trainer.py:191-216). Same conventions as oron_b200.h: device pointers, caller-owned
 * buffers, enqueue-only on `stream`, 0 on success. The reference has no FFI: each entry point replaces the
 * autograd node(s) torch would run for the cited forward lines.
 *
 * The dense contractions of the backward pass (dgrad: dX = dY W, wgrad: dW = dY^T X) run on oron_gemm_bf16 with
 * transposed operand copies produced by oron_transpose_bf16; everything else is below.
 */
#ifndef ORON_B200_TRAIN_H_
#define ORON_B200_TRAIN_H_

#include "oron_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

/* bf16 [nbatch*rows_per_batch, ld_in] (columns [0, C)) -> bf16 [C, ld_out] (row c = column c of the input over all
 * rows). Rows t >= seq_lens[b] are written as zeros when seq_lens != NULL. colsum (f32 [C], optional) += column
 * sums of the (masked) input: the bias gradient of the Linear whose output gradient is being transposed. */
int oron_transpose_bf16(const void* in, int64_t ld_in, int32_t rows_per_batch, int32_t nbatch, int32_t C,
                        const int32_t* seq_lens, void* out, int64_t ld_out, float* colsum, oron_stream_t stream);

/* out[c] += sum over rows of in[r, c] (bf16 [rows, ld_in], columns [0, C)): the bias gradient of a Linear from its
 * output gradient, when the weight gradient is taken straight from the row-major operands (a_mn_major / b_mn_major). */
int oron_colsum_bf16(const void* in, int64_t ld_in, int64_t rows, int32_t C, float* out, oron_stream_t stream);

/* Backward of oron_ln_modulate (modules.py:218, 234, 341 and the affine LayerNorms modules.py:169):
 *   y = LN(x) * (add_one + scale[b]) + shift[b]
 *   dx (+)= LN-backward(dy * (add_one + scale)) ; dscale[b or 0] += sum_t dy * xhat ; dshift += sum_t dy
 * x f32 [rows, ldx], dy bf16 [rows, lddy]; scale address scale + b*mod_ld (mod_ld = 0: per-channel affine weight);
 * dscale/dshift address + b*dmod_ld. Rows t >= seq_lens[b] contribute nothing (dx row zeroed unless accumulate).
 * C in {64, 128, 256, 512, 1024}. */
int oron_ln_bwd(const float* x, int64_t ldx, const void* dy_bf16, int64_t lddy, int32_t rows_per_batch,
                int32_t nbatch, int32_t C, float eps, const float* scale, int64_t mod_ld, int32_t add_one,
                const int32_t* seq_lens, float* dx, int64_t lddx, int32_t accumulate, float* dscale,
                float* dshift, int64_t dmod_ld, oron_stream_t stream);

/* Elementwise activation, forward (out = act(in)) and backward (out = dy * act'(pre)); act = oron_act or
 * ORON_ACT_MISH (= 4). in/pre/dy/out are [rows, C] with leading dimensions; *_f32 flags select f32 (1) or bf16 (0).
 * With seq_lens != NULL rows t >= seq_lens[row / rows_per_batch] are written as zeros (the masks of
 * ConvPositionEmbedding, modules.py:136-140).
 * dropout_p > 0: nn.Dropout after the activation (FeedForward, modules.py:297) -- element (r, c) survives iff
 * hash(dropout_seed, r * C + c) >= p * 2^32 and is scaled by 1 / (1 - p); the backward recomputes the same mask. */
#define ORON_ACT_MISH 4
int oron_act_fwd(const void* in, int32_t in_f32, int64_t ld_in, int64_t rows, int32_t C, int32_t act, void* out,
                 int32_t out_f32, int64_t ld_out, int32_t rows_per_batch, const int32_t* seq_lens, float dropout_p,
                 uint64_t dropout_seed, oron_stream_t stream);
int oron_act_bwd(const void* dy, int32_t dy_f32, int64_t ld_dy, const void* pre, int32_t pre_f32, int64_t ld_pre,
                 int64_t rows, int32_t C, int32_t act, void* out, int32_t out_f32, int64_t ld_out,
                 int32_t rows_per_batch, const int32_t* seq_lens, float dropout_p, uint64_t dropout_seed,
                 oron_stream_t stream);

/* Gated residual of DiTBlock (modules.py:338, 343) un-fused for training:
 *   fwd: out[r, :] = x[r, :] + gate[b, :] * y[r, :]   (rows t >= seq_lens[b]: y taken as 0 when mask_rows,
 *        modules.py:281-282); out may be x (in place) or the next saved residual-stream buffer
 *   bwd: dy[r, :] = gate[b, :] * dx[r, :] (bf16; zero rows beyond seq_lens) ; dgate[b, :] += sum_t dx * y ;
 *        dbias[:] += sum over all rows of dy (optional: the bias gradient of the Linear that produced y)
 * dropout_p > 0: y is first passed through nn.Dropout (Attention.to_out[1], modules.py:253), same stateless mask as above. */
int oron_gate_resid(const float* x, int64_t ldx, const void* y_bf16, int64_t ldy, int32_t rows_per_batch, int32_t nbatch,
                    int32_t C, const float* gate, int64_t gate_ld, const int32_t* seq_lens, int32_t mask_rows,
                    float dropout_p, uint64_t dropout_seed, float* out, int64_t ldo, oron_stream_t stream);
int oron_gate_bwd(const float* dx, int64_t lddx, const void* y_bf16, int64_t ldy, int32_t rows_per_batch,
                  int32_t nbatch, int32_t C, const float* gate, int64_t gate_ld, const int32_t* seq_lens,
                  void* dy_bf16, int64_t lddy, float* dgate, int64_t dgate_ld, float* dbias, float dropout_p,
                  uint64_t dropout_seed, oron_stream_t stream);

/* Depthwise Conv1d(k=7, pad=3, groups=C) over frames, un-fused (modules.py:178; backward data path with flip=1):
 *   out[t, c] (+)= bias[c] + sum_k w[c, flip ? 6-k : k] * x[t + k - 3, c], x = 0 outside [0, seq_lens[b]).
 * and its weight/bias gradient: dw[c, k] += sum_t dy[t, c] * x[t + k - 3, c] ; db[c] += sum_t dy[t, c]. */
int oron_dwconv7(const float* x, int64_t ldx, int32_t rows_per_batch, int32_t nbatch, int32_t C,
                 const int32_t* seq_lens, const float* w, const float* bias, int32_t flip, float* out, int64_t ldo,
                 int32_t accumulate, oron_stream_t stream);
int oron_dwconv7_wgrad(const float* x, int64_t ldx, const float* dy, int64_t lddy, int32_t rows_per_batch,
                       int32_t nbatch, int32_t C, const int32_t* seq_lens, float* dw, float* db,
                       oron_stream_t stream);

/* Backward of GELU(erf) -> GRN (modules.py:153-156, 182-183) of ConvNeXtV2Block, from the saved pre-activation:
 *   h = gelu(pre) ; y = gamma * h * nx + beta + h ; nx[b, c] = gx / (mean_c gx + 1e-6), gx = ||h[b, :, c]||_2 over frames
 * Pass 1 (reduce): A[b, c] = sum_t dy * h, dbeta[c] += sum dy. Pass 2 (apply), after oron_grn_bwd_coef turned A and
 * the forward's gx2 into coef[b, c] (and accumulated dgamma): dpre = (dy * (gamma * nx + 1) + coef * h) * gelu'(pre). */
int oron_grn_bwd_reduce(const void* dy_bf16, int64_t lddy, const void* pre_bf16, int64_t ldpre,
                        int32_t rows_per_batch, int32_t nb, int32_t C, const int32_t* seq_lens, float* A,
                        float* dbeta, oron_stream_t stream);
int oron_grn_bwd_coef(const float* A, const float* gx2, int32_t nb, int32_t C, const float* gamma, float* coef,
                      float* nx, float* dgamma, oron_stream_t stream);
int oron_grn_bwd_apply(const void* dy_bf16, int64_t lddy, const void* pre_bf16, int64_t ldpre,
                       int32_t rows_per_batch, int32_t nb, int32_t C, const int32_t* seq_lens, const float* gamma,
                       const float* nx, const float* coef, void* dpre_bf16, int64_t ldo, oron_stream_t stream);

/* Backward of the TextEmbedding front end (encoder.py:68-91): dtable[id] += dx[row] for non-filler rows (id = 0 when
 * the text of the batch element was dropped), ids already +1 shifted. */
int oron_text_embed_bwd(const int32_t* ids, const uint8_t* drop, const float* dx, int64_t lddx,
                        int32_t rows_per_batch, int32_t nb, int32_t C, float* dtable, oron_stream_t stream);

/* Skinny (M = nb <= 64 rows) f32 matmuls of the timestep conditioning path (modules.py:60-62, 214, 232):
 *   dgrad: dX[b, k] (+)= sum_n dY[b, n] * W[n, k]    W bf16 [N, ldw]
 *   wgrad: dW[n, k] (+)= sum_b dY[b, n] * X[b, k] ;  db[n] (+)= sum_b dY[b, n] (optional) */
int oron_skinny_dgrad(const float* dY, int64_t lddy, int32_t nb, int32_t N, const void* W_bf16, int64_t ldw,
                      int32_t K, float* dX, int64_t lddx, oron_stream_t stream);
int oron_skinny_wgrad(const float* dY, int64_t lddy, const float* X, int64_t ldx, int32_t nb, int32_t N, int32_t K,
                      float* dW, int64_t lddw, float* db, int32_t accumulate, oron_stream_t stream);

/* Grouped Conv1d(k = taps, pad = taps/2) weight gradient of ConvPositionEmbedding (modules.py:120-124):
 *   dw[co, ci, k] += sum_{b,t} dy[b, t, co] * x[b, t + k - pad, g(co)*cg + ci]   (x = 0 outside [0, seq_lens[b]))
 *   db[co] += sum dy. x, dy bf16 [rows, C]; dw f32 [C, cg, taps] (the nn.Conv1d layout). cg must divide 64. */
int oron_gconv_wgrad(const void* x_bf16, int64_t ldx, const void* dy_bf16, int64_t lddy, int32_t rows_per_batch,
                     int32_t nbatch, int32_t C, int32_t cg, int32_t taps, const int32_t* seq_lens, float* dw,
                     float* db, oron_stream_t stream);
/* Same gradient on the tensor core (tcgen05, operands read MN-major as stored; csrc/gconv_wgrad_tcgen05.cuh): dw += ...
 * Needs C % 128 == 0, cg dividing 64, rows_per_batch % 64 == 0, and rows t >= seq_len of x and dy already ZERO (no seq_lens
 * argument: the forward masks x, oron_act_bwd zeroes dy). No bias gradient: use oron_colsum_bf16 on dy. */
int oron_gconv_wgrad_tc(const void* x_bf16, int64_t ldx, const void* dy_bf16, int64_t lddy, int32_t rows_per_batch,
                        int32_t nbatch, int32_t C, int32_t cg, int32_t taps, float* dw, oron_stream_t stream);


/* Masked MSE of CFM.forward (flow.py:156-159) and its gradient:
 *   loss_sum += sum_{rows with span[r]} sum_c (pred - flow)^2 ;  dpred[r, c] = span[r] ? 2 (pred - flow) / (count * n_mels) : 0
 * pred f32 [rows, ldp]; flow f32 [rows, n_mels]; span uint8 [rows]; count: device int32 (number of span rows);
 * dpred bf16 [rows, ldd] (columns [n_mels, ldd) are zeroed: K padding of the GEMMs that consume it). */
int oron_cfm_loss(const float* pred, int64_t ldp, const float* flow, const uint8_t* span, const int32_t* count,
                  int64_t rows, int32_t n_mels, float* loss_sum, void* dpred_bf16, int64_t ldd, oron_stream_t stream);

/* sumsq += sum g^2 (global gradient norm, trainer.py:171-177) over a flat f32 arena. */
int oron_sumsq(const float* g, int64_t n, float* sumsq, oron_stream_t stream);
/* Fused clip_grad_norm_ + AdamW (trainer.py:76-80, 206-211; torch.optim.AdamW semantics) over flat arenas:
 *   coef = min(1, max_norm / (sqrt(*sumsq) * grad_scale + 1e-6)) ; g = g * grad_scale * coef
 *   p *= 1 - lr * wd ; m = b1 m + (1 - b1) g ; v = b2 v + (1 - b2) g^2 ; p -= lr / bc1 * m / (sqrt(v) / sqrt(bc2) + eps)
 * and the bf16 copy pb = bf16(p). A non-finite norm skips the update (trainer.py:195-204) and sets *skipped = 1.
 * grad_scale folds the 1/world_size of the data-parallel mean into the same pass. */
int oron_adamw_clip(float* p, const float* g, float* m, float* v, void* p_bf16, int64_t n, const float* sumsq,
                    float grad_scale, float max_norm, float lr, float beta1, float beta2, float eps, float wd,
                    float bc1, float bc2, int32_t* skipped, oron_stream_t stream);

/* x[r, :] = 0 where row_valid[r] == 0: the masked_fill of TextEmbedding (encoder.py:86-87, 95) applied to a gradient. */
int oron_mask_rows_f32(float* x, int64_t ldx, int64_t rows, int32_t C, const uint8_t* row_valid, oron_stream_t stream);

/* IEEE f16 -> bf16 copy of [rows, C] (the V columns the QKV GEMM writes as f16 for the forward attention kernel). */
int oron_f16_to_bf16(const void* in, int64_t ld_in, int64_t rows, int32_t C, void* out, int64_t ld_out,
                     oron_stream_t stream);

/* oron_attention_bf16 for the training forward: also writes lse[(b*heads + h)*rows_per_batch + t] =
 * log2(sum_k exp2(s_tk * scale * log2 e)) of every query row, which oron_attention_bwd takes with have_lse = 1.
 * workspace (optional, as for oron_attention_bf16: oron_attention_workspace_bytes / oron_attention_plan): the planned
 * schedule (equal shares of the key-tile list per CTA); NULL: one CTA per (batch, head, query tile). */
int oron_attention_fwd_lse(const void* qkv, int64_t ld_qkv, void* out, int64_t ldo, int32_t nbatch,
                           int32_t rows_per_batch, int32_t heads, const int32_t* seq_lens, float scale, float* lse,
                           void* workspace, int64_t workspace_bytes, oron_stream_t stream);

/*
 * Backward of oron_attention_bf16 (F.scaled_dot_product_attention + key-padding mask, modules.py:271-278) with
 * the RoPE of q and k (modules.py:96-104) inverted on the way out. head_dim 64; tcgen05 MMAs, TMEM accumulators.
 *   qk:  bf16 [rows, ld_qk]: q | k (post-RoPE, as written by the QKV GEMM) at columns 0 | H*64
 *   v:   bf16 [rows, ld_v]; o: bf16 [rows, ld_o] (forward output); d_o: bf16 [rows, ld_do]
 *   dqkv: bf16 [rows, ld_dqkv]: dq | dk | dv w.r.t. the PRE-RoPE projections; rows t >= seq_lens[b] are zeroed
 *   lse, delta: f32 [nbatch*heads*rows_per_batch] workspaces (log2-domain log-sum-exp, rowsum(dO * O));
 *   have_lse = 1: lse holds the values written by oron_attention_fwd_lse (otherwise a pre-pass recomputes them)
 */
int oron_attention_bwd(const void* qk, int64_t ld_qk, const void* v, int64_t ld_v, const void* o, int64_t ld_o,
                       const void* d_o, int64_t ld_do, void* dqkv, int64_t ld_dqkv, int32_t nbatch,
                       int32_t rows_per_batch, int32_t heads, const int32_t* seq_lens, float scale,
                       const float* rope_cos, const float* rope_sin, float* lse, float* delta, int32_t have_lse,
                       oron_stream_t stream);

/* Profiling aid: int64 [2 * grid, 16] device buffer that receives clock64 stamps of the two attention-backward launches
 * (grid = tiles * heads * nbatch; dQ launch first); NULL disables. */
void oron_debug_set_attention_bwd_stamps(void* buf);

#ifdef __cplusplus
}
#endif
#endif /* ORON_B200_TRAIN_H_ */
