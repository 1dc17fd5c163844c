/*
 * oron_b200.h — C ABI of the B200 (sm_100a) kernels behind the OronTTS inference hot path.
 *
 * The reference (btseee/oron-tts) is pure Python/PyTorch and has no FFI of its own; each entry
 * point below replaces the torch library calls of one reference call site (cited per function,
 * paths relative to the reference tree). The Python host code in oron_tts_b200/ binds this
 * header with ctypes.
 *
 * Conventions
 *   - Every pointer is a CUDA device pointer unless stated otherwise. The caller owns all
 *     buffers (outputs and workspaces included); the library never allocates device memory.
 *   - Activations are "frame-major": [nbatch * rows_per_batch, C] row-major, leading dimension
 *     given in elements. bf16 = __nv_bfloat16, f32 = float.
 *   - Every call only enqueues work on `stream` (a cudaStream_t); nothing synchronises, so all
 *     entry points are legal inside CUDA-graph capture.
 *   - Return value: 0 on success, a positive cudaError_t, or a negative argument-check code.
 *     oron_last_error() returns a thread-local, human-readable message for the last failure.
 */
#ifndef ORON_B200_H_
#define ORON_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef void* oron_stream_t; /* cudaStream_t */

#define ORON_ABI_VERSION 1

enum oron_status {
  ORON_OK = 0,
  ORON_ERR_BAD_ARG = -1,
  ORON_ERR_UNSUPPORTED = -2,
  ORON_ERR_NO_DRIVER = -3,
};

/* Epilogue fused into the tensor-core GEMM (what follows the Linear/Conv1d in the reference). */
enum oron_epilogue {
  ORON_EPI_BF16 = 0,            /* out_bf16 = act(acc + bias) */
  ORON_EPI_F32 = 1,             /* out_f32 = acc + bias (+ addend[row, col]) */
  ORON_EPI_QKV_ROPE = 2,        /* bias, RoPE rotate-half on cols < rope_cols (modules.py:96-104, 264-269) */
  ORON_EPI_GATE_RESID = 3,      /* out_f32[row] += gate[b] * (acc + bias)   (modules.py:338, :343, :281-282) */
  ORON_EPI_EMBED_DUAL = 4,      /* v = valid ? acc + addend : 0 -> out_f32 and out2_bf16 (dit.py:53, modules.py:136) */
  ORON_EPI_MISH_MASK_BF16 = 5,  /* out_bf16 = valid ? mish(acc + bias) : 0 (modules.py:137-140) */
  ORON_EPI_MISH_MASK_RESID = 6, /* out_f32 = (valid ? mish(acc + bias) : 0) + addend (dit.py:54) */
  ORON_EPI_SCALE_RESID = 7,     /* out_f32 = valid ? addend + colscale * (acc + bias) : 0 (modules.py:185, encoder.py:94) */
  /* training-step fusions of FeedForward (modules.py:294-297), two_sm only; the dropout mask is the stateless hash of
   * oron_b200_train.h over element index row * N + col (dropout_p = 0: none) */
  ORON_EPI_GELU_DROP_DUAL = 8,  /* out2_bf16 = pre = acc + bias ; out_bf16 = dropout(gelu_tanh(pre)) */
  ORON_EPI_GELU_DROP_BWD = 9,   /* out_bf16 = acc * gelu_tanh'(out2_bf16[row, col]) * mask  (out2 is read: the saved pre-activation) */
  ORON_EPI_GATE_RESID_DUAL = 10, /* y = acc + bias -> out2_bf16 ; out_f32 = addend + gate[b] * dropout(valid ? y : 0): the gated residuals of
                                  * DiTBlock (modules.py:338, 343) writing the next saved residual-stream buffer and keeping y for the backward */
};
enum oron_act { ORON_ACT_NONE = 0, ORON_ACT_GELU_TANH = 1, ORON_ACT_GELU_ERF = 2, ORON_ACT_SILU = 3 };

/*
 * D[rows, N] = A[rows, K] * W[N, K]^T with a fused epilogue; bf16 operands, fp32 accumulate in TMEM.
 * Replaces: nn.Linear calls of Attention/FeedForward/InputEmbedding/proj_out/AdaLN/TimestepEmbedding
 * (modules.py:60-62, 214, 232, 264-266, 279, 294-299; dit.py:53, 234), the pointwise Linear layers of
 * ConvNeXtV2Block (modules.py:181, 184) and of the Vocos backbone/head, and — with taps > 1 — the
 * Conv1d layers as implicit GEMMs: ConvPositionEmbedding's grouped k=31 convs (modules.py:120-124)
 * and the Vocos embed conv (k=7).
 *
 *   A: bf16 [nbatch*rows_per_batch, lda]; columns [0, a_cols) are read (zero beyond).
 *   W: bf16 [N, ldw]; K = w_cols must be a multiple of 64 for taps > 1.
 *   taps == 1: plain GEMM, K = w_cols.
 *   taps  > 1: W is laid out [N, taps * cin_blocks * 64] (tap-major); A rows are shifted by
 *              (tap - pad) with zero fill outside [0, rows_per_batch) of each batch element;
 *              grouped = G > 0 (a multiple of 64, == cin_blocks*64, block_n == 64): grouped conv whose
 *              groups are G channels wide; narrower groups are packed block-diagonally into 64-wide
 *              blocks by the caller (zero weights across groups).
 */
typedef struct oron_gemm_desc {
  const void* A;
  int64_t lda;
  int32_t a_cols;
  const void* W;
  int64_t ldw;
  int32_t w_cols;
  int32_t rows_per_batch;
  int32_t nbatch;
  int32_t N;
  int32_t taps;
  int32_t cin_blocks;
  int32_t pad;
  int32_t grouped;
  int32_t block_n; /* 64, 128 or 256 */
  int32_t epilogue;
  int32_t act;
  const float* bias; /* [N] or NULL */
  void* out;
  int64_t ldo;
  void* out2;
  int64_t ldo2;
  const float* addend; /* f32 [rows, ld_add] */
  int64_t ld_add;
  const float* gate; /* modulation table (GATE_RESID) or per-column scale (SCALE_RESID) */
  int64_t gate_ld;
  int32_t gate_nb;
  int64_t gate_step_stride;
  const int32_t* step_ptr;   /* device-side ODE step counter or NULL */
  const float* rope_cos;     /* f32 [rows_per_batch, 32] */
  const float* rope_sin;
  int32_t rope_cols;
  const int32_t* seq_lens;   /* [nbatch] or NULL */
  const uint8_t* row_valid;  /* [rows] or NULL */
  int32_t mask_rows;
  int32_t max_ctas;          /* 0 = one CTA per SM */
  int32_t two_sm;            /* 1: 2-SM (cta_group::2) kernel, 256 x block_n tile per SM pair */
  int32_t f16_from_col;      /* QKV_ROPE: columns >= this (> 0) are stored as IEEE f16, not bf16 (the V operand of attention) */
  int32_t stream_k;          /* 1 (two_sm + ORON_EPI_GATE_RESID or ORON_EPI_F32, taps == 1 only): cut the flat (tile, k-block) list into equal
                              * shares per SM pair; partial sums are added to `out` with f32 vector reductions (sum order, hence the last
                              * bit, may vary run to run). ORON_EPI_F32: out += A W^T (+ bias once), no addend -- the weight-gradient GEMMs,
                              * whose few output tiles (N_out x K_in / 256^2) would otherwise leave most SM pairs idle */
  void* debug_stamps;        /* NULL, or int64 [grid, 16] device buffer for per-CTA clock64 stamps (profiling aid) */
  /* Backward-pass operand layouts (two_sm = 1, taps = 1, K = w_cols a multiple of 64): the contraction runs over the
   * ROWS of the operand as stored, so no transposed copy is needed (TMA boxes of 64 x 64, MN-major UMMA descriptors):
   *   a_mn_major: A points to bf16 [K, lda], columns [0, rows_per_batch) -> D[m, n] = sum_k A[k, m] * ... (nbatch = 1)
   *   b_mn_major: W points to bf16 [K, ldw], columns [0, N)             -> D[m, n] = sum_k ... * W[k, n]
   * weight gradient dW = dY^T X: both; data gradient dX = dY W: b_mn_major only. */
  int32_t a_mn_major;
  int32_t b_mn_major;
  float dropout_p;       /* ORON_EPI_GELU_DROP_*: dropout probability (0: off) and seed */
  uint64_t dropout_seed;
} oron_gemm_desc;

int oron_gemm_bf16(const oron_gemm_desc* desc, oron_stream_t stream);

/*
 * oron_gemm_bf16 (two_sm = 1, ORON_EPI_GATE_RESID, block_n 256 or 192, K-major operands, taps == 1; stream_k allowed) followed,
 * INSIDE the launch, by the LayerNorm + AdaLN modulation of the updated residual rows: the attention out-projection + gated
 * residual + the block's second norm (modules.py:279, 338-341), and the FeedForward down-projection + gated residual + the next
 * block's first norm / AdaLayerNormFinal (modules.py:299, 343; 218, 234). Replaces the oron_ln_modulate launch that would follow:
 *   out_bf16[r, :] = LN(desc->out[r, :], eps) * (add_one + scale) + shift      for every row r, N = C in {128, 256, 512, 768, 1024},
 * scale / shift addressed like oron_ln_modulate with desc->step_ptr as the step counter. Rows are normalised as soon as the
 * tiles of their 256-row block have landed (per-block arrival counters), by all SMs of the launch.
 * counters: int32 [oron_gemm_ln_counters(rows_per_batch, nbatch)], zeroed ONCE by the caller; every launch leaves them zeroed.
 * One counter buffer must not be used by two launches that may run concurrently. Same arithmetic as oron_ln_modulate: with
 * stream_k = 0 the result is bit-identical to the two separate launches.
 */
typedef struct oron_ln_tail {
  const float* scale;
  const float* shift;  /* or NULL */
  int64_t mod_ld;
  int32_t mod_nb;
  int64_t step_stride;
  float eps;
  int32_t add_one;
  void* out_bf16;
  int64_t ldo;
  int32_t* counters;
  int32_t n_counters;
} oron_ln_tail;

int32_t oron_gemm_ln_counters(int32_t rows_per_batch, int32_t nbatch);
int oron_gemm_ln_bf16(const oron_gemm_desc* desc, const oron_ln_tail* ln, oron_stream_t stream);

/*
 * FeedForward of a DiTBlock in ONE launch (modules.py:294-299 and the gated residual modules.py:343; replaces the two
 * oron_gemm_bf16 calls `up` then `down`):   H = act(A W1^T + b1) ; resid += gate * (H W2^T + b2).
 *   up:   ORON_EPI_BF16 (act ORON_ACT_GELU_TANH or NONE), two_sm = 1, block_n = 256, N % 256 == 0; writes H (bf16) to up->out
 *   down: ORON_EPI_GATE_RESID, two_sm = 1, block_n = 256, A == up->out, w_cols == up->N
 * Every SM pair runs its share of up-projection tiles and then an equal share of the down-projection's (tile, k-block)
 * list; a down-projection k-block is fetched as soon as the H tile it reads is complete (global flags in `workspace`),
 * partial K sums are added to resid with f32 vector reductions (like stream_k: the last bit may vary run to run).
 * workspace: oron_ffn_workspace_bytes() bytes, 16-byte aligned, zeroed ONCE by the caller; the kernel leaves it zeroed.
 * One workspace must not be used by two launches that may run concurrently.
 */
int64_t oron_ffn_workspace_bytes(int32_t rows_per_batch, int32_t nbatch, int32_t ff_dim);
int oron_ffn_bf16(const oron_gemm_desc* up, const oron_gemm_desc* down, void* workspace, int64_t workspace_bytes,
                  oron_stream_t stream);

/*
 * softmax(Q K^T * scale + key_padding_mask) V over the fused QKV activation; head_dim 64.
 * Replaces F.scaled_dot_product_attention + mask (modules.py:271-278). RoPE is already applied
 * by the QKV GEMM epilogue.
 *   qkv: 16-bit [nbatch*rows_per_batch, ld_qkv], q | k | v at column offsets 0 | H*64 | 2*H*64;
 *        q and k are bf16, v is IEEE f16 (written so by the QKV GEMM with f16_from_col = 2*H*64): the
 *        probabilities come out of the packed f16x2 exp2 unit and P*V runs as an f16 x f16 MMA.
 *   out: bf16 [nbatch*rows_per_batch, ldo], head h at columns [64h, 64h+64). Query tiles that lie
 *        entirely beyond seq_lens[b] are not written.
 */
int oron_attention_bf16(const void* qkv, int64_t ld_qkv, void* out, int64_t ldo, int32_t nbatch,
                        int32_t rows_per_batch, int32_t heads, const int32_t* seq_lens, float scale,
                        void* workspace, int64_t workspace_bytes, oron_stream_t stream);
/*
 * Optional workspace of oron_attention_bf16 (16-byte aligned). With it -- and once oron_attention_plan has filled
 * it for the current sequence lengths -- calls with more (batch, head, query-tile) items than resident CTA slots
 * (2 per SM) run a balanced persistent schedule: every CTA owns an equal share of the flat (item, key tile) list,
 * items that straddle two shares are split along the keys, their normalised partial results (f16) are staged in the
 * workspace and combined by a small merge kernel launched right after. Passing NULL / 0 is always valid (one CTA per
 * item). The same workspace may be used by any number of calls with the planned shape and lengths.
 */
int64_t oron_attention_workspace_bytes(int32_t nbatch, int32_t rows_per_batch, int32_t heads);
/* Writes the schedule for these sequence lengths (device array, or NULL = all rows valid) into the workspace; enqueue
 * it on the same stream before the first oron_attention_bf16 that uses the workspace, and again whenever the lengths
 * change. An attention call whose workspace holds no plan for its shape traps with a message. */
int oron_attention_plan(const int32_t* seq_lens, int32_t nbatch, int32_t rows_per_batch, int32_t heads,
                        void* workspace, int64_t workspace_bytes, oron_stream_t stream);

/*
 * y = LayerNorm(x) * (add_one + scale) + shift, fp32 statistics, biased variance.
 * Replaces AdaLayerNorm / AdaLayerNormFinal / DiTBlock.ff_norm modulation (modules.py:218, 234, 341)
 * and affine nn.LayerNorm (modules.py:169; Vocos norms) with add_one = 0, mod_ld = 0.
 * scale/shift address: ptr + step*step_stride + (b % mod_nb)*mod_ld. C in {512, 1024}.
 */
int oron_ln_modulate(const float* x, int64_t ldx, int32_t rows_per_batch, int32_t nbatch, int32_t C,
                     float eps, const float* scale, const float* shift, int64_t mod_ld, int32_t mod_nb,
                     int64_t step_stride, const int32_t* step_ptr, int32_t add_one, void* out_bf16,
                     float* out_f32, int64_t ldo, oron_stream_t stream);

/*
 * One ODE update with classifier-free guidance (flow.py:266-267, 295-299):
 *   v = v_c + (v_c - v_u) * cfg ; x += v * dt[*step_ptr] ; traj[*step_ptr + 1] = x ; ++*step_ptr
 * and refresh of the bf16 operand of the next evaluation's input projection.
 * method 0: Euler (the reference). method 1: explicit midpoint rule behind the same call (the device counter then counts
 * velocity evaluations e, interval i = e / 2): even e writes only the operand x + v * dt[i] / 2, odd e does
 * x += v * dt[i] and fills trajectory slot i + 1. One launch per evaluation either way (CUDA-graph replayable).
 */
int oron_cfg_euler_step(float* x, const float* v, int64_t ldv, int32_t nb, int32_t rows_per_batch,
                        int32_t n_mels, int32_t has_uncond, float cfg, const float* dt, int32_t* step_ptr,
                        void* xb_bf16, int64_t ldxb, float* traj, float* v_out, int32_t method, oron_stream_t stream);

/* fp32 [rows, C] -> bf16 [reps*rows, ldo] (replicated `reps` times along rows). */
int oron_cast_rows_bf16(const float* x, int64_t ldx, int64_t rows, int32_t C, void* out_bf16, int64_t ldo,
                        int32_t reps, oron_stream_t stream);

/* SinusoidalEmbedding(256) of n timesteps (modules.py:39-45) -> bf16 [n, ldo]. */
int oron_time_sinusoid(const float* t, int32_t n, void* out_bf16, int64_t ldo, oron_stream_t stream);

/*
 * TextEmbedding front end (encoder.py:68-91): embedding gather + absolute sinusoid + filler zeroing.
 *   ids: int32 [nb*rows_per_batch], already +1 shifted (0 = filler / padding)
 *   drop: uint8 [nb]: text dropped for this batch element (CFG unconditional branch)
 *   table: f32 [vocab+1, C]; pos_table: f32 [>= rows_per_batch, C] (precompute_freqs_cis)
 */
int oron_text_embed_front(const int32_t* ids, const uint8_t* drop, const float* table, const float* pos_table,
                          int32_t rows_per_batch, int32_t nb, int32_t C, float* x, int64_t ldx,
                          uint8_t* row_valid, oron_stream_t stream);

/*
 * Depthwise Conv1d(k=7, pad=3, groups=C) over frames fused with the affine LayerNorm that follows
 * (modules.py:178-180; Vocos ConvNeXtBlock). x f32 -> out bf16. C == 512.
 */
int oron_dwconv7_ln(const float* x, int64_t ldx, int32_t rows_per_batch, int32_t nbatch, int32_t C,
                    const int32_t* seq_lens, const float* w, const float* wb, const float* ln_w,
                    const float* ln_b, float eps, void* out_bf16, int64_t ldo, oron_stream_t stream);

/*
 * GRN (modules.py:153-156) in place on bf16 h [rows, ldh]; gx2 is an f32 [nb, C] workspace.
 * The L2 norm runs over frames t < seq_lens[b] of each sequence.
 */
int oron_grn(void* h_bf16, int64_t ldh, int32_t rows_per_batch, int32_t nb, int32_t C,
             const int32_t* seq_lens, const float* gamma, const float* beta, float* gx2,
             oron_stream_t stream);

/*
 * log-mel front end (src/utils/audio.py:94-110 == torchaudio MelSpectrogram(center, reflect, hann,
 * power=1, HTK, norm=None) + log(clamp(., 1e-5))), fused: frames never leave the SM.
 *   wav: f32 [nb, n_samples] (ld_wav elements between clips); fb: f32 [n_fft/2+1, n_mels];
 *   window: f32 [n_fft]; out: f32 [nb, n_mels, n_frames], n_frames = 1 + n_samples / hop.
 * n_fft = 1024, hop = 256 only.
 *   bands: the non-zero band of every triangular filter, prepared ONCE per filterbank by oron_logmel_bands into a
 *   caller-owned, 16-byte aligned device buffer of oron_logmel_bands_bytes() bytes (the filterbank is banded: 1-31
 *   non-zero bins per filter, SURVEY appendix A.1 -- the projection is a short banded sum, not a 513 x 100 GEMM).
 */
int64_t oron_logmel_bands_bytes(void);
int oron_logmel_bands(const float* fb, int32_t n_mels, void* bands, oron_stream_t stream);
int oron_logmel(const float* wav, int64_t ld_wav, int32_t nb, int32_t n_samples, const float* window,
                const float* fb, const void* bands, int32_t n_mels, float clip, float* out, oron_stream_t stream);

/*
 * Vocos ISTFTHead tail (vocos heads.py ISTFTHead.forward + spectral_ops.py ISTFT "center"):
 *   mag = min(exp(h[:, :513]), 100) ; S = mag * (cos p + i sin p), p = h[:, 513:1026]
 *   frames = irfft(S, 1024) * window ; overlap-add at hop 256 ; / window-envelope ; trim 512 each side
 * (mode 1: h = interleaved-free real|imag halves, normalized=True, as src/models/decoder.py:86-102).
 *   h: f32 [nb*rows_per_batch, ldh]; n_frames valid frames per clip;
 *   out: f32 [nb, (n_frames-1)*256] with ld_out elements between clips.
 */
int oron_istft_head(const float* h, int64_t ldh, int32_t rows_per_batch, int32_t nb, int32_t n_frames,
                    const float* window, int32_t mode, float* out, int64_t ld_out, oron_stream_t stream);

/* Peak normalisation (audio.py:73-77): x / (max|x| + 1e-7) clamped to [-1, 1]; silent clips pass through.
 * scratch: f32 [nb]. */
int oron_peak_normalize(const float* x, int64_t ldx, int32_t nb, int32_t n, float* out, int64_t ldo,
                        float* scratch, oron_stream_t stream);

/* Profiling aid: int64 [n_ctas, 16] device buffer that receives per-CTA clock64 stamps of the attention kernel; NULL disables. */
void oron_debug_set_attention_stamps(void* buf);
/* Test aid: attention schedule override. -1 = automatic (default), 0 = one CTA per item, 1 = balanced whenever a
 * planned workspace is passed (exercises the split / merge path on small shapes). */
void oron_debug_set_attention_schedule(int32_t mode);
/* A/B aid: 4 = the current attention kernel (attn_fwd4.cuh, default), 3 = the round-1 kernel (attn_tcgen05.cuh). A
 * workspace must be re-planned (oron_attention_plan) after switching. Env ORON_ATT_VERSION sets the initial value. */
void oron_debug_set_attention_version(int32_t version);

int oron_abi_version(void);
const char* oron_last_error(void);
/* Number of kernels this library has launched in the calling process (bench.py "gpu_launches"). */
uint64_t oron_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif /* ORON_B200_H_ */
