"""ctypes binding of the training-step entry points of liboron_b200.so (include/oron_b200_train.h)."""

from __future__ import annotations

from ctypes import c_float, c_int32, c_int64, c_uint64, c_void_p

import torch

from ._lib import _check, _ld, _ptr, _stream, lib

ACT_MISH = 4
BF16, F32 = torch.bfloat16, torch.float32

TRAIN_SYMBOLS = (
    "oron_transpose_bf16", "oron_ln_bwd", "oron_act_fwd", "oron_act_bwd", "oron_gate_resid", "oron_gate_bwd",
    "oron_dwconv7", "oron_dwconv7_wgrad", "oron_grn_bwd_reduce", "oron_grn_bwd_coef", "oron_grn_bwd_apply",
    "oron_text_embed_bwd", "oron_skinny_dgrad", "oron_skinny_wgrad", "oron_gconv_wgrad", "oron_gconv_wgrad_tc", "oron_cfm_loss", "oron_sumsq",
    "oron_adamw_clip", "oron_f16_to_bf16", "oron_attention_bwd", "oron_mask_rows_f32", "oron_attention_fwd_lse", "oron_colsum_bf16",
)

_P, _I, _L, _F, _U = c_void_p, c_int32, c_int64, c_float, c_uint64
_ARGTYPES = {
    "oron_transpose_bf16": [_P, _L, _I, _I, _I, _P, _P, _L, _P, _P],
    "oron_ln_bwd": [_P, _L, _P, _L, _I, _I, _I, _F, _P, _L, _I, _P, _P, _L, _I, _P, _P, _L, _P],
    "oron_act_fwd": [_P, _I, _L, _L, _I, _I, _P, _I, _L, _I, _P, _F, _U, _P],
    "oron_act_bwd": [_P, _I, _L, _P, _I, _L, _L, _I, _I, _P, _I, _L, _I, _P, _F, _U, _P],
    "oron_gate_resid": [_P, _L, _P, _L, _I, _I, _I, _P, _L, _P, _I, _F, _U, _P, _L, _P],
    "oron_gate_bwd": [_P, _L, _P, _L, _I, _I, _I, _P, _L, _P, _P, _L, _P, _L, _P, _F, _U, _P],
    "oron_dwconv7": [_P, _L, _I, _I, _I, _P, _P, _P, _I, _P, _L, _I, _P],
    "oron_dwconv7_wgrad": [_P, _L, _P, _L, _I, _I, _I, _P, _P, _P, _P],
    "oron_grn_bwd_reduce": [_P, _L, _P, _L, _I, _I, _I, _P, _P, _P, _P],
    "oron_grn_bwd_coef": [_P, _P, _I, _I, _P, _P, _P, _P, _P],
    "oron_grn_bwd_apply": [_P, _L, _P, _L, _I, _I, _I, _P, _P, _P, _P, _P, _L, _P],
    "oron_text_embed_bwd": [_P, _P, _P, _L, _I, _I, _I, _P, _P],
    "oron_skinny_dgrad": [_P, _L, _I, _I, _P, _L, _I, _P, _L, _P],
    "oron_skinny_wgrad": [_P, _L, _P, _L, _I, _I, _I, _P, _L, _P, _I, _P],
    "oron_gconv_wgrad": [_P, _L, _P, _L, _I, _I, _I, _I, _I, _P, _P, _P, _P],
    "oron_gconv_wgrad_tc": [_P, _L, _P, _L, _I, _I, _I, _I, _I, _P, _P],
    "oron_cfm_loss": [_P, _L, _P, _P, _P, _L, _I, _P, _P, _L, _P],
    "oron_sumsq": [_P, _L, _P, _P],
    "oron_adamw_clip": [_P, _P, _P, _P, _P, _L, _P, _F, _F, _F, _F, _F, _F, _F, _F, _F, _P, _P],
    "oron_f16_to_bf16": [_P, _L, _L, _I, _P, _L, _P],
    "oron_mask_rows_f32": [_P, _L, _L, _I, _P, _P],
    "oron_colsum_bf16": [_P, _L, _L, _I, _P, _P],
    "oron_attention_bwd": [_P, _L, _P, _L, _P, _L, _P, _L, _P, _L, _I, _I, _I, _P, _F, _P, _P, _P, _P, _I, _P],
    "oron_attention_fwd_lse": [_P, _L, _P, _L, _I, _I, _I, _P, _F, _P, _P, _L, _P],
}
_bound = False


def tlib():
    global _bound
    L = lib()
    if not _bound:
        for name, at in _ARGTYPES.items():
            getattr(L, name).argtypes = at
        _bound = True
    return L


def _is32(t: torch.Tensor) -> int:
    if t.dtype == F32:
        return 1
    if t.dtype == BF16:
        return 0
    raise TypeError(f"expected float32 or bfloat16, got {t.dtype}")


def transpose(x: torch.Tensor, out: torch.Tensor, *, rows_per_batch: int, nbatch: int, seq_lens: torch.Tensor | None = None,
              colsum: torch.Tensor | None = None) -> None:
    """x bf16 [R, C] -> out bf16 [C, >= R]; optional colsum f32 [C] += column sums (bias gradient)."""
    _check(tlib().oron_transpose_bf16(_ptr(x, BF16, "x"), _ld(x), rows_per_batch, nbatch, x.shape[1],
                                      _ptr(seq_lens, torch.int32, "seq_lens"), _ptr(out, BF16, "out"), _ld(out),
                                      _ptr(colsum, F32, "colsum"), _stream()), "oron_transpose_bf16")


def ln_bwd(x: torch.Tensor, dy: torch.Tensor, *, rows_per_batch: int, nbatch: int, eps: float, scale: torch.Tensor,
           mod_ld: int, add_one: bool, seq_lens: torch.Tensor | None, dx: torch.Tensor, accumulate: bool,
           dscale: torch.Tensor | None, dshift: torch.Tensor | None, dmod_ld: int) -> None:
    _check(tlib().oron_ln_bwd(_ptr(x, F32, "x"), _ld(x), _ptr(dy, BF16, "dy"), _ld(dy), rows_per_batch, nbatch, x.shape[1],
                              float(eps), _ptr(scale, F32, "scale"), int(mod_ld), int(bool(add_one)),
                              _ptr(seq_lens, torch.int32, "seq_lens"), _ptr(dx, F32, "dx"), _ld(dx), int(bool(accumulate)),
                              _ptr(dscale, F32, "dscale"), _ptr(dshift, F32, "dshift"), int(dmod_ld), _stream()),
           "oron_ln_bwd")


def act_fwd(x: torch.Tensor, out: torch.Tensor, act: int, *, rows_per_batch: int = 0,
            seq_lens: torch.Tensor | None = None, dropout_p: float = 0.0, dropout_seed: int = 0) -> None:
    _check(tlib().oron_act_fwd(_ptr(x), _is32(x), _ld(x), x.shape[0], x.shape[1], act, _ptr(out), _is32(out), _ld(out),
                               int(rows_per_batch), _ptr(seq_lens, torch.int32, "seq_lens"), float(dropout_p),
                               int(dropout_seed), _stream()), "oron_act_fwd")


def act_bwd(dy: torch.Tensor, pre: torch.Tensor, out: torch.Tensor, act: int, *, rows_per_batch: int = 0,
            seq_lens: torch.Tensor | None = None, dropout_p: float = 0.0, dropout_seed: int = 0) -> None:
    _check(tlib().oron_act_bwd(_ptr(dy), _is32(dy), _ld(dy), _ptr(pre), _is32(pre), _ld(pre), dy.shape[0], dy.shape[1], act,
                               _ptr(out), _is32(out), _ld(out), int(rows_per_batch), _ptr(seq_lens, torch.int32, "seq_lens"),
                               float(dropout_p), int(dropout_seed), _stream()), "oron_act_bwd")


def gate_resid(x: torch.Tensor, y: torch.Tensor, *, rows_per_batch: int, nbatch: int, gate: torch.Tensor, gate_ld: int,
               seq_lens: torch.Tensor | None, mask_rows: bool, dropout_p: float = 0.0, dropout_seed: int = 0,
               out: torch.Tensor | None = None) -> None:
    """out (default: x, in place) = x + gate[b] * dropout(y)."""
    o = x if out is None else out
    _check(tlib().oron_gate_resid(_ptr(x, F32, "x"), _ld(x), _ptr(y, BF16, "y"), _ld(y), rows_per_batch, nbatch, x.shape[1],
                                  _ptr(gate, F32, "gate"), int(gate_ld), _ptr(seq_lens, torch.int32, "seq_lens"),
                                  int(bool(mask_rows)), float(dropout_p), int(dropout_seed), _ptr(o, F32, "out"), _ld(o),
                                  _stream()), "oron_gate_resid")


def gate_bwd(dx: torch.Tensor, y: torch.Tensor, *, rows_per_batch: int, nbatch: int, gate: torch.Tensor, gate_ld: int,
             seq_lens: torch.Tensor | None, dy: torch.Tensor, dgate: torch.Tensor | None, dgate_ld: int,
             dropout_p: float = 0.0, dropout_seed: int = 0, dbias: torch.Tensor | None = None) -> None:
    _check(tlib().oron_gate_bwd(_ptr(dx, F32, "dx"), _ld(dx), _ptr(y, BF16, "y"), _ld(y), rows_per_batch, nbatch,
                                dx.shape[1], _ptr(gate, F32, "gate"), int(gate_ld), _ptr(seq_lens, torch.int32, "seq_lens"),
                                _ptr(dy, BF16, "dy"), _ld(dy), _ptr(dgate, F32, "dgate"), int(dgate_ld),
                                _ptr(dbias, F32, "dbias"), float(dropout_p), int(dropout_seed), _stream()), "oron_gate_bwd")


def dwconv7(x: torch.Tensor, out: torch.Tensor, *, rows_per_batch: int, nbatch: int, seq_lens: torch.Tensor | None,
            w: torch.Tensor, bias: torch.Tensor | None, flip: bool = False, accumulate: bool = False) -> None:
    _check(tlib().oron_dwconv7(_ptr(x, F32, "x"), _ld(x), rows_per_batch, nbatch, x.shape[1],
                               _ptr(seq_lens, torch.int32, "seq_lens"), _ptr(w, F32, "w"), _ptr(bias, F32, "bias"),
                               int(bool(flip)), _ptr(out, F32, "out"), _ld(out), int(bool(accumulate)), _stream()),
           "oron_dwconv7")


def dwconv7_wgrad(x: torch.Tensor, dy: torch.Tensor, *, rows_per_batch: int, nbatch: int, seq_lens: torch.Tensor | None,
                  dw: torch.Tensor, db: torch.Tensor | None) -> None:
    _check(tlib().oron_dwconv7_wgrad(_ptr(x, F32, "x"), _ld(x), _ptr(dy, F32, "dy"), _ld(dy), rows_per_batch, nbatch,
                                     x.shape[1], _ptr(seq_lens, torch.int32, "seq_lens"), _ptr(dw, F32, "dw"),
                                     _ptr(db, F32, "db"), _stream()), "oron_dwconv7_wgrad")


def grn_bwd(dy: torch.Tensor, pre: torch.Tensor, dpre: torch.Tensor, *, rows_per_batch: int, nb: int,
            seq_lens: torch.Tensor | None, gamma: torch.Tensor, gx2: torch.Tensor, A: torch.Tensor, nx: torch.Tensor,
            coef: torch.Tensor, dgamma: torch.Tensor, dbeta: torch.Tensor) -> None:
    """Backward of GELU(erf) -> GRN from the saved pre-activation (three launches: reduce, coefficients, apply)."""
    L = tlib()
    C = dy.shape[1]
    sl = _ptr(seq_lens, torch.int32, "seq_lens")
    _check(L.oron_grn_bwd_reduce(_ptr(dy, BF16, "dy"), _ld(dy), _ptr(pre, BF16, "pre"), _ld(pre), rows_per_batch, nb, C, sl,
                                 _ptr(A, F32, "A"), _ptr(dbeta, F32, "dbeta"), _stream()), "oron_grn_bwd_reduce")
    _check(L.oron_grn_bwd_coef(_ptr(A, F32), _ptr(gx2, F32, "gx2"), nb, C, _ptr(gamma, F32, "gamma"), _ptr(coef, F32, "coef"),
                               _ptr(nx, F32, "nx"), _ptr(dgamma, F32, "dgamma"), _stream()), "oron_grn_bwd_coef")
    _check(L.oron_grn_bwd_apply(_ptr(dy, BF16), _ld(dy), _ptr(pre, BF16), _ld(pre), rows_per_batch, nb, C, sl,
                                _ptr(gamma, F32), _ptr(nx, F32), _ptr(coef, F32), _ptr(dpre, BF16, "dpre"), _ld(dpre),
                                _stream()), "oron_grn_bwd_apply")


def text_embed_bwd(ids: torch.Tensor, drop: torch.Tensor, dx: torch.Tensor, dtable: torch.Tensor, *, rows_per_batch: int,
                   nb: int) -> None:
    _check(tlib().oron_text_embed_bwd(_ptr(ids, torch.int32, "ids"), _ptr(drop, torch.uint8, "drop"), _ptr(dx, F32, "dx"),
                                      _ld(dx), rows_per_batch, nb, dx.shape[1], _ptr(dtable, F32, "dtable"), _stream()),
           "oron_text_embed_bwd")


def skinny_dgrad(dY: torch.Tensor, W: torch.Tensor, dX: torch.Tensor) -> None:
    """dX[b, k] += sum_n dY[b, n] W[n, k]; dY f32 [nb, N], W bf16 [N, K], dX f32 [nb, K] (accumulates)."""
    _check(tlib().oron_skinny_dgrad(_ptr(dY, F32, "dY"), _ld(dY), dY.shape[0], W.shape[0], _ptr(W, BF16, "W"), _ld(W),
                                    W.shape[1], _ptr(dX, F32, "dX"), _ld(dX), _stream()), "oron_skinny_dgrad")


def skinny_wgrad(dY: torch.Tensor, X: torch.Tensor, dW: torch.Tensor, db: torch.Tensor | None, *, accumulate: bool) -> None:
    """dW[n, k] (+)= sum_b dY[b, n] X[b, k]; db[n] (+)= sum_b dY[b, n]."""
    _check(tlib().oron_skinny_wgrad(_ptr(dY, F32, "dY"), _ld(dY), _ptr(X, F32, "X"), _ld(X), dY.shape[0], dY.shape[1],
                                    X.shape[1], _ptr(dW, F32, "dW"), _ld(dW), _ptr(db, F32, "db"), int(bool(accumulate)),
                                    _stream()), "oron_skinny_wgrad")


def gconv_wgrad(x: torch.Tensor, dy: torch.Tensor, *, rows_per_batch: int, nbatch: int, cg: int, taps: int,
                seq_lens: torch.Tensor | None, dw: torch.Tensor, db: torch.Tensor | None) -> None:
    _check(tlib().oron_gconv_wgrad(_ptr(x, BF16, "x"), _ld(x), _ptr(dy, BF16, "dy"), _ld(dy), rows_per_batch, nbatch,
                                   x.shape[1], cg, taps, _ptr(seq_lens, torch.int32, "seq_lens"), _ptr(dw, F32, "dw"),
                                   _ptr(db, F32, "db"), _stream()), "oron_gconv_wgrad")


def gconv_wgrad_tc(x: torch.Tensor, dy: torch.Tensor, *, rows_per_batch: int, nbatch: int, cg: int, taps: int, dw: torch.Tensor) -> None:
    """Tensor-core variant: dw += ...; rows beyond each sequence's length must already be zero in x and dy; no bias gradient."""
    _check(tlib().oron_gconv_wgrad_tc(_ptr(x, BF16, "x"), _ld(x), _ptr(dy, BF16, "dy"), _ld(dy), rows_per_batch, nbatch,
                                      x.shape[1], cg, taps, _ptr(dw, F32, "dw"), _stream()), "oron_gconv_wgrad_tc")


def cfm_loss(pred: torch.Tensor, flow: torch.Tensor, span: torch.Tensor, count: torch.Tensor, loss_sum: torch.Tensor,
             dpred: torch.Tensor, *, n_mels: int) -> None:
    _check(tlib().oron_cfm_loss(_ptr(pred, F32, "pred"), _ld(pred), _ptr(flow, F32, "flow"), _ptr(span, torch.uint8, "span"),
                                _ptr(count, torch.int32, "count"), pred.shape[0], n_mels, _ptr(loss_sum, F32, "loss_sum"),
                                _ptr(dpred, BF16, "dpred"), _ld(dpred), _stream()), "oron_cfm_loss")


def sumsq(g: torch.Tensor, out: torch.Tensor) -> None:
    _check(tlib().oron_sumsq(_ptr(g, F32, "g"), g.numel(), _ptr(out, F32, "out"), _stream()), "oron_sumsq")


def adamw_clip(p: torch.Tensor, g: torch.Tensor, m: torch.Tensor, v: torch.Tensor, pb: torch.Tensor | None,
               sumsq_t: torch.Tensor, *, grad_scale: float, max_norm: float, lr: float, beta1: float, beta2: float,
               eps: float, wd: float, step: int, skipped: torch.Tensor | None = None) -> None:
    bc1, bc2 = 1.0 - beta1 ** step, 1.0 - beta2 ** step
    _check(tlib().oron_adamw_clip(_ptr(p, F32, "p"), _ptr(g, F32, "g"), _ptr(m, F32, "m"), _ptr(v, F32, "v"),
                                  _ptr(pb, BF16, "pb"), p.numel(), _ptr(sumsq_t, F32, "sumsq"), grad_scale, max_norm, lr,
                                  beta1, beta2, eps, wd, bc1, bc2, _ptr(skipped, torch.int32, "skipped"), _stream()),
           "oron_adamw_clip")


def colsum(x: torch.Tensor, out: torch.Tensor) -> None:
    """out f32 [C] += column sums of x bf16 [rows, C]."""
    _check(tlib().oron_colsum_bf16(_ptr(x, BF16, "x"), _ld(x), x.shape[0], x.shape[1], _ptr(out, F32, "out"), _stream()),
           "oron_colsum_bf16")


def mask_rows(x: torch.Tensor, row_valid: torch.Tensor) -> None:
    _check(tlib().oron_mask_rows_f32(_ptr(x, F32, "x"), _ld(x), x.shape[0], x.shape[1], _ptr(row_valid, torch.uint8, "row_valid"),
                                     _stream()), "oron_mask_rows_f32")


def f16_to_bf16(x: torch.Tensor, out: torch.Tensor) -> None:
    """x: 16-bit [rows, C] holding IEEE f16 bit patterns (any 16-bit torch dtype) -> out bf16."""
    _check(tlib().oron_f16_to_bf16(_ptr(x), _ld(x), x.shape[0], x.shape[1], _ptr(out, BF16, "out"), _ld(out), _stream()),
           "oron_f16_to_bf16")


def attention_bwd(qk: torch.Tensor, v: torch.Tensor, o: torch.Tensor, d_o: torch.Tensor, dqkv: torch.Tensor, *, nbatch: int,
                  rows_per_batch: int, heads: int, seq_lens: torch.Tensor | None, scale: float, rope_cos: torch.Tensor,
                  rope_sin: torch.Tensor, lse: torch.Tensor, delta: torch.Tensor, have_lse: bool = False) -> None:
    _check(tlib().oron_attention_bwd(_ptr(qk, BF16, "qk"), _ld(qk), _ptr(v, BF16, "v"), _ld(v), _ptr(o, BF16, "o"), _ld(o),
                                     _ptr(d_o, BF16, "d_o"), _ld(d_o), _ptr(dqkv, BF16, "dqkv"), _ld(dqkv), nbatch,
                                     rows_per_batch, heads, _ptr(seq_lens, torch.int32, "seq_lens"), float(scale),
                                     _ptr(rope_cos, F32, "rope_cos"), _ptr(rope_sin, F32, "rope_sin"), _ptr(lse, F32, "lse"),
                                     _ptr(delta, F32, "delta"), int(bool(have_lse)), _stream()), "oron_attention_bwd")


def attention_fwd_lse(qkv: torch.Tensor, out: torch.Tensor, lse: torch.Tensor, *, nbatch: int, rows_per_batch: int, heads: int,
                      seq_lens: torch.Tensor | None, scale: float, workspace: torch.Tensor | None = None) -> None:
    """`workspace`: a planned attention workspace (_lib.attention_workspace / attention_plan for these seq_lens), or None."""
    _check(tlib().oron_attention_fwd_lse(_ptr(qkv, BF16, "qkv"), _ld(qkv), _ptr(out, BF16, "out"), _ld(out), nbatch,
                                         rows_per_batch, heads, _ptr(seq_lens, torch.int32, "seq_lens"), float(scale),
                                         _ptr(lse, F32, "lse"), _ptr(workspace, torch.uint8, "workspace"),
                                         0 if workspace is None else workspace.numel(), _stream()), "oron_attention_fwd_lse")
