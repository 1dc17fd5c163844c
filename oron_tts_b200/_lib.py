"""ctypes binding of liboron_b200.so (the C ABI declared in include/oron_b200.h).

The library is built in-tree by ``__graft_entry__.build()`` (nvcc, sm_100a). There is no CPU
fallback: every wrapper raises if the shared object is missing or a tensor is not on a CUDA
device.
"""

from __future__ import annotations

import ctypes
import os
from ctypes import POINTER, c_char_p, c_float, c_int32, c_int64, c_uint8, c_uint64, c_void_p

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("ORON_LIB_PATH") or os.path.join(_HERE, "liboron_b200.so")

# ---- enums (include/oron_b200.h) -------------------------------------------------------------
EPI_BF16, EPI_F32, EPI_QKV_ROPE, EPI_GATE_RESID = 0, 1, 2, 3
EPI_EMBED_DUAL, EPI_MISH_MASK_BF16, EPI_MISH_MASK_RESID, EPI_SCALE_RESID = 4, 5, 6, 7
EPI_GELU_DROP_DUAL, EPI_GELU_DROP_BWD, EPI_GATE_RESID_DUAL = 8, 9, 10
ACT_NONE, ACT_GELU_TANH, ACT_GELU_ERF, ACT_SILU = 0, 1, 2, 3

EXPORTED_SYMBOLS = (
    "oron_gemm_bf16",
    "oron_ffn_bf16",
    "oron_ffn_workspace_bytes",
    "oron_gemm_ln_bf16",
    "oron_gemm_ln_counters",
    "oron_attention_bf16",
    "oron_attention_workspace_bytes",
    "oron_attention_plan",
    "oron_ln_modulate",
    "oron_cfg_euler_step",
    "oron_cast_rows_bf16",
    "oron_time_sinusoid",
    "oron_text_embed_front",
    "oron_dwconv7_ln",
    "oron_grn",
    "oron_logmel",
    "oron_logmel_bands",
    "oron_logmel_bands_bytes",
    "oron_istft_head",
    "oron_peak_normalize",
    "oron_debug_set_attention_stamps",
    "oron_debug_set_attention_schedule",
    "oron_debug_set_attention_version",
    "oron_abi_version",
    "oron_last_error",
    "oron_launch_count",
)


class GemmDesc(ctypes.Structure):
    """Mirror of ``struct oron_gemm_desc`` — field order and types must match the header."""

    _fields_ = [
        ("A", c_void_p),
        ("lda", c_int64),
        ("a_cols", c_int32),
        ("W", c_void_p),
        ("ldw", c_int64),
        ("w_cols", c_int32),
        ("rows_per_batch", c_int32),
        ("nbatch", c_int32),
        ("N", c_int32),
        ("taps", c_int32),
        ("cin_blocks", c_int32),
        ("pad", c_int32),
        ("grouped", c_int32),
        ("block_n", c_int32),
        ("epilogue", c_int32),
        ("act", c_int32),
        ("bias", c_void_p),
        ("out", c_void_p),
        ("ldo", c_int64),
        ("out2", c_void_p),
        ("ldo2", c_int64),
        ("addend", c_void_p),
        ("ld_add", c_int64),
        ("gate", c_void_p),
        ("gate_ld", c_int64),
        ("gate_nb", c_int32),
        ("gate_step_stride", c_int64),
        ("step_ptr", c_void_p),
        ("rope_cos", c_void_p),
        ("rope_sin", c_void_p),
        ("rope_cols", c_int32),
        ("seq_lens", c_void_p),
        ("row_valid", c_void_p),
        ("mask_rows", c_int32),
        ("max_ctas", c_int32),
        ("two_sm", c_int32),
        ("f16_from_col", c_int32),
        ("stream_k", c_int32),
        ("debug_stamps", c_void_p),
        ("a_mn_major", c_int32),
        ("b_mn_major", c_int32),
        ("dropout_p", c_float),
        ("dropout_seed", c_uint64),
    ]


class LnTail(ctypes.Structure):
    """Mirror of ``struct oron_ln_tail``."""

    _fields_ = [
        ("scale", c_void_p),
        ("shift", c_void_p),
        ("mod_ld", c_int64),
        ("mod_nb", c_int32),
        ("step_stride", c_int64),
        ("eps", c_float),
        ("add_one", c_int32),
        ("out_bf16", c_void_p),
        ("ldo", c_int64),
        ("counters", c_void_p),
        ("n_counters", c_int32),
    ]


_lib = None


def lib() -> ctypes.CDLL:
    """Load the shared object once; fail loudly when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(nvcc, sm_100a). oron_tts_b200 has no CPU / PyTorch fallback."
        )
    L = ctypes.CDLL(LIB_PATH)
    L.oron_last_error.restype = c_char_p
    L.oron_launch_count.restype = c_uint64
    L.oron_abi_version.restype = c_int32
    L.oron_debug_set_attention_stamps.argtypes = [c_void_p]
    L.oron_debug_set_attention_stamps.restype = None
    L.oron_debug_set_attention_schedule.argtypes = [c_int32]
    L.oron_debug_set_attention_schedule.restype = None
    L.oron_debug_set_attention_version.argtypes = [c_int32]
    L.oron_debug_set_attention_version.restype = None
    L.oron_gemm_bf16.argtypes = [POINTER(GemmDesc), c_void_p]
    L.oron_ffn_bf16.argtypes = [POINTER(GemmDesc), POINTER(GemmDesc), c_void_p, c_int64, c_void_p]
    L.oron_gemm_ln_bf16.argtypes = [POINTER(GemmDesc), POINTER(LnTail), c_void_p]
    L.oron_gemm_ln_counters.argtypes = [c_int32, c_int32]
    L.oron_gemm_ln_counters.restype = c_int32
    L.oron_ffn_workspace_bytes.argtypes = [c_int32, c_int32, c_int32]
    L.oron_ffn_workspace_bytes.restype = c_int64
    L.oron_attention_bf16.argtypes = [c_void_p, c_int64, c_void_p, c_int64, c_int32, c_int32, c_int32,
                                      c_void_p, c_float, c_void_p, c_int64, c_void_p]
    L.oron_attention_workspace_bytes.argtypes = [c_int32, c_int32, c_int32]
    L.oron_attention_workspace_bytes.restype = c_int64
    L.oron_attention_plan.argtypes = [c_void_p, c_int32, c_int32, c_int32, c_void_p, c_int64, c_void_p]
    L.oron_ln_modulate.argtypes = [c_void_p, c_int64, c_int32, c_int32, c_int32, c_float, c_void_p, c_void_p,
                                   c_int64, c_int32, c_int64, c_void_p, c_int32, c_void_p, c_void_p, c_int64,
                                   c_void_p]
    L.oron_cfg_euler_step.argtypes = [c_void_p, c_void_p, c_int64, c_int32, c_int32, c_int32, c_int32, c_float,
                                      c_void_p, c_void_p, c_void_p, c_int64, c_void_p, c_void_p, c_int32, c_void_p]
    L.oron_cast_rows_bf16.argtypes = [c_void_p, c_int64, c_int64, c_int32, c_void_p, c_int64, c_int32, c_void_p]
    L.oron_time_sinusoid.argtypes = [c_void_p, c_int32, c_void_p, c_int64, c_void_p]
    L.oron_text_embed_front.argtypes = [c_void_p, c_void_p, c_void_p, c_void_p, c_int32, c_int32, c_int32,
                                        c_void_p, c_int64, c_void_p, c_void_p]
    L.oron_dwconv7_ln.argtypes = [c_void_p, c_int64, c_int32, c_int32, c_int32, c_void_p, c_void_p, c_void_p,
                                  c_void_p, c_void_p, c_float, c_void_p, c_int64, c_void_p]
    L.oron_grn.argtypes = [c_void_p, c_int64, c_int32, c_int32, c_int32, c_void_p, c_void_p, c_void_p, c_void_p,
                           c_void_p]
    L.oron_logmel.argtypes = [c_void_p, c_int64, c_int32, c_int32, c_void_p, c_void_p, c_void_p, c_int32, c_float,
                              c_void_p, c_void_p]
    L.oron_logmel_bands.argtypes = [c_void_p, c_int32, c_void_p, c_void_p]
    L.oron_logmel_bands_bytes.argtypes = []
    L.oron_logmel_bands_bytes.restype = c_int64
    L.oron_istft_head.argtypes = [c_void_p, c_int64, c_int32, c_int32, c_int32, c_void_p, c_int32, c_void_p,
                                  c_int64, c_void_p]
    L.oron_peak_normalize.argtypes = [c_void_p, c_int64, c_int32, c_int32, c_void_p, c_int64, c_void_p, c_void_p]
    if L.oron_abi_version() != 1:
        raise RuntimeError("liboron_b200.so ABI version mismatch; rebuild")
    _lib = L
    return L


def launch_count() -> int:
    return int(lib().oron_launch_count())


def _check(rc: int, what: str) -> None:
    if rc != 0:
        msg = lib().oron_last_error().decode("utf-8", "replace")
        raise RuntimeError(f"{what} failed (code {rc}): {msg}")


def _ptr(t: torch.Tensor | None, dtype: torch.dtype | None = None, name: str = "tensor") -> int | None:
    if t is None:
        return None
    if not t.is_cuda:
        raise RuntimeError(f"{name} must live on a CUDA device (oron_tts_b200 has no CPU path)")
    if dtype is not None and t.dtype != dtype:
        raise TypeError(f"{name}: expected {dtype}, got {t.dtype}")
    return t.data_ptr()


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _ld(t: torch.Tensor) -> int:
    """Leading dimension (elements) of a 2-D row-major view."""
    assert t.dim() == 2 and t.stride(1) == 1, "expected a row-major 2-D tensor"
    return t.stride(0)


# ---- wrappers ---------------------------------------------------------------------------------
def gemm(
    A: torch.Tensor,
    W: torch.Tensor,
    out: torch.Tensor,
    *,
    epilogue: int,
    bias: torch.Tensor | None = None,
    act: int = ACT_NONE,
    rows_per_batch: int | None = None,
    nbatch: int = 1,
    n: int | None = None,
    a_cols: int | None = None,
    w_cols: int | None = None,
    taps: int = 1,
    cin_blocks: int = 0,
    pad: int = 0,
    grouped: int = 0,
    block_n: int = 128,
    out2: torch.Tensor | None = None,
    addend: torch.Tensor | None = None,
    gate: torch.Tensor | None = None,
    gate_ld: int = 0,
    gate_nb: int = 1,
    gate_step_stride: int = 0,
    step_ptr: torch.Tensor | None = None,
    rope_cos: torch.Tensor | None = None,
    rope_sin: torch.Tensor | None = None,
    rope_cols: int = 0,
    seq_lens: torch.Tensor | None = None,
    row_valid: torch.Tensor | None = None,
    mask_rows: bool = False,
    max_ctas: int = 0,
    two_sm: bool = False,
    debug_stamps: torch.Tensor | None = None,
    f16_from_col: int = 0,
    stream_k: bool = False,
    a_mn: bool = False,
    b_mn: bool = False,
    k: int | None = None,
    dropout_p: float = 0.0,
    dropout_seed: int = 0,
    desc_only: bool = False,
):
    """out = epilogue(A @ W.T) on the tcgen05 GEMM. A [rows, lda] bf16, W [N, ldw] bf16.
    ``desc_only``: build and return the descriptor without launching (operand of `ffn`).

    ``a_mn`` / ``b_mn`` (2-SM kernel): the operand is given as stored for the backward pass, with the contraction over
    its ROWS -- A as [K, M] (out = A.T @ ...), W as [K, N] (out = ... @ W); pass ``rows_per_batch`` = M and ``n`` = N."""
    if a_mn or b_mn:
        kk = int(k if k is not None else (A.shape[0] if a_mn else A.shape[1]))
        if rows_per_batch is None:
            rows_per_batch = A.shape[1] if a_mn else A.shape[0]
        if n is None:
            n = W.shape[1] if b_mn else W.shape[0]
        a_cols = kk if a_cols is None else a_cols
        w_cols = kk
    d = GemmDesc()
    d.A = _ptr(A, torch.bfloat16, "A")
    d.lda = _ld(A)
    d.a_cols = int(a_cols if a_cols is not None else A.shape[1])
    d.W = _ptr(W, torch.bfloat16, "W")
    d.ldw = _ld(W)
    d.w_cols = int(w_cols if w_cols is not None else W.shape[1])
    d.rows_per_batch = int(rows_per_batch if rows_per_batch is not None else A.shape[0])
    d.nbatch = int(nbatch)
    d.N = int(n if n is not None else W.shape[0])
    d.taps, d.cin_blocks, d.pad, d.grouped = int(taps), int(cin_blocks), int(pad), int(grouped)
    d.block_n, d.epilogue, d.act = int(block_n), int(epilogue), int(act)
    d.bias = _ptr(bias, torch.float32, "bias")
    d.out = _ptr(out, None, "out")
    d.ldo = _ld(out)
    d.out2 = _ptr(out2, torch.bfloat16, "out2")
    d.ldo2 = _ld(out2) if out2 is not None else 0
    d.addend = _ptr(addend, torch.float32, "addend")
    d.ld_add = _ld(addend) if addend is not None else 0
    d.gate = _ptr(gate, torch.float32, "gate")
    d.gate_ld, d.gate_nb, d.gate_step_stride = int(gate_ld), int(gate_nb), int(gate_step_stride)
    d.step_ptr = _ptr(step_ptr, torch.int32, "step_ptr")
    d.rope_cos = _ptr(rope_cos, torch.float32, "rope_cos")
    d.rope_sin = _ptr(rope_sin, torch.float32, "rope_sin")
    d.rope_cols = int(rope_cols)
    d.seq_lens = _ptr(seq_lens, torch.int32, "seq_lens")
    d.row_valid = _ptr(row_valid, torch.uint8, "row_valid")
    d.mask_rows = int(bool(mask_rows))
    d.max_ctas = int(max_ctas)
    d.two_sm = int(bool(two_sm))
    d.f16_from_col = int(f16_from_col)
    d.stream_k = int(bool(stream_k))
    d.debug_stamps = _ptr(debug_stamps, torch.int64, "debug_stamps")
    d.a_mn_major, d.b_mn_major = int(bool(a_mn)), int(bool(b_mn))
    d.dropout_p, d.dropout_seed = float(dropout_p), int(dropout_seed)
    want = torch.float32 if epilogue in (EPI_F32, EPI_GATE_RESID, EPI_EMBED_DUAL, EPI_MISH_MASK_RESID,
                                         EPI_SCALE_RESID, EPI_GATE_RESID_DUAL) else torch.bfloat16
    if out.dtype != want:
        raise TypeError(f"gemm epilogue {epilogue}: out must be {want}, got {out.dtype}")
    if desc_only:
        return d
    _check(lib().oron_gemm_bf16(ctypes.byref(d), _stream()), "oron_gemm_bf16")


def gemm_ln_counters(rows_per_batch: int, nbatch: int, device) -> torch.Tensor:
    """Zeroed arrival counters of `gemm_ln` (every launch leaves them zeroed; one buffer serves launches in stream order)."""
    return torch.zeros(int(lib().oron_gemm_ln_counters(rows_per_batch, nbatch)), dtype=torch.int32, device=device)


def gemm_ln(desc, counters: torch.Tensor, *, eps: float, scale: torch.Tensor, shift: torch.Tensor | None, out_bf16: torch.Tensor,
            mod_ld: int = 0, mod_nb: int = 1, step_stride: int = 0, add_one: bool = True) -> None:
    """`desc` (from ``gemm(..., epilogue=EPI_GATE_RESID, two_sm=True, desc_only=True)``) + LayerNorm / modulation of the updated
    residual rows into ``out_bf16`` inside the same launch (oron_gemm_ln_bf16); the step counter is the descriptor's."""
    t = LnTail()
    t.scale = _ptr(scale, torch.float32, "scale")
    t.shift = _ptr(shift, torch.float32, "shift")
    t.mod_ld, t.mod_nb, t.step_stride = int(mod_ld), int(mod_nb), int(step_stride)
    t.eps, t.add_one = float(eps), int(bool(add_one))
    t.out_bf16 = _ptr(out_bf16, torch.bfloat16, "out_bf16")
    t.ldo = _ld(out_bf16)
    t.counters = _ptr(counters, torch.int32, "counters")
    t.n_counters = counters.numel()
    _check(lib().oron_gemm_ln_bf16(ctypes.byref(desc), ctypes.byref(t), _stream()), "oron_gemm_ln_bf16")


def ffn_workspace(rows_per_batch: int, nbatch: int, ff_dim: int, device) -> torch.Tensor:
    """Zeroed flag buffer of `ffn` (the kernel leaves it zeroed; one buffer serves any number of launches in stream order)."""
    n = int(lib().oron_ffn_workspace_bytes(rows_per_batch, nbatch, ff_dim))
    return torch.zeros(n, dtype=torch.uint8, device=device)


def ffn(up, down, workspace: torch.Tensor) -> None:
    """FeedForward up-projection + activation + down-projection + gated residual in one launch (oron_ffn_bf16);
    `up` / `down` are descriptors from `gemm(..., desc_only=True)`."""
    _check(lib().oron_ffn_bf16(ctypes.byref(up), ctypes.byref(down), _ptr(workspace, torch.uint8, "workspace"),
                               workspace.numel(), _stream()), "oron_ffn_bf16")


def attention_workspace(nbatch: int, rows_per_batch: int, heads: int, device, seq_lens: torch.Tensor | None = None) -> torch.Tensor:
    """Workspace for the balanced schedule of `attention`, planned for `seq_lens` (call `attention_plan` again when the
    lengths change)."""
    n = int(lib().oron_attention_workspace_bytes(nbatch, rows_per_batch, heads))
    ws = torch.zeros(n, dtype=torch.uint8, device=device)
    attention_plan(ws, nbatch=nbatch, rows_per_batch=rows_per_batch, heads=heads, seq_lens=seq_lens)
    return ws


def attention_plan(workspace: torch.Tensor, *, nbatch: int, rows_per_batch: int, heads: int,
                   seq_lens: torch.Tensor | None) -> None:
    _check(
        lib().oron_attention_plan(_ptr(seq_lens, torch.int32, "seq_lens"), nbatch, rows_per_batch, heads,
                                  _ptr(workspace, torch.uint8, "workspace"), workspace.numel(), _stream()),
        "oron_attention_plan",
    )


def attention(qkv: torch.Tensor, out: torch.Tensor, *, nbatch: int, rows_per_batch: int, heads: int,
              seq_lens: torch.Tensor | None, scale: float, workspace: torch.Tensor | None = None) -> None:
    _check(
        lib().oron_attention_bf16(_ptr(qkv, torch.bfloat16, "qkv"), _ld(qkv), _ptr(out, torch.bfloat16, "out"),
                                  _ld(out), nbatch, rows_per_batch, heads, _ptr(seq_lens, torch.int32, "seq_lens"),
                                  float(scale), _ptr(workspace, torch.uint8, "workspace"),
                                  workspace.numel() if workspace is not None else 0, _stream()),
        "oron_attention_bf16",
    )


def ln_modulate(x: torch.Tensor, *, rows_per_batch: int, nbatch: int, eps: float, scale: torch.Tensor,
                shift: torch.Tensor | None, mod_ld: int = 0, mod_nb: int = 1, step_stride: int = 0,
                step_ptr: torch.Tensor | None = None, add_one: bool = True,
                out_bf16: torch.Tensor | None = None, out_f32: torch.Tensor | None = None) -> None:
    o = out_bf16 if out_bf16 is not None else out_f32
    _check(
        lib().oron_ln_modulate(_ptr(x, torch.float32, "x"), _ld(x), rows_per_batch, nbatch, x.shape[1], float(eps),
                               _ptr(scale, torch.float32, "scale"), _ptr(shift, torch.float32, "shift"), int(mod_ld),
                               int(mod_nb), int(step_stride), _ptr(step_ptr, torch.int32, "step_ptr"),
                               int(bool(add_one)), _ptr(out_bf16, torch.bfloat16, "out_bf16"),
                               _ptr(out_f32, torch.float32, "out_f32"), _ld(o), _stream()),
        "oron_ln_modulate",
    )


def cfg_euler_step(x: torch.Tensor, v: torch.Tensor, *, nb: int, rows_per_batch: int, n_mels: int,
                   has_uncond: bool, cfg: float, dt: torch.Tensor, step_ptr: torch.Tensor, xb: torch.Tensor,
                   traj: torch.Tensor | None = None, v_out: torch.Tensor | None = None, method: int = 0) -> None:
    _check(
        lib().oron_cfg_euler_step(_ptr(x, torch.float32, "x"), _ptr(v, torch.float32, "v"), _ld(v), nb,
                                  rows_per_batch, n_mels, int(bool(has_uncond)), float(cfg),
                                  _ptr(dt, torch.float32, "dt"), _ptr(step_ptr, torch.int32, "step_ptr"),
                                  _ptr(xb, torch.bfloat16, "xb"), _ld(xb), _ptr(traj, torch.float32, "traj"),
                                  _ptr(v_out, torch.float32, "v_out"), int(method), _stream()),
        "oron_cfg_euler_step",
    )


def cast_rows_bf16(x: torch.Tensor, out: torch.Tensor, *, reps: int = 1) -> None:
    _check(
        lib().oron_cast_rows_bf16(_ptr(x, torch.float32, "x"), _ld(x), x.shape[0], x.shape[1],
                                  _ptr(out, torch.bfloat16, "out"), _ld(out), reps, _stream()),
        "oron_cast_rows_bf16",
    )


def time_sinusoid(t: torch.Tensor, out: torch.Tensor) -> None:
    _check(lib().oron_time_sinusoid(_ptr(t, torch.float32, "t"), t.numel(), _ptr(out, torch.bfloat16, "out"),
                                    _ld(out), _stream()), "oron_time_sinusoid")


def text_embed_front(ids: torch.Tensor, drop: torch.Tensor, table: torch.Tensor, pos_table: torch.Tensor, *,
                     rows_per_batch: int, nb: int, x: torch.Tensor, row_valid: torch.Tensor) -> None:
    _check(
        lib().oron_text_embed_front(_ptr(ids, torch.int32, "ids"), _ptr(drop, torch.uint8, "drop"),
                                    _ptr(table, torch.float32, "table"), _ptr(pos_table, torch.float32, "pos"),
                                    rows_per_batch, nb, x.shape[1], _ptr(x, torch.float32, "x"), _ld(x),
                                    _ptr(row_valid, torch.uint8, "row_valid"), _stream()),
        "oron_text_embed_front",
    )


def dwconv7_ln(x: torch.Tensor, *, rows_per_batch: int, nbatch: int, seq_lens: torch.Tensor | None,
               w: torch.Tensor, wb: torch.Tensor, ln_w: torch.Tensor, ln_b: torch.Tensor, eps: float,
               out: torch.Tensor) -> None:
    _check(
        lib().oron_dwconv7_ln(_ptr(x, torch.float32, "x"), _ld(x), rows_per_batch, nbatch, x.shape[1],
                              _ptr(seq_lens, torch.int32, "seq_lens"), _ptr(w, torch.float32, "w"),
                              _ptr(wb, torch.float32, "wb"), _ptr(ln_w, torch.float32, "ln_w"),
                              _ptr(ln_b, torch.float32, "ln_b"), float(eps), _ptr(out, torch.bfloat16, "out"),
                              _ld(out), _stream()),
        "oron_dwconv7_ln",
    )


def grn(h: torch.Tensor, *, rows_per_batch: int, nb: int, seq_lens: torch.Tensor | None, gamma: torch.Tensor,
        beta: torch.Tensor, gx2: torch.Tensor) -> None:
    _check(
        lib().oron_grn(_ptr(h, torch.bfloat16, "h"), _ld(h), rows_per_batch, nb, h.shape[1],
                       _ptr(seq_lens, torch.int32, "seq_lens"), _ptr(gamma, torch.float32, "gamma"),
                       _ptr(beta, torch.float32, "beta"), _ptr(gx2, torch.float32, "gx2"), _stream()),
        "oron_grn",
    )


def logmel_bands(fb: torch.Tensor) -> torch.Tensor:
    """The non-zero band of every mel filter of `fb` [513, n_mels], laid out for `logmel` (once per filterbank)."""
    L = lib()
    bands = torch.empty(int(L.oron_logmel_bands_bytes()), dtype=torch.uint8, device=fb.device)
    _check(L.oron_logmel_bands(_ptr(fb, torch.float32, "fb"), fb.shape[1], _ptr(bands, torch.uint8, "bands"), _stream()),
           "oron_logmel_bands")
    return bands


def logmel(wav: torch.Tensor, window: torch.Tensor, fb: torch.Tensor, out: torch.Tensor, *, clip: float,
           bands: torch.Tensor | None = None) -> None:
    """wav f32 [nb, S] -> out f32 [nb, n_mels, 1 + S // 256]. `bands` = logmel_bands(fb) (computed here when omitted)."""
    if bands is None:
        bands = logmel_bands(fb)
    _check(
        lib().oron_logmel(_ptr(wav, torch.float32, "wav"), _ld(wav), wav.shape[0], wav.shape[1],
                          _ptr(window, torch.float32, "window"), _ptr(fb, torch.float32, "fb"),
                          _ptr(bands, torch.uint8, "bands"), fb.shape[1],
                          float(clip), _ptr(out, torch.float32, "out"), _stream()),
        "oron_logmel",
    )


def istft_head(h: torch.Tensor, window: torch.Tensor, out: torch.Tensor, *, rows_per_batch: int, nb: int,
               n_frames: int, mode: int = 0) -> None:
    _check(
        lib().oron_istft_head(_ptr(h, torch.float32, "h"), _ld(h), rows_per_batch, nb, n_frames,
                              _ptr(window, torch.float32, "window"), mode, _ptr(out, torch.float32, "out"),
                              _ld(out), _stream()),
        "oron_istft_head",
    )


def peak_normalize(x: torch.Tensor, out: torch.Tensor, scratch: torch.Tensor) -> None:
    _check(
        lib().oron_peak_normalize(_ptr(x, torch.float32, "x"), _ld(x), x.shape[0], x.shape[1],
                                  _ptr(out, torch.float32, "out"), _ld(out), _ptr(scratch, torch.float32, "scratch"),
                                  _stream()),
        "oron_peak_normalize",
    )
