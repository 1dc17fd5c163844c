"""DiT backbone with the reference's module tree (=> identical ``state_dict()`` keys and shapes,
SURVEY.md §8b) whose forward runs on the sm_100a engine instead of torch ops.

The nn.Module classes below are parameter holders: they reproduce the attribute names of
src/models/modules.py, src/models/encoder.py and src/models/dit.py so a reference checkpoint loads
with strict=True, but none of them implements a torch forward — DiT.forward drives
engine.DiTEngine. Packed bf16 weight copies live outside the state dict and are rebuilt whenever
parameters change.
"""

from __future__ import annotations

import torch
import torch.nn as nn

from .engine import Branch, DiTEngine, DiTWeights, TILE, _rup


def _no_forward(self, *a, **k):  # pragma: no cover
    raise RuntimeError(f"{type(self).__name__} is a parameter holder; call DiT.forward / CFM.sample")


class _Holder(nn.Module):
    forward = _no_forward


class TimestepEmbedding(_Holder):  # modules.py:48-58
    def __init__(self, dim: int, freq_embed_dim: int = 256):
        super().__init__()
        self.time_mlp = nn.Sequential(nn.Linear(freq_embed_dim, dim), nn.SiLU(), nn.Linear(dim, dim))


class RotaryEmbedding(_Holder):  # modules.py:68-74 — inv_freq is a persistent buffer
    def __init__(self, dim: int):
        super().__init__()
        self.register_buffer("inv_freq", 1.0 / (10000 ** (torch.arange(0, dim, 2).float() / dim)))


class GRN(_Holder):  # modules.py:147-151
    def __init__(self, dim: int):
        super().__init__()
        self.gamma = nn.Parameter(torch.zeros(1, 1, dim))
        self.beta = nn.Parameter(torch.zeros(1, 1, dim))


class ConvNeXtV2Block(_Holder):  # modules.py:162-173
    def __init__(self, dim: int, intermediate_dim: int):
        super().__init__()
        self.dwconv = nn.Conv1d(dim, dim, kernel_size=7, padding=3, groups=dim)
        self.norm = nn.LayerNorm(dim, eps=1e-6)
        self.pwconv1 = nn.Linear(dim, intermediate_dim)
        self.grn = GRN(intermediate_dim)
        self.pwconv2 = nn.Linear(intermediate_dim, dim)


class TextEmbedding(_Holder):  # encoder.py:27-50
    def __init__(self, vocab_size: int, text_dim: int, conv_layers: int = 0, conv_mult: int = 2):
        super().__init__()
        self.text_embed = nn.Embedding(vocab_size + 1, text_dim)
        self.extra_modeling = conv_layers > 0
        if conv_layers > 0:
            self.text_blocks = nn.Sequential(*[ConvNeXtV2Block(text_dim, text_dim * conv_mult) for _ in range(conv_layers)])


class ConvPositionEmbedding(_Holder):  # modules.py:117-125
    def __init__(self, dim: int, kernel_size: int = 31, groups: int = 16):
        super().__init__()
        self.conv1d = nn.Sequential(
            nn.Conv1d(dim, dim, kernel_size, groups=groups, padding=kernel_size // 2), nn.Mish(),
            nn.Conv1d(dim, dim, kernel_size, groups=groups, padding=kernel_size // 2), nn.Mish(),
        )


class InputEmbedding(_Holder):  # dit.py:30-33
    def __init__(self, mel_dim: int, text_dim: int, out_dim: int):
        super().__init__()
        self.proj = nn.Linear(mel_dim * 2 + text_dim, out_dim)
        self.conv_pos_embed = ConvPositionEmbedding(dim=out_dim)


class AdaLayerNorm(_Holder):  # modules.py:205-209
    def __init__(self, dim: int, n_chunks: int = 6):
        super().__init__()
        self.linear = nn.Linear(dim, dim * n_chunks)


class Attention(_Holder):  # modules.py:241-253
    def __init__(self, dim: int, heads: int, dim_head: int = 64, dropout: float = 0.0):
        super().__init__()
        self.heads = heads
        self.inner_dim = dim_head * heads
        self.to_q = nn.Linear(dim, self.inner_dim)
        self.to_k = nn.Linear(dim, self.inner_dim)
        self.to_v = nn.Linear(dim, self.inner_dim)
        self.to_out = nn.Sequential(nn.Linear(self.inner_dim, dim), nn.Dropout(dropout))


class FeedForward(_Holder):  # modules.py:291-299
    def __init__(self, dim: int, mult: int = 4, dropout: float = 0.0):
        super().__init__()
        inner = int(dim * mult)
        self.ff = nn.Sequential(nn.Linear(dim, inner), nn.GELU(approximate="tanh"), nn.Dropout(dropout), nn.Linear(inner, dim))


class DiTBlock(_Holder):  # modules.py:311-324
    def __init__(self, dim: int, heads: int, dim_head: int = 64, ff_mult: int = 4, dropout: float = 0.1):
        super().__init__()
        self.attn_norm = AdaLayerNorm(dim, 6)
        self.attn = Attention(dim=dim, heads=heads, dim_head=dim_head, dropout=dropout)
        self.ff = FeedForward(dim=dim, mult=ff_mult, dropout=dropout)


class DiT(nn.Module):
    """Drop-in for src/models/dit.py:58-234 (same ctor kwargs, attributes, forward signature)."""

    def __init__(self, *, dim: int = 1024, depth: int = 22, heads: int = 16, dim_head: int = 64, ff_mult: int = 4,
                 dropout: float = 0.1, mel_dim: int = 100, vocab_size: int = 65, text_dim: int = 512,
                 conv_layers: int = 4, gradient_checkpointing: bool = False) -> None:
        super().__init__()
        self.dim, self.depth = dim, depth
        self.mel_dim = mel_dim
        self.gradient_checkpointing = gradient_checkpointing
        self.time_embed = TimestepEmbedding(dim)
        self.text_embed = TextEmbedding(vocab_size=vocab_size, text_dim=text_dim, conv_layers=conv_layers)
        self.text_cond: torch.Tensor | None = None
        self.text_uncond: torch.Tensor | None = None
        self.input_embed = InputEmbedding(mel_dim, text_dim, dim)
        self.rotary_embed = RotaryEmbedding(dim_head)
        self.transformer_blocks = nn.ModuleList(
            [DiTBlock(dim=dim, heads=heads, dim_head=dim_head, ff_mult=ff_mult, dropout=dropout) for _ in range(depth)])
        self.norm_out = AdaLayerNorm(dim, 2)
        self.proj_out = nn.Linear(dim, mel_dim)
        self._initialize_weights()
        self.__dict__["_engine"] = None
        self.__dict__["_engine_sig"] = None

    def _initialize_weights(self) -> None:
        # dit.py:119-129: AdaLN projections and the output head start at zero
        for blk in self.transformer_blocks:
            nn.init.zeros_(blk.attn_norm.linear.weight)
            nn.init.zeros_(blk.attn_norm.linear.bias)
        for m in (self.norm_out.linear, self.proj_out):
            nn.init.zeros_(m.weight)
            nn.init.zeros_(m.bias)

    # ---- engine / packed weights ---------------------------------------------------------------
    def _signature(self):
        def ver(p):  # tensors created under torch.inference_mode() have no version counter
            try:
                return p._version
            except RuntimeError:
                return -1

        # nn.Module.parameters() walks the module tree (2 ms per call for the Base model, with the GPU idle at the start of
        # every synthesize): keep the (owner module, name) slots and re-read the Parameter objects from them, so replaced
        # parameters (load_state_dict(assign=True), .to()) are still seen; the slot list is rebuilt after _apply /
        # load_state_dict / a sub-module assignment on this module
        slots = self.__dict__.get("_param_slots")
        if slots is None:
            slots = [(m, n) for m in self.modules() for n in m._parameters if m._parameters[n] is not None]
            self.__dict__["_param_slots"] = slots
        return tuple((p.data_ptr(), ver(p)) for p in (m._parameters[n] for m, n in slots))

    def _apply(self, fn, *args, **kwargs):
        self.__dict__["_param_slots"] = None
        return super()._apply(fn, *args, **kwargs)

    def load_state_dict(self, *args, **kwargs):
        self.__dict__["_param_slots"] = None
        return super().load_state_dict(*args, **kwargs)

    def __setattr__(self, name, value):
        if isinstance(value, nn.Module):
            self.__dict__["_param_slots"] = None
        super().__setattr__(name, value)

    def engine(self) -> DiTEngine:
        """Packed-weight engine for the current parameters (rebuilt after load_state_dict / .to())."""
        p0 = next(self.parameters())
        if not p0.is_cuda:
            raise RuntimeError("oron_tts_b200.DiT runs only on a CUDA device (no CPU fallback); call .to('cuda')")
        sig = self._signature()
        if self.__dict__["_engine"] is None or self.__dict__["_engine_sig"] != sig:
            sd = {k: v for k, v in self.state_dict().items()}
            self.__dict__["_engine"] = DiTEngine(DiTWeights(sd, p0.device))
            self.__dict__["_engine_sig"] = sig
        return self.__dict__["_engine"]

    def precise(self):
        """fp32-mode engine for the current parameters (rebuilt when they change)."""
        from .precise import PreciseDiT

        sig = self._signature()
        if self.__dict__.get("_precise") is None or self.__dict__.get("_precise_sig") != sig:
            self.__dict__["_precise"] = PreciseDiT(self)
            self.__dict__["_precise_sig"] = sig
        return self.__dict__["_precise"]

    def clear_cache(self) -> None:
        self.text_cond = None
        self.text_uncond = None

    # ---- forward -------------------------------------------------------------------------------
    @torch.no_grad()
    def forward(self, x: torch.Tensor, cond: torch.Tensor, text: torch.Tensor, time: torch.Tensor,
                mask: torch.Tensor | None = None, drop_audio_cond: bool = False, drop_text: bool = False,
                cfg_infer: bool = False, cache: bool = False, precision: str = "bf16") -> torch.Tensor:
        """Velocity field [B, N, mel] (or [2B, N, mel] with cfg_infer) — dit.py:165-234.

        ``mask`` must be a prefix mask (frames [0, len_b) valid), which is what CFM builds from lengths.
        ``cache`` is accepted for signature parity; text embeddings are recomputed per call here (the
        CFM.sample fast path hoists them out of the ODE loop instead).
        """
        if precision == "fp32":
            # parity / debugging mode (extra kwarg): fp32 activations, 3-way bf16-split GEMMs, fp32 attention: within
            # 1e-6 of the fp32 reference, an order of magnitude slower (oron_tts_b200/precise.py)
            return self.precise().forward(x, cond, text, time, mask=mask, drop_audio_cond=drop_audio_cond, drop_text=drop_text,
                                          cfg_infer=cfg_infer)
        if precision != "bf16":
            raise ValueError(f"precision must be 'bf16' or 'fp32', got {precision!r}")
        if torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters()) and self.training:
            raise NotImplementedError("DiT.forward alone builds no autograd graph: train through CFM.forward / F5TTS.forward "
                                      "(an autograd node over the sm_100a training engine) or oron_tts_b200.train.TrainEngine")
        eng = self.engine()
        w = eng.w
        B, T, _ = x.shape
        if time.ndim == 0:
            time = time.repeat(B)
        branches = [Branch(False, False), Branch(True, True)] if cfg_infer else [Branch(drop_audio_cond, drop_text)]
        if mask is None:
            durations = [T] * B
        else:
            durations = [int(v) for v in mask.sum(dim=-1).tolist()]
        tpad = _rup(T, TILE)
        ws = eng.workspace(B, B * len(branches), tpad, max(B, 1), False)
        eng.load_sequences(ws, text=text, durations=durations, seq_len=T, branches=branches, text_len=T)
        eng.text_embed(ws)
        cond_in = cond.to(torch.float32)
        eng.static_embed(ws, cond_in, branches)
        eng.modulation_table(ws, time.to(torch.float32))
        xv = ws.x.view(B, tpad, w.n_mels)
        xv.zero_()
        xv[:, :T].copy_(x)
        from . import _lib as L
        L.cast_rows_bf16(ws.x, ws.xb[:, : w.n_mels], reps=len(branches))
        eng.velocity(ws, mod_nb=B, use_step=False)
        return ws.v.view(B * len(branches), tpad, w.n_mels)[:, :T].clone()
