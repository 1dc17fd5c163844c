"""fp32 ("1e-4") mode of DiT.forward (north_star: per-NFE-step velocity within 1e-4 relative L2 of the fp32 reference).

A parity / debugging mode, slow by design: every activation stays fp32, every non-GEMM op runs in fp32 kernels with
libm-accurate transcendentals (include/oron_b200_precise.h), and every Linear / Conv1d still runs on the tcgen05 GEMM
with its fp32 operands split into three bf16 terms (24 mantissa bits): six accumulating passes hi*hi, hi*mid, mid*hi,
hi*lo, lo*hi, mid*mid through the f32 epilogues. Semantics follow the reference's batched DiT.forward exactly
(dit.py:165-234; TextEmbedding over the whole padded batch, encoder.py:68-96).

    v = PreciseDiT(model.cfm.backbone).forward(x, cond, text, time, mask=mask, cfg_infer=True)
"""

from __future__ import annotations

import math
from ctypes import c_float, c_int32, c_int64, c_void_p

import torch

from . import _lib as L
from . import _lib_train as T
from ._lib import _check, _ld, _ptr, _stream, lib
from .engine import BF16, F32, TILE, _rup, pack_conv_pos

PRECISE_SYMBOLS = ("oron_split3_bf16", "oron_time_sinusoid_f32", "oron_rope_f32", "oron_attention_f32", "oron_grn_f32",
                   "oron_act_f32_precise", "oron_add_f32")
_P, _I, _L, _F = c_void_p, c_int32, c_int64, c_float
_ARGTYPES = {
    "oron_split3_bf16": [_P, _L, _L, _I, _P, _P, _P, _L, _P],
    "oron_time_sinusoid_f32": [_P, _I, _P, _L, _P],
    "oron_rope_f32": [_P, _L, _I, _I, _I, _P, _P, _P],
    "oron_attention_f32": [_P, _L, _P, _L, _I, _I, _I, _P, _F, _P],
    "oron_grn_f32": [_P, _L, _I, _I, _I, _P, _P, _P, _P, _P],
    "oron_act_f32_precise": [_P, _L, _L, _I, _I, _P, _L, _I, _P, _P],
    "oron_add_f32": [_P, _L, _P, _L, _L, _I, _P, _L, _P],
}
_bound = False
# (activation term, weight term) of the six passes; dropped: mid*lo, lo*mid, lo*lo (< 2^-24 relative)
PASSES = ((0, 0), (0, 1), (1, 0), (0, 2), (2, 0), (1, 1))


def plib():
    global _bound
    Lb = lib()
    if not _bound:
        for name, at in _ARGTYPES.items():
            getattr(Lb, name).argtypes = at
        _bound = True
    return Lb


def split3_host(w: torch.Tensor, kpad: int | None = None) -> list[torch.Tensor]:
    """Weight [N, K] f32 -> three bf16 terms [N, kpad] (zero padded), exact to 24 bits. Load-time layout work."""
    w = w.detach().float()
    k = w.shape[1] if kpad is None else kpad
    out = []
    rest = w
    for _ in range(3):
        part = rest.to(BF16)
        buf = torch.zeros(w.shape[0], k, device=w.device, dtype=BF16)
        buf[:, : w.shape[1]] = part
        out.append(buf)
        rest = rest - part.float()
    return out


class PreciseDiT:
    def __init__(self, dit: torch.nn.Module):
        p0 = next(dit.parameters())
        if not p0.is_cuda:
            raise RuntimeError("PreciseDiT runs only on a CUDA device (oron_tts_b200 has no CPU fallback)")
        sd = {k: v.detach().float() for k, v in dit.state_dict().items()}
        dev = p0.device
        self.dev = dev
        self.D = D = sd["proj_out.weight"].shape[1]
        self.M = sd["proj_out.weight"].shape[0]
        self.C = C = sd["text_embed.text_embed.weight"].shape[1]
        self.depth = 1 + max(int(k.split(".")[1]) for k in sd if k.startswith("transformer_blocks."))
        tb = [int(k.split(".")[2]) for k in sd if k.startswith("text_embed.text_blocks.")]
        self.conv_layers = (1 + max(tb)) if tb else 0
        if 2 * sd["rotary_embed.inv_freq"].shape[0] != 64:
            raise NotImplementedError("the fp32 mode handles head_dim 64 only")
        self.heads = D // 64
        self.inv_freq = sd["rotary_embed.inv_freq"].contiguous()
        f = lambda k: sd[k].contiguous()  # noqa: E731
        self.t0, self.t0_b = split3_host(sd["time_embed.time_mlp.0.weight"]), f("time_embed.time_mlp.0.bias")
        self.t2, self.t2_b = split3_host(sd["time_embed.time_mlp.2.weight"]), f("time_embed.time_mlp.2.bias")
        ws = [sd[f"transformer_blocks.{i}.attn_norm.linear.weight"] for i in range(self.depth)] + [sd["norm_out.linear.weight"]]
        bs = [sd[f"transformer_blocks.{i}.attn_norm.linear.bias"] for i in range(self.depth)] + [sd["norm_out.linear.bias"]]
        self.ada, self.ada_b = split3_host(torch.cat(ws, 0)), torch.cat(bs, 0).contiguous()
        self.ada_n = self.ada_b.numel()
        self.table = f("text_embed.text_embed.weight")
        self.text_blocks = []
        for i in range(self.conv_layers):
            p = f"text_embed.text_blocks.{i}."
            self.text_blocks.append(dict(
                dw_w=sd[p + "dwconv.weight"].reshape(C, 7).contiguous(), dw_b=f(p + "dwconv.bias"), ln_w=f(p + "norm.weight"),
                ln_b=f(p + "norm.bias"), w1=split3_host(sd[p + "pwconv1.weight"]), b1=f(p + "pwconv1.bias"),
                gamma=sd[p + "grn.gamma"].reshape(-1).contiguous(), beta=sd[p + "grn.beta"].reshape(-1).contiguous(),
                w2=split3_host(sd[p + "pwconv2.weight"]), b2=f(p + "pwconv2.bias")))
        self.kin = _rup(2 * self.M + C, 64)
        self.win, self.in_b = split3_host(sd["input_embed.proj.weight"], self.kin), f("input_embed.proj.bias")
        self.conv_pos = []
        for idx in (0, 2):
            w = sd[f"input_embed.conv_pos_embed.conv1d.{idx}.weight"]
            parts, rest, gsz = [], w, None
            for _ in range(3):
                part = rest.to(BF16).float()
                pk = pack_conv_pos(part, D)
                parts.append(pk["w"])
                gsz = pk["gsz"]
                rest = rest - part
            self.conv_pos.append(dict(w=parts, b=f(f"input_embed.conv_pos_embed.conv1d.{idx}.bias"), taps=w.shape[2], gsz=gsz))
        self.blocks = []
        for i in range(self.depth):
            p = f"transformer_blocks.{i}."
            self.blocks.append(dict(
                wqkv=split3_host(torch.cat([sd[p + "attn.to_q.weight"], sd[p + "attn.to_k.weight"], sd[p + "attn.to_v.weight"]], 0)),
                bqkv=torch.cat([sd[p + "attn.to_q.bias"], sd[p + "attn.to_k.bias"], sd[p + "attn.to_v.bias"]], 0).contiguous(),
                wo=split3_host(sd[p + "attn.to_out.0.weight"]), bo=f(p + "attn.to_out.0.bias"),
                w1=split3_host(sd[p + "ff.ff.0.weight"]), b1=f(p + "ff.ff.0.bias"),
                w2=split3_host(sd[p + "ff.ff.3.weight"]), b2=f(p + "ff.ff.3.bias")))
        self.wp, self.bp = split3_host(sd["proj_out.weight"]), f("proj_out.bias")

    # ---- helpers ---------------------------------------------------------------------------------------------------
    def _split(self, x: torch.Tensor, kpad: int | None = None) -> list[torch.Tensor]:
        rows, c = x.shape
        k = _rup(c, 64) if kpad is None else kpad
        parts = [torch.zeros(rows, k, device=self.dev, dtype=BF16) for _ in range(3)]
        _check(plib().oron_split3_bf16(_ptr(x, F32, "x"), _ld(x), rows, c, _ptr(parts[0]), _ptr(parts[1]), _ptr(parts[2]), k,
                                       _stream()), "oron_split3_bf16")
        return parts

    @staticmethod
    def _bn(n: int) -> int:
        return 256 if n % 256 == 0 else (128 if n >= 128 else 64)

    def _gemm6(self, a3, w3, out, *, epilogue, bias=None, addend=None, block_n=None, **kw) -> None:
        n = w3[0].shape[0]
        bn = self._bn(n) if block_n is None else block_n
        for idx, (i, j) in enumerate(PASSES):
            first = idx == 0
            extra = dict(kw)
            if epilogue == L.EPI_F32:
                extra["addend"] = addend if first else out
            elif epilogue == L.EPI_SCALE_RESID:
                extra["addend"] = addend if first else out
            L.gemm(a3[i], w3[j], out, epilogue=epilogue, bias=bias if first else None, block_n=bn, **extra)

    def _act(self, x: torch.Tensor, out: torch.Tensor, act: int, rpb: int = 0, seq_lens: torch.Tensor | None = None) -> None:
        _check(plib().oron_act_f32_precise(_ptr(x, F32, "x"), _ld(x), x.shape[0], x.shape[1], act, _ptr(out, F32, "out"), _ld(out),
                                           rpb, _ptr(seq_lens, torch.int32, "seq_lens"), _stream()), "oron_act_f32_precise")

    # ---- forward (dit.py:165-234) --------------------------------------------------------------------------------------
    @torch.no_grad()
    def forward(self, x: torch.Tensor, cond: torch.Tensor, text: torch.Tensor, time: torch.Tensor,
                mask: torch.Tensor | None = None, drop_audio_cond: bool = False, drop_text: bool = False,
                cfg_infer: bool = False) -> torch.Tensor:
        dev, D, C, M = self.dev, self.D, self.C, self.M
        B, Tn, _ = x.shape
        if time.ndim == 0:
            time = time.repeat(B)
        branches = [(False, False), (True, True)] if cfg_infer else [(drop_audio_cond, drop_text)]
        nbp, tpad = B * len(branches), _rup(Tn, TILE)
        R = nbp * tpad
        z = lambda *s, dt=F32: torch.zeros(*s, device=dev, dtype=dt)  # noqa: E731
        durations = [Tn] * B if mask is None else [int(v) for v in mask.sum(dim=-1).tolist()]
        seq_lens = torch.tensor(durations * len(branches), device=dev, dtype=torch.int32)
        text_lens = torch.full((nbp,), Tn, device=dev, dtype=torch.int32)
        ids = (text.to(torch.int64) + 1)[:, :Tn]
        ids2 = z(B, tpad, dt=torch.int32)
        ids2[:, : ids.shape[1]] = ids.to(torch.int32)
        ids_all = ids2.repeat(len(branches), 1).reshape(-1).contiguous()
        drop = torch.tensor([int(dt_) for (_, dt_) in branches for _ in range(B)], device=dev, dtype=torch.uint8)
        row_valid = z(R, dt=torch.uint8)
        common = dict(rows_per_batch=tpad, nbatch=nbp)
        # -- text embedding
        n = max(8192, tpad)
        freqs = 1.0 / (10000 ** (torch.arange(0, C, 2)[: (C // 2)].float() / C))
        ang = torch.outer(torch.arange(n), freqs).float()
        pos_table = torch.cat([torch.cos(ang), torch.sin(ang)], dim=-1).to(dev).contiguous()
        xt = z(R, C)
        L.text_embed_front(ids_all, drop, self.table, pos_table, rows_per_batch=tpad, nb=nbp, x=xt, row_valid=row_valid)
        for blk in self.text_blocks:
            conv, nrm, pre, nxt = z(R, C), z(R, C), z(R, 2 * C), z(R, C)
            T.dwconv7(xt, conv, seq_lens=text_lens, w=blk["dw_w"], bias=blk["dw_b"], **common)
            L.ln_modulate(conv, eps=1e-6, scale=blk["ln_w"], shift=blk["ln_b"], add_one=False, out_f32=nrm, **common)
            self._gemm6(self._split(nrm), blk["w1"], pre, epilogue=L.EPI_F32, bias=blk["b1"], **common)
            self._act(pre, pre, L.ACT_GELU_ERF)
            gx2 = z(nbp, 2 * C)
            _check(plib().oron_grn_f32(_ptr(pre, F32), _ld(pre), tpad, nbp, 2 * C, _ptr(text_lens, torch.int32),
                                       _ptr(blk["gamma"], F32), _ptr(blk["beta"], F32), _ptr(gx2, F32), _stream()), "oron_grn_f32")
            self._gemm6(self._split(pre), blk["w2"], nxt, epilogue=L.EPI_SCALE_RESID, bias=blk["b2"], addend=xt,
                        row_valid=row_valid, **common)
            xt = nxt
        # -- input embedding: Linear(cat[x, cond, text_embed]) (dit.py:53)
        a_in = z(R, 2 * M + C)
        av = a_in.view(nbp, tpad, 2 * M + C)
        for bi, (da, _) in enumerate(branches):
            av[bi * B:(bi + 1) * B, :Tn, :M] = x.float()
            if not da:
                av[bi * B:(bi + 1) * B, :Tn, M:2 * M] = cond.float()
        av[:, :, 2 * M:] = xt.view(nbp, tpad, C)
        h_raw, h0 = z(R, D), z(R, D)
        self._gemm6(self._split(a_in, self.kin), self.win, h_raw, epilogue=L.EPI_F32, bias=self.in_b, **common)
        self._act(h_raw, h0, 0, tpad, seq_lens)  # masked copy (modules.py:136)
        c1, c2 = self.conv_pos
        conv = lambda cp: dict(taps=cp["taps"], cin_blocks=cp["gsz"] // 64, pad=cp["taps"] // 2, grouped=cp["gsz"],  # noqa: E731
                               block_n=64, **common)
        z1, z2 = z(R, D), z(R, D)
        self._gemm6(self._split(h0), c1["w"], z1, epilogue=L.EPI_F32, bias=c1["b"], **conv(c1))
        self._act(z1, z1, 4, tpad, seq_lens)
        self._gemm6(self._split(z1), c2["w"], z2, epilogue=L.EPI_F32, bias=c2["b"], **conv(c2))
        self._act(z2, z2, 4, tpad, seq_lens)
        xres = z(R, D)
        _check(plib().oron_add_f32(_ptr(h0, F32), _ld(h0), _ptr(z2, F32), _ld(z2), R, D, _ptr(xres, F32), _ld(xres), _stream()),
               "oron_add_f32")
        # -- timestep conditioning and all AdaLN projections
        tf = z(B, 256)
        tvals = time.to(dev, F32).contiguous()
        _check(plib().oron_time_sinusoid_f32(_ptr(tvals, F32), B, _ptr(tf, F32), _ld(tf), _stream()), "oron_time_sinusoid_f32")
        pre0, pre2, table = z(B, D), z(B, D), z(B, self.ada_n)
        self._gemm6(self._split(tf), self.t0, pre0, epilogue=L.EPI_F32, bias=self.t0_b)
        self._act(pre0, pre0, L.ACT_SILU)
        self._gemm6(self._split(pre0), self.t2, pre2, epilogue=L.EPI_F32, bias=self.t2_b)
        self._act(pre2, pre2, L.ACT_SILU)
        self._gemm6(self._split(pre2), self.ada, table, epilogue=L.EPI_F32, bias=self.ada_b)
        # -- transformer blocks
        tab, an = table.view(-1), self.ada_n
        t = torch.arange(tpad, device=dev).float()
        angr = torch.outer(t, self.inv_freq)
        cos, sin = angr.cos().contiguous(), angr.sin().contiguous()
        mod = dict(mod_ld=an, mod_nb=B, add_one=True, eps=1e-6, **common)
        gate = dict(gate_ld=an, gate_nb=B, seq_lens=seq_lens, **common)
        nrm, qkv, ao, hid = z(R, D), z(R, 3 * D), z(R, D), z(R, 4 * D)
        for i, blk in enumerate(self.blocks):
            o = i * 6 * D
            L.ln_modulate(xres, scale=tab[o + D:], shift=tab[o:], out_f32=nrm, **mod)
            self._gemm6(self._split(nrm), blk["wqkv"], qkv, epilogue=L.EPI_F32, bias=blk["bqkv"], **common)
            _check(plib().oron_rope_f32(_ptr(qkv, F32), _ld(qkv), tpad, nbp, self.heads, _ptr(cos, F32), _ptr(sin, F32), _stream()),
                   "oron_rope_f32")
            _check(plib().oron_attention_f32(_ptr(qkv, F32), _ld(qkv), _ptr(ao, F32), _ld(ao), nbp, tpad, self.heads,
                                             _ptr(seq_lens, torch.int32), 1.0 / math.sqrt(64.0), _stream()), "oron_attention_f32")
            self._gemm6(self._split(ao), blk["wo"], xres, epilogue=L.EPI_GATE_RESID, bias=blk["bo"], gate=tab[o + 2 * D:],
                        mask_rows=True, **gate)
            L.ln_modulate(xres, scale=tab[o + 4 * D:], shift=tab[o + 3 * D:], out_f32=nrm, **mod)
            hv = hid[:, : blk["w1"][0].shape[0]]
            self._gemm6(self._split(nrm), blk["w1"], hv, epilogue=L.EPI_F32, bias=blk["b1"], **common)
            self._act(hv, hv, L.ACT_GELU_TANH)
            self._gemm6(self._split(hv), blk["w2"], xres, epilogue=L.EPI_GATE_RESID, bias=blk["b2"], gate=tab[o + 5 * D:],
                        mask_rows=False, **gate)
        o = self.depth * 6 * D
        L.ln_modulate(xres, scale=tab[o:], shift=tab[o + D:], out_f32=nrm, **mod)
        v = z(R, _rup(M, 4))
        self._gemm6(self._split(nrm), self.wp, v, epilogue=L.EPI_F32, bias=self.bp, block_n=128, n=M, **common)
        return v.view(nbp, tpad, -1)[:, :Tn, :M].clone()
