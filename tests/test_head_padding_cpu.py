"""Host logic of the head-width padding (engine.DiTWeights): heads narrower than 64 are packed into the kernels' 64-wide
layout. The packed QKV / out-projection operands must reproduce the reference attention block (modules.py:92-104 rotate-half
RoPE, 264-282 attention) when evaluated the way the kernels evaluate them: RoPE on pairs (i, i + 32) of every 64-wide head
with the packed frequency table, softmax scale 1/sqrt(true head width). Pure torch on CPU: no kernel is called."""

import math
import os
import sys

import pytest
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))
import weights as GW  # noqa: E402

from oron_tts_b200.engine import DiTWeights  # noqa: E402
from oron_tts_b200.f5tts import F5TTS  # noqa: E402


def _reference_attention(sd, x, heads, dh):
    """The reference block on the original weights: q/k/v projections, rotate-half RoPE over the true head width, SDPA."""
    p = "transformer_blocks.0.attn."
    T = x.shape[0]
    q = x @ sd[p + "to_q.weight"].t() + sd[p + "to_q.bias"]
    k = x @ sd[p + "to_k.weight"].t() + sd[p + "to_k.bias"]
    v = x @ sd[p + "to_v.weight"].t() + sd[p + "to_v.bias"]
    ang = torch.outer(torch.arange(T).float(), sd["rotary_embed.inv_freq"].float())
    cos, sin = torch.cat([ang, ang], -1).cos(), torch.cat([ang, ang], -1).sin()

    def rope(t):
        t = t.view(T, heads, dh)
        t1, t2 = t[..., : dh // 2], t[..., dh // 2:]
        rot = torch.cat([-t2, t1], -1)
        return t * cos[:, None, :] + rot * sin[:, None, :]

    q, k, v = rope(q), rope(k), v.view(T, heads, dh)
    att = torch.softmax(torch.einsum("thd,shd->hts", q, k) / math.sqrt(dh), dim=-1)
    o = torch.einsum("hts,shd->thd", att, v).reshape(T, heads * dh)
    return o @ sd[p + "to_out.0.weight"].t() + sd[p + "to_out.0.bias"]


def _packed_attention(w, x):
    """The same block on the packed operands, evaluated like the kernels: 64-wide heads, RoPE pairs (i, i + 32)."""
    T = x.shape[0]
    blk = w.blocks[0]
    qkv = x @ blk["wqkv"].float().t() + blk["bqkv"]
    H, inner = w.heads, w.inner
    cos, sin = w.rope(T)  # [T, 32]

    def rope(t):
        t = t.reshape(T, H, 64)
        a, b = t[..., :32], t[..., 32:]
        return torch.cat([a * cos[:, None, :] - b * sin[:, None, :], b * cos[:, None, :] + a * sin[:, None, :]], -1)

    q, k = rope(qkv[:, :inner]), rope(qkv[:, inner:2 * inner])
    v = qkv[:, 2 * inner:].reshape(T, H, 64)
    att = torch.softmax(torch.einsum("thd,shd->hts", q, k) / math.sqrt(w.dim_head), dim=-1)
    o = torch.einsum("hts,shd->thd", att, v).reshape(T, inner)
    return o @ blk["wo"].float().t() + blk["bo"]


@pytest.mark.parametrize("cfg", ["micro", "tiny", dict(model=dict(dim=128, depth=1, heads=8, ff_mult=2, text_dim=32, conv_layers=1))])
def test_packed_heads_reproduce_reference_attention(cfg):
    conf = GW.CONFIGS[cfg] if isinstance(cfg, str) else cfg
    m = F5TTS.from_config(conf)
    bb = m.cfm.backbone
    sd = GW.fill_state_dict(bb.state_dict(), 7)
    # keep the weights bf16-representable so that the packed (bf16) copies are exact
    sd = {k: (v.to(torch.bfloat16).float() if v.is_floating_point() and "inv_freq" not in k else v) for k, v in sd.items()}
    w = DiTWeights(sd, torch.device("cpu"))
    D = conf["model"]["dim"]
    heads = conf["model"]["heads"]
    dh = D // heads
    assert w.heads == heads and w.dim_head == dh and w.inner == heads * 64
    x = torch.randn(37, D, generator=torch.Generator().manual_seed(1))
    ref = _reference_attention(sd, x, heads, dh)
    got = _packed_attention(w, x)
    assert float((got - ref).abs().max()) < 2e-5 * float(ref.abs().max() + 1)
    if dh < 64:  # pad channels are exact zeros: q / k / v rows and biases, out-projection columns
        rows = torch.ones(w.inner, dtype=torch.bool)
        rows[w._head_rows.cpu()] = False
        blk = w.blocks[0]
        for part in range(3):
            assert float(blk["wqkv"][part * w.inner:(part + 1) * w.inner][rows].abs().max()) == 0.0
            assert float(blk["bqkv"][part * w.inner:(part + 1) * w.inner][rows].abs().max()) == 0.0
        assert float(blk["wo"][:, rows].abs().max()) == 0.0


def test_wide_heads_are_rejected():
    m = F5TTS.from_config(dict(model=dict(dim=256, depth=1, heads=2, ff_mult=2, text_dim=32, conv_layers=1)))
    with pytest.raises(NotImplementedError):
        DiTWeights(m.cfm.backbone.state_dict(), torch.device("cpu"))
