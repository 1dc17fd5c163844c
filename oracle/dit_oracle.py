"""ORACLE — TEST INFRASTRUCTURE ONLY (tests/, __graft_entry__.smoke() and bench.py's CPU legs).

Plain-PyTorch fp32, CPU, functional restatement of the reference's CFM.sample -> DiT.forward path,
driven directly by a reference-layout ``state_dict`` (key names of SURVEY.md §8b). Nothing in
oron_tts_b200/ imports this module; the product path has no CPU fallback.

Pinned against the live reference (imported from /root/reference in the build container) by
tests/golden/make_golden.py -> tests/golden/*.pt and checked in tests/test_oracle_golden.py.
Each function cites the reference lines it restates (paths relative to the reference tree).
"""

from __future__ import annotations

import math

import torch
import torch.nn.functional as F

Tensor = torch.Tensor
SD = dict  # state dict: str -> Tensor

BB = "cfm.backbone."


def _lin(sd: SD, name: str, x: Tensor) -> Tensor:
    return F.linear(x, sd[name + ".weight"], sd[name + ".bias"])


def model_dims(sd: SD) -> dict:
    """Recover the architecture hyper-parameters from tensor shapes alone."""
    dim = sd[BB + "proj_out.weight"].shape[1]
    n_mels = sd[BB + "proj_out.weight"].shape[0]
    text_dim = sd[BB + "text_embed.text_embed.weight"].shape[1]
    depth = 1 + max(int(k.split(".")[3]) for k in sd if k.startswith(BB + "transformer_blocks."))
    convs = [int(k.split(".")[4]) for k in sd if k.startswith(BB + "text_embed.text_blocks.")]
    dim_head = 2 * sd[BB + "rotary_embed.inv_freq"].shape[0]
    return dict(dim=dim, n_mels=n_mels, text_dim=text_dim, depth=depth, conv_layers=(1 + max(convs)) if convs else 0,
                dim_head=dim_head, heads=dim // dim_head)


# --------------------------------------------------------------------------------------------
# timestep conditioning — src/models/modules.py:39-45 (sinusoid), :60-62 (MLP)
# --------------------------------------------------------------------------------------------
def time_embedding(sd: SD, t: Tensor) -> Tensor:
    half = 128
    freqs = torch.exp(torch.arange(half, dtype=torch.float32) * -(math.log(10000) / (half - 1)))
    ang = 1000.0 * t[:, None] * freqs[None, :]
    feat = torch.cat([ang.sin(), ang.cos()], dim=-1)
    h = _lin(sd, BB + "time_embed.time_mlp.0", feat)
    return _lin(sd, BB + "time_embed.time_mlp.2", F.silu(h))


# --------------------------------------------------------------------------------------------
# text embedding — src/models/encoder.py:68-96, ConvNeXtV2 block modules.py:175-185, GRN :153-156
# --------------------------------------------------------------------------------------------
def text_pos_table(text_dim: int, length: int) -> Tensor:
    # modules.py:191-196
    inv = 1.0 / (10000 ** (torch.arange(0, text_dim, 2)[: text_dim // 2].float() / text_dim))
    ang = torch.outer(torch.arange(length), inv).float()
    return torch.cat([ang.cos(), ang.sin()], dim=-1)


def text_embedding(sd: SD, text: Tensor, seq_len: int, drop_text: bool) -> Tensor:
    """text: int64 [B, Nt] raw ids (-1 = padding). Returns [B, seq_len, text_dim]."""
    d = model_dims(sd)
    ids = text + 1
    ids = ids[:, :seq_len]
    ids = F.pad(ids, (0, seq_len - ids.shape[1]), value=0)
    filler = ids == 0  # decided before the text is dropped (encoder.py:77-80)
    if drop_text:
        ids = torch.zeros_like(ids)
    x = sd[BB + "text_embed.text_embed.weight"][ids]
    if d["conv_layers"] == 0:
        return x
    x = x + text_pos_table(d["text_dim"], seq_len)[None]
    x = x.masked_fill(filler[..., None], 0.0)
    for i in range(d["conv_layers"]):
        p = f"{BB}text_embed.text_blocks.{i}."
        y = F.conv1d(x.transpose(1, 2), sd[p + "dwconv.weight"], sd[p + "dwconv.bias"], padding=3,
                     groups=x.shape[-1]).transpose(1, 2)
        y = F.layer_norm(y, (y.shape[-1],), sd[p + "norm.weight"], sd[p + "norm.bias"], eps=1e-6)
        y = F.gelu(_lin(sd, p + "pwconv1", y))
        gx = torch.linalg.vector_norm(y, ord=2, dim=1, keepdim=True)       # over frames
        nx = gx / (gx.mean(dim=-1, keepdim=True) + 1e-6)
        y = sd[p + "grn.gamma"] * (y * nx) + sd[p + "grn.beta"] + y
        x = x + _lin(sd, p + "pwconv2", y)
        x = x.masked_fill(filler[..., None], 0.0)
    return x


# --------------------------------------------------------------------------------------------
# input embedding — src/models/dit.py:35-55, ConvPositionEmbedding modules.py:131-141
# --------------------------------------------------------------------------------------------
def input_embedding(sd: SD, x: Tensor, cond: Tensor, text_embed: Tensor, drop_audio_cond: bool, mask: Tensor | None) -> Tensor:
    if drop_audio_cond:
        cond = torch.zeros_like(cond)
    h = _lin(sd, BB + "input_embed.proj", torch.cat([x, cond, text_embed], dim=-1))
    groups = 16
    pad = sd[BB + "input_embed.conv_pos_embed.conv1d.0.weight"].shape[-1] // 2
    y = h.transpose(1, 2)
    keep = None if mask is None else mask[:, None, :]
    if keep is not None:
        y = y * keep
    for idx in (0, 2):
        p = f"{BB}input_embed.conv_pos_embed.conv1d.{idx}."
        y = F.conv1d(y, sd[p + "weight"], sd[p + "bias"], padding=pad, groups=groups)
        if keep is not None:
            y = y * keep
        y = F.mish(y)
    return y.transpose(1, 2) + h


# --------------------------------------------------------------------------------------------
# transformer block — modules.py:214-219 (AdaLN), :264-283 (attention), :294-299 (FFN), :334-345
# --------------------------------------------------------------------------------------------
def rope_tables(sd: SD, length: int) -> tuple[Tensor, Tensor]:
    ang = torch.outer(torch.arange(length).float(), sd[BB + "rotary_embed.inv_freq"].float())
    ang = torch.cat([ang, ang], dim=-1)
    return ang.cos(), ang.sin()


def _rope(x: Tensor, cos: Tensor, sin: Tensor) -> Tensor:
    half = x.shape[-1] // 2
    rot = torch.cat([-x[..., half:], x[..., :half]], dim=-1)
    return x * cos + rot * sin


def attention(sd: SD, p: str, x: Tensor, mask: Tensor | None, heads: int, cos: Tensor, sin: Tensor) -> Tensor:
    B, T, D = x.shape
    dh = D // heads
    q = _lin(sd, p + "to_q", x).view(B, T, heads, dh).transpose(1, 2)
    k = _lin(sd, p + "to_k", x).view(B, T, heads, dh).transpose(1, 2)
    v = _lin(sd, p + "to_v", x).view(B, T, heads, dh).transpose(1, 2)
    q, k = _rope(q, cos, sin), _rope(k, cos, sin)
    s = (q @ k.transpose(-1, -2)) / math.sqrt(dh)
    if mask is not None:
        s = s.masked_fill(~mask[:, None, None, :], float("-inf"))
    o = (s.softmax(dim=-1) @ v).transpose(1, 2).reshape(B, T, D)
    o = _lin(sd, p + "to_out.0", o)
    if mask is not None:
        o = o * mask[..., None]
    return o


def dit_block(sd: SD, i: int, x: Tensor, t_emb: Tensor, mask: Tensor | None, heads: int, cos: Tensor, sin: Tensor) -> Tensor:
    p = f"{BB}transformer_blocks.{i}."
    D = x.shape[-1]
    mod = _lin(sd, p + "attn_norm.linear", F.silu(t_emb))
    shift_a, scale_a, gate_a, shift_m, scale_m, gate_m = mod.split(D, dim=1)
    n = F.layer_norm(x, (D,), eps=1e-6) * (1 + scale_a[:, None]) + shift_a[:, None]
    x = x + gate_a[:, None] * attention(sd, p + "attn.", n, mask, heads, cos, sin)
    n = F.layer_norm(x, (D,), eps=1e-6) * (1 + scale_m[:, None]) + shift_m[:, None]
    ff = _lin(sd, p + "ff.ff.3", F.gelu(_lin(sd, p + "ff.ff.0", n), approximate="tanh"))
    return x + gate_m[:, None] * ff


# --------------------------------------------------------------------------------------------
# DiT.forward — src/models/dit.py:165-234 (cfg_infer doubles the batch: [cond ; uncond])
# --------------------------------------------------------------------------------------------
def dit_forward(sd: SD, x: Tensor, cond: Tensor, text: Tensor, time: Tensor, mask: Tensor | None = None,
                drop_audio_cond: bool = False, drop_text: bool = False, cfg_infer: bool = False,
                text_cache: dict | None = None) -> Tensor:
    d = model_dims(sd)
    B, T, _ = x.shape
    if time.ndim == 0:
        time = time.repeat(B)
    t_emb = time_embedding(sd, time.float())

    def embed(drop_a: bool, drop_t: bool) -> Tensor:
        key = "uncond" if drop_t else "cond"
        if text_cache is not None and key in text_cache:
            te = text_cache[key]
        else:
            te = text_embedding(sd, text, T, drop_t)
            if text_cache is not None:
                text_cache[key] = te
        return input_embedding(sd, x, cond, te, drop_a, mask)

    if cfg_infer:
        h = torch.cat([embed(False, False), embed(True, True)], dim=0)
        t_emb = torch.cat([t_emb, t_emb], dim=0)
        m = None if mask is None else torch.cat([mask, mask], dim=0)
    else:
        h = embed(drop_audio_cond, drop_text)
        m = mask
    cos, sin = rope_tables(sd, T)
    for i in range(d["depth"]):
        h = dit_block(sd, i, h, t_emb, m, d["heads"], cos, sin)
    mod = _lin(sd, BB + "norm_out.linear", F.silu(t_emb))
    scale, shift = mod.split(d["dim"], dim=1)
    h = F.layer_norm(h, (d["dim"],), eps=1e-6) * (1 + scale)[:, None] + shift[:, None]
    return _lin(sd, BB + "proj_out", h)


# --------------------------------------------------------------------------------------------
# CFM.sample — src/models/flow.py:161-306
# --------------------------------------------------------------------------------------------
def sway_schedule(steps: int, sway: float | None) -> Tensor:
    t = torch.linspace(0, 1, steps + 1, dtype=torch.float32)
    if sway is not None:
        t = t + sway * (torch.cos(torch.pi / 2 * t) - 1 + t)
    return t


def seeded_noise(durations: list[int], n_mels: int, seed: int | None) -> Tensor:
    """Per-sample sequential draws from one CPU generator, zero padded (flow.py:270-283)."""
    gen = None if seed is None else torch.Generator(device="cpu").manual_seed(seed)
    ys = [torch.randn(d, n_mels, generator=gen) for d in durations]
    return torch.nn.utils.rnn.pad_sequence(ys, padding_value=0.0, batch_first=True)


@torch.inference_mode()
def cfm_sample(sd: SD, cond: Tensor, text_ids: Tensor, duration: Tensor | int, *, lens: Tensor | None = None,
               steps: int = 32, cfg_strength: float = 1.0, sway_sampling_coef: float | None = None,
               seed: int | None = None, y0: Tensor | None = None, return_velocity: bool = False, method: str = "euler"):
    """method "midpoint": explicit midpoint rule on the same schedule (an extension over the reference's Euler loop,
    flow.py:290-299) -- x_{i+1} = x_i + dt_i v(x_i + dt_i/2 v(x_i, t_i), t_i + dt_i/2)."""
    d = model_dims(sd)
    B, T_ref, _ = cond.shape
    lens = torch.full((B,), T_ref, dtype=torch.long) if lens is None else lens.long()
    duration = torch.full((B,), duration, dtype=torch.long) if isinstance(duration, int) else duration.long()
    T = int(duration.max())
    idx = torch.arange(T)
    cond_mask = idx[None, :] < lens[:, None]
    cond = F.pad(cond, (0, 0, 0, T - T_ref))
    step_cond = torch.where(cond_mask[..., None], cond, torch.zeros_like(cond))
    attn_mask = idx[None, :] < duration[:, None]
    if y0 is None:
        y0 = seeded_noise([int(v) for v in duration], d["n_mels"], seed)
    t = sway_schedule(steps, sway_sampling_coef)
    cache: dict = {}
    x = y0
    traj, vels = [y0], []
    def velocity(xx: Tensor, tt: Tensor) -> Tensor:
        tb = tt.expand(B)
        if cfg_strength < 1e-5:
            return dit_forward(sd, xx, step_cond, text_ids, tb, attn_mask, text_cache=cache)
        both = dit_forward(sd, xx, step_cond, text_ids, tb, attn_mask, cfg_infer=True, text_cache=cache)
        vc, vu = both[:B], both[B:]
        return vc + (vc - vu) * cfg_strength

    for i in range(steps):
        dt = t[i + 1] - t[i]
        v = velocity(x, t[i])
        if method == "midpoint":
            v = velocity(x + v * (0.5 * dt), 0.5 * (t[i] + t[i + 1]))
        vels.append(v)
        x = x + v * dt
        traj.append(x)
    out = torch.where(cond_mask[..., None], cond, x)
    if return_velocity:
        return out, traj, vels
    return out, traj


# --------------------------------------------------------------------------------------------
# CFM.forward (training objective) — src/models/flow.py:69-159. Differentiable (plain torch ops), so
# torch.autograd over this function is the gradient oracle of the training step.
# --------------------------------------------------------------------------------------------
def cfm_eval_draws(x1: Tensor, lens: Tensor, frac_lengths_mask: tuple[float, float] = (0.7, 1.0)) -> dict:
    """The deterministic eval-mode choices (flow.py:113-128, 136-138): centred span, t = 0.5, seed-0 noise."""
    B, T, _ = x1.shape
    mask = torch.arange(T)[None, :] < lens[:, None]
    mid = sum(frac_lengths_mask) / 2
    span_len = (torch.full((B,), mid).float() * lens).long()
    start = ((lens - span_len) // 2).clamp(min=0)
    pos = torch.arange(T)
    span = (pos[None, :] >= start[:, None]) & (pos[None, :] < (start + span_len)[:, None]) & mask
    x0 = torch.randn(x1.shape, generator=torch.Generator().manual_seed(0), dtype=x1.dtype)
    return dict(x1=x1, x0=x0, time=torch.full((B,), 0.5, dtype=x1.dtype), span=span, drop_audio=False, drop_text=False)


def cfm_loss(sd: SD, draws: dict, text: Tensor, lens: Tensor) -> Tensor:
    x1, x0, time, span = draws["x1"], draws["x0"], draws["time"], draws["span"]
    T = x1.shape[1]
    mask = torch.arange(T)[None, :] < lens[:, None]
    t = time[:, None, None]
    phi = (1 - t) * x0 + t * x1
    cond = torch.where(span[..., None], torch.zeros_like(x1), x1)
    pred = dit_forward(sd, phi, cond, text, time, mask, drop_audio_cond=draws["drop_audio"], drop_text=draws["drop_text"])
    return F.mse_loss(pred, x1 - x0, reduction="none")[span].mean()


def cfm_loss_and_grads(sd: SD, draws: dict, text: Tensor, lens: Tensor) -> tuple[Tensor, dict]:
    """Loss and d loss / d parameter for every floating-point entry of the state dict (inv_freq excluded)."""
    leaves = {k: v.detach().clone().requires_grad_(True) for k, v in sd.items() if v.is_floating_point() and "inv_freq" not in k}
    full = dict(sd)
    full.update(leaves)
    loss = cfm_loss(full, draws, text, lens)
    grads = torch.autograd.grad(loss, list(leaves.values()), allow_unused=True)
    return loss.detach(), {k: (g if g is not None else torch.zeros_like(leaves[k])) for k, g in zip(leaves, grads)}


def adamw_clip_step(params: dict, grads: dict, state: dict, *, lr: float, step: int, max_norm: float = 1.0,
                    betas: tuple[float, float] = (0.9, 0.999), eps: float = 1e-8, weight_decay: float = 0.01) -> float:
    """clip_grad_norm_ + torch.optim.AdamW update restated (trainer.py:76-80, 206-211). In place on `params`;
    `state` holds exp_avg / exp_avg_sq per key. Returns the pre-clip global norm."""
    total = math.sqrt(sum(float((g.double() ** 2).sum()) for g in grads.values()))
    coef = min(1.0, max_norm / (total + 1e-6))
    b1, b2 = betas
    for k, p in params.items():
        g = grads[k] * coef
        m = state.setdefault(k + ".m", torch.zeros_like(p))
        v = state.setdefault(k + ".v", torch.zeros_like(p))
        p.mul_(1 - lr * weight_decay)
        m.mul_(b1).add_(g, alpha=1 - b1)
        v.mul_(b2).addcmul_(g, g, value=1 - b2)
        denom = (v.sqrt() / math.sqrt(1 - b2 ** step)).add_(eps)
        p.addcdiv_(m, denom, value=-lr / (1 - b1 ** step))
    return total
