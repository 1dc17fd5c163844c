"""Deterministic, shape-driven weights shared by the golden generator (reference side, build container)
and the tests (oracle / CUDA side, any box): same torch CPU generator stream => identical tensors.

Default reference init gives velocity == 0 (src/models/dit.py:119-129 zero-inits every AdaLN projection
and the output head), which would make parity vacuous; here every tensor is re-drawn, with the
zero-initialised families at sigma = 0.02 as in SURVEY.md §0.
"""

from __future__ import annotations

import math

import torch

SMALL_TENSORS = ("attn_norm.linear", "norm_out.linear", "proj_out", "grn.gamma", "grn.beta")


def draw(name: str, shape: tuple, gen: torch.Generator) -> torch.Tensor:
    if any(tag in name for tag in SMALL_TENSORS):
        return torch.randn(shape, generator=gen) * 0.02
    if name.endswith("norm.weight") or name.endswith("final_layer_norm.weight"):
        return 1.0 + 0.02 * torch.randn(shape, generator=gen)
    if name.endswith("norm.bias") or name.endswith("final_layer_norm.bias"):
        return 0.02 * torch.randn(shape, generator=gen)
    if name.endswith("gamma"):  # Vocos layer scale
        return 0.3 + 0.05 * torch.randn(shape, generator=gen)
    if name.endswith("text_embed.weight"):
        return torch.randn(shape, generator=gen)
    fan_in = max(1, math.prod(shape[1:])) if len(shape) >= 2 else None
    if fan_in is None:  # bias of a linear / conv layer
        return (torch.rand(shape, generator=gen) * 2 - 1) * 0.05
    bound = 1.0 / math.sqrt(fan_in)
    return (torch.rand(shape, generator=gen) * 2 - 1) * bound


def fill_state_dict(reference_sd: dict, seed: int) -> dict:
    """New state dict with the same keys/shapes; buffers that are not parameters (inv_freq, window) are kept."""
    gen = torch.Generator(device="cpu").manual_seed(seed)
    out = {}
    for name in sorted(reference_sd):
        t = reference_sd[name]
        if name.endswith("inv_freq") or name.endswith("window") or not t.is_floating_point():
            out[name] = t.detach().clone()
        else:
            out[name] = draw(name, tuple(t.shape), gen).to(torch.float32)
    return out


CONFIGS = {
    "tiny": {"model": dict(dim=128, depth=2, heads=2, text_dim=64, conv_layers=2)},
    # the reference's own test configuration (tests/test_checkpoint.py:9-24): head_dim 32, dim 64, text_dim 32
    "micro": {"model": dict(vocab_size=65, dim=64, depth=1, heads=2, ff_mult=2, text_dim=32, conv_layers=1)},
    "small": {"model": dict(dim=512, depth=12, heads=8, text_dim=256, conv_layers=4)},
    "base": {"model": dict(dim=1024, depth=22, heads=16, text_dim=512, conv_layers=4)},
}
SEEDS = {"tiny": 1234, "micro": 1234, "small": 1234, "base": 1234}
