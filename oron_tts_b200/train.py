"""OT-CFM training step on the sm_100a kernels.

Drop-in for what ``F5Trainer.train_step`` / ``_optimizer_step`` do with torch autograd (src/training/trainer.py:191-262):
``CFM.forward`` (flow.py:69-159) -> backward through DiT.forward -> global-norm clip -> AdamW, with the data-parallel
gradient mean over ``torch.distributed`` (NCCL) between backward and the optimizer.

Design:
  * master parameters, gradients, Adam moments and the bf16 GEMM copies live in five flat arenas with one layout;
    ``nn.Parameter.data`` / ``.grad`` are views, so ``state_dict()`` / checkpoints are unchanged. The arena order makes
    every fused operand of the kernels (Wqkv, the stacked AdaLN projection) a plain contiguous view, so the bf16
    copy written by the fused AdamW kernel IS the next step's GEMM operand;
  * the forward pass is the inference engine's kernel sequence with the two fused epilogues that hide a needed
    intermediate (gated residual, activations) un-fused and every activation kept (180 GB HBM: no recomputation);
  * backward: dgrad / wgrad GEMMs on the same tcgen05 GEMM reading the row-major activations / weights as stored
    (MN-major operand variants, stream-K for the few-tile weight gradients), bias gradients as column sums, GELU +
    dropout and the gated residuals fused into GEMM epilogues, tcgen05 attention backward, row-wise backward kernels
    (include/oron_b200_train.h);
  * gradients of each transformer block are all-reduced (async, NCCL stream) as soon as its backward is enqueued.

Train-mode randomness (t, span, noise, CFG drops) is drawn with torch as the reference does. Dropout (p_dropout,
modules.py:253, 297) is a stateless mask recomputed by the backward kernels from (seed, element index); the torch
Philox stream cannot be reproduced, so parity is defined on the deterministic objective and dropout is
checked statistically (tests/test_train_kernels_gpu.py::test_dropout_masks).
"""

from __future__ import annotations

import math
from random import random

import os

import torch

from . import _lib as L
from . import _lib_train as T
from .engine import BF16, F32, TILE, DiTWeights, _rup, pack_conv_pos

I32 = torch.int32


# ----------------------------------------------------------------------------------------------------------
# flat parameter / gradient / optimizer-state arenas
# ----------------------------------------------------------------------------------------------------------
class ParamArena:
    ALIGN = 64  # elements; every fused view stays 128-byte aligned for TMA in both f32 and bf16

    def __init__(self, module: torch.nn.Module, prefix: str | None = None):
        if prefix is None:  # F5TTS (cfm.backbone.*) or CFM (backbone.*)
            prefix = "cfm.backbone." if hasattr(module, "cfm") else "backbone."
        named = dict(module.named_parameters())
        keys = [k for k in named if k.startswith(prefix)]
        if len(keys) != len(named):
            raise RuntimeError("every trainable parameter is expected under " + prefix)
        bb = prefix
        depth = 1 + max(int(k[len(bb):].split(".")[1]) for k in keys if k.startswith(bb + "transformer_blocks."))
        order: list[str] = []
        # stacked AdaLN projections first (one contiguous [depth*6D + 2D, D] operand), then their biases
        order += [f"{bb}transformer_blocks.{i}.attn_norm.linear.weight" for i in range(depth)] + [bb + "norm_out.linear.weight"]
        order += [f"{bb}transformer_blocks.{i}.attn_norm.linear.bias" for i in range(depth)] + [bb + "norm_out.linear.bias"]
        self.block_ranges: list[list[tuple[int, int]]] = []
        per_block = ["attn.to_q.weight", "attn.to_k.weight", "attn.to_v.weight", "attn.to_q.bias", "attn.to_k.bias",
                     "attn.to_v.bias", "attn.to_out.0.weight", "attn.to_out.0.bias", "ff.ff.0.weight", "ff.ff.0.bias",
                     "ff.ff.3.weight", "ff.ff.3.bias"]
        block_first = {}
        for i in range(depth):
            block_first[i] = len(order)
            order += [f"{bb}transformer_blocks.{i}.{n}" for n in per_block]
        rest = [k for k in keys if k not in set(order)]
        order += sorted(rest)
        assert sorted(order) == sorted(keys)
        dev = named[order[0]].device
        self.offsets: dict[str, int] = {}
        off = 0
        for k in order:
            self.offsets[k] = off
            off += _rup(named[k].numel(), self.ALIGN)
        self.numel_used = off
        off = _rup(off, 8 * 1024)  # every arena splits evenly over 1/2/4/8 ranks with 4 KB-aligned shards (sharded optimizer)
        self.numel = off
        for i in range(depth):  # what is final once block i's backward has run: its own tensors + its AdaLN projection
            first = f"{bb}transformer_blocks.{i}.{per_block[0]}"
            last = f"{bb}transformer_blocks.{i}.{per_block[-1]}"
            aw, ab = f"{bb}transformer_blocks.{i}.attn_norm.linear.weight", f"{bb}transformer_blocks.{i}.attn_norm.linear.bias"
            self.block_ranges.append([(self.offsets[first], self.offsets[last] + _rup(named[last].numel(), self.ALIGN)),
                                      (self.offsets[aw], self.offsets[aw] + _rup(named[aw].numel(), self.ALIGN)),
                                      (self.offsets[ab], self.offsets[ab] + _rup(named[ab].numel(), self.ALIGN))])
        self.p = torch.zeros(off, device=dev, dtype=F32)
        self.g = torch.zeros(off, device=dev, dtype=F32)
        self.m = torch.zeros(off, device=dev, dtype=F32)
        self.v = torch.zeros(off, device=dev, dtype=F32)
        self.pb = torch.zeros(off, device=dev, dtype=BF16)
        self.named, self.order, self.prefix, self.depth = named, order, prefix, depth
        with torch.no_grad():
            for k in order:
                prm = named[k]
                o, n = self.offsets[k], prm.numel()
                self.p[o:o + n].copy_(prm.detach().reshape(-1))
                prm.data = self.p[o:o + n].view(prm.shape)
                prm.grad = self.g[o:o + n].view(prm.shape)
            self.pb.copy_(self.p)

    def view(self, arena: torch.Tensor, key: str, shape: tuple | None = None, span: int | None = None) -> torch.Tensor:
        """View of `arena` at parameter `key`; `span` elements (default: the parameter's own) reshaped to `shape`."""
        prm = self.named[self.prefix + key]
        o = self.offsets[self.prefix + key]
        n = prm.numel() if span is None else span
        return arena[o:o + n].view(prm.shape if shape is None else shape)


class GradReducer:
    """Sum of the flat gradient arena over the data-parallel ranks (trainer.py:70-71 wraps the model in DDP for this).
    The ranges of the transformer blocks are all-reduced asynchronously as soon as the backward pass reports them final
    (``block_done``), overlapping NCCL with the backward kernels of the earlier blocks; ``finish`` waits for those and
    reduces everything outside the block ranges (embeddings, AdaLN projections, output head). The division by the world
    size is folded into the optimizer kernel. No-op without an initialised process group."""

    def __init__(self, g: torch.Tensor, block_ranges: list[list[tuple[int, int]]], overlap: bool = True):
        self.g, self.block_ranges, self.overlap = g, block_ranges, overlap
        self._pending: list = []

    @staticmethod
    def _world() -> int:
        import torch.distributed as dist

        return dist.get_world_size() if (dist.is_available() and dist.is_initialized()) else 1

    def block_done(self, i: int) -> None:
        import torch.distributed as dist

        if self._world() > 1 and self.overlap:
            for lo, hi in self.block_ranges[i]:
                self._pending.append((dist.all_reduce(self.g[lo:hi], op=dist.ReduceOp.SUM, async_op=True), lo, hi))

    def drain(self) -> None:
        """Wait for every all-reduce in flight (before the gradient arena is zeroed or rewritten)."""
        for h, _, _ in self._pending:
            h.wait()

    def finish(self) -> None:
        import torch.distributed as dist

        if self._world() == 1:
            return
        done = sorted((lo, hi) for _, lo, hi in self._pending)
        for h, _, _ in self._pending:
            h.wait()
        self._pending.clear()
        cur = 0
        for lo, hi in done + [(self.g.numel(), self.g.numel())]:  # everything outside the already reduced ranges
            if lo > cur:
                dist.all_reduce(self.g[cur:lo], op=dist.ReduceOp.SUM)
            cur = max(cur, hi)


# Parameters the kernels (or TrainWeights.refresh) read in fp32 straight from the master arena: everything except the big
# GEMM operands, which are consumed through the bf16 copy only.
_BF16_ONLY = ("attn_norm.linear.weight", "norm_out.linear.weight", "attn.to_q.weight", "attn.to_k.weight", "attn.to_v.weight",
              "attn.to_out.0.weight", "ff.ff.0.weight", "ff.ff.3.weight", "pwconv1.weight", "pwconv2.weight",
              "time_embed.time_mlp.0.weight", "time_embed.time_mlp.2.weight")


class ShardedOptimizer:
    """ZeRO-1 style optimizer step for the data-parallel training step (config 5): every rank owns 1 / world of the flat
    arenas. Per step: reduce-scatter of the fp32 gradients (half the bytes of the all-reduce it replaces), global gradient
    norm from the shard sums, clip + AdamW on the own shard only (1 / world of the optimizer's HBM traffic), all-gather of
    the BF16 operand copy (half the bytes of an fp32 all-gather) and a small fp32 exchange for the parameters the kernels
    read in fp32 (biases, norms, embeddings, conv-position weights: a few MB). Master weights and Adam moments stay fp32,
    so the arithmetic is the reference's (trainer.py:191-216); on the ranks that do not own them, the fp32 masters of the
    bf16-only operands go stale between ``consolidate()`` calls (called by the state-dict / EMA exports)."""

    def __init__(self, arena: ParamArena):
        import torch.distributed as dist

        self.arena = arena
        self.world, self.rank = dist.get_world_size(), dist.get_rank()
        self.nccl = dist.get_backend() == "nccl"
        n = arena.numel
        assert n % self.world == 0
        self.shard = n // self.world
        self.lo, self.hi = self.rank * self.shard, (self.rank + 1) * self.shard
        idx = []
        for k in arena.order:
            if not k.endswith(_BF16_ONLY):
                o = arena.offsets[k]
                idx.append(torch.arange(o, o + arena.named[k].numel(), device=arena.p.device))
        self.small_idx = torch.cat(idx)
        mine = (self.small_idx >= self.lo) & (self.small_idx < self.hi)
        self.small_mine = mine
        self.small_idx_mine = self.small_idx[mine]
        self.small_buf = torch.zeros(self.small_idx.numel(), device=arena.p.device, dtype=F32)
        self.stale = False

    def _reduce_scatter(self, t: torch.Tensor) -> None:
        """Sum over the ranks, result only in this rank's shard of `t` (in place)."""
        import torch.distributed as dist

        if self.nccl:
            dist.reduce_scatter_tensor(t[self.lo:self.hi], t, op=dist.ReduceOp.SUM)
        else:  # gloo (CPU tests): no reduce-scatter on tensors
            dist.all_reduce(t, op=dist.ReduceOp.SUM)

    def _all_gather(self, t: torch.Tensor) -> None:
        """Every rank's shard of `t` to everybody (in place)."""
        import torch.distributed as dist

        if self.nccl:
            dist.all_gather_into_tensor(t, t[self.lo:self.hi])
        else:
            parts = [torch.empty_like(t[: self.shard]) for _ in range(self.world)]
            dist.all_gather(parts, t[self.lo:self.hi].contiguous())
            for r, prt in enumerate(parts):
                t[r * self.shard:(r + 1) * self.shard].copy_(prt)

    def reduce_gradients(self) -> None:
        self._reduce_scatter(self.arena.g)

    def exchange_after_step(self) -> None:
        """bf16 operands of every rank's shard to everybody + the fp32 values of the small parameters."""
        import torch.distributed as dist

        a = self.arena
        self._all_gather(a.pb)
        self.small_buf.zero_()
        self.small_buf[self.small_mine] = a.p[self.small_idx_mine]
        dist.all_reduce(self.small_buf, op=dist.ReduceOp.SUM)  # exactly one rank contributes a non-zero value per entry
        a.p[self.small_idx] = self.small_buf
        self.stale = True

    def consolidate(self) -> None:
        """All ranks get the complete fp32 masters and Adam moments (before a state-dict export)."""
        if self.stale:
            a = self.arena
            for t in (a.p, a.m, a.v):
                self._all_gather(t)
            self.stale = False


class TrainWeights(DiTWeights):
    """The engine's packed-weight record, as views into the arenas (bf16 operands, f32 biases / row-wise parameters)
    plus the few re-laid-out copies (input projection split, conv-position taps, transposes for the data gradients)
    that ``refresh()`` rebuilds after every optimizer step."""

    def __init__(self, arena: ParamArena, inv_freq: torch.Tensor):  # noqa: super().__init__ deliberately not called
        a = arena
        P, PB = a.p, a.pb
        pw = a.named[a.prefix + "proj_out.weight"]
        self.arena = a
        self.device = pw.device
        self.dim = D = pw.shape[1]
        self.n_mels = M = pw.shape[0]
        self.text_dim = C = a.named[a.prefix + "text_embed.text_embed.weight"].shape[1]
        self.depth = a.depth
        tb = [int(k[len(a.prefix):].split(".")[2]) for k in a.order if k.startswith(a.prefix + "text_embed.text_blocks.")]
        self.conv_layers = (1 + max(tb)) if tb else 0
        self.dim_head = 2 * inv_freq.shape[0]
        self.heads = D // self.dim_head
        if self.dim_head != 64 or D % 128 or C % 64:
            raise NotImplementedError("training kernels need head_dim 64, dim % 128 == 0, text_dim % 64 == 0")
        self.inv_freq = inv_freq.detach().to(self.device, F32).contiguous()
        self.t0_w, self.t0_b = a.view(PB, "time_embed.time_mlp.0.weight"), a.view(P, "time_embed.time_mlp.0.bias")
        self.t2_w, self.t2_b = a.view(PB, "time_embed.time_mlp.2.weight"), a.view(P, "time_embed.time_mlp.2.bias")
        self.ada_n = self.depth * 6 * D + 2 * D
        self.ada_w = a.view(PB, "transformer_blocks.0.attn_norm.linear.weight", (self.ada_n, D), self.ada_n * D)
        self.ada_b = a.view(P, "transformer_blocks.0.attn_norm.linear.bias", (self.ada_n,), self.ada_n)
        self.text_table = a.view(P, "text_embed.text_embed.weight")
        self.text_blocks = []
        for i in range(self.conv_layers):
            p = f"text_embed.text_blocks.{i}."
            self.text_blocks.append(dict(
                dw_w=a.view(P, p + "dwconv.weight", (C, 7)), dw_b=a.view(P, p + "dwconv.bias"),
                ln_w=a.view(P, p + "norm.weight"), ln_b=a.view(P, p + "norm.bias"),
                w1=a.view(PB, p + "pwconv1.weight"), b1=a.view(P, p + "pwconv1.bias"),
                gamma=a.view(P, p + "grn.gamma", (2 * C,)), beta=a.view(P, p + "grn.beta", (2 * C,)),
                w2=a.view(PB, p + "pwconv2.weight"), b2=a.view(P, p + "pwconv2.bias"), key=p))
        self.kx, self.kct = _rup(M, 64), _rup(M + C, 64)
        self.in_b = a.view(P, "input_embed.proj.bias")
        self.blocks = []
        for i in range(self.depth):
            p = f"transformer_blocks.{i}."
            self.blocks.append(dict(
                wqkv=a.view(PB, p + "attn.to_q.weight", (3 * D, D), 3 * D * D), bqkv=a.view(P, p + "attn.to_q.bias", (3 * D,), 3 * D),
                wo=a.view(PB, p + "attn.to_out.0.weight"), bo=a.view(P, p + "attn.to_out.0.bias"),
                w1=a.view(PB, p + "ff.ff.0.weight"), b1=a.view(P, p + "ff.ff.0.bias"),
                w2=a.view(PB, p + "ff.ff.3.weight"), b2=a.view(P, p + "ff.ff.3.bias"), key=p))
        self.ff_dim = self.blocks[0]["w1"].shape[0]
        self.wp, self.bp = a.view(PB, "proj_out.weight"), a.view(P, "proj_out.bias")
        self._pos_table, self._rope = None, {}
        dev = self.device
        H = self.ff_dim
        z = lambda *s: torch.zeros(*s, device=dev, dtype=BF16)  # noqa: E731
        self.wx, self.wct = z(D, self.kx), z(D, self.kct)
        self.wpT = z(D, _rup(M, 64))
        self.conv_pos = [None, None]
        self.conv_posT = [None, None]
        self.refresh()

    @torch.no_grad()
    def refresh(self) -> None:
        """Rebuild the re-laid-out copies from the arenas (after an optimizer step)."""
        a, D, M, C = self.arena, self.dim, self.n_mels, self.text_dim
        W = a.view(a.p, "input_embed.proj.weight")
        self.wx[:, :M].copy_(W[:, :M])
        self.wct[:, :M + C].copy_(W[:, M:])
        self.wpT[:, :M].copy_(a.view(a.p, "proj_out.weight").t())
        for j, idx in enumerate((0, 2)):
            w = a.view(a.p, f"input_embed.conv_pos_embed.conv1d.{idx}.weight")
            b = a.view(a.p, f"input_embed.conv_pos_embed.conv1d.{idx}.bias")
            cg, ks = w.shape[1], w.shape[2]
            fwd = pack_conv_pos(w, D)
            # data gradient = the same grouped conv with in/out swapped inside each group and the taps reversed
            wt = w.view(D // cg, cg, cg, ks).permute(0, 2, 1, 3).flip(-1).reshape(D, cg, ks)
            bwd = pack_conv_pos(wt, D)
            if self.conv_pos[j] is None:
                self.conv_pos[j] = dict(w=fwd["w"], b=b, taps=ks, gsz=fwd["gsz"], cg=cg, idx=idx)
                self.conv_posT[j] = dict(w=bwd["w"], taps=ks, gsz=bwd["gsz"])
            else:
                self.conv_pos[j]["w"].copy_(fwd["w"])
                self.conv_posT[j]["w"].copy_(bwd["w"])


# ----------------------------------------------------------------------------------------------------------
# per-shape activation store
# ----------------------------------------------------------------------------------------------------------
class TrainWorkspace:
    def __init__(self, w: TrainWeights, nb: int, tpad: int):
        dev = w.device
        D, C, M, H = w.dim, w.text_dim, w.n_mels, w.ff_dim
        R = nb * tpad
        z = lambda *s, dt=F32: torch.zeros(*s, device=dev, dtype=dt)  # noqa: E731
        self.nb, self.tpad, self.R = nb, tpad, R
        self.ids, self.drop, self.row_valid = z(R, dt=I32), z(nb, dt=torch.uint8), z(R, dt=torch.uint8)
        self.seq_lens, self.text_lens = z(nb, dt=I32), z(nb, dt=I32)
        self.attn_ws = torch.zeros(int(L.lib().oron_attention_workspace_bytes(nb, tpad, w.heads)), dtype=torch.uint8, device=dev)
        nl = w.conv_layers
        self.xt = [z(R, C) for _ in range(nl + 1)]
        self.conv = [z(R, C) for _ in range(nl)]
        self.tn = [z(R, C, dt=BF16) for _ in range(nl)]
        self.tpre = [z(R, 2 * C, dt=BF16) for _ in range(nl)]
        self.thg = [z(R, 2 * C, dt=BF16) for _ in range(nl)]
        self.gx2 = [z(nb, 2 * C) for _ in range(nl)]
        self.a_ct, self.xb = z(R, w.kct, dt=BF16), z(R, w.kx, dt=BF16)
        self.c0, self.h0, self.h0b = z(R, D), z(R, D), z(R, D, dt=BF16)
        self.z1, self.c1, self.z2, self.m2 = (z(R, D, dt=BF16) for _ in range(4))
        self.ones = torch.ones(D, device=dev, dtype=F32)
        nd = w.depth
        self.xin = [z(R, D) for _ in range(nd + 1)]  # residual stream entering block i; [nd] = after the last block
        self.xres = self.xin[nd]
        self.xmid = [z(R, D) for _ in range(nd)]
        self.nrm1 = [z(R, D, dt=BF16) for _ in range(nd)]
        self.nrm2 = [z(R, D, dt=BF16) for _ in range(nd)]
        self.qkv = [z(R, 3 * D, dt=BF16) for _ in range(nd)]
        self.ao = [z(R, D, dt=BF16) for _ in range(nd)]
        self.y1 = [z(R, D, dt=BF16) for _ in range(nd)]
        self.y2 = [z(R, D, dt=BF16) for _ in range(nd)]
        self.hpre = [z(R, H, dt=BF16) for _ in range(nd)]
        self.hid = [z(R, H, dt=BF16) for _ in range(nd)]
        self.nrmf, self.v = z(R, D, dt=BF16), z(R, M)
        self.tvals, self.tfeat = z(nb), z(nb, 256, dt=BF16)
        self.pre0, self.pre2 = z(nb, D), z(nb, D)
        self.th, self.ts, self.ts32 = z(nb, D, dt=BF16), z(nb, D, dt=BF16), z(nb, D)
        self.table = z(nb, w.ada_n)
        # loss
        self.flow, self.span, self.count = z(R, M), z(R, dt=torch.uint8), z(1, dt=I32)
        self.loss_sum = z(1)
        self.dpred = z(R, _rup(M, 64), dt=BF16)
        # backward
        self.dx = z(R, D)
        self.g_d, self.g_ao = z(R, D, dt=BF16), z(R, D, dt=BF16)
        self.g_h, self.g_qkv = z(R, H, dt=BF16), z(R, 3 * D, dt=BF16)
        self.vb = z(R, D, dt=BF16)
        self.lse = [z(nb * w.heads * tpad) for _ in range(nd)]
        self.delta = z(nb * w.heads * tpad)
        self.dtab = z(nb, w.ada_n)
        self.dts, self.dth, self.dpre2, self.dpre0 = z(nb, D), z(nb, D), z(nb, D), z(nb, D)
        self.dxt, self.dconv = z(R, C), z(R, C)
        self.g_c, self.g_h2 = z(R, C, dt=BF16), z(R, 2 * C, dt=BF16)
        self.grn_A, self.grn_nx, self.grn_coef = z(nb, 2 * C), z(nb, 2 * C), z(nb, 2 * C)


# ----------------------------------------------------------------------------------------------------------
class TrainEngine:
    """forward + backward of the OT-CFM objective and the optimizer step for one F5TTS model on one GPU."""

    def __init__(self, model: torch.nn.Module, *, lr: float = 1e-4, betas: tuple[float, float] = (0.9, 0.999), eps: float = 1e-8,
                 weight_decay: float = 0.01, max_grad_norm: float = 1.0):
        p0 = next(model.parameters())
        if not p0.is_cuda:
            raise RuntimeError("TrainEngine runs only on a CUDA device (oron_tts_b200 has no CPU fallback)")
        self.model = model
        self.cfm = model.cfm if hasattr(model, "cfm") else model
        self.arena = ParamArena(model)
        self.w = TrainWeights(self.arena, self.cfm.backbone.rotary_embed.inv_freq)
        self._seen_version = self._versions()
        self.lr, self.betas, self.eps, self.wd, self.max_norm = lr, betas, eps, weight_decay, max_grad_norm
        self.step_count = 0
        self._ws: dict = {}
        self.sumsq = torch.zeros(1, device=p0.device, dtype=F32)
        self.skipped = torch.zeros(1, device=p0.device, dtype=I32)
        self.reducer = GradReducer(self.arena.g, self.arena.block_ranges)
        # data-parallel runs: sharded optimizer (ZeRO-1) unless ORON_ZERO1=0 (then: all-reduce + replicated AdamW)
        self.sharded = None
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1 and os.environ.get("ORON_ZERO1", "1") != "0":
            self.sharded = ShardedOptimizer(self.arena)
        drops = [m.p for m in self.cfm.backbone.modules() if isinstance(m, torch.nn.Dropout)]
        self.dropout_p = float(drops[0]) if drops else 0.0  # Attention.to_out[1] / FeedForward.ff[2] share p_dropout

    def workspace(self, nb: int, tpad: int) -> TrainWorkspace:
        key = (nb, tpad)
        if key not in self._ws:
            self._ws.clear()
            self._ws[key] = TrainWorkspace(self.w, nb, tpad)
        return self._ws[key]

    # ---- GEMM helpers ----------------------------------------------------------------------------------
    @staticmethod
    def _bn(n: int) -> int:
        return 256 if n % 256 == 0 else (128 if n % 128 == 0 or n > 128 else 64)

    def _dgrad(self, ws, dy, w, out):
        """out[R, K_in] (bf16) = dy[R, N_out] @ w, w = the Linear's weight as stored ([N_out, K_in], read MN-major)."""
        nn = w.shape[1]
        bn = 256 if nn % 256 == 0 else 128
        L.gemm(dy, w, out, epilogue=L.EPI_BF16, rows_per_batch=ws.tpad, nbatch=ws.nb, block_n=bn, b_mn=True, two_sm=True)

    def _wgrad(self, dy, x, out, *, accumulate=False):  # noqa: ARG002 (accumulation is implicit, see below)
        """out[N_out, K_in] (f32, a gradient-arena view) += dy[R, N_out]^T @ x[R, K_in], both operands as stored."""
        nn = x.shape[1]
        bn = 256 if nn % 256 == 0 else 128
        # stream-K: the output has few 256 x 256 tiles (16 for a dim x dim weight) but a long reduction (K = rows), so every
        # SM pair takes an equal share of the (tile, k-block) list and adds its partial sum into the gradient arena, which
        # the step zeroed (or keeps accumulating into): `accumulate` needs nothing extra
        L.gemm(dy, x, out, epilogue=L.EPI_F32, block_n=bn, a_mn=True, b_mn=True, two_sm=True, stream_k=True)

    def _gconv_wgrad(self, x, dy, cp, sl, gw, gb, common):
        """Weight + bias gradient of one ConvPositionEmbedding conv. Tensor-core kernel when the shape allows (dim % 128 == 0;
        rows beyond each length are zero in x and dy here: the forward masks x, act_bwd zeroes dy); ORON_GCONV_TC=0 keeps the
        CUDA-core kernel (0.95 ms per conv at config 5)."""
        C = x.shape[1]
        if C % 128 == 0 and 64 % cp["cg"] == 0 and common["rows_per_batch"] % 64 == 0 and os.environ.get("ORON_GCONV_TC", "1") != "0":
            T.gconv_wgrad_tc(x, dy, cg=cp["cg"], taps=cp["taps"], dw=gw, **common)
            T.colsum(dy, gb)
        else:
            T.gconv_wgrad(x, dy, cg=cp["cg"], taps=cp["taps"], seq_lens=sl, dw=gw, db=gb, **common)

    def _linear_bwd(self, ws, dy, x_saved, w, gw, gb, dx_out, *, acc=False):
        """Backward of y = x W^T + b for [R, .] activations: bias and weight gradients into the arena, data gradient.
        No transposed copies: the tcgen05 GEMM reads the row-major operands MN-major (oron_gemm_desc.a/b_mn_major)."""
        if gb is not None:  # None: the kernel that produced dy already summed its columns
            T.colsum(dy, gb)
        self._wgrad(dy, x_saved, gw, accumulate=acc)
        if dx_out is not None:
            self._dgrad(ws, dy, w, dx_out)

    # ---- objective ---------------------------------------------------------------------------------------
    def draw(self, mel: torch.Tensor, lens: torch.Tensor, training: bool = True) -> dict:
        """The random (train) or fixed (eval) choices of CFM.forward (flow.py:101-138) for one batch."""
        cfm = self.cfm
        x1 = mel.transpose(1, 2) if (mel.ndim == 3 and mel.shape[1] == cfm.n_mels) else mel
        B, Tn, dev = x1.shape[0], x1.shape[1], x1.device
        mask = torch.arange(Tn, device=dev)[None, :] < lens[:, None]
        pos = torch.arange(Tn, device=dev)
        if training:
            frac = torch.zeros(B, device=dev).float().uniform_(*cfm.frac_lengths_mask)
            lengths = (frac * lens).long()
            start = ((lens - lengths) * torch.rand_like(frac)).long().clamp(min=0)
            span = (pos[None, :] >= start[:, None]) & (pos[None, :] < (start + lengths)[:, None]) & mask
            time = torch.rand(B, dtype=x1.dtype, device=dev)
            drop_audio = random() < cfm.audio_drop_prob
            drop_text = random() < cfm.cond_drop_prob
            drop_audio = drop_audio or drop_text
            x0 = torch.randn_like(x1)
        else:
            mid = sum(cfm.frac_lengths_mask) / 2
            sp = (torch.full((B,), mid, device=dev).float() * lens).long()
            start = ((lens - sp) // 2).clamp(min=0)
            span = (pos[None, :] >= start[:, None]) & (pos[None, :] < (start + sp)[:, None]) & mask
            time = torch.full((B,), 0.5, dtype=x1.dtype, device=dev)
            drop_audio = drop_text = False
            x0 = torch.randn(x1.shape, generator=torch.Generator(device=dev).manual_seed(0), device=dev, dtype=x1.dtype)
        dseed = int(torch.randint(0, 2 ** 62, ()).item()) if (training and self.dropout_p > 0) else None
        return dict(x1=x1, x0=x0, time=time, span=span, drop_audio=drop_audio, drop_text=drop_text, dropout_seed=dseed)

    @torch.no_grad()
    def loss_and_grad(self, mel: torch.Tensor, text_ids: torch.Tensor, lens: torch.Tensor | None = None, *,
                      draws: dict | None = None, training: bool = True, accumulate: bool = False,
                      reduce: bool = True) -> torch.Tensor:
        """CFM.forward + backward: returns the scalar loss (device tensor) with every ``param.grad`` filled
        (added to, when ``accumulate``). ``draws``: output of ``draw`` to inject the batch's randomness.

        ``reduce``: issue this engine's per-block gradient all-reduces during the backward pass (data-parallel runs).
        Pass ``False`` for every micro-batch of a gradient-accumulation window except the last one (the ranges are summed
        in place, so reducing them before the window is complete would count the earlier micro-batches ``world`` times),
        and whenever something else owns the reduction (the autograd bridge under DDP, trainer.py:70-71)."""
        x1 = mel.transpose(1, 2) if (mel.ndim == 3 and mel.shape[1] == self.cfm.n_mels) else mel
        B, Tn, dev = x1.shape[0], x1.shape[1], x1.device
        if lens is None:
            lens = torch.full((B,), Tn, device=dev, dtype=torch.long)
        d = draws if draws is not None else self.draw(mel, lens, training)
        x1, x0, time, span = d["x1"].float(), d["x0"].float(), d["time"].float(), d["span"]
        tt = time[:, None, None]
        phi = (1 - tt) * x0 + tt * x1
        cond = torch.where(span[..., None], torch.zeros_like(x1), x1)
        tpad = _rup(Tn, TILE)
        ws = self.workspace(B, tpad)
        self.reducer.drain()  # nothing may still be reducing the ranges this pass zeroes / adds to
        if accumulate and self.reducer._pending:
            raise RuntimeError("loss_and_grad(accumulate=True) after a micro-batch that already reduced its gradients: pass "
                               "reduce=False for every micro-batch of the window except the last")
        if not accumulate:
            self.reducer._pending.clear()
            self.arena.g.zero_()
        self._reduce_in_backward = bool(reduce)
        dseed = d.get("dropout_seed")
        ws.drop_p = self.dropout_p if dseed is not None else 0.0
        ws.drop_seed = int(dseed) if dseed is not None else 0
        self._forward(ws, phi, cond, text_ids, time, lens, d["drop_audio"], d["drop_text"])
        # loss + d(pred)
        M = self.w.n_mels
        fl = ws.flow.view(B, tpad, M)
        fl.zero_()
        fl[:, :Tn].copy_(x1 - x0)
        sp = ws.span.view(B, tpad)
        sp.zero_()
        sp[:, :Tn].copy_(span)
        ws.count.copy_(span.sum().to(I32).reshape(1))
        ws.loss_sum.zero_()
        T.cfm_loss(ws.v, ws.flow, ws.span, ws.count, ws.loss_sum, ws.dpred, n_mels=M)
        self._backward(ws, acc=accumulate)
        self._check_ids_end()
        return (ws.loss_sum / (ws.count.clamp(min=1).float() * M)).reshape(())

    # ---- forward (DiT.forward dit.py:165-234 with every intermediate kept) -----------------------------------
    def _forward(self, ws: TrainWorkspace, x: torch.Tensor, cond: torch.Tensor, text: torch.Tensor, time: torch.Tensor,
                 lens: torch.Tensor, drop_audio: bool, drop_text: bool) -> None:
        w = self.w
        nb, tpad, D, C, M, H = ws.nb, ws.tpad, w.dim, w.text_dim, w.n_mels, w.ff_dim
        Tn = x.shape[1]
        ids = (text.to(torch.int64) + 1)[:, :Tn]
        iv = ws.ids.view(nb, tpad)
        iv.zero_()
        iv[:, : ids.shape[1]].copy_(ids.clamp(0, w.text_table.shape[0] - 1))  # the kernels never index outside the table
        ws.drop.fill_(1 if drop_text else 0)
        self._check_ids_begin(ids)
        ws.seq_lens.copy_(lens)
        # the attention kernel's work plan for these lengths (equal shares of the key-tile list per CTA), once per pass
        L.attention_plan(ws.attn_ws, nbatch=nb, rows_per_batch=tpad, heads=w.heads, seq_lens=ws.seq_lens)
        ws.text_lens.fill_(Tn)  # TextEmbedding runs over the whole padded batch (encoder.py:68-96), fillers included
        common = dict(rows_per_batch=tpad, nbatch=nb)
        # -- text embedding
        L.text_embed_front(ws.ids, ws.drop, w.text_table, w.pos_table(tpad), rows_per_batch=tpad, nb=nb, x=ws.xt[0],
                           row_valid=ws.row_valid)
        for j, blk in enumerate(w.text_blocks):
            T.dwconv7(ws.xt[j], ws.conv[j], seq_lens=ws.text_lens, w=blk["dw_w"], bias=blk["dw_b"], **common)
            L.ln_modulate(ws.conv[j], eps=1e-6, scale=blk["ln_w"], shift=blk["ln_b"], add_one=False, out_bf16=ws.tn[j], **common)
            L.gemm(ws.tn[j], blk["w1"], ws.tpre[j], epilogue=L.EPI_BF16, bias=blk["b1"], block_n=128, **common)
            T.act_fwd(ws.tpre[j], ws.thg[j], L.ACT_GELU_ERF)
            L.grn(ws.thg[j], rows_per_batch=tpad, nb=nb, seq_lens=ws.text_lens, gamma=blk["gamma"], beta=blk["beta"], gx2=ws.gx2[j])
            L.gemm(ws.thg[j], blk["w2"], ws.xt[j + 1], epilogue=L.EPI_SCALE_RESID, bias=blk["b2"], addend=ws.xt[j],
                   row_valid=ws.row_valid, block_n=128 if C % 128 == 0 else 64, **common)
        xt = ws.xt[-1]
        # -- input embedding (dit.py:35-55)
        cv = torch.zeros(nb, tpad, M, device=x.device, dtype=F32)
        if not drop_audio:
            cv[:, :Tn].copy_(cond)
        L.cast_rows_bf16(cv.view(-1, M), ws.a_ct[:, :M])
        L.cast_rows_bf16(xt, ws.a_ct[:, M:M + C])
        xv = torch.zeros(nb, tpad, M, device=x.device, dtype=F32)
        xv[:, :Tn].copy_(x)
        L.cast_rows_bf16(xv.view(-1, M), ws.xb[:, :M])
        L.gemm(ws.a_ct, w.wct, ws.c0, epilogue=L.EPI_F32, bias=w.in_b, block_n=self._bn(D), **common)
        L.gemm(ws.xb, w.wx, ws.h0, epilogue=L.EPI_EMBED_DUAL, addend=ws.c0, seq_lens=ws.seq_lens, out2=ws.h0b, block_n=128,
               **common)
        c1, c2 = w.conv_pos
        conv = lambda cp: dict(taps=cp["taps"], cin_blocks=cp["gsz"] // 64, pad=cp["taps"] // 2, grouped=cp["gsz"],  # noqa: E731
                               block_n=64, **common)
        L.gemm(ws.h0b, c1["w"], ws.z1, epilogue=L.EPI_BF16, bias=c1["b"], **conv(c1))
        T.act_fwd(ws.z1, ws.c1, T.ACT_MISH, rows_per_batch=tpad, seq_lens=ws.seq_lens)
        L.gemm(ws.c1, c2["w"], ws.z2, epilogue=L.EPI_BF16, bias=c2["b"], **conv(c2))
        T.act_fwd(ws.z2, ws.m2, T.ACT_MISH, rows_per_batch=tpad, seq_lens=ws.seq_lens)
        T.gate_resid(ws.h0, ws.m2, gate=ws.ones, gate_ld=0, seq_lens=None, mask_rows=False, out=ws.xin[0], **common)
        # -- timestep conditioning (modules.py:39-62) and every AdaLN projection (modules.py:214, 232)
        ws.tvals.copy_(time)
        L.time_sinusoid(ws.tvals, ws.tfeat)
        L.gemm(ws.tfeat, w.t0_w, ws.pre0, epilogue=L.EPI_F32, bias=w.t0_b)
        T.act_fwd(ws.pre0, ws.th, L.ACT_SILU)
        L.gemm(ws.th, w.t2_w, ws.pre2, epilogue=L.EPI_F32, bias=w.t2_b)
        T.act_fwd(ws.pre2, ws.ts, L.ACT_SILU)
        T.act_fwd(ws.pre2, ws.ts32, L.ACT_SILU)
        L.gemm(ws.ts, w.ada_w, ws.table, epilogue=L.EPI_F32, bias=w.ada_b, block_n=256 if w.ada_n % 256 == 0 else 128)
        # -- transformer blocks (modules.py:326-345)
        tab = ws.table.view(-1)
        an = w.ada_n
        cos, sin = w.rope(tpad)
        bn_big = 256 if D % 256 == 0 else 128
        mod = dict(mod_ld=an, mod_nb=nb, add_one=True, eps=1e-6, **common)
        for i, blk in enumerate(w.blocks):
            o = i * 6 * D
            # the residual stream is never copied: each gated residual writes the next saved buffer
            L.ln_modulate(ws.xin[i], scale=tab[o + D:], shift=tab[o:], out_bf16=ws.nrm1[i], **mod)
            L.gemm(ws.nrm1[i], blk["wqkv"], ws.qkv[i], epilogue=L.EPI_QKV_ROPE, bias=blk["bqkv"], rope_cos=cos, rope_sin=sin,
                   rope_cols=2 * D, f16_from_col=2 * D, block_n=bn_big, two_sm=True, **common)
            T.attention_fwd_lse(ws.qkv[i], ws.ao[i], ws.lse[i], nbatch=nb, rows_per_batch=tpad, heads=w.heads,
                                seq_lens=ws.seq_lens, scale=1.0 / math.sqrt(w.dim_head),
                                workspace=ws.attn_ws if os.environ.get("ORON_TRAIN_ATT_PLAN", "1") != "0" else None)
            # out-projection with the gated residual in the epilogue: xmid = xin + gate_msa * dropout(mask(y1)), y1 kept
            L.gemm(ws.ao[i], blk["wo"], ws.xmid[i], epilogue=L.EPI_GATE_RESID_DUAL, bias=blk["bo"], out2=ws.y1[i], addend=ws.xin[i],
                   gate=tab[o + 2 * D:], gate_ld=an, gate_nb=nb, seq_lens=ws.seq_lens, mask_rows=True, block_n=bn_big, two_sm=True,
                   dropout_p=ws.drop_p, dropout_seed=ws.drop_seed + 4 * i, **common)
            L.ln_modulate(ws.xmid[i], scale=tab[o + 4 * D:], shift=tab[o + 3 * D:], out_bf16=ws.nrm2[i], **mod)
            # FFN up-projection: the epilogue keeps the pre-activation (for the backward) and writes dropout(gelu(pre))
            L.gemm(ws.nrm2[i], blk["w1"], ws.hid[i], epilogue=L.EPI_GELU_DROP_DUAL, bias=blk["b1"], out2=ws.hpre[i],
                   block_n=256 if H % 256 == 0 else 128, two_sm=True, dropout_p=ws.drop_p, dropout_seed=ws.drop_seed + 4 * i + 1,
                   **common)
            L.gemm(ws.hid[i], blk["w2"], ws.xin[i + 1], epilogue=L.EPI_GATE_RESID_DUAL, bias=blk["b2"], out2=ws.y2[i],
                   addend=ws.xmid[i], gate=tab[o + 5 * D:], gate_ld=an, gate_nb=nb, mask_rows=False, block_n=bn_big, two_sm=True,
                   **common)
        o = w.depth * 6 * D
        L.ln_modulate(ws.xres, scale=tab[o:], shift=tab[o + D:], out_bf16=ws.nrmf, **mod)
        L.gemm(ws.nrmf, w.wp, ws.v, epilogue=L.EPI_F32, bias=w.bp, block_n=128, **common)

    # ---- backward ----------------------------------------------------------------------------------------------
    def _backward(self, ws: TrainWorkspace, acc: bool) -> None:
        w, a = self.w, self.arena
        G = a.g
        nb, tpad, D, C, M, H = ws.nb, ws.tpad, w.dim, w.text_dim, w.n_mels, w.ff_dim
        an = w.ada_n
        common = dict(rows_per_batch=tpad, nbatch=nb)
        tab, dtab = ws.table.view(-1), ws.dtab.view(-1)
        ws.dtab.zero_()
        sl = ws.seq_lens
        # -- proj_out (dit.py:234) and the final AdaLN (modules.py:232-234: chunks (scale, shift))
        T.colsum(ws.dpred[:, :M], a.view(G, "proj_out.bias"))
        self._wgrad(ws.dpred[:, :M], ws.nrmf, a.view(G, "proj_out.weight"), accumulate=acc)
        # K = n_mels is not a multiple of 64: this one data gradient keeps a (tiny) transposed, zero-padded weight copy
        L.gemm(ws.dpred, w.wpT, ws.g_d, epilogue=L.EPI_BF16, block_n=256 if D % 256 == 0 else 128, two_sm=True, **common)
        o = w.depth * 6 * D
        lnb = dict(eps=1e-6, mod_ld=an, add_one=True, seq_lens=sl, dx=ws.dx, dmod_ld=an, **common)
        T.ln_bwd(ws.xres, ws.g_d, scale=tab[o:], accumulate=False, dscale=dtab[o:], dshift=dtab[o + D:], **lnb)
        gada_w = a.view(G, "transformer_blocks.0.attn_norm.linear.weight", (an, D), an * D)
        gada_b = a.view(G, "transformer_blocks.0.attn_norm.linear.bias", (an,), an)
        T.skinny_wgrad(ws.dtab[:, o:], ws.ts32, gada_w[o:], gada_b[o:], accumulate=acc)  # norm_out.linear (modules.py:232)
        cos, sin = w.rope(tpad)
        for i in reversed(range(w.depth)):
            blk = w.blocks[i]
            p = blk["key"]
            o = i * 6 * D
            # FFN branch: x += gate_mlp * (W2 gelu(W1 n + b1) + b2)
            T.gate_bwd(ws.dx, ws.y2[i], gate=tab[o + 5 * D:], gate_ld=an, seq_lens=sl, dy=ws.g_d, dgate=dtab[o + 5 * D:],
                       dgate_ld=an, dbias=a.view(G, p + "ff.ff.3.bias"), **common)
            self._wgrad(ws.g_d, ws.hid[i], a.view(G, p + "ff.ff.3.weight"), accumulate=acc)
            # data gradient with the GELU derivative and the dropout mask applied in the epilogue (reads the saved pre-activation)
            L.gemm(ws.g_d, blk["w2"], ws.g_h, epilogue=L.EPI_GELU_DROP_BWD, out2=ws.hpre[i], b_mn=True, two_sm=True,
                   block_n=256 if H % 256 == 0 else 128, dropout_p=ws.drop_p, dropout_seed=ws.drop_seed + 4 * i + 1, **common)
            self._linear_bwd(ws, ws.g_h, ws.nrm2[i], blk["w1"], a.view(G, p + "ff.ff.0.weight"), a.view(G, p + "ff.ff.0.bias"),
                             ws.g_d, acc=acc)
            T.ln_bwd(ws.xmid[i], ws.g_d, scale=tab[o + 4 * D:], accumulate=True, dscale=dtab[o + 4 * D:], dshift=dtab[o + 3 * D:],
                     **lnb)
            # attention branch: x += gate_msa * mask(Wo attn(...) + bo)
            T.gate_bwd(ws.dx, ws.y1[i], gate=tab[o + 2 * D:], gate_ld=an, seq_lens=sl, dy=ws.g_d, dgate=dtab[o + 2 * D:],
                       dgate_ld=an, dbias=a.view(G, p + "attn.to_out.0.bias"), dropout_p=ws.drop_p,
                       dropout_seed=ws.drop_seed + 4 * i, **common)
            self._linear_bwd(ws, ws.g_d, ws.ao[i], blk["wo"], a.view(G, p + "attn.to_out.0.weight"), None, ws.g_ao, acc=acc)
            T.f16_to_bf16(ws.qkv[i][:, 2 * D:], ws.vb)
            T.attention_bwd(ws.qkv[i][:, : 2 * D], ws.vb, ws.ao[i], ws.g_ao, ws.g_qkv, nbatch=nb, rows_per_batch=tpad,
                            heads=w.heads, seq_lens=sl, scale=1.0 / math.sqrt(w.dim_head), rope_cos=cos, rope_sin=sin,
                            lse=ws.lse[i], delta=ws.delta, have_lse=True)
            self._linear_bwd(ws, ws.g_qkv, ws.nrm1[i], blk["wqkv"], a.view(G, p + "attn.to_q.weight", (3 * D, D), 3 * D * D),
                             a.view(G, p + "attn.to_q.bias", (3 * D,), 3 * D), ws.g_d, acc=acc)
            T.ln_bwd(ws.xin[i], ws.g_d, scale=tab[o + D:], accumulate=True, dscale=dtab[o + D:], dshift=dtab[o:], **lnb)
            # this block's six modulation gradients are final: its AdaLN projection (modules.py:214) gets its gradient now,
            # so that it travels with the block's all-reduce bucket
            T.skinny_wgrad(ws.dtab[:, o:o + 6 * D], ws.ts32, gada_w[o:o + 6 * D], gada_b[o:o + 6 * D], accumulate=acc)
            self._block_done(i)
        # -- ConvPositionEmbedding (modules.py:131-141) + residual (dit.py:54): xres = h0 + mask * mish(z2)
        c1, c2 = w.conv_pos
        t1, t2 = w.conv_posT
        conv = lambda cp: dict(taps=cp["taps"], cin_blocks=cp["gsz"] // 64, pad=cp["taps"] // 2, grouped=cp["gsz"],  # noqa: E731
                               block_n=64, **common)
        gkey = "input_embed.conv_pos_embed.conv1d."
        T.act_bwd(ws.dx, ws.z2, ws.g_d, T.ACT_MISH, rows_per_batch=tpad, seq_lens=sl)
        self._gconv_wgrad(ws.c1, ws.g_d, c2, sl, a.view(G, gkey + "2.weight"), a.view(G, gkey + "2.bias"), common)
        L.gemm(ws.g_d, t2["w"], ws.g_ao, epilogue=L.EPI_BF16, **conv(t2))
        T.act_bwd(ws.g_ao, ws.z1, ws.g_ao, T.ACT_MISH, rows_per_batch=tpad, seq_lens=sl)
        self._gconv_wgrad(ws.h0b, ws.g_ao, c1, sl, a.view(G, gkey + "0.weight"), a.view(G, gkey + "0.bias"), common)
        L.gemm(ws.g_ao, t1["w"], ws.dx, epilogue=L.EPI_SCALE_RESID, addend=ws.dx, seq_lens=sl, out2=ws.g_d, **conv(t1))
        # -- InputEmbedding.proj (dit.py:53): h0 = [x | cond | text] W^T + b
        Gw = a.view(G, "input_embed.proj.weight")
        T.colsum(ws.g_d, a.view(G, "input_embed.proj.bias"))
        self._wgrad(ws.g_d, ws.xb[:, :M], Gw[:, :M], accumulate=acc)
        self._wgrad(ws.g_d, ws.a_ct[:, : M + C], Gw[:, M:], accumulate=acc)
        # d text_embed = dh0 @ W[:, 2M:] (the text columns of the projection, read in place from the bf16 arena)
        L.gemm(ws.g_d, a.view(a.pb, "input_embed.proj.weight")[:, 2 * M:], ws.dxt, epilogue=L.EPI_F32,
               block_n=256 if C % 256 == 0 else 128, b_mn=True, two_sm=True, **common)
        # -- TextEmbedding (encoder.py:68-96)
        tl = ws.text_lens
        for j in reversed(range(w.conv_layers)):
            blk = w.text_blocks[j]
            p = blk["key"]
            T.mask_rows(ws.dxt, ws.row_valid)
            L.cast_rows_bf16(ws.dxt, ws.g_c)
            self._linear_bwd(ws, ws.g_c, ws.thg[j], blk["w2"], a.view(G, p + "pwconv2.weight"), a.view(G, p + "pwconv2.bias"),
                             ws.g_h2, acc=acc)
            T.grn_bwd(ws.g_h2, ws.tpre[j], ws.g_h2, rows_per_batch=tpad, nb=nb, seq_lens=tl, gamma=blk["gamma"], gx2=ws.gx2[j],
                      A=ws.grn_A, nx=ws.grn_nx, coef=ws.grn_coef, dgamma=a.view(G, p + "grn.gamma", (2 * C,)),
                      dbeta=a.view(G, p + "grn.beta", (2 * C,)))
            self._linear_bwd(ws, ws.g_h2, ws.tn[j], blk["w1"], a.view(G, p + "pwconv1.weight"), a.view(G, p + "pwconv1.bias"),
                             ws.g_c, acc=acc)
            T.ln_bwd(ws.conv[j], ws.g_c, eps=1e-6, scale=blk["ln_w"], mod_ld=0, add_one=False, seq_lens=tl, dx=ws.dconv,
                     accumulate=False, dscale=a.view(G, p + "norm.weight"), dshift=a.view(G, p + "norm.bias"), dmod_ld=0, **common)
            T.dwconv7_wgrad(ws.xt[j], ws.dconv, seq_lens=tl, dw=a.view(G, p + "dwconv.weight", (C, 7)),
                            db=a.view(G, p + "dwconv.bias"), **common)
            T.dwconv7(ws.dconv, ws.dxt, seq_lens=tl, w=blk["dw_w"], bias=None, flip=True, accumulate=True, **common)
        T.mask_rows(ws.dxt, ws.row_valid)
        T.text_embed_bwd(ws.ids, ws.drop, ws.dxt, a.view(G, "text_embed.text_embed.weight"), rows_per_batch=tpad, nb=nb)
        # -- AdaLN projections (stacked) and the timestep MLP
        ws.dts.zero_()
        T.skinny_dgrad(ws.dtab, w.ada_w, ws.dts)
        T.act_bwd(ws.dts, ws.pre2, ws.dpre2, L.ACT_SILU)
        T.skinny_wgrad(ws.dpre2, ws.th.float(), a.view(G, "time_embed.time_mlp.2.weight"), a.view(G, "time_embed.time_mlp.2.bias"),
                       accumulate=acc)
        ws.dth.zero_()
        T.skinny_dgrad(ws.dpre2, w.t2_w, ws.dth)
        T.act_bwd(ws.dth, ws.pre0, ws.dpre0, L.ACT_SILU)
        T.skinny_wgrad(ws.dpre0, ws.tfeat.float(), a.view(G, "time_embed.time_mlp.0.weight"), a.view(G, "time_embed.time_mlp.0.bias"),
                       accumulate=acc)

    # ---- data-parallel gradient mean + optimizer ---------------------------------------------------------------
    def _block_done(self, i: int) -> None:
        if getattr(self, "_reduce_in_backward", True) and getattr(self, "sharded", None) is None:
            self.reducer.block_done(i)  # sharded optimizer: one reduce-scatter of the whole arena in optimizer_step instead

    # token ids outside [-1, vocab) (nn.Embedding raises IndexError for them, encoder.py:68-75): the range test runs on the
    # device at the start of the pass, the kernels see clamped ids, and the verdict travels to pinned host memory on a SIDE
    # stream right away (a copy on the compute stream would be ordered behind the whole pass and make the host wait for
    # the backward: no run-ahead into the next step). It is read once the pass has been enqueued, by which time the copy
    # has long finished. The same read tells whether the PREVIOUS optimizer step was skipped on the device (non-finite
    # gradients), in which case its Adam step number is given back.
    def _check_ids_begin(self, ids_shifted: torch.Tensor):
        vocab1 = self.w.text_table.shape[0]
        self._id_flags = torch.stack([((ids_shifted < 0) | (ids_shifted >= vocab1)).any().to(I32), self.skipped.reshape(())])
        if self._id_flags.is_cuda:
            if getattr(self, "_flag_stream", None) is None:
                self._flag_stream = torch.cuda.Stream(device=self._id_flags.device)
                self._flag_host = torch.zeros(2, dtype=I32).pin_memory()
                self._flag_event = torch.cuda.Event()
            main = torch.cuda.current_stream(self._id_flags.device)
            self._flag_stream.wait_stream(main)
            with torch.cuda.stream(self._flag_stream):
                self._flag_host.copy_(self._id_flags, non_blocking=True)
                self._flag_event.record(self._flag_stream)
            self._id_flags.record_stream(self._flag_stream)
        return self._id_flags

    def _check_ids_end(self) -> None:
        if self._id_flags.is_cuda:
            self._flag_event.synchronize()
            bad, skipped_prev = (int(v) for v in self._flag_host.tolist())
        else:
            bad, skipped_prev = (int(v) for v in self._id_flags.tolist())
        if skipped_prev and getattr(self, "_step_pending_skip_check", False):
            self.step_count -= 1
        self._step_pending_skip_check = False
        if bad:
            raise IndexError(f"text_ids out of range: ids must lie in [-1, {self.w.text_table.shape[0] - 2}] "
                             "(index out of range in the text embedding table)")

    def reduce_gradients(self) -> None:
        self.reducer.finish()

    @torch.no_grad()
    def optimizer_step(self, lr: float | None = None, accum_steps: int = 1) -> None:
        """clip_grad_norm_(max_grad_norm) + AdamW + bf16 operand refresh (trainer.py:191-216), gradients averaged over
        the ranks -- and over the ``accum_steps`` micro-batches accumulated into them (trainer.py:236 divides each
        micro-batch loss by grad_accum) -- first. Non-finite gradient norm: the update is skipped on the device
        (``self.skipped`` = 1) and the Adam step number is given back at the next pass."""
        import torch.distributed as dist

        world = dist.get_world_size() if (dist.is_available() and dist.is_initialized()) else 1
        a = self.arena
        self.step_count += 1
        self._step_pending_skip_check = True
        self.sumsq.zero_()
        self.skipped.zero_()
        kw = dict(grad_scale=1.0 / (world * max(int(accum_steps), 1)), max_norm=self.max_norm,
                  lr=self.lr if lr is None else lr, beta1=self.betas[0], beta2=self.betas[1], eps=self.eps, wd=self.wd,
                  step=self.step_count, skipped=self.skipped)
        if self.sharded is not None:
            so = self.sharded
            so.reduce_gradients()                       # this rank's shard of g now holds the sum over the ranks
            lo, hi = so.lo, so.hi
            T.sumsq(a.g[lo:hi], self.sumsq)
            dist.all_reduce(self.sumsq, op=dist.ReduceOp.SUM)  # global norm: every rank clips (or skips) alike
            T.adamw_clip(a.p[lo:hi], a.g[lo:hi], a.m[lo:hi], a.v[lo:hi], a.pb[lo:hi], self.sumsq, **kw)
            so.exchange_after_step()
        else:
            self.reduce_gradients()
            T.sumsq(a.g, self.sumsq)
            T.adamw_clip(a.p, a.g, a.m, a.v, a.pb, self.sumsq, **kw)
        self.w.refresh()
        self._seen_version = self._versions()

    @torch.no_grad()
    def sync_params(self) -> None:
        """Master weights changed outside ``optimizer_step`` (a torch optimizer stepping the parameter views,
        ``load_state_dict``): refresh the bf16 operands and the re-laid-out copies."""
        v = self._versions()
        if v != self._seen_version:
            self.arena.pb.copy_(self.arena.p)
            self.w.refresh()
            self._seen_version = v

    def _versions(self) -> int:
        # ``param.data = view`` gives every parameter its own version counter: in-place updates by a torch optimizer
        # bump these, not the arena's
        return sum(prm._version for prm in self.arena.named.values())

    def detach_grads(self) -> None:
        """Autograd-bridge mode: ``param.grad`` must not alias the gradient arena (autograd accumulates into it)."""
        lo = self.arena.g.data_ptr()
        hi = lo + self.arena.g.numel() * 4
        for prm in self.arena.named.values():
            if prm.grad is not None and lo <= prm.grad.data_ptr() < hi:
                prm.grad = None

    # ---- checkpoint / resume and EMA (trainer.py:98-102, 214-215, 507-526) -----------------------------------------
    def optimizer_state_dict(self) -> dict:
        """The Adam moments in ``torch.optim.AdamW.state_dict()`` format (parameter order = ``model.parameters()``), so
        that ``CheckpointManager.save`` / a later ``torch.optim.AdamW.load_state_dict`` work unchanged."""
        self.consolidate()
        a = self.arena
        state, order = {}, []
        for idx, (name, prm) in enumerate(self.model.named_parameters()):
            o, n = a.offsets[name], prm.numel()
            state[idx] = {"step": torch.tensor(float(self.step_count)), "exp_avg": a.m[o:o + n].view(prm.shape).clone(),
                          "exp_avg_sq": a.v[o:o + n].view(prm.shape).clone()}
            order.append(idx)
        group = dict(lr=self.lr, betas=tuple(self.betas), eps=self.eps, weight_decay=self.wd, amsgrad=False, maximize=False,
                     foreach=None, capturable=False, differentiable=False, fused=None, decoupled_weight_decay=True, params=order)
        return {"state": state, "param_groups": [group]}

    @torch.no_grad()
    def load_optimizer_state_dict(self, sd: dict) -> None:
        a = self.arena
        names = [n for n, _ in self.model.named_parameters()]
        for idx, st in sd["state"].items():
            name = names[int(idx)]
            o, n = a.offsets[name], a.named[name].numel()
            a.m[o:o + n].copy_(st["exp_avg"].reshape(-1))
            a.v[o:o + n].copy_(st["exp_avg_sq"].reshape(-1))
            self.step_count = int(float(st["step"]))
        g = sd["param_groups"][0]
        self.lr, self.betas, self.eps, self.wd = g["lr"], tuple(g["betas"]), g["eps"], g["weight_decay"]

    @torch.no_grad()
    def consolidate(self) -> None:
        """Sharded optimizer: complete fp32 master weights / Adam moments on every rank (a collective: call it on all ranks
        before ``state_dict()`` / ``optimizer_state_dict()`` / an EMA update). No-op otherwise."""
        if getattr(self, "sharded", None) is not None:
            self.sharded.consolidate()

    def enable_ema(self, decay: float = 0.9999) -> None:
        """Exponential moving average of the parameters with torch_ema's update rule (trainer.py:98-102)."""
        self.ema, self.ema_decay, self.ema_updates = self.arena.p.clone(), decay, 0

    @torch.no_grad()
    def ema_update(self) -> None:
        self.ema_updates += 1
        d = min(self.ema_decay, (1 + self.ema_updates) / (10 + self.ema_updates))
        self.consolidate()
        self.ema.lerp_(self.arena.p, 1.0 - d)  # shadow -= (1 - d) * (shadow - param), one pass over the flat arena

    def ema_state_dict(self) -> dict:
        """``model.state_dict()`` with the EMA weights in place of the parameters: the ``ema_state_dict`` entry that
        scripts/infer.py prefers (infer.py:20-28, trainer.py:511-514)."""
        a = self.arena
        out = {k: v.detach().clone() for k, v in self.model.state_dict().items()}
        for name, prm in self.model.named_parameters():
            o, n = a.offsets[name], prm.numel()
            out[name] = self.ema[o:o + n].view(prm.shape).clone()
        return out

    def grad_norm(self) -> torch.Tensor:
        """Global L2 norm of the current gradients (trainer.py:171-177), device scalar."""
        s = torch.zeros(1, device=self.arena.g.device, dtype=F32)
        T.sumsq(self.arena.g, s)
        return s.sqrt().reshape(())

    def train_step(self, mel: torch.Tensor, text_ids: torch.Tensor, lens: torch.Tensor | None = None, *, lr: float | None = None,
                   draws: dict | None = None, training: bool = True) -> torch.Tensor:
        loss = self.loss_and_grad(mel, text_ids, lens, draws=draws, training=training)
        self.optimizer_step(lr)
        return loss


def reference_lr(update: int, *, base_lr: float = 1e-4, warmup_steps: int = 1000, total_steps: int = 100000,
                 start_factor: float = 1e-4, eta_min: float = 1e-6) -> float:
    """Learning rate in effect for optimizer update number ``update`` (0-based) under the reference's schedule
    (trainer.py:88-96): SequentialLR([LinearLR(start_factor -> 1 over warmup_steps), CosineAnnealingLR(T_max =
    max(total_steps - warmup_steps, 1), eta_min)], milestones=[warmup_steps]), stepped once per update."""
    if update < warmup_steps:
        return base_lr * (start_factor + (1.0 - start_factor) * update / max(warmup_steps, 1))
    t_max = max(total_steps - warmup_steps, 1)
    e = update - warmup_steps
    return eta_min + (base_lr - eta_min) * (1.0 + math.cos(math.pi * e / t_max)) / 2.0


class OTCFMLoss(torch.autograd.Function):
    """``loss = CFM.forward(...)`` as an autograd node, for hosts that drive training the reference's way
    (``loss.backward(); optimizer.step()``, trainer.py:236-262, DDP / torch.optim included): forward runs the whole
    forward + backward on the kernels, backward hands the parameter gradients (times the incoming scalar) to autograd.
    ``TrainEngine.train_step`` is the fast path (no per-tensor gradient copies, fused optimizer)."""

    @staticmethod
    def forward(ctx, eng: TrainEngine, mel, text_ids, lens, *params):
        eng.sync_params()
        eng.detach_grads()
        loss = eng.loss_and_grad(mel, text_ids, lens, training=True, reduce=False)  # DDP / the caller owns the reduction
        ctx.eng = eng
        ctx.keys = [k for k in eng.arena.order]
        return loss.clone()

    @staticmethod
    def backward(ctx, gout):
        a = ctx.eng.arena
        grads = []
        for k in a.order:
            prm = a.named[k]
            o, n = a.offsets[k], prm.numel()
            grads.append(a.g[o:o + n].view(prm.shape) * gout)
        return (None, None, None, None, *grads)
