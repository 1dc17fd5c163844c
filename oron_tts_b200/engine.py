"""Host-side driver of the DiT forward / Euler-ODE step on the sm_100a kernels.

Data layout in HBM: every activation of the DiT is a frame-major matrix
[nb' * Tpad, C] where nb' = B (no CFG) or 2B ([conditional ; unconditional] halves, dit.py:200-215)
and Tpad = ceil(max_duration / 128) * 128, so that 128-row GEMM / attention tiles never straddle two
sequences. Frames t >= duration[b] are padding: zeroed where the reference masks them, ignored
otherwise. The fp32 residual stream, the bf16 GEMM operands and the per-step modulation table all
live in one workspace that is reused by every ODE step, which is what makes the whole step
CUDA-graph replayable (the only per-step state is a device-side step counter).

Everything that is arithmetic runs in liboron_b200.so; torch is used for allocation, the seeded
noise draw (flow.py:270-283 semantics depend on torch.randn) and tiny integer/index preparation.
"""

from __future__ import annotations

import math
import os
from dataclasses import dataclass

import torch

from . import _lib as L

# The FFN-down GEMM (N = dim, K = 4 dim) has 44 tiles for 74 SM pairs at config 2; stream-K gives every pair an equal
# share of k-blocks instead (partial sums land in the fp32 residual stream with vector reductions, so the last bit
# of a run depends on arrival order): 39.8 -> 33.8 us per call. The out-projection (K = dim) is too short to gain
# (21.5 -> 23.6 us). ORON_STREAM_K=0 restores whole tiles.
STREAM_K = os.environ.get("ORON_STREAM_K", "1") != "0"
# FeedForward up + down projection as one persistent launch (csrc/ffn_tcgen05.cuh, oron_ffn_bf16): both phases share one
# equal split of k-block units over the SM pairs and the down-projection starts per H tile. Measured at config 2
# (tools/kernel_bench.py ffn, sustained): 50.9-62 us per block against 55.0-55.6 us for the two launches, run-to-run
# unstable (the main loop is bound by the 64 B/clk per-SM TMA ingest and by the power cap, not by the
# launch boundary), so it stays opt-in: ORON_FFN_FUSED=1.
FFN_FUSED = os.environ.get("ORON_FFN_FUSED", "0") == "1"
SMALL_ROWS = int(os.environ.get("ORON_SMALL_ROWS", "600"))
# Whole-tile N = dim GEMMs with a residual epilogue (attention out-projection; FFN down-projection without stream-K) may use
# 192-wide tiles of the 2-SM kernel: at config 2 the out-projection is 44 tiles of 256 columns on 74 SM pairs -- one wave either
# way, but a 192-wide k-block is 28 KB of TMA ingest per SM instead of 32 KB and the exposed epilogue is a quarter shorter:
# 13.8 -> 12.6 us per launch (tools/kernel_bench.py outproj), bit-identical results. ORON_BN192=0 keeps 256 everywhere.
BN192 = os.environ.get("ORON_BN192", "1") != "0"
# 224-wide tiles for the FeedForward up-projection where the same cost model prefers them (up_tile_width): opt-in, measured no
# gain at config 2 (same-box A/B 2.878 vs 2.889 ms per NFE: the 19th column tile carries 64 useful columns at the full k-loop)
BN224 = os.environ.get("ORON_BN224", "0") == "1"
# Opt-in (ORON_LN_TAIL=1), a measured NEGATIVE result: the LayerNorm + modulation that follows each gated-residual GEMM (a
# block's second norm after the out-projection; the next block's first norm / AdaLayerNormFinal after the down-projection) run
# as the TAIL of that GEMM's launch (oron_gemm_ln_bf16: rows are normalised by all SMs as soon as the tiles of their 256-row
# block have landed; 44 of the 45 ln_modulate launches of a Base NFE disappear; bit-identical without stream-K). Same-box A/B at
# config 2: 2.905 -> 3.01 ms per NFE. The separate ln_modulate launch needs no shared memory, so it is resident next to the
# GEMM's CTAs before they finish and -- more important -- the NEXT GEMM's CTAs move in and run their prologue (barriers, TMEM,
# cluster sync) while it executes; a tail keeps the SM's shared memory until the norm is done.
LN_TAIL = os.environ.get("ORON_LN_TAIL", "0") == "1"


def tile_width(rows_per_batch: int, nbatch: int, n: int, bn_big: int, pairs: int, narrow: int) -> int:
    """Tile width (256 or `narrow`) of a whole-tile 2-SM GEMM with N = n: waves x per-k-block TMA ingest of one SM (the bound of
    the main loop, DESIGN 5.9: 16 KB of A rows + bn / 2 weight rows of 128 bytes), the narrower tile only when strictly cheaper."""
    if not BN192 or bn_big != 256 or pairs <= 0:
        return bn_big
    tiles_mp = (((rows_per_batch + 127) // 128) * nbatch + 1) // 2
    cost = {}
    for bn in (256, narrow):
        waves = -(-(tiles_mp * -(-n // bn)) // pairs)
        cost[bn] = waves * (16384 + bn * 64)
    return narrow if cost[narrow] < cost[256] else 256


def resid_tile_width(rows_per_batch: int, nbatch: int, n: int, bn_big: int, pairs: int) -> int:
    """Gated-residual GEMMs (out-projection, whole-tile down-projection): 256 or 192 columns."""
    return tile_width(rows_per_batch, nbatch, n, bn_big, pairs, 192)


def up_tile_width(rows_per_batch: int, nbatch: int, n: int, bn_big: int, pairs: int) -> int:
    """FeedForward up-projection: 256 or 224 columns (config 2: 16 x 11 = 176 tiles of 256 are 2.38 waves on 74 SM pairs, i.e.
    three; 19 x 11 = 209 tiles of 224 are three waves as well, of k-blocks that cost 30 KB of ingest instead of 32)."""
    return tile_width(rows_per_batch, nbatch, n, bn_big, pairs, 224) if BN224 else bn_big


BF16 = torch.bfloat16
F32 = torch.float32
TILE = 128


def _rup(x: int, m: int) -> int:
    return (x + m - 1) // m * m


def pick_block_n(rows: int, n: int) -> int:
    """Tile width that minimises the number of 148-SM waves (ties -> wider tile, less smem traffic)."""
    best, best_cost = 128, None
    tiles_m = (rows + TILE - 1) // TILE
    for bn in (256, 128):
        if n % bn and bn == 256 and n < 256:
            continue
        tiles = tiles_m * ((n + bn - 1) // bn)
        waves = (tiles + 147) // 148
        cost = waves * bn  # time ~ waves * (tile work ~ bn)
        if best_cost is None or cost < best_cost:
            best, best_cost = bn, cost
    return best


def pack_conv_pos(w: torch.Tensor, D: int) -> dict:
    """Grouped Conv1d weight [D, cg, taps] (modules.py:120-124) -> tap-major implicit-GEMM operand [D, taps * gsz] (bf16),
    gsz = max(64, cg); groups narrower than 64 channels are packed block-diagonally into 64-channel blocks."""
    cg, ks = w.shape[1], w.shape[2]
    gsz = max(64, cg)
    if gsz % 64:
        raise NotImplementedError("conv_pos group width must divide or be a multiple of 64")
    device = w.device
    o = torch.arange(D, device=device)
    off = (o // cg) * cg - (o // gsz) * gsz
    w2 = torch.zeros(D, gsz, ks, device=device, dtype=F32)
    w2[o[:, None], off[:, None] + torch.arange(cg, device=device)[None, :], :] = w.to(F32)
    return dict(w=w2.permute(0, 2, 1).reshape(D, ks * gsz).to(BF16).contiguous(), gsz=gsz)


# ------------------------------------------------------------------------------------------------
# packed weights
# ------------------------------------------------------------------------------------------------
class DiTWeights:
    """bf16 / fp32 device copies of a DiT state dict in the layouts the kernels consume.

    Lives outside ``state_dict()``: rebuilt whenever the owning module's parameters
    change (load_state_dict / .to()).  ``sd`` uses keys relative to the DiT module
    (e.g. ``transformer_blocks.0.attn.to_q.weight``).
    """

    def __init__(self, sd: dict, device: torch.device):
        def bf(t):
            return t.detach().to(device=device, dtype=BF16).contiguous()

        def f32(t):
            return t.detach().to(device=device, dtype=F32).contiguous()

        self.device = device
        self.dim = D = sd["proj_out.weight"].shape[1]
        self.n_mels = n_mels = sd["proj_out.weight"].shape[0]
        self.text_dim = C = sd["text_embed.text_embed.weight"].shape[1]
        self.depth = 1 + max(int(k.split(".")[1]) for k in sd if k.startswith("transformer_blocks."))
        tb = [int(k.split(".")[2]) for k in sd if k.startswith("text_embed.text_blocks.")]
        self.conv_layers = (1 + max(tb)) if tb else 0
        self.dim_head = dh = 2 * sd["rotary_embed.inv_freq"].shape[0]
        self.heads = H = sd["transformer_blocks.0.attn.to_q.weight"].shape[0] // dh
        # The attention kernel and the RoPE epilogue work on 64-wide heads (pairs (i, i + 32)). Narrower heads (the
        # reference's own test configuration has 2 heads of 32, tests/test_checkpoint.py:9-24) are zero-padded to 64 when the
        # weights are packed: channel c of a head goes to slot c (c < dh/2) or 32 + c - dh/2, so the rotate-half partner
        # (c, c + dh/2) of modules.py:92-104 sits 32 slots away, q and k are permuted alike (q.k unchanged), the padded q / k /
        # v channels are exact zeros and the out-projection reads zero columns for them. The softmax scale stays 1/sqrt(dh).
        if dh > 64 or dh % 2:
            raise NotImplementedError("the sm_100a attention kernel handles head_dim <= 64 (even)")
        if D % 32 or C % 32:
            raise NotImplementedError("dim and text_dim must be multiples of 32")
        self.inner = H * 64  # padded attention width
        half = dh // 2
        slot = torch.cat([torch.arange(half), 32 + torch.arange(half)])  # slot of channel c inside its 64-wide head
        self._head_rows = (torch.arange(H)[:, None] * 64 + slot[None, :]).reshape(-1).to(device)  # [H * dh] -> padded row

        def pad_rows(t):  # [H * dh, ...] -> [H * 64, ...]
            if dh == 64:
                return t
            out = torch.zeros(self.inner, *t.shape[1:], device=t.device, dtype=t.dtype)
            out[self._head_rows.to(t.device)] = t
            return out

        def pad_cols(t):  # [n, H * dh] -> [n, H * 64]
            return t if dh == 64 else pad_rows(t.t()).t()

        inv = torch.zeros(32)
        inv[:half] = sd["rotary_embed.inv_freq"].detach().float().cpu()
        self.inv_freq = inv.to(device)

        # timestep MLP (modules.py:54-58)
        self.t0_w, self.t0_b = bf(sd["time_embed.time_mlp.0.weight"]), f32(sd["time_embed.time_mlp.0.bias"])
        self.t2_w, self.t2_b = bf(sd["time_embed.time_mlp.2.weight"]), f32(sd["time_embed.time_mlp.2.bias"])

        # all AdaLN projections stacked: one GEMM produces the whole modulation table
        ws = [sd[f"transformer_blocks.{i}.attn_norm.linear.weight"] for i in range(self.depth)]
        bs = [sd[f"transformer_blocks.{i}.attn_norm.linear.bias"] for i in range(self.depth)]
        ws.append(sd["norm_out.linear.weight"])
        bs.append(sd["norm_out.linear.bias"])
        self.ada_w = bf(torch.cat(ws, 0))
        self.ada_b = f32(torch.cat(bs, 0))
        self.ada_n = self.ada_w.shape[0]  # depth*6D + 2D

        # text embedding (encoder.py)
        self.text_table = f32(sd["text_embed.text_embed.weight"])
        self.text_blocks = []
        for i in range(self.conv_layers):
            p = f"text_embed.text_blocks.{i}."
            self.text_blocks.append(dict(
                dw_w=f32(sd[p + "dwconv.weight"]).view(C, 7).contiguous(), dw_b=f32(sd[p + "dwconv.bias"]),
                ln_w=f32(sd[p + "norm.weight"]), ln_b=f32(sd[p + "norm.bias"]),
                w1=bf(sd[p + "pwconv1.weight"]), b1=f32(sd[p + "pwconv1.bias"]),
                gamma=f32(sd[p + "grn.gamma"]).view(-1).contiguous(), beta=f32(sd[p + "grn.beta"]).view(-1).contiguous(),
                w2=bf(sd[p + "pwconv2.weight"]), b2=f32(sd[p + "pwconv2.bias"]),
            ))

        # input projection split by source (dit.py:53): x | cond | text
        W = sd["input_embed.proj.weight"].detach().to(device=device, dtype=F32)
        self.kx = _rup(n_mels, 64)
        wx = torch.zeros(D, self.kx, device=device, dtype=F32)
        wx[:, :n_mels] = W[:, :n_mels]
        self.wx = wx.to(BF16).contiguous()
        self.kct = _rup(n_mels + C, 64)
        wct = torch.zeros(D, self.kct, device=device, dtype=F32)
        wct[:, : n_mels + C] = W[:, n_mels:]
        self.wct = wct.to(BF16).contiguous()
        self.in_b = f32(sd["input_embed.proj.bias"])

        # ConvPositionEmbedding (modules.py:120-124): grouped k=31 convs as tap-major implicit-GEMM weights,
        # narrow groups packed block-diagonally into 64-channel blocks
        self.conv_pos = []
        for idx in (0, 2):
            w = sd[f"input_embed.conv_pos_embed.conv1d.{idx}.weight"].detach().to(device=device, dtype=F32)
            pk = pack_conv_pos(w, D)
            self.conv_pos.append(dict(w=pk["w"], b=f32(sd[f"input_embed.conv_pos_embed.conv1d.{idx}.bias"]), taps=w.shape[2],
                                      gsz=pk["gsz"]))

        # transformer blocks
        self.blocks = []
        for i in range(self.depth):
            p = f"transformer_blocks.{i}."
            self.blocks.append(dict(
                wqkv=bf(torch.cat([pad_rows(sd[p + "attn.to_q.weight"]), pad_rows(sd[p + "attn.to_k.weight"]),
                                   pad_rows(sd[p + "attn.to_v.weight"])], 0)),
                bqkv=f32(torch.cat([pad_rows(sd[p + "attn.to_q.bias"]), pad_rows(sd[p + "attn.to_k.bias"]),
                                    pad_rows(sd[p + "attn.to_v.bias"])], 0)),
                wo=bf(pad_cols(sd[p + "attn.to_out.0.weight"])), bo=f32(sd[p + "attn.to_out.0.bias"]),
                w1=bf(sd[p + "ff.ff.0.weight"]), b1=f32(sd[p + "ff.ff.0.bias"]),
                w2=bf(sd[p + "ff.ff.3.weight"]), b2=f32(sd[p + "ff.ff.3.bias"]),
            ))
        self.ff_dim = self.blocks[0]["w1"].shape[0]
        self.wp, self.bp = bf(sd["proj_out.weight"]), f32(sd["proj_out.bias"])

        # host-built constant tables, bit-identical to the reference buffers
        self._pos_table = None
        self._rope = {}

    def pos_table(self, length: int) -> torch.Tensor:
        """precompute_freqs_cis(text_dim, 8192) (modules.py:191-196), built on CPU like the reference buffer."""
        if self._pos_table is None or self._pos_table.shape[0] < length:
            n = max(8192, length)
            C = self.text_dim
            freqs = 1.0 / (10000 ** (torch.arange(0, C, 2)[: (C // 2)].float() / C))
            ang = torch.outer(torch.arange(n), freqs).float()
            self._pos_table = torch.cat([torch.cos(ang), torch.sin(ang)], dim=-1).to(self.device).contiguous()
        return self._pos_table

    def rope(self, length: int):
        """cos/sin [length, 32] of RotaryEmbedding._build_cache (modules.py:80-89), computed on device as there."""
        if length not in self._rope:
            t = torch.arange(length, device=self.device).float()
            ang = torch.outer(t, self.inv_freq)
            self._rope[length] = (ang.cos().contiguous(), ang.sin().contiguous())
        return self._rope[length]


# ------------------------------------------------------------------------------------------------
# per-shape workspace
# ------------------------------------------------------------------------------------------------
class Workspace:
    def __init__(self, w: DiTWeights, nb: int, nbp: int, tpad: int, steps: int, keep_traj: bool):
        dev = w.device
        D, C, M = w.dim, w.text_dim, w.n_mels
        R, Rb = nbp * tpad, nb * tpad

        def z(*s, dt=F32):
            # plain (non-inference) tensors: the workspace outlives the inference_mode block that creates it
            with torch.inference_mode(False):
                return torch.zeros(*s, device=dev, dtype=dt)

        self.nb, self.nbp, self.tpad, self.steps = nb, nbp, tpad, steps
        self.ids = z(R, dt=torch.int32)
        self.drop = z(nbp, dt=torch.uint8)
        self.row_valid = z(R, dt=torch.uint8)
        self.seq_lens = z(nbp, dt=torch.int32)
        self.text_lens = z(nbp, dt=torch.int32)  # frames the TextEmbedding ConvNeXt blocks / GRN run over (load_sequences)
        self.xt = z(R, C)
        self.xt_n = z(R, C, dt=BF16)
        self.xt_h = z(R, 2 * C, dt=BF16)
        self.gx2 = z(nbp, 2 * C)
        self.a_ct = z(R, w.kct, dt=BF16)
        self.cond = z(Rb, M)
        self.c0 = z(R, D)
        self.x = z(Rb, M)
        self.xb = z(R, w.kx, dt=BF16)
        self.h0 = z(R, D)
        self.h0b = z(R, D, dt=BF16)
        self.c1 = z(R, D, dt=BF16)
        self.xres = z(R, D)
        self.nrm = z(R, D, dt=BF16)
        self.qkv = z(R, 3 * w.inner, dt=BF16)
        self.ao = z(R, w.inner, dt=BF16)
        # Workspace of the balanced attention schedule (plan + partial results of the items split between two CTAs);
        # planned in `DiTEngine` once the sequence lengths of the call are known (attention_plan).
        with torch.inference_mode(False):
            self.attn_ws = torch.zeros(int(L.lib().oron_attention_workspace_bytes(nbp, tpad, w.heads)), dtype=torch.uint8, device=dev)
        self.hid = z(R, w.ff_dim, dt=BF16)
        with torch.inference_mode(False):
            self.ffn_ws = L.ffn_workspace(tpad, nbp, w.ff_dim, dev)
            self.ln_cnt = L.gemm_ln_counters(tpad, nbp, dev)  # arrival counters of the GEMM + LayerNorm launches (self-resetting)
        self.v = z(R, M)
        self.vg = z(Rb, M)
        self.step = z(1, dt=torch.int32)
        self.tvals = z(steps)
        self.dt = z(steps)
        self.tfeat = z(steps, 256, dt=BF16)
        self.th = z(steps, D, dt=BF16)
        self.ts = z(steps, D, dt=BF16)
        self.table = z(steps, w.ada_n)
        self.traj = z(steps + 1, Rb, M) if keep_traj else None
        self.graph = None
        self.graph_key = None
        self.graph_launches = 0


@dataclass
class Branch:
    drop_audio: bool
    drop_text: bool


class DiTEngine:
    """Runs DiT.forward (dit.py:165-234) and the CFG Euler loop (flow.py:244-299) on the CUDA kernels."""

    def __init__(self, weights: DiTWeights):
        self.w = weights
        self._ws: dict = {}
        self.use_graph = True
        # True: every reduction runs in a fixed order (the stream-K split of the FFN down-projection is switched off), so two
        # runs from the same noise agree bit for bit; set per call by CFM.sample(deterministic=...), part of the graph key
        self.deterministic = False
        self.replayed_launches = 0  # kernels executed through CUDA-graph replays (not seen by oron_launch_count)
        dev = weights.device
        self.sm_pairs = torch.cuda.get_device_properties(dev).multi_processor_count // 2 if dev.type == "cuda" else 0

    # -------------------------------------------------------------------------------------------
    def workspace(self, nb: int, nbp: int, tpad: int, steps: int, keep_traj: bool) -> Workspace:
        key = (nb, nbp, tpad, steps, keep_traj)
        ws = self._ws.pop(key, None)
        if ws is not None:
            self._ws[key] = ws  # most recently used last
        if ws is None:
            if len(self._ws) >= 24:  # drop the least recently used shape (each holds its buffers and its CUDA graph)
                self._ws.pop(next(iter(self._ws)))
            ws = Workspace(self.w, nb, nbp, tpad, steps, keep_traj)
            self._ws[key] = ws
        return ws

    # -------------------------------------------------------------------------------------------
    def load_sequences(self, ws: Workspace, *, text: torch.Tensor, durations: list[int], seq_len: int,
                       branches: list[Branch], text_len: int | None = None) -> None:
        """Token ids (+1 shift, crop / 0-pad to seq_len: encoder.py:68-74), drop flags and lengths.

        ``text_len``: frames the text ConvNeXt blocks / GRN statistics cover. The reference's TextEmbedding runs over the whole
        padded length of the batch (encoder.py:68-96), so filler rows of shorter sequences enter the GRN norm: DiT.forward and
        the eval-mode CFM.forward pass ``text_len = seq_len`` and reproduce that. CFM.sample leaves it None = each sequence's
        own duration, which makes a batched sample equal to the per-utterance calls the reference's synthesize makes (B = 1
        there, f5tts.py:301-320) whatever the batch composition."""
        nb, tpad = ws.nb, ws.tpad
        ids = (text.to(torch.int64) + 1)[:, :seq_len]
        vocab1 = self.w.text_table.shape[0]
        # the range check reads the ids on the host: free for a CPU tensor (what F5TTS.synthesize passes), a host sync that
        # waits for everything enqueued so far for a CUDA tensor
        if ids.numel() and (int(ids.min()) < 0 or int(ids.max()) >= vocab1):  # nn.Embedding raises here too (encoder.py:75)
            raise IndexError(f"text ids out of range: ids must lie in [-1, {vocab1 - 2}] (index out of range in the text embedding table)")
        dev = self.w.device

        def h2d(dst, src):  # host values through pinned memory: an async copy instead of copy_'s stream synchronisation
            dst.copy_(src.pin_memory(), non_blocking=True)

        nbr = len(branches)
        if ids.is_cuda:
            ids2 = torch.zeros(nb, tpad, device=dev, dtype=torch.int32)
            ids2[:, : ids.shape[1]] = ids.to(torch.int32)
            ws.ids.view(ws.nbp, tpad).copy_(ids2.repeat(nbr, 1))
        else:
            ids2 = torch.zeros(nb, tpad, dtype=torch.int32)
            ids2[:, : ids.shape[1]] = ids.to(torch.int32)
            h2d(ws.ids.view(ws.nbp, tpad), ids2.repeat(nbr, 1))
        h2d(ws.drop, torch.tensor([int(br.drop_text) for br in branches for _ in range(nb)], dtype=torch.uint8))
        h2d(ws.seq_lens, torch.tensor(durations * nbr, dtype=torch.int32))
        h2d(ws.text_lens, torch.tensor((durations if text_len is None else [int(text_len)] * nb) * nbr, dtype=torch.int32))
        # schedule of the attention kernel for these lengths (once per call, outside the per-NFE graph)
        L.attention_plan(ws.attn_ws, nbatch=ws.nbp, rows_per_batch=tpad, heads=self.w.heads, seq_lens=ws.seq_lens)

    def text_embed(self, ws: Workspace) -> None:
        """TextEmbedding.forward for all branches at once -> ws.xt (fp32 [R, text_dim])."""
        w = self.w
        L.text_embed_front(ws.ids, ws.drop, w.text_table, w.pos_table(ws.tpad), rows_per_batch=ws.tpad, nb=ws.nbp,
                           x=ws.xt, row_valid=ws.row_valid)
        for blk in w.text_blocks:
            L.dwconv7_ln(ws.xt, rows_per_batch=ws.tpad, nbatch=ws.nbp, seq_lens=ws.text_lens, w=blk["dw_w"],
                         wb=blk["dw_b"], ln_w=blk["ln_w"], ln_b=blk["ln_b"], eps=1e-6, out=ws.xt_n)
            L.gemm(ws.xt_n, blk["w1"], ws.xt_h, epilogue=L.EPI_BF16, bias=blk["b1"], act=L.ACT_GELU_ERF,
                   rows_per_batch=ws.tpad, nbatch=ws.nbp, block_n=128)
            L.grn(ws.xt_h, rows_per_batch=ws.tpad, nb=ws.nbp, seq_lens=ws.text_lens, gamma=blk["gamma"],
                  beta=blk["beta"], gx2=ws.gx2)
            L.gemm(ws.xt_h, blk["w2"], ws.xt, epilogue=L.EPI_SCALE_RESID, bias=blk["b2"], rows_per_batch=ws.tpad,
                   nbatch=ws.nbp, addend=ws.xt, row_valid=ws.row_valid,
                   block_n=128 if w.text_dim % 128 == 0 else 64)

    def static_embed(self, ws: Workspace, cond: torch.Tensor, branches: list[Branch]) -> None:
        """C0 = [cond | text_embed] @ W[:, n_mels:]^T + b — the step-invariant part of InputEmbedding.proj."""
        w = self.w
        nb, tpad, M = ws.nb, ws.tpad, w.n_mels
        Rb = nb * tpad
        T = cond.shape[1]
        ws.cond.view(nb, tpad, M)[:, :T].copy_(cond)
        if T < tpad:
            ws.cond.view(nb, tpad, M)[:, T:].zero_()
        for i, br in enumerate(branches):
            dst = ws.a_ct[i * Rb:(i + 1) * Rb]
            if br.drop_audio:
                dst[:, :M].zero_()
            else:
                L.cast_rows_bf16(ws.cond, dst[:, :M])
        L.cast_rows_bf16(ws.xt, ws.a_ct[:, M:M + w.text_dim])
        L.gemm(ws.a_ct, w.wct, ws.c0, epilogue=L.EPI_F32, bias=w.in_b, rows_per_batch=ws.tpad, nbatch=ws.nbp,
               block_n=pick_block_n(ws.nbp * ws.tpad, w.dim))

    def modulation_table(self, ws: Workspace, times: torch.Tensor) -> None:
        """time_embed -> SiLU -> every AdaLN projection, for all rows of `times` (fp32 [n <= ws.steps])."""
        w = self.w
        n = times.numel()
        ws.tvals[:n].copy_(times)
        L.time_sinusoid(ws.tvals[:n], ws.tfeat[:n])
        L.gemm(ws.tfeat[:n], w.t0_w, ws.th[:n], epilogue=L.EPI_BF16, bias=w.t0_b, act=L.ACT_SILU)
        L.gemm(ws.th[:n], w.t2_w, ws.ts[:n], epilogue=L.EPI_BF16, bias=w.t2_b, act=L.ACT_SILU)
        L.gemm(ws.ts[:n], w.ada_w, ws.table[:n], epilogue=L.EPI_F32, bias=w.ada_b, block_n=256)

    # -------------------------------------------------------------------------------------------
    def velocity(self, ws: Workspace, *, mod_nb: int, use_step: bool) -> None:
        """One DiT forward from ws.xb (bf16 copy of the ODE state) to ws.v (fp32 velocity, all branches)."""
        w = self.w
        D, tpad, nbp = w.dim, ws.tpad, ws.nbp
        R = nbp * tpad
        cos, sin = w.rope(tpad)
        tab = ws.table.view(-1)
        step_ptr = ws.step if use_step else None
        # modulation rows: per ODE step one row shared by every sequence (mod_nb == 1), or one row per
        # batch element (generic DiT.forward with per-sample times)
        sstride = w.ada_n if use_step else 0
        mld = 0 if mod_nb == 1 else w.ada_n
        common = dict(rows_per_batch=tpad, nbatch=nbp)
        # transformer GEMMs run on the 2-SM kernel: 256 x 256 tile per SM pair (256 x 128 for narrow models)
        bn_big = 256 if D % 256 == 0 else 128
        # few rows (utterances up to ~3 s with CFG): launch-latency bound, one wave either way; the 1-SM kernel has no cluster
        # launch / cluster barriers in its prologue and epilogue. tools/short_utt_bench.py, T = 143: Small 724 -> 687 us per NFE,
        # Base 1595 -> 1565; at 1024 rows Small gains (811 -> 715) but Base loses (1725 -> 1894), so the switch stays at 600 rows
        # (ORON_SMALL_ROWS; 0 = always the 2-SM kernel)
        two = R > SMALL_ROWS
        if not two:
            bn_big = 128
        bn_res = resid_tile_width(tpad, nbp, D, bn_big, self.sm_pairs) if two else bn_big
        bn_up = up_tile_width(tpad, nbp, w.ff_dim, bn_big, self.sm_pairs) if two and not FFN_FUSED else bn_big

        L.gemm(ws.xb, w.wx, ws.h0, epilogue=L.EPI_EMBED_DUAL, addend=ws.c0, seq_lens=ws.seq_lens, out2=ws.h0b,
               block_n=128, **common)
        c1, c2 = w.conv_pos
        L.gemm(ws.h0b, c1["w"], ws.c1, epilogue=L.EPI_MISH_MASK_BF16, bias=c1["b"], taps=c1["taps"],
               cin_blocks=c1["gsz"] // 64, pad=c1["taps"] // 2, grouped=c1["gsz"], block_n=64, seq_lens=ws.seq_lens,
               **common)
        L.gemm(ws.c1, c2["w"], ws.xres, epilogue=L.EPI_MISH_MASK_RESID, bias=c2["b"], taps=c2["taps"],
               cin_blocks=c2["gsz"] // 64, pad=c2["taps"] // 2, grouped=c2["gsz"], block_n=64, seq_lens=ws.seq_lens,
               addend=ws.h0, **common)

        ln = dict(eps=1e-6, mod_ld=mld, mod_nb=mod_nb, step_stride=sstride, add_one=True, out_bf16=ws.nrm)
        # norm that follows a gated-residual GEMM: inside its launch (tail) when the 2-SM kernel runs and N = dim fits
        tail = LN_TAIL and two and not FFN_FUSED and D in (128, 256, 512, 768, 1024) and bn_big == 256
        of = w.depth * 6 * D  # AdaLayerNormFinal: (scale, shift) — modules.py:233
        L.ln_modulate(ws.xres, scale=tab[D:], shift=tab, step_ptr=step_ptr, **ln, **common)  # block 0, first norm
        for i, blk in enumerate(w.blocks):
            o = i * 6 * D  # (shift_msa, scale_msa, gate_msa, shift_mlp, scale_mlp, gate_mlp) — modules.py:215-217
            last = i == w.depth - 1
            L.gemm(ws.nrm, blk["wqkv"], ws.qkv, epilogue=L.EPI_QKV_ROPE, bias=blk["bqkv"], rope_cos=cos, rope_sin=sin,
                   rope_cols=2 * w.inner, f16_from_col=2 * w.inner, block_n=bn_big, two_sm=two, **common)
            L.attention(ws.qkv, ws.ao, nbatch=nbp, rows_per_batch=tpad, heads=w.heads, seq_lens=ws.seq_lens,
                        scale=1.0 / math.sqrt(w.dim_head), workspace=ws.attn_ws)
            d = L.gemm(ws.ao, blk["wo"], ws.xres, epilogue=L.EPI_GATE_RESID, bias=blk["bo"], gate=tab[o + 2 * D:],
                       gate_ld=mld, gate_nb=mod_nb, gate_step_stride=sstride, step_ptr=step_ptr, seq_lens=ws.seq_lens,
                       mask_rows=True, block_n=bn_res, two_sm=two, desc_only=tail, **common)  # K = dim: too short for stream-K to pay
            if tail:
                L.gemm_ln(d, ws.ln_cnt, scale=tab[o + 4 * D:], shift=tab[o + 3 * D:], **ln)
            else:
                L.ln_modulate(ws.xres, scale=tab[o + 4 * D:], shift=tab[o + 3 * D:], step_ptr=step_ptr, **ln, **common)
            fused = FFN_FUSED and STREAM_K and two and not self.deterministic and bn_big == 256 and w.ff_dim % 256 == 0
            sk_down = STREAM_K and two and not self.deterministic
            up = L.gemm(ws.nrm, blk["w1"], ws.hid, epilogue=L.EPI_BF16, bias=blk["b1"], act=L.ACT_GELU_TANH,
                        block_n=bn_up, two_sm=two, desc_only=fused, **common)
            down = L.gemm(ws.hid, blk["w2"], ws.xres, epilogue=L.EPI_GATE_RESID, bias=blk["b2"], gate=tab[o + 5 * D:],
                          gate_ld=mld, gate_nb=mod_nb, gate_step_stride=sstride, step_ptr=step_ptr, mask_rows=False,
                          block_n=bn_big if sk_down else bn_res, two_sm=two, stream_k=sk_down, desc_only=fused or tail, **common)
            if fused:
                L.ffn(up, down, ws.ffn_ws)
            # next block's first norm (shift, scale at o + 6D) or the final norm (scale, shift)
            nsc, nsh = (tab[of:], tab[of + D:]) if last else (tab[o + 7 * D:], tab[o + 6 * D:])
            if tail:
                L.gemm_ln(down, ws.ln_cnt, scale=nsc, shift=nsh, **ln)
            else:
                L.ln_modulate(ws.xres, scale=nsc, shift=nsh, step_ptr=step_ptr, **ln, **common)
        L.gemm(ws.nrm, w.wp, ws.v, epilogue=L.EPI_F32, bias=w.bp, block_n=128, **common)

    def euler(self, ws: Workspace, *, cfg: float, has_uncond: bool, method: int = 0) -> None:
        L.cfg_euler_step(ws.x, ws.v, nb=ws.nb, rows_per_batch=ws.tpad, n_mels=self.w.n_mels, has_uncond=has_uncond,
                         cfg=cfg, dt=ws.dt, step_ptr=ws.step, xb=ws.xb, traj=ws.traj, v_out=ws.vg, method=method)

    # -------------------------------------------------------------------------------------------
    def run_ode(self, ws: Workspace, *, steps: int, cfg: float, has_uncond: bool, method: int = 0) -> None:
        """`steps` x (DiT forward + CFG / ODE update): `steps` counts velocity evaluations (2 per interval for the
        midpoint rule, method 1). One evaluation is captured once into a CUDA graph and replayed."""

        def one_step():
            self.velocity(ws, mod_nb=1, use_step=True)
            self.euler(ws, cfg=cfg, has_uncond=has_uncond, method=method)

        if not self.use_graph:
            for _ in range(steps):
                one_step()
            return
        key = (cfg, has_uncond, method, self.deterministic)
        if ws.graph is None or ws.graph_key != key:
            # warm-up outside capture (sets kernel attributes, builds tables), then rewind the state it touched
            x_save, xb_save = ws.x.clone(), ws.xb.clone()
            one_step()
            ws.x.copy_(x_save)
            ws.xb.copy_(xb_save)
            ws.step.zero_()
            torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            n0 = L.launch_count()
            with torch.cuda.graph(g):
                one_step()
            ws.graph_launches = L.launch_count() - n0  # kernels replayed per ODE step
            ws.x.copy_(x_save)
            ws.xb.copy_(xb_save)
            ws.step.zero_()
            ws.graph, ws.graph_key = g, key
        for _ in range(steps):
            ws.graph.replay()
        self.replayed_launches += steps * ws.graph_launches
