"""Vocos vocoder (mel -> waveform) on the sm_100a kernels.

Call-site contract (src/models/f5tts.py:196-202, 416): ``Vocos.from_pretrained(repo).eval().to(device)``
and ``.decode(mel [B, 100, T]) -> wav [B, (T-1)*256]``. The module tree mirrors upstream vocos 0.1.x
(``backbone.embed / norm / convnext.{i}.{dwconv,norm,pwconv1,pwconv2,gamma} / final_layer_norm``,
``head.out``, ``head.istft.window``) so the pretrained ``charactr/vocos-mel-24khz`` state dict loads
unchanged. The upstream package and weights are not reachable in the build environment, so parity is
pinned on random-init weights against oracle/audio_oracle.py ("parity unpinned" for the
pretrained weights).

Kernels: embed Conv1d(k=7) as a dense implicit GEMM; per block depthwise-conv+LayerNorm (one fused
row-wise kernel), two tcgen05 GEMMs (GELU and layer-scale+residual epilogues); head GEMM; fused
exp/cos/sin + irFFT + window + overlap-add + envelope kernel.
"""

from __future__ import annotations

import os

import torch
import torch.nn as nn

from . import _lib as L

BF16, F32 = torch.bfloat16, torch.float32


def _version_of(p: torch.Tensor) -> int:
    """In-place update counter; tensors created under torch.inference_mode() (the vocoder is built inside
    ``F5TTS.synthesize``) do not have one."""
    try:
        return p._version
    except RuntimeError:
        return -1


class _ConvNeXtBlock(nn.Module):
    def __init__(self, dim: int, intermediate_dim: int, layer_scale_init_value: float):
        super().__init__()
        self.dwconv = nn.Conv1d(dim, dim, kernel_size=7, padding=3, groups=dim)
        self.norm = nn.LayerNorm(dim, eps=1e-6)
        self.pwconv1 = nn.Linear(dim, intermediate_dim)
        self.pwconv2 = nn.Linear(intermediate_dim, dim)
        self.gamma = nn.Parameter(layer_scale_init_value * torch.ones(dim))


class _Backbone(nn.Module):
    def __init__(self, input_channels: int, dim: int, intermediate_dim: int, num_layers: int):
        super().__init__()
        self.embed = nn.Conv1d(input_channels, dim, kernel_size=7, padding=3)
        self.norm = nn.LayerNorm(dim, eps=1e-6)
        self.convnext = nn.ModuleList([_ConvNeXtBlock(dim, intermediate_dim, 1.0 / num_layers) for _ in range(num_layers)])
        self.final_layer_norm = nn.LayerNorm(dim, eps=1e-6)
        for m in self.modules():
            if isinstance(m, (nn.Conv1d, nn.Linear)):
                nn.init.trunc_normal_(m.weight, std=0.02)
                nn.init.constant_(m.bias, 0)


class _ISTFT(nn.Module):
    def __init__(self, n_fft: int):
        super().__init__()
        self.register_buffer("window", torch.hann_window(n_fft))


class _Head(nn.Module):
    def __init__(self, dim: int, n_fft: int):
        super().__init__()
        self.out = nn.Linear(dim, n_fft + 2)
        self.istft = _ISTFT(n_fft)


class Vocos(nn.Module):
    def __init__(self, n_mels: int = 100, dim: int = 512, intermediate_dim: int = 1536, num_layers: int = 8,
                 n_fft: int = 1024, hop_length: int = 256) -> None:
        super().__init__()
        if (n_fft, hop_length) != (1024, 256) or dim != 512:
            raise NotImplementedError("kernels are specialised for the vocos-mel-24khz geometry (dim 512, n_fft 1024, hop 256)")
        self.n_mels, self.dim, self.n_fft, self.hop_length = n_mels, dim, n_fft, hop_length
        self.backbone = _Backbone(n_mels, dim, intermediate_dim, num_layers)
        self.head = _Head(dim, n_fft)
        self.__dict__["_packed"] = None
        self.__dict__["_packed_sig"] = None

    # ---- loading -----------------------------------------------------------------------------------
    @classmethod
    def from_pretrained(cls, repo_id: str, revision: str | None = None) -> "Vocos":
        """Load ``pytorch_model.bin`` of an upstream Vocos checkpoint from a local directory or the HF cache."""
        path = repo_id
        if not os.path.isdir(path):
            from huggingface_hub import hf_hub_download  # raises offline unless cached

            path = os.path.dirname(hf_hub_download(repo_id=repo_id, filename="pytorch_model.bin", revision=revision))
        sd = torch.load(os.path.join(path, "pytorch_model.bin"), map_location="cpu")
        sd = {k: v for k, v in sd.items() if not k.startswith("feature_extractor.")}
        model = cls()
        model.load_state_dict(sd, strict=True)
        return model.eval()

    def _apply(self, fn, *args, **kwargs):
        self.__dict__["_param_slots"] = None
        return super()._apply(fn, *args, **kwargs)

    def load_state_dict(self, *args, **kwargs):
        self.__dict__["_param_slots"] = None
        return super().load_state_dict(*args, **kwargs)

    # ---- packed weights ------------------------------------------------------------------------------
    def _pack(self):
        p0 = next(self.parameters())
        if not p0.is_cuda:
            raise RuntimeError("oron_tts_b200.Vocos runs only on a CUDA device (no CPU fallback)")
        slots = self.__dict__.get("_param_slots")  # see DiT._signature: parameters() walks the module tree on every decode
        if slots is None:
            slots = [(m, n) for m in self.modules() for n in m._parameters if m._parameters[n] is not None]
            self.__dict__["_param_slots"] = slots
        sig = tuple((p.data_ptr(), _version_of(p)) for p in (m._parameters[n] for m, n in slots))
        if self.__dict__["_packed"] is not None and self.__dict__["_packed_sig"] == sig:
            return self.__dict__["_packed"]
        dev = p0.device
        bf = lambda t: t.detach().to(dev, BF16).contiguous()
        f32 = lambda t: t.detach().to(dev, F32).contiguous()
        bb = self.backbone
        cin_pad = (self.n_mels + 63) // 64 * 64
        we = torch.zeros(self.dim, 7, cin_pad, device=dev, dtype=F32)
        we[:, :, : self.n_mels] = bb.embed.weight.detach().float().permute(0, 2, 1)
        pk = dict(
            cin_pad=cin_pad, embed_w=we.reshape(self.dim, 7 * cin_pad).to(BF16).contiguous(), embed_b=f32(bb.embed.bias),
            norm_w=f32(bb.norm.weight), norm_b=f32(bb.norm.bias),
            fin_w=f32(bb.final_layer_norm.weight), fin_b=f32(bb.final_layer_norm.bias),
            head_w=bf(self.head.out.weight), head_b=f32(self.head.out.bias), window=f32(self.head.istft.window),
            blocks=[dict(dw_w=f32(b.dwconv.weight).view(self.dim, 7).contiguous(), dw_b=f32(b.dwconv.bias),
                         ln_w=f32(b.norm.weight), ln_b=f32(b.norm.bias), w1=bf(b.pwconv1.weight), b1=f32(b.pwconv1.bias),
                         w2=bf(b.pwconv2.weight), b2=f32(b.pwconv2.bias), gamma=f32(b.gamma)) for b in bb.convnext],
        )
        self.__dict__["_packed"], self.__dict__["_packed_sig"] = pk, sig
        return pk

    # ---- decode -----------------------------------------------------------------------------------------
    @torch.no_grad()
    def decode(self, features_input: torch.Tensor) -> torch.Tensor:
        """mel [B, n_mels, T] (fp32, CUDA) -> waveform [B, (T-1)*hop]."""
        pk = self._pack()
        mel = features_input
        if mel.dim() == 2:
            mel = mel.unsqueeze(0)
        B, M, T = mel.shape
        if T < 2:
            raise ValueError("Vocos.decode needs at least 2 frames")
        dev = mel.device
        R = B * T
        D, H = self.dim, pk["blocks"][0]["w1"].shape[0]
        # frame-major bf16 operand for the embed conv (channels zero-padded to a multiple of 64)
        a0 = torch.zeros(R, pk["cin_pad"], device=dev, dtype=BF16)
        L.cast_rows_bf16(mel.transpose(1, 2).reshape(R, M).float().contiguous(), a0[:, :M])
        x = torch.empty(R, D, device=dev, dtype=F32)
        # tile choice (tools/cfg4_bench.py): 256-wide tiles win once there are enough row tiles to fill 148 SMs;
        # short utterances keep 128-wide tiles for more CTAs
        big = R >= 8192
        L.gemm(a0, pk["embed_w"], x, epilogue=L.EPI_F32, bias=pk["embed_b"], rows_per_batch=T, nbatch=B, taps=7,
               cin_blocks=pk["cin_pad"] // 64, pad=3, block_n=256 if big else 128)
        L.ln_modulate(x, rows_per_batch=T, nbatch=B, eps=1e-6, scale=pk["norm_w"], shift=pk["norm_b"], add_one=False,
                      out_f32=x)
        n = torch.empty(R, D, device=dev, dtype=BF16)
        h = torch.empty(R, H, device=dev, dtype=BF16)
        for blk in pk["blocks"]:
            L.dwconv7_ln(x, rows_per_batch=T, nbatch=B, seq_lens=None, w=blk["dw_w"], wb=blk["dw_b"], ln_w=blk["ln_w"],
                         ln_b=blk["ln_b"], eps=1e-6, out=n)
            L.gemm(n, blk["w1"], h, epilogue=L.EPI_BF16, bias=blk["b1"], act=L.ACT_GELU_ERF, rows_per_batch=T, nbatch=B,
                   block_n=256 if H % 256 == 0 else 128, two_sm=big and H % 256 == 0)
            L.gemm(h, blk["w2"], x, epilogue=L.EPI_SCALE_RESID, bias=blk["b2"], rows_per_batch=T, nbatch=B, addend=x,
                   gate=blk["gamma"], block_n=256 if big else 128, two_sm=big)
        L.ln_modulate(x, rows_per_batch=T, nbatch=B, eps=1e-6, scale=pk["fin_w"], shift=pk["fin_b"], add_one=False,
                      out_bf16=n)
        nh = pk["head_w"].shape[0]
        ldh = (nh + 31) // 32 * 32
        hs = torch.empty(R, ldh, device=dev, dtype=F32)
        L.gemm(n, pk["head_w"], hs, epilogue=L.EPI_F32, bias=pk["head_b"], rows_per_batch=T, nbatch=B,
               block_n=256 if big else 128, n=nh, two_sm=big)
        wav = torch.empty(B, (T - 1) * self.hop_length, device=dev, dtype=F32)
        L.istft_head(hs, pk["window"], wav, rows_per_batch=T, nb=B, n_frames=T, mode=0)
        return wav

    def forward(self, features_input: torch.Tensor) -> torch.Tensor:
        return self.decode(features_input)
