"""ORACLE — TEST INFRASTRUCTURE ONLY (see oracle/dit_oracle.py header).

CPU fp32 restatements of
  * AudioProcessor.mel_spectrogram / normalize_audio  (src/utils/audio.py:28-30, 50-58, 73-77, 94-110),
    i.e. torchaudio MelSpectrogram(center=True, reflect, periodic Hann, power=1, HTK, norm=None) + safe log —
    written out explicitly (framing, rFFT, hand-built HTK filterbank) so it does not call torchaudio;
    pinned bit-for-bit against the live reference in tests/golden.
  * the Vocos decoder. The pretrained `vocos` package is NOT in /root/reference nor installed here
    (pyproject.toml:50-52 pins only `vocos>=0.1.0`; weights `charactr/vocos-mel-24khz` need the network):
    PARITY UNPINNED for upstream weights. The restatement follows upstream vocos 0.1.0
    (models.VocosBackbone, modules.ConvNeXtBlock, heads.ISTFTHead, spectral_ops.ISTFT padding="center") as
    summarised in SURVEY.md §8(a16); its shared sub-ops (dwconv k7, LayerNorm, pointwise MLP, irfft +
    overlap-add) are pinned against the importable in-repo analogue src/models/decoder.py:8-103.
"""

from __future__ import annotations

import math

import torch
import torch.nn.functional as F

Tensor = torch.Tensor


# --------------------------------------------------------------------------------------------
# log-mel
# --------------------------------------------------------------------------------------------
def htk_filterbank(n_freqs: int = 513, n_mels: int = 100, sample_rate: int = 24000,
                   f_min: float = 0.0, f_max: float | None = None) -> Tensor:
    """Triangular HTK mel filterbank [n_freqs, n_mels], no area normalisation."""
    f_max = sample_rate / 2 if f_max is None else f_max
    bins = torch.linspace(0, sample_rate // 2, n_freqs)
    m_lo = 2595.0 * math.log10(1.0 + f_min / 700.0)
    m_hi = 2595.0 * math.log10(1.0 + f_max / 700.0)
    m_pts = torch.linspace(m_lo, m_hi, n_mels + 2)
    f_pts = 700.0 * (10.0 ** (m_pts / 2595.0) - 1.0)
    width = f_pts[1:] - f_pts[:-1]
    slopes = f_pts[None, :] - bins[:, None]
    down = -slopes[:, :-2] / width[:-1]
    up = slopes[:, 2:] / width[1:]
    return torch.clamp(torch.minimum(down, up), min=0.0)


def log_mel(wav: Tensor, n_fft: int = 1024, hop: int = 256, n_mels: int = 100, sample_rate: int = 24000,
            clip: float = 1e-5) -> Tensor:
    """wav [S] or [B, S] -> [n_mels, T] or [B, n_mels, T], T = 1 + S // hop."""
    squeeze = wav.dim() == 1
    x = wav[None] if squeeze else wav
    xp = F.pad(x[:, None, :], (n_fft // 2, n_fft // 2), mode="reflect")[:, 0]
    frames = xp.unfold(-1, n_fft, hop)                                  # [B, T, n_fft]
    w = torch.hann_window(n_fft, periodic=True, dtype=x.dtype)
    mag = torch.fft.rfft(frames * w, dim=-1).abs()                      # [B, T, 513]
    mel = mag @ htk_filterbank(n_fft // 2 + 1, n_mels, sample_rate).to(x.dtype)
    out = torch.log(torch.clamp(mel, min=clip)).transpose(1, 2)
    return out[0] if squeeze else out


def peak_normalize(x: Tensor) -> Tensor:
    mx = x.abs().max()
    if mx < 1e-8:
        return x
    return torch.clamp(x / (mx + 1e-7), -1.0, 1.0)


# --------------------------------------------------------------------------------------------
# iSTFT ("center" padding): irfft, window, overlap-add, divide by the window envelope, trim n_fft/2
# --------------------------------------------------------------------------------------------
def istft_center(spec: Tensor, n_fft: int = 1024, hop: int = 256, normalized: bool = False) -> Tensor:
    """spec complex [B, n_fft/2+1, T] -> [B, hop * (T - 1)]."""
    B, _, T = spec.shape
    w = torch.hann_window(n_fft, periodic=True, dtype=torch.float32)
    if normalized:
        spec = spec * math.sqrt(n_fft)
    fr = torch.fft.irfft(spec.transpose(1, 2), n=n_fft, dim=-1) * w      # [B, T, n_fft]
    length = n_fft + hop * (T - 1)
    y = torch.zeros(B, length)
    env = torch.zeros(length)
    for t in range(T):
        y[:, t * hop:t * hop + n_fft] += fr[:, t]
        env[t * hop:t * hop + n_fft] += w * w
    y = y / env
    return y[:, n_fft // 2:length - n_fft // 2]


# --------------------------------------------------------------------------------------------
# Vocos (upstream layout): backbone.embed / norm / convnext.{i}.{dwconv,norm,pwconv1,pwconv2,gamma} /
# final_layer_norm ; head.out ; head.istft.window
# --------------------------------------------------------------------------------------------
def vocos_backbone(sd: dict, mel: Tensor) -> Tensor:
    """mel [B, n_mels, T] -> features [B, T, dim]."""
    x = F.conv1d(mel, sd["backbone.embed.weight"], sd["backbone.embed.bias"], padding=3)
    x = F.layer_norm(x.transpose(1, 2), (x.shape[1],), sd["backbone.norm.weight"], sd["backbone.norm.bias"], eps=1e-6)
    n_layers = 1 + max(int(k.split(".")[2]) for k in sd if k.startswith("backbone.convnext."))
    for i in range(n_layers):
        p = f"backbone.convnext.{i}."
        C = x.shape[-1]
        y = F.conv1d(x.transpose(1, 2), sd[p + "dwconv.weight"], sd[p + "dwconv.bias"], padding=3, groups=C)
        y = F.layer_norm(y.transpose(1, 2), (C,), sd[p + "norm.weight"], sd[p + "norm.bias"], eps=1e-6)
        y = F.linear(F.gelu(F.linear(y, sd[p + "pwconv1.weight"], sd[p + "pwconv1.bias"])),
                     sd[p + "pwconv2.weight"], sd[p + "pwconv2.bias"])
        if p + "gamma" in sd:
            y = sd[p + "gamma"] * y
        x = x + y
    C = x.shape[-1]
    return F.layer_norm(x, (C,), sd["backbone.final_layer_norm.weight"], sd["backbone.final_layer_norm.bias"], eps=1e-6)


def vocos_decode(sd: dict, mel: Tensor) -> Tensor:
    """mel [B, n_mels, T] -> wav [B, 256 (T-1)]; ISTFTHead: exp-magnitude (clipped at 1e2) and phase halves."""
    feat = vocos_backbone(sd, mel)
    h = F.linear(feat, sd["head.out.weight"], sd["head.out.bias"])       # [B, T, n_fft + 2]
    nb = h.shape[-1] // 2
    mag = torch.clip(torch.exp(h[..., :nb]), max=1e2)
    ph = h[..., nb:]
    spec = torch.complex(mag * torch.cos(ph), mag * torch.sin(ph)).transpose(1, 2)
    return istft_center(spec, n_fft=2 * (nb - 1), hop=(2 * (nb - 1)) // 4, normalized=False)
