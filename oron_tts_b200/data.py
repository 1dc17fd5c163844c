"""Training data path on the GPU: what ``TTSDataset.__getitem__`` + ``TTSCollator`` do on the host per
item (src/data/dataset.py:187-222, 334-362) — peak normalisation, log-mel, token ids stretched to the mel length,
zero / -1 padding — as one call that takes the raw waveforms of a batch and leaves the training batch on the device,
with the audio arithmetic in the fused sm_100a kernels; plus the frame-budget batch plan of ``DynamicBatchSampler``
(dataset.py:365-423), which is pure index logic.

The reference computes every mel on CPU DataLoader workers (dataset.py:212) and ships [B, 100, T] floats to the GPU;
here 4 bytes per sample go up and the spectrogram never exists on the host.
"""

from __future__ import annotations

import torch

from .audio import AudioProcessor
from .f5tts import _stretch_text_to_len
from .text import TextCleaner


class GpuBatcher:
    def __init__(self, sample_rate: int = 24000, n_mels: int = 100, min_duration_s: float = 1.0, device: str = "cuda"):
        self.audio = AudioProcessor(sample_rate=sample_rate, n_mels=n_mels)
        self.text = TextCleaner()
        self.sample_rate, self.n_mels, self.device = sample_rate, n_mels, device
        self.min_audio_len = int(min_duration_s * sample_rate)

    @torch.no_grad()
    def __call__(self, waveforms: list[torch.Tensor], texts: list[str], langs: list[str] | None = None,
                 attr_tokens: list[list[str]] | None = None) -> dict[str, torch.Tensor]:
        """waveforms: 1-D float tensors (host, ideally pinned, or device) at ``sample_rate``. Returns the collated batch
        {"mel" [B, n_mels, T], "text_ids" [B, T] (-1 padded), "mask" [B, T], "mel_lengths" [B]} on the device."""
        n = len(waveforms)
        if n == 0 or n != len(texts):
            raise ValueError("waveforms and texts must be non-empty lists of equal length")
        langs = langs or ["mn"] * n
        attr_tokens = attr_tokens or [[] for _ in range(n)]
        hop = self.audio.hop_length
        frames = [1 + int(w.numel()) // hop for w in waveforms]
        t_max = max(frames)
        dev = self.device
        mel = torch.zeros(n, self.n_mels, t_max, device=dev)
        ids = torch.full((n, t_max), -1, dtype=torch.long)
        for i, w in enumerate(waveforms):
            if w.numel() < self.min_audio_len:
                raise ValueError(f"Audio too short at sample {i}: {w.numel() / self.sample_rate:.2f}s")
            x = w.reshape(-1).to(dev, non_blocking=True).float()
            x = self.audio.normalize_audio(x)                       # peak-normalise kernel (audio.py:73-77)
            # one launch per clip: the reflect padding of the STFT belongs to the clip's own end, not to the batch's
            mel[i, :, : frames[i]] = self.audio.mel_spectrogram(x)  # fused frame + window + FFT + mel + log kernel
            raw = self.text.text_to_sequence(texts[i], lang=langs[i], attr_tokens=attr_tokens[i])
            ids[i, : frames[i]] = torch.tensor(_stretch_text_to_len(raw, frames[i]), dtype=torch.long)
        lengths = torch.tensor(frames, dtype=torch.long)
        mask = torch.arange(t_max)[None, :] < lengths[:, None]
        if not bool(torch.isfinite(mel).all()):
            raise ValueError("Invalid audio values in the batch")
        return {"mel": mel, "text_ids": ids.to(dev), "mask": mask.to(dev), "mel_lengths": lengths.to(dev)}


def dynamic_batches(durations: list[float], frames_threshold: int, max_samples: int = 0, sample_rate: int = 24000,
                    hop_length: int = 256, drop_last: bool = False) -> list[list[int]]:
    """The batch plan of DynamicBatchSampler (dataset.py:375-411): indices sorted by frame length, greedily packed while
    the frame total stays within ``frames_threshold`` (and at most ``max_samples`` per batch when > 0)."""
    frame_lens = [d * sample_rate / hop_length for d in durations]
    batches: list[list[int]] = []
    cur: list[int] = []
    total = 0.0
    for idx, fl in sorted(enumerate(frame_lens), key=lambda x: x[1]):
        if total + fl <= frames_threshold and (max_samples == 0 or len(cur) < max_samples):
            cur.append(idx)
            total += fl
        else:
            if cur:
                batches.append(cur)
            cur, total = [idx], fl
    if cur and not drop_last:
        batches.append(cur)
    return batches


def epoch_order(n_batches: int, epoch: int) -> list[int]:
    """Batch order of an epoch (dataset.py:416-420): torch.randperm seeded with the epoch number."""
    g = torch.Generator()
    g.manual_seed(epoch)
    return torch.randperm(n_batches, generator=g).tolist()
