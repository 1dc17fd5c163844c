"""AudioProcessor with the reference's interface (src/utils/audio.py:33-113); the log-mel front end
runs in the fused sm_100a kernel (frame + window + rFFT + |.| + mel + log in one pass, spectrum never
written to HBM) instead of torchaudio's MelSpectrogram.
"""

from __future__ import annotations

import math
from pathlib import Path

import torch

from . import _lib as L

DEFAULT_SAMPLE_RATE = 24000
DEFAULT_N_MELS = 100
DEFAULT_N_FFT = 1024
DEFAULT_HOP_LENGTH = 256
DEFAULT_WIN_LENGTH = 1024


def mel_filterbank(n_freqs: int, n_mels: int, sample_rate: int, f_min: float = 0.0, f_max: float | None = None) -> torch.Tensor:
    """HTK triangular filters [n_freqs, n_mels], norm=None — what torchaudio MelScale builds by default
    (audio.py:50-58 passes no f_min/f_max/norm/mel_scale)."""
    f_max = float(sample_rate // 2) if f_max is None else f_max
    all_freqs = torch.linspace(0, sample_rate // 2, n_freqs)
    m_min = 2595.0 * math.log10(1.0 + f_min / 700.0)
    m_max = 2595.0 * math.log10(1.0 + f_max / 700.0)
    f_pts = 700.0 * (10 ** (torch.linspace(m_min, m_max, n_mels + 2) / 2595.0) - 1.0)
    f_diff = f_pts[1:] - f_pts[:-1]
    slopes = f_pts.unsqueeze(0) - all_freqs.unsqueeze(1)
    down = (-1.0 * slopes[:, :-2]) / f_diff[:-1]
    up = slopes[:, 2:] / f_diff[1:]
    return torch.max(torch.zeros(1), torch.min(down, up))


class AudioProcessor:
    def __init__(self, sample_rate: int = DEFAULT_SAMPLE_RATE, n_fft: int = DEFAULT_N_FFT,
                 hop_length: int = DEFAULT_HOP_LENGTH, win_length: int = DEFAULT_WIN_LENGTH,
                 n_mels: int = DEFAULT_N_MELS) -> None:
        self.sample_rate = sample_rate
        self.n_fft = n_fft
        self.hop_length = hop_length
        self.win_length = win_length
        self.n_mels = n_mels
        if (n_fft, hop_length, win_length) != (1024, 256, 1024):
            raise NotImplementedError("the sm_100a STFT kernels are specialised for n_fft=1024, hop=256, win=1024")
        self._fb_cpu = mel_filterbank(n_fft // 2 + 1, n_mels, sample_rate).contiguous()
        self._win_cpu = torch.hann_window(win_length, periodic=True)
        self._dev_cache: dict = {}

    def _tables(self, device: torch.device):
        key = str(device)
        if key not in self._dev_cache:
            fb = self._fb_cpu.to(device)
            with torch.inference_mode(False):
                bands = L.logmel_bands(fb)  # the banded form of the filterbank, once per device
            self._dev_cache[key] = (self._win_cpu.to(device), fb, bands)
        return self._dev_cache[key]

    # ---- I/O helpers: unchanged behaviour, not on the GPU path ------------------------------------
    def load_audio(self, path: str | Path) -> tuple[torch.Tensor, int]:
        import torchaudio

        waveform, sr = torchaudio.load(str(path))
        if sr != self.sample_rate:
            waveform = torchaudio.functional.resample(waveform, sr, self.sample_rate)
        if waveform.shape[0] > 1:
            waveform = waveform.mean(dim=0, keepdim=True)
        return waveform.squeeze(0), self.sample_rate

    def save_audio(self, path: str | Path, audio) -> None:
        import soundfile as sf

        if isinstance(audio, torch.Tensor):
            audio = audio.cpu().numpy()
        sf.write(str(path), audio, self.sample_rate)

    def trim_silence(self, audio: torch.Tensor, top_db: float = 20.0, frame_length: int = 2048,
                     hop_length: int = 512) -> torch.Tensor:
        import librosa

        trimmed, _ = librosa.effects.trim(audio.cpu().numpy(), top_db=top_db, frame_length=frame_length,
                                          hop_length=hop_length)
        return torch.from_numpy(trimmed)

    def get_audio_duration(self, audio: torch.Tensor) -> float:
        return len(audio) / self.sample_rate

    # ---- GPU path ----------------------------------------------------------------------------------
    def normalize_audio(self, audio: torch.Tensor) -> torch.Tensor:
        """Peak normalisation (audio.py:73-77). CUDA tensors use the kernel; CPU tensors (the reference calls
        this before ``.to(device)``, f5tts.py:357) are tiny host-side prep and stay in torch."""
        if not audio.is_cuda:
            mx = audio.abs().max()
            return audio if mx < 1e-8 else torch.clamp(audio / (mx + 1e-7), -1.0, 1.0)
        x = audio.reshape(1, -1).contiguous().float()
        out = torch.empty_like(x)
        L.peak_normalize(x, out, torch.empty(1, device=x.device))
        return out.reshape(audio.shape)

    def mel_spectrogram(self, audio: torch.Tensor) -> torch.Tensor:
        """Waveform [T] / [1, T] / [B, T] on a CUDA device -> log-mel [n_mels, frames] ([B, n_mels, frames])."""
        if not audio.is_cuda:
            raise RuntimeError("AudioProcessor.mel_spectrogram runs only on CUDA tensors (no CPU fallback); "
                               "move the waveform to the GPU first, as F5TTS._synthesize_segment does")
        x = audio.unsqueeze(0) if audio.dim() == 1 else audio
        x = x.contiguous().float()
        window, fb, bands = self._tables(x.device)
        frames = 1 + x.shape[1] // self.hop_length
        out = torch.empty(x.shape[0], self.n_mels, frames, device=x.device, dtype=torch.float32)
        L.logmel(x, window, fb, out, clip=1e-5, bands=bands)
        return out.squeeze(0)
