"""F5TTS top level: text -> ids -> (reference mel) -> CFM.sample -> Vocos, with the reference's public
API (src/models/f5tts.py:111-444) so that scripts/infer.py runs unchanged. Host orchestration stays
Python; every tensor op on the path (log-mel, DiT, ODE, vocoder) runs in the sm_100a kernels.

Integer contracts that must match the reference bit for bit (north_star): chunking
(f5tts.py:43-75), token stretching (:95-108), target length / T_total rules (:366-377).
"""

from __future__ import annotations

import logging
import re
from pathlib import Path
from typing import Any

import torch
import torch.nn as nn
import torch.nn.functional as F

from .audio import AudioProcessor
from .dit import DiT
from .flow import CFM
from .text import TextCleaner, validate_language

_log = logging.getLogger(__name__)

_KAZAKH_ONLY = frozenset("әғқңұһі")
_MAX_CHARS = 120
_PAUSE_S = 0.25
_BREAK_TIERS = (".!?…", ",;:", " ")


def _normalise_synthesis_text(text: str) -> str:
    return re.sub(r"\s+", " ", text).strip()


def _find_split_index(text: str, max_chars: int) -> int:
    """Last sentence break, else clause break, else space inside (0.55*max, max]; else hard cut."""
    hi = min(max_chars, len(text))
    lo = max(1, int(max_chars * 0.55))
    for tier in _BREAK_TIERS:
        for cut in range(hi, lo, -1):
            if text[cut - 1] in tier:
                return cut
    return hi


def split_text_for_synthesis(text: str, max_chars: int) -> list[str]:
    rest = _normalise_synthesis_text(text)
    if not rest:
        return []
    if max_chars < 1:
        return [rest]
    pieces: list[str] = []
    while len(rest) > max_chars:
        cut = _find_split_index(rest, max_chars)
        head = rest[:cut].strip()
        if head:
            pieces.append(head)
        rest = rest[cut:].strip()
    if rest:
        pieces.append(rest)
    return pieces


def _concat_with_pause(waveforms: list[torch.Tensor], sample_rate: int, pause_s: float) -> torch.Tensor:
    if not waveforms:
        return torch.empty(0)
    gap = int(sample_rate * pause_s) if (len(waveforms) > 1 and pause_s > 0) else 0
    if gap <= 0:
        return torch.cat(waveforms)
    silence = torch.zeros(gap, dtype=waveforms[0].dtype)
    seq: list[torch.Tensor] = [waveforms[0]]
    for wav in waveforms[1:]:
        seq += [silence, wav]
    return torch.cat(seq)


def _stretch_text_to_len(token_ids: list[int], target_len: int) -> list[int]:
    """Frame i takes token floor(i*n/T); too many tokens are truncated; none -> all filler (-1)."""
    n = len(token_ids)
    if n == 0:
        return [-1] * target_len
    if n >= target_len:
        return token_ids[:target_len]
    return [token_ids[int(i * n / target_len)] for i in range(target_len)]


def estimate_target_len(*, target_duration_s: float | None, sample_rate: int, hop_length: int, ref_len: int,
                        n_ref_ids: int, n_target_ids: int, text: str, speed: float) -> int:
    """Target frame count (f5tts.py:366-375)."""
    if target_duration_s is not None:
        return max(1, int(target_duration_s * sample_rate / hop_length))
    if ref_len > 0 and n_ref_ids:
        return max(50, int(ref_len * n_target_ids / n_ref_ids / speed))
    chars = max(1, len(text.replace(" ", "")))
    return max(50, int(chars * 13 / speed))


def _to_device(module: nn.Module, device: str | torch.device) -> nn.Module:
    """``module.to(device)``, skipped when the module already lives there (judged by its first parameter)."""
    want = torch.device(device)
    if want.type == "cuda" and want.index is None and torch.cuda.is_available():
        want = torch.device("cuda", torch.cuda.current_device())
    first = next(module.parameters(), None)
    if first is not None and first.device == want:
        return module
    return module.to(device)


class F5TTS(nn.Module):
    def __init__(self, n_mels: int = 100, vocab_size: int = 65, dim: int = 1024, depth: int = 22, heads: int = 16,
                 dim_head: int = 64, ff_mult: int = 4, text_dim: int = 512, conv_layers: int = 4, p_dropout: float = 0.1,
                 audio_drop_prob: float = 0.3, cond_drop_prob: float = 0.2,
                 frac_lengths_mask: tuple[float, float] = (0.7, 1.0), sample_rate: int = 24000, n_fft: int = 1024,
                 hop_length: int = 256, gradient_checkpointing: bool = False) -> None:
        super().__init__()
        self.n_mels = n_mels
        self.sample_rate = sample_rate
        self.hop_length = hop_length
        self._text_cleaner = TextCleaner()
        self._audio_processor = AudioProcessor(sample_rate=sample_rate, n_fft=n_fft, hop_length=hop_length, n_mels=n_mels)
        backbone = DiT(dim=dim, depth=depth, heads=heads, dim_head=dim_head, ff_mult=ff_mult, dropout=p_dropout,
                       mel_dim=n_mels, vocab_size=vocab_size, text_dim=text_dim, conv_layers=conv_layers,
                       gradient_checkpointing=gradient_checkpointing)
        self.cfm = CFM(backbone, audio_drop_prob=audio_drop_prob, cond_drop_prob=cond_drop_prob,
                       frac_lengths_mask=frac_lengths_mask, n_mels=n_mels)

    def forward(self, mel: torch.Tensor, text_ids: torch.Tensor, lens: torch.Tensor | None = None) -> torch.Tensor:
        if lens is not None and lens.dtype == torch.bool and lens.ndim == 2:
            lens = lens.sum(dim=-1).long()
        return self.cfm(mel, text_ids, lens=lens)

    def _ready(self, device: str) -> None:
        """``self.eval(); self.to(device)`` (f5tts.py:243-244) without the two sweeps over ~2 k submodules when nothing
        would change: they cost ~4 ms of host time per call during which the GPU has nothing to do."""
        if self.training:
            self.eval()
        _to_device(self, device)

    # ---- vocoder: cached outside the parameter tree (f5tts.py:190-202) ---------------------------------
    def _get_vocos(self, device: str) -> Any:
        vocos = self.__dict__.get("_vocos_cache")
        if vocos is None:
            from .vocos import Vocos

            vocos = Vocos.from_pretrained("charactr/vocos-mel-24khz").eval()
            object.__setattr__(self, "_vocos_cache", vocos)
        return _to_device(vocos, device)

    def set_vocoder(self, vocos: Any) -> None:
        """Install a vocoder instance explicitly (offline boxes: random-init or locally stored weights)."""
        object.__setattr__(self, "_vocos_cache", vocos)

    @staticmethod
    def _warn_lang_contamination(text: str, lang: str) -> None:
        if validate_language(lang) != "mn":
            return
        odd = sorted({c for c in text.lower() if c in _KAZAKH_ONLY})
        if odd:
            _log.warning("Mongolian input contains Kazakh-only characters %s; the model was conditioned with "
                         "[LANG_MN] and may produce out-of-distribution audio.", odd)

    # ---- inference -----------------------------------------------------------------------------------------
    @torch.inference_mode()
    def synthesize(self, text: str, lang: str = "mn", ref_audio_path: str | Path | None = None,
                   ref_text: str | None = None, n_steps: int = 32, cfg_strength: float = 2.0,
                   sway_sampling_coef: float | None = -1.0, speed: float = 1.0, target_duration_s: float | None = None,
                   max_chars_per_chunk: int | None = _MAX_CHARS, pause_s: float = _PAUSE_S, seed: int | None = None,
                   device: str = "cuda") -> torch.Tensor:
        lang = validate_language(lang)
        if n_steps < 1:
            raise ValueError(f"n_steps must be >= 1, got {n_steps}")
        if cfg_strength < 0:
            raise ValueError(f"cfg_strength must be >= 0, got {cfg_strength}")
        if speed <= 0:
            raise ValueError(f"speed must be > 0, got {speed}")
        if target_duration_s is not None and target_duration_s <= 0:
            raise ValueError(f"target_duration_s must be > 0, got {target_duration_s}")
        if max_chars_per_chunk is not None and max_chars_per_chunk < 0:
            raise ValueError(f"max_chars_per_chunk must be >= 0, got {max_chars_per_chunk}")
        if pause_s < 0:
            raise ValueError(f"pause_s must be >= 0, got {pause_s}")
        self._ready(device)
        self._warn_lang_contamination(text, lang)
        if ref_text:
            self._warn_lang_contamination(ref_text, lang)

        limit = max_chars_per_chunk or 0
        chunks = [c for c in (split_text_for_synthesis(text, limit) if limit > 0 else [text.strip()]) if c]
        if not chunks:
            raise ValueError("text must not be empty")
        common = dict(lang=lang, ref_audio_path=ref_audio_path, ref_text=ref_text, n_steps=n_steps,
                      cfg_strength=cfg_strength, sway_sampling_coef=sway_sampling_coef, speed=speed, device=device)
        if len(chunks) == 1:
            return self._synthesize_segment(text=chunks[0], target_duration_s=target_duration_s, seed=seed, **common)

        _log.info("Splitting long synthesis request into %d chunks", len(chunks))
        weights = [max(1, len(c.replace(" ", ""))) for c in chunks]
        total = sum(weights)
        waves: list[torch.Tensor] = []
        for i, chunk in enumerate(chunks):
            dur = None if target_duration_s is None else target_duration_s * weights[i] / total
            waves.append(self._synthesize_segment(text=chunk, target_duration_s=dur,
                                                  seed=None if seed is None else seed + i, **common))
        return _concat_with_pause(waves, self.sample_rate, pause_s)

    # ---- batched inference (extension over the reference) -----------------------------------
    @torch.inference_mode()
    def synthesize_batch(self, texts: list[str], lang: str = "mn", ref_audio_path: str | Path | None = None,
                         ref_text: str | None = None, n_steps: int = 32, cfg_strength: float = 2.0,
                         sway_sampling_coef: float | None = -1.0, speed: float = 1.0,
                         target_durations_s: list[float | None] | None = None,
                         max_chars_per_chunk: int | None = _MAX_CHARS, pause_s: float = _PAUSE_S,
                         seeds: list[int | None] | None = None, max_rows_per_batch: int = 8192,
                         device: str = "cuda") -> list[torch.Tensor]:
        """``[synthesize(t, ...) for t in texts]`` with the segments of all requests packed into length-sorted
        batches (the reference loops B = 1, f5tts.py:301-320). Every segment keeps the semantics of a B = 1 call:
        its own seeded noise draw (flow.py:270-283), its own frame counts, GRN / convolutions over its own frames
        only, so the waveforms equal the one-by-one results up to the order of fp32 partial sums.
        One shared reference clip (``ref_audio_path`` / ``ref_text``) conditions every request."""
        lang = validate_language(lang)
        if n_steps < 1:
            raise ValueError(f"n_steps must be >= 1, got {n_steps}")
        if cfg_strength < 0:
            raise ValueError(f"cfg_strength must be >= 0, got {cfg_strength}")
        if speed <= 0:
            raise ValueError(f"speed must be > 0, got {speed}")
        if max_chars_per_chunk is not None and max_chars_per_chunk < 0:
            raise ValueError(f"max_chars_per_chunk must be >= 0, got {max_chars_per_chunk}")
        if pause_s < 0:
            raise ValueError(f"pause_s must be >= 0, got {pause_s}")
        n = len(texts)
        durs = list(target_durations_s) if target_durations_s is not None else [None] * n
        sds = list(seeds) if seeds is not None else [None] * n
        if len(durs) != n or len(sds) != n:
            raise ValueError("target_durations_s and seeds must have one entry per text")
        if any(d is not None and d <= 0 for d in durs):
            raise ValueError("target_duration_s must be > 0")
        from .shard import plan_batches

        self._ready(device)
        ap = self._audio_processor
        ref_mel_raw = None
        if ref_audio_path is not None:
            if isinstance(ref_audio_path, torch.Tensor):
                wav = ap.normalize_audio(ref_audio_path.reshape(-1).to(device, non_blocking=True))
            else:
                wav, _ = ap.load_audio(ref_audio_path)
                wav = ap.normalize_audio(wav).to(device)
            ref_mel_raw = ap.mel_spectrogram(wav)  # [n_mels, T_ref]
        limit = max_chars_per_chunk or 0
        segs: list[dict] = []  # one per (request, chunk)
        owners: list[list[int]] = []
        for r, text in enumerate(texts):
            self._warn_lang_contamination(text, lang)
            chunks = [c for c in (split_text_for_synthesis(text, limit) if limit > 0 else [text.strip()]) if c]
            if not chunks:
                raise ValueError("text must not be empty")
            weights = [max(1, len(c.replace(" ", ""))) for c in chunks]
            total_w = sum(weights)
            mine = []
            for i, chunk in enumerate(chunks):
                if len(chunks) == 1:
                    dur, seed = durs[r], sds[r]
                else:
                    dur = None if durs[r] is None else durs[r] * weights[i] / total_w
                    seed = None if sds[r] is None else sds[r] + i
                plan = self.prepare_segment(chunk, lang, ref_mel_raw, ref_text, speed, dur)
                plan["seed"] = seed
                mine.append(len(segs))
                segs.append(plan)
            owners.append(mine)

        ref_len = 0 if ref_mel_raw is None else int(ref_mel_raw.shape[-1])
        ref_rows = None if ref_mel_raw is None else ref_mel_raw.transpose(0, 1)  # [T_ref, n_mels]
        vocos = self._get_vocos(device)
        waves: list[torch.Tensor | None] = [None] * len(segs)
        frames = [s["T_total"] for s in segs]
        for batch in plan_batches(frames, max_rows=max_rows_per_batch):
            tmax = max(frames[i] for i in batch)
            B = len(batch)
            ids = torch.full((B, tmax), -1, dtype=torch.long)
            for j, i in enumerate(batch):
                ids[j, : frames[i]] = torch.tensor(segs[i]["full_ids"], dtype=torch.long)
            cond = torch.zeros(B, tmax, self.n_mels, device=device)
            if ref_rows is not None:
                cond[:, :ref_len] = ref_rows
            y0 = torch.zeros(B, tmax, self.n_mels, device=device)
            for j, i in enumerate(batch):  # the B = 1 noise stream of every segment (flow.py:270-283)
                g = None
                if segs[i]["seed"] is not None:
                    g = torch.Generator(device=device).manual_seed(segs[i]["seed"])
                y0[j, : frames[i]] = torch.randn(frames[i], self.n_mels, device=device, generator=g)
            # ids / duration / lens stay on the host: CFM.sample validates them without a device read-back (no host sync)
            mel, _ = self.cfm.sample(cond=cond, text_ids=ids,
                                     duration=torch.tensor([frames[i] for i in batch], dtype=torch.long),
                                     lens=torch.full((B,), ref_len, dtype=torch.long), steps=n_steps,
                                     cfg_strength=cfg_strength, sway_sampling_coef=sway_sampling_coef, y0=y0)
            for j, i in enumerate(batch):
                target_mel = mel[j:j + 1, ref_len:frames[i], :].transpose(1, 2)
                waves[i] = vocos.decode(target_mel).squeeze(0)
        host = [w.cpu() for w in waves]
        return [host[m[0]] if len(m) == 1 else _concat_with_pause([host[i] for i in m], self.sample_rate, pause_s)
                for m in owners]

    def prepare_segment(self, text: str, lang: str, ref_mel: torch.Tensor | None, ref_text: str | None, speed: float,
                        target_duration_s: float | None) -> dict:
        """Integer side of one segment: ids, frame counts, stretched text (all host-side, bit-exact contract)."""
        target_ids = self._text_cleaner.text_to_sequence(text, lang=lang)
        ref_len = 0 if ref_mel is None else int(ref_mel.shape[-1])
        ref_ids: list[int] = []
        if ref_mel is not None and ref_text is not None:
            ref_ids = self._text_cleaner.text_to_sequence(ref_text, lang=lang)
        target_len = estimate_target_len(target_duration_s=target_duration_s, sample_rate=self.sample_rate,
                                         hop_length=self.hop_length, ref_len=ref_len, n_ref_ids=len(ref_ids),
                                         n_target_ids=len(target_ids), text=text, speed=speed)
        total = ref_len + target_len
        if ref_len > 0:
            full = _stretch_text_to_len(ref_ids, ref_len) + _stretch_text_to_len(target_ids, target_len)
        else:
            full = _stretch_text_to_len(target_ids, total)
        return dict(target_ids=target_ids, ref_ids=ref_ids, ref_len=ref_len, target_len=target_len, T_total=total,
                    full_ids=full)

    def _synthesize_segment(self, text: str, lang: str, ref_audio_path: str | Path | None, ref_text: str | None,
                            n_steps: int, cfg_strength: float, sway_sampling_coef: float | None, speed: float,
                            target_duration_s: float | None, seed: int | None, device: str) -> torch.Tensor:
        ap = self._audio_processor
        ref_mel_raw = None
        if ref_audio_path is not None:
            if not ref_text:
                _log.warning("ref_audio_path was provided without ref_text; duration will fall back to the ref-free "
                             "estimate and the reference region will use filler text.")
            if isinstance(ref_audio_path, torch.Tensor):
                # extension over the reference: an in-memory 24 kHz mono waveform (e.g. pinned host memory)
                wav = ref_audio_path.reshape(-1)
                wav = ap.normalize_audio(wav.to(device, non_blocking=True))
            else:
                wav, _ = ap.load_audio(ref_audio_path)
                wav = ap.normalize_audio(wav).to(device)
            ref_mel_raw = ap.mel_spectrogram(wav)  # [n_mels, T_ref]
        plan = self.prepare_segment(text, lang, ref_mel_raw, ref_text, speed, target_duration_s)
        ref_len, total = plan["ref_len"], plan["T_total"]

        ids = torch.tensor([plan["full_ids"]], dtype=torch.long)  # host side: validated without a device read-back
        if ref_mel_raw is not None:
            cond = F.pad(ref_mel_raw.unsqueeze(0).transpose(1, 2), (0, 0, 0, total - ref_len), value=0.0)
        else:
            cond = torch.zeros(1, total, self.n_mels, device=device)
        mel, _ = self.cfm.sample(cond=cond, text_ids=ids,
                                 duration=torch.tensor([total], dtype=torch.long),
                                 lens=torch.tensor([ref_len], dtype=torch.long),
                                 steps=n_steps, cfg_strength=cfg_strength, sway_sampling_coef=sway_sampling_coef, seed=seed)
        target_mel = mel[:, ref_len:, :].transpose(1, 2)  # [1, n_mels, target_len]
        return self._get_vocos(device).decode(target_mel).squeeze(0).cpu()

    @classmethod
    def from_config(cls, config: dict[str, Any]) -> "F5TTS":
        m = config.get("model", {})
        dim, heads = m.get("dim", 1024), m.get("heads", 16)
        return cls(
            n_mels=config.get("n_mels", 100), vocab_size=m.get("vocab_size", 65), dim=dim, depth=m.get("depth", 22),
            heads=heads, dim_head=dim // heads, ff_mult=m.get("ff_mult", 4), text_dim=m.get("text_dim", 512),
            conv_layers=m.get("conv_layers", 4), p_dropout=m.get("p_dropout", 0.1),
            audio_drop_prob=m.get("audio_drop_prob", 0.3), cond_drop_prob=m.get("cond_drop_prob", 0.2),
            frac_lengths_mask=tuple(m.get("frac_lengths_mask", [0.7, 1.0])), sample_rate=config.get("sample_rate", 24000),
            n_fft=config.get("n_fft", 1024), hop_length=config.get("hop_length", 256),
            gradient_checkpointing=config.get("gradient_checkpointing", False),
        )
