"""Checkpoint I/O contract used by scripts/infer.py and scripts/train.py (src/utils/checkpoint.py:24-59, 62-228):
file names, dict keys, config.json sibling, the ``._orig_mod.`` key adaptation for backbones that were wrapped by
torch.compile, and the shape-filtered pretrained loader (``load_pretrained_f5tts``, pinned by the reference's
tests/test_checkpoint.py:58-86, mirrored in tests/test_checkpoint_contract.py). The Hugging Face upload/download half of
the reference class is control plane and out of scope here."""

from __future__ import annotations

import json
import re
from collections.abc import Mapping
from pathlib import Path
from typing import Any

import torch

_COMPILED = "._orig_mod."


def _plain(key: str) -> str:
    return key.replace(_COMPILED, ".")


def adapt_state_dict_to_model(state_dict: Mapping[str, torch.Tensor], model: torch.nn.Module) -> dict[str, torch.Tensor]:
    """Rename keys so that eager and torch.compile-wrapped (``_orig_mod``) layouts load into either model."""
    wanted = {_plain(k): k for k in model.state_dict()}
    return {wanted.get(_plain(k), k): v for k, v in state_dict.items()}


def _is_step_checkpoint(path: str, model_name: str) -> bool:
    return re.fullmatch(rf"{re.escape(model_name)}_step_\d+\.pt", Path(path).name) is not None


def stale_remote_checkpoint_paths(remote_paths: list[str], local_paths: list[str], model_name: str) -> list[str]:
    """Step checkpoints present remotely but rotated away locally (src/utils/checkpoint.py:24-36)."""
    keep = {Path(p).name for p in local_paths if _is_step_checkpoint(p, model_name)}
    return [p for p in remote_paths if _is_step_checkpoint(p, model_name) and Path(p).name not in keep]


class CheckpointManager:
    def __init__(self, checkpoint_dir: str | Path, model_name: str = "f5tts", max_checkpoints: int = 5) -> None:
        self.checkpoint_dir = Path(checkpoint_dir)
        self.checkpoint_dir.mkdir(parents=True, exist_ok=True)
        self.model_name = model_name
        self.max_checkpoints = max_checkpoints

    def _step_path(self, step: int) -> Path:
        return self.checkpoint_dir / f"{self.model_name}_step_{step:08d}.pt"

    def _best_path(self) -> Path:
        return self.checkpoint_dir / f"{self.model_name}_best.pt"

    def _steps(self) -> list[Path]:
        return sorted(self.checkpoint_dir.glob(f"{self.model_name}_step_*.pt"), key=lambda p: int(p.stem.split("_")[-1]))

    def save(self, step: int, model: torch.nn.Module, optimizer: torch.optim.Optimizer, scheduler=None,
             ema_state: dict[str, Any] | None = None, loss: float | None = None, config: dict[str, Any] | None = None,
             is_best: bool = False, extra_state: dict[str, Any] | None = None) -> Path:
        blob: dict[str, Any] = {
            "step": step,
            "model_state_dict": {_plain(k): v for k, v in model.state_dict().items()},
            "optimizer_state_dict": optimizer.state_dict(),
            "loss": loss,
        }
        if scheduler is not None:
            blob["scheduler_state_dict"] = scheduler.state_dict()
        if ema_state is not None:
            blob["ema_state_dict"] = {_plain(k): v for k, v in ema_state.items()}
        blob.update(extra_state or {})
        path = self._step_path(step)
        torch.save(blob, path)
        if config is not None:
            (self.checkpoint_dir / "config.json").write_text(json.dumps(config, indent=2))
        if is_best:
            torch.save(blob, self._best_path())
        steps = self._steps()
        while len(steps) > self.max_checkpoints:
            steps.pop(0).unlink()
        return path

    def load(self, model: torch.nn.Module, optimizer=None, scheduler=None, path: str | Path | None = None,
             load_best: bool = False, device: str = "cpu") -> dict[str, Any]:
        if path is None:
            steps = self._steps()
            path = self._best_path() if load_best else (steps[-1] if steps else None)
        if path is None or not Path(path).exists():
            return {"step": 0, "loss": None, "ema_state_dict": None}
        blob = torch.load(path, map_location=device, weights_only=False)
        model.load_state_dict(adapt_state_dict_to_model(blob["model_state_dict"], model))
        if optimizer is not None and "optimizer_state_dict" in blob:
            optimizer.load_state_dict(blob["optimizer_state_dict"])
        if scheduler is not None and "scheduler_state_dict" in blob:
            scheduler.load_state_dict(blob["scheduler_state_dict"])
        return {"step": blob.get("step", 0), "loss": blob.get("loss"), "ema_state_dict": blob.get("ema_state_dict"),
                "epoch": blob.get("epoch", 0), "best_val": blob.get("best_val", float("inf"))}

    def load_pretrained_f5tts(self, model: torch.nn.Module, checkpoint_path: str | Path, device: str = "cpu",
                              strict: bool = False) -> dict[str, Any]:
        """Load a pretrained F5-TTS checkpoint (src/utils/checkpoint.py:153-205): ``.safetensors`` or a torch file whose
        ``ema_state_dict`` / ``ema_model_state_dict`` / ``model_state_dict`` entry is preferred in that order; with
        ``strict=False`` tensors whose shape differs from the model's (the extended Cyrillic text embedding) are skipped
        and reported."""
        path = Path(checkpoint_path)
        if path.suffix == ".safetensors":
            try:
                from safetensors.torch import load_file
            except ImportError as exc:
                raise ImportError("Install safetensors: pip install safetensors") from exc
            state = load_file(str(path), device=device)
        else:
            state = torch.load(path, map_location=device, weights_only=True)
            for key in ("ema_state_dict", "ema_model_state_dict", "model_state_dict"):
                if key in state:
                    state = state[key]
                    break
        state = adapt_state_dict_to_model(state, model)
        if strict:
            missing, unexpected = model.load_state_dict(state, strict=True)
            return {"missing_keys": missing, "unexpected_keys": unexpected, "skipped_keys": []}
        have = model.state_dict()
        skipped = [k for k, v in state.items() if k in have and have[k].shape != v.shape]
        missing, unexpected = model.load_state_dict({k: v for k, v in state.items() if k not in skipped}, strict=False)
        return {"missing_keys": missing, "unexpected_keys": unexpected, "skipped_keys": skipped}

    def load_config(self) -> dict[str, Any] | None:
        p = self.checkpoint_dir / "config.json"
        return json.loads(p.read_text()) if p.exists() else None
