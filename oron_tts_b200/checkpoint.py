"""Checkpoint I/O contract used by scripts/infer.py (src/utils/checkpoint.py:39-59, 62-151, 214-228):
file names, dict keys, config.json sibling and the ``._orig_mod.`` key adaptation for backbones that were
wrapped by torch.compile. The Hugging Face upload/download half of the reference class is control plane
and out of scope here."""

from __future__ import annotations

import json
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

    def load_config(self) -> dict[str, Any] | None:
        p = self.checkpoint_dir / "config.json"
        return json.loads(p.read_text()) if p.exists() else None
