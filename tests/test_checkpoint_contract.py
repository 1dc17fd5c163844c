"""The reference's own checkpoint tests (tests/test_checkpoint.py:9-106 of btseee/oron-tts), restated against the
``src.`` import-path shim: same tiny config (dim 64, heads 2 -> dim_head 32), same assertions. They pin the drop-in
boundary of SURVEY section 8(b): ``CheckpointManager.load`` accepts torch.compile-style ``_orig_mod`` keys,
``load_pretrained_f5tts`` skips shape-incompatible tensors and prefers the EMA weights, and the rotation helper keeps
local step files. CPU only: nothing here launches a kernel."""
from pathlib import Path

import torch

from src.models.f5tts import F5TTS
from src.utils.checkpoint import CheckpointManager, stale_remote_checkpoint_paths


def _tiny_config():
    return {"sample_rate": 24000, "n_fft": 1024, "hop_length": 256, "n_mels": 100,
            "model": {"vocab_size": 65, "dim": 64, "depth": 1, "heads": 2, "ff_mult": 2, "text_dim": 32, "conv_layers": 1}}


def _compiled_keys(model):
    return {(k.replace("cfm.backbone.", "cfm.backbone._orig_mod.", 1) if k.startswith("cfm.backbone.") else k): v.clone()
            for k, v in model.state_dict().items()}


def test_load_accepts_compiled_backbone_checkpoint(tmp_path: Path):
    source, target = F5TTS.from_config(_tiny_config()), F5TTS.from_config(_tiny_config())
    for p in source.parameters():  # default init leaves whole families at zero: make every tensor distinctive
        torch.nn.init.normal_(p, std=0.1)
    path = tmp_path / "compiled.pt"
    torch.save({"step": 12, "loss": 0.25, "model_state_dict": _compiled_keys(source)}, path)
    info = CheckpointManager(tmp_path).load(target, path=path)
    assert info["step"] == 12 and info["loss"] == 0.25
    for key, value in source.state_dict().items():
        assert torch.equal(target.state_dict()[key], value), key


def test_load_pretrained_skips_incompatible_shapes(tmp_path: Path):
    model = F5TTS.from_config(_tiny_config())
    state = {k: v.clone() for k, v in model.state_dict().items()}
    state["cfm.backbone.text_embed.text_embed.weight"] = torch.randn(10, 32)
    path = tmp_path / "pretrained.pt"
    torch.save({"model_state_dict": state}, path)
    result = CheckpointManager(tmp_path).load_pretrained_f5tts(model, path, strict=False)
    assert result["skipped_keys"] == ["cfm.backbone.text_embed.text_embed.weight"]
    assert result["unexpected_keys"] == []


def test_load_pretrained_prefers_ema_state_dict(tmp_path: Path):
    source, target = F5TTS.from_config(_tiny_config()), F5TTS.from_config(_tiny_config())
    raw = {k: v.clone() for k, v in source.state_dict().items()}
    ema = {k: v.clone() for k, v in source.state_dict().items()}
    key = next(k for k, v in ema.items() if v.is_floating_point())
    ema[key] = torch.full_like(ema[key], 0.25)
    raw[key] = torch.full_like(raw[key], -0.25)
    path = tmp_path / "oron_best.pt"
    torch.save({"model_state_dict": raw, "ema_state_dict": ema}, path)
    CheckpointManager(tmp_path).load_pretrained_f5tts(target, path, strict=False)
    assert torch.equal(target.state_dict()[key], ema[key])


def test_load_pretrained_strict_and_safetensors(tmp_path: Path):
    from safetensors.torch import save_file

    source, target = F5TTS.from_config(_tiny_config()), F5TTS.from_config(_tiny_config())
    for p in source.parameters():
        torch.nn.init.normal_(p, std=0.1)
    path = tmp_path / "model.safetensors"
    save_file({k: v.contiguous() for k, v in _compiled_keys(source).items()}, str(path))
    result = CheckpointManager(tmp_path).load_pretrained_f5tts(target, path, strict=True)
    assert result == {"missing_keys": [], "unexpected_keys": [], "skipped_keys": []}
    for key, value in source.state_dict().items():
        assert torch.equal(target.state_dict()[key], value), key


def test_stale_remote_checkpoint_paths_keeps_local_rotation():
    remote = ["README.md", "config.json", "f5tts_best.pt", "f5tts_step_00000010.pt", "f5tts_step_00000020.pt",
              "f5tts_step_00000030.pt", "tb_logs/events.out.tfevents.test"]
    local = ["f5tts_step_00000020.pt", "f5tts_step_00000030.pt"]
    assert stale_remote_checkpoint_paths(remote, local, "f5tts") == ["f5tts_step_00000010.pt"]
