"""Drop-in test of the UNMODIFIED reference script scripts/infer.py: the copy under
baseline/_ref/scripts/ (tools/make_baseline_ref.py; git-ignored, shipped with the gpurun snapshot) is executed with runpy
against THIS repository's `src` import-path shim, a saved checkpoint + config.json written by CheckpointManager, and a
Vocos `pytorch_model.bin` found through the Hugging Face cache layout -- the code paths a user of the reference takes:
`load_model` (EMA weights preferred, `_orig_mod` adaptation, strict=False), `F5TTS.from_config`, `model.to(device)`,
`synthesize`, `Vocos.from_pretrained("charactr/vocos-mel-24khz")`, `soundfile.write`. Nothing here reads /root/reference."""
import os
import runpy
import sys
import wave

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
INFER = os.path.join(ROOT, "baseline", "_ref", "scripts", "infer.py")
STUBS = os.path.join(ROOT, "baseline", "_ref", "_stubs")

pytestmark = pytest.mark.gpu

CONFIG = {"sample_rate": 24000, "n_fft": 1024, "hop_length": 256, "n_mels": 100,
          "model": {"vocab_size": 65, "dim": 128, "depth": 2, "heads": 2, "ff_mult": 2, "text_dim": 64, "conv_layers": 1}}


def _hf_cache_with_vocos(cache_dir):
    """A Hugging Face hub cache holding charactr/vocos-mel-24khz/pytorch_model.bin with synthetic weights (the real
    ones are not reachable offline), including the feature_extractor.* tensors the upstream file carries."""
    from oron_tts_b200.vocos import Vocos

    torch.manual_seed(7)
    voc = Vocos()
    sd = {k: v.clone() for k, v in voc.state_dict().items()}
    sd["feature_extractor.mel_spec.spectrogram.window"] = torch.hann_window(1024)
    sd["feature_extractor.mel_spec.mel_scale.fb"] = torch.zeros(513, 100)
    rev = "0" * 40
    repo = os.path.join(cache_dir, "models--charactr--vocos-mel-24khz")
    os.makedirs(os.path.join(repo, "snapshots", rev))
    os.makedirs(os.path.join(repo, "refs"))
    with open(os.path.join(repo, "refs", "main"), "w") as f:
        f.write(rev)
    torch.save(sd, os.path.join(repo, "snapshots", rev, "pytorch_model.bin"))
    return sd


@pytest.mark.skipif(not os.path.exists(INFER), reason="baseline/_ref/scripts/infer.py missing (run tools/make_baseline_ref.py where /root/reference exists)")
def test_unmodified_infer_script_runs_against_the_shim(tmp_path, monkeypatch, capsys):
    if ROOT not in sys.path:
        sys.path.insert(0, ROOT)
    from src.models.f5tts import F5TTS
    from src.utils.checkpoint import CheckpointManager

    # 1. a training-style checkpoint: raw weights + EMA weights (distinct), compiled-backbone key style, config.json
    torch.manual_seed(0)
    model = F5TTS.from_config(CONFIG)
    for p in model.parameters():
        torch.nn.init.normal_(p, std=0.05)
    ema = {k.replace("cfm.backbone.", "cfm.backbone._orig_mod.", 1): v.clone() for k, v in model.state_dict().items()}
    ckpt_dir = tmp_path / "ckpt"
    cm = CheckpointManager(ckpt_dir)
    opt = torch.optim.AdamW(model.parameters(), lr=1e-4)
    cm.save(step=10, model=model, optimizer=opt, ema_state=ema, loss=0.5, config=CONFIG, is_best=True)
    assert (ckpt_dir / "f5tts_best.pt").exists() and (ckpt_dir / "config.json").exists()

    # 2. the vocoder weights where `Vocos.from_pretrained("charactr/vocos-mel-24khz")` looks for them
    cache = tmp_path / "hf"
    os.makedirs(cache)
    _hf_cache_with_vocos(str(cache))
    monkeypatch.setenv("HF_HUB_CACHE", str(cache))
    monkeypatch.setenv("HF_HUB_OFFLINE", "1")
    import huggingface_hub.constants as hc

    monkeypatch.setattr(hc, "HF_HUB_CACHE", str(cache), raising=False)
    monkeypatch.setattr(hc, "HF_HUB_OFFLINE", True, raising=False)

    # 3. run the script exactly as a user would
    out = tmp_path / "out" / "hello.wav"
    monkeypatch.syspath_prepend(STUBS)   # `import soundfile` (not installed in this image)
    monkeypatch.setattr(sys, "argv", ["infer.py", "--checkpoint", str(ckpt_dir / "f5tts_best.pt"), "--text", "Сайн байна уу",
                                       "--lang", "mn", "--output", str(out), "--steps", "4", "--seed", "0", "--device", "cuda"])
    runpy.run_path(INFER, run_name="__main__")
    printed = capsys.readouterr().out
    assert "Loading EMA weights" in printed and "Saved:" in printed

    # 4. the waveform: 11 characters -> max(50, int(11 * 13 / 1.0)) = 143 frames -> (143 - 1) * 256 samples (f5tts.py:374-375, 413-416)
    with wave.open(str(out), "rb") as w:
        n, sr = w.getnframes(), w.getframerate()
        pcm = np.frombuffer(w.readframes(n), dtype="<i2")
    assert sr == 24000 and n == (143 - 1) * 256
    assert np.isfinite(pcm.astype(np.float32)).all() and np.abs(pcm).max() > 0
