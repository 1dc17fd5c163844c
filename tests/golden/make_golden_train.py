"""Golden fixture for the OT-CFM training step, generated from the LIVE reference.

Run in the build container only:  python tests/golden/make_golden_train.py
Output: train_tiny.pt — tiny F5TTS (weights.py), eval-mode (deterministic: t = 0.5, centred span, seed-0 noise,
no dropout; flow.py:113-128, 136-138) loss of one batch, its gradient w.r.t. every parameter (autograd through the
reference), the pre-clip global gradient norm as F5Trainer._grad_norm computes it (trainer.py:171-177), and the
parameters after two optimizer steps of the reference recipe: clip_grad_norm_(1.0) + AdamW(lr 1e-4 x LinearLR
warm-up factor, betas (0.9, 0.999), weight_decay 0.01) (trainer.py:76-96, 191-216).
"""

from __future__ import annotations

import importlib.util
import os
import sys
import tempfile
import types

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference"
_stub = types.ModuleType("soundfile")
_stub.write = lambda *a, **k: None
sys.modules["soundfile"] = _stub
sys.path = [REF] + [p for p in sys.path if os.path.abspath(p or ".") != os.path.dirname(os.path.dirname(HERE))]
os.chdir(tempfile.gettempdir())

import torch  # noqa: E402

from src.models.f5tts import F5TTS  # noqa: E402

assert os.path.realpath(sys.modules["src"].__path__[0]).startswith(REF)
_spec = importlib.util.spec_from_file_location("golden_weights", os.path.join(HERE, "weights.py"))
W = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(W)


def main() -> None:
    model = F5TTS.from_config(W.CONFIGS["tiny"]).eval()
    model.load_state_dict(W.fill_state_dict(model.state_dict(), W.SEEDS["tiny"]), strict=True)
    gen = torch.Generator().manual_seed(17)
    B, T = 2, 150
    lens = torch.tensor([150, 97])
    mel = torch.randn(B, 100, T, generator=gen) * 1.5 - 3.0
    text = torch.randint(4, 65, (B, T), generator=gen)
    text[0, 120:] = -1
    text[1, 97:] = -1
    text[1, 30:35] = -1
    out = dict(mel=mel, text=text, lens=lens)
    params = dict(model.named_parameters())
    opt = torch.optim.AdamW(model.parameters(), lr=1e-4, betas=(0.9, 0.999), weight_decay=0.01)
    sched = torch.optim.lr_scheduler.LinearLR(opt, start_factor=0.5, end_factor=1.0, total_iters=4)
    for step in range(2):
        opt.zero_grad(set_to_none=True)
        loss = model(mel, text, lens)  # eval mode: deterministic objective
        loss.backward()
        gn = torch.norm(torch.stack([torch.norm(p.grad.detach()) for p in model.parameters() if p.grad is not None]))
        if step == 0:
            out["loss"] = loss.detach().clone()
            out["grads"] = {k: p.grad.detach().clone() for k, p in params.items() if p.grad is not None}
            out["no_grad"] = [k for k, p in params.items() if p.grad is None]
            out["grad_norm"] = gn.clone()
        out[f"loss_step{step}"] = loss.detach().clone()
        out[f"lr_step{step}"] = opt.param_groups[0]["lr"]
        torch.nn.utils.clip_grad_norm_(model.parameters(), 1.0)
        opt.step()
        sched.step()
    out["params_after_2_steps"] = {k: p.detach().clone() for k, p in params.items()}
    torch.save(out, os.path.join(HERE, "train_tiny.pt"))
    print("loss", float(out["loss"]), "grad_norm", float(out["grad_norm"]), "tensors", len(out["grads"]),
          "no_grad", out["no_grad"], "lrs", out["lr_step0"], out["lr_step1"])


if __name__ == "__main__":
    main()
