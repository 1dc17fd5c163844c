"""smoke(): one small invocation of the hot path on cuda:0 checked against the CPU oracle.
(The oracle is test infrastructure; this module is only reached from __graft_entry__.smoke().)"""

from __future__ import annotations

import os
import sys

import torch


def run_smoke() -> None:
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, os.path.join(root, "tests", "golden"))
    if root not in sys.path:
        sys.path.insert(0, root)
    import weights as GW
    from oracle import dit_oracle as DO

    from . import _lib as L
    from .f5tts import F5TTS, _stretch_text_to_len
    from .vocos import Vocos

    dev = "cuda:0"
    torch.cuda.set_device(0)
    model = F5TTS.from_config(GW.CONFIGS["tiny"])
    sd = GW.fill_state_dict(model.state_dict(), 1234)
    model.load_state_dict(sd, strict=True)
    model = model.to(dev).eval()
    ids = model._text_cleaner.text_to_sequence("Сайн байна уу", lang="mn")
    T = 143
    full = torch.tensor([_stretch_text_to_len(ids, T)])
    launches0 = L.launch_count()
    o_mel, o_traj = DO.cfm_sample(sd, torch.zeros(1, T, 100), full, torch.tensor([T]), lens=torch.tensor([0]), steps=4,
                                  cfg_strength=2.0, sway_sampling_coef=-1.0, seed=0)
    mel, traj = model.cfm.sample(torch.zeros(1, T, 100, device=dev), full.to(dev), torch.tensor([T], device=dev),
                                 lens=torch.tensor([0], device=dev), steps=4, cfg_strength=2.0, sway_sampling_coef=-1.0,
                                 y0=o_traj[0])
    err = float((mel.cpu() - o_mel).norm() / o_mel.norm())
    voc = Vocos()
    voc.load_state_dict(GW.fill_state_dict(voc.state_dict(), 4321))
    wav = voc.to(dev).eval().decode(mel.transpose(1, 2))
    torch.cuda.synchronize()
    n = L.launch_count() - launches0
    print(f"[smoke] tiny CFM.sample (4 NFE, CFG) rel-L2 vs oracle = {err:.3e}; wav {tuple(wav.shape)}; "
          f"{n} kernel launches issued through liboron_b200.so")
    if not (err < 1e-2) or not torch.isfinite(wav).all():
        raise RuntimeError(f"smoke parity failed: rel-L2 {err}")
