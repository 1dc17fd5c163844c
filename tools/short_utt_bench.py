"""ms per NFE of CFM.sample (32 NFE, CFG 2.0, batch 1) for short utterances, where the ODE step is launch-bound:
Small and Base DiT at T = 143 / 400 / 800 frames (1.5 / 4.3 / 8.5 s of audio). Used to compare launch-count
reductions (ORON_FFN_FUSED=1 vs 0):   python tools/short_utt_bench.py"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
import torch
import weights as GW

from oron_tts_b200.f5tts import F5TTS

dev = torch.device("cuda")
for name in ("small", "base"):
    m = F5TTS.from_config(GW.CONFIGS[name])
    m.load_state_dict(GW.fill_state_dict(m.state_dict(), GW.SEEDS[name]), strict=True)
    m = m.to(dev).eval()
    for T in (143, 400, 800):
        ids = torch.randint(4, 65, (1, T), device=dev)
        cond, lens = torch.zeros(1, T, 100, device=dev), torch.tensor([0], device=dev)
        fn = lambda: m.cfm.sample(cond, ids, T, lens=lens, steps=32, cfg_strength=2.0, sway_sampling_coef=-1.0)
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            fn()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        print(f"{name:5s} T={T:4d}: {ms:7.2f} ms per utterance, {ms / 32 * 1e3:7.1f} us per NFE, {(T - 1) * 256 / 24000 / (ms / 1e3):6.1f} audio-s/s "
              f"(ORON_FFN_FUSED={os.environ.get('ORON_FFN_FUSED', '0')})", flush=True)
    del m
    torch.cuda.empty_cache()
