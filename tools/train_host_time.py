"""Host time of one config-5 TrainEngine.train_step (no synchronisation inside the timed calls) next to the GPU time per
step: if the two are close the step is (partly) bound by the Python / ctypes launch path rather than by the kernels."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
import torch  # noqa: E402
import weights as GW  # noqa: E402

from oron_tts_b200.f5tts import F5TTS  # noqa: E402
from oron_tts_b200.train import TrainEngine  # noqa: E402

dev = "cuda"
torch.manual_seed(0)
with torch.device(dev):
    m = F5TTS.from_config(GW.CONFIGS["base"])
    for k, p in m.named_parameters():
        if float(p.detach().abs().max()) == 0.0:
            torch.nn.init.normal_(p, std=0.02)
eng = TrainEngine(m.train())
B, Tn = 8, 1024
g = torch.Generator(device=dev).manual_seed(1)
mel = torch.randn(B, 100, Tn, device=dev, generator=g) * 1.5 - 3.0
text = torch.randint(4, 65, (B, Tn), device=dev, generator=g)
lens = torch.full((B,), Tn, device=dev, dtype=torch.long)
for _ in range(4):
    eng.train_step(mel, text, lens)
torch.cuda.synchronize()
host = []
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    t0 = time.perf_counter()
    eng.train_step(mel, text, lens)
    host.append((time.perf_counter() - t0) * 1e3)
e1.record()
torch.cuda.synchronize()
print("host ms per step (enqueue only):", [round(h, 1) for h in host])
print(f"gpu ms per step: {e0.elapsed_time(e1) / 10:.2f}")
