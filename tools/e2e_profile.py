"""Host-side profile of the end-to-end synthesize() call of the bench (config 2): cProfile over a few calls, plus the
GPU-idle estimate (e2e wall time per call minus the device-resident time per call)."""
import cProfile
import os
import pstats
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench as B  # noqa: E402

dev = torch.device("cuda", 0)
from oron_tts_b200 import _lib as L  # noqa: E402

L.lib()
model, voc, ref_mel, ids, ref_wav = B.build_workload(dev, seed=100)


def step():
    return model.synthesize(B.BENCH_TEXT, lang="mn", ref_audio_path=ref_wav, ref_text=B.BENCH_REF_TEXT, n_steps=B.STEPS_NFE,
                            cfg_strength=B.CFG, sway_sampling_coef=B.SWAY, target_duration_s=10.0, seed=None, device=str(dev))


for _ in range(3):
    step()
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(5):
    step()
torch.cuda.synchronize()
print(f"e2e wall per call: {(time.perf_counter() - t0) / 5 * 1e3:.2f} ms")
pr = cProfile.Profile()
pr.enable()
for _ in range(5):
    step()
pr.disable()
st = pstats.Stats(pr)
st.sort_stats("cumulative").print_stats(45)
