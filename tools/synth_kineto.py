"""Per-kernel GPU time of ONE end-to-end synthesize() call of the bench (config 2) with torch.profiler (kineto): what runs outside
the 32 CUDA-graph replays of the ODE step (text / time embedding, modulation table, log-mel of the reference, Vocos decode)?
  python tools/synth_kineto.py"""
import os
import sys
from collections import defaultdict

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench as B  # noqa: E402

dev = torch.device("cuda", 0)
model, voc, ref_mel, ids, ref_wav = B.build_workload(dev, seed=100)


def step():
    return model.synthesize(B.BENCH_TEXT, lang="mn", ref_audio_path=ref_wav, ref_text=B.BENCH_REF_TEXT, n_steps=B.STEPS_NFE,
                            cfg_strength=B.CFG, sway_sampling_coef=B.SWAY, target_duration_s=10.0, seed=None, device=str(dev))


for _ in range(3):
    step()
torch.cuda.synchronize()
from torch.profiler import ProfilerActivity, profile  # noqa: E402

with profile(activities=[ProfilerActivity.CUDA]) as prof:
    step()
    torch.cuda.synchronize()
ev = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
tot, cnt = defaultdict(float), defaultdict(int)
for e in ev:
    tot[e.name[:90]] += e.device_time if hasattr(e, "device_time") else e.cuda_time
    cnt[e.name[:90]] += 1
t0 = min(e.time_range.start for e in ev)
t1 = max(e.time_range.end for e in ev)
busy = sum(tot.values())
print(f"GPU span of one call: {(t1 - t0) / 1e3:.2f} ms; sum of kernel / memcpy durations: {busy / 1e3:.2f} ms; {len(ev)} device events")
for k, v in sorted(tot.items(), key=lambda kv: -kv[1])[:40]:
    print(f"{k:90s} {cnt[k]:6d} {v / 1e3:9.3f} ms")
