"""Config-5 training step driver for profiling: Base DiT, per-GPU batch 8 x 1024 frames, random-init weights.

  python tools/train_profile.py [--steps N] [--small]          # prints ms per step (CUDA events)
  ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/train_launches.csv \
      python tools/train_profile.py --steps 1 --warm 1          # per-launch list; tools/train_profile.py --summarise <csv>
"""
import csv
import os
import sys
from collections import defaultdict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))


def summarise(path):
    rows = list(csv.reader(l for l in open(path, errors="replace") if l.startswith('"')))
    hdr = rows[0]
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    tot, cnt = defaultdict(float), defaultdict(int)
    for r in rows[1:]:
        if len(r) <= vi:
            continue
        v = float(r[vi].replace(",", ""))
        v = v / 1e3 if r[ui] in ("ns", "nsecond") else (v * 1e3 if r[ui] in ("ms", "msecond") else v)  # -> us
        name = r[ki].split("(")[0]
        tot[name] += v
        cnt[name] += 1
    total = sum(tot.values())
    print(f"total kernel time {total / 1e3:.2f} ms over {sum(cnt.values())} launches")
    for k, v in sorted(tot.items(), key=lambda kv: -kv[1]):
        print(f"{v / 1e3:9.3f} ms {100 * v / total:5.1f} %  x{cnt[k]:5d}  avg {v / cnt[k]:8.1f} us  {k}")


if __name__ == "__main__":
    if "--summarise" in sys.argv:
        summarise(sys.argv[sys.argv.index("--summarise") + 1])
        sys.exit(0)
    import torch
    import weights as GW

    from oron_tts_b200.f5tts import F5TTS
    from oron_tts_b200.train import TrainEngine

    steps = int(sys.argv[sys.argv.index("--steps") + 1]) if "--steps" in sys.argv else 3
    warm = int(sys.argv[sys.argv.index("--warm") + 1]) if "--warm" in sys.argv else 2
    name = "small" if "--small" in sys.argv else "base"
    dev = "cuda"
    torch.manual_seed(0)
    with torch.device(dev):
        m = F5TTS.from_config(GW.CONFIGS[name])
        for k, p in m.named_parameters():  # zero-initialised families would make the step degenerate
            if float(p.detach().abs().max()) == 0.0:
                torch.nn.init.normal_(p, std=0.02)
    eng = TrainEngine(m.train())
    B, Tn = 8, 1024
    g = torch.Generator(device=dev).manual_seed(1)
    mel = torch.randn(B, 100, Tn, device=dev, generator=g) * 1.5 - 3.0
    text = torch.randint(4, 65, (B, Tn), device=dev, generator=g)
    lens = torch.full((B,), Tn, device=dev, dtype=torch.long)
    for _ in range(warm):
        eng.train_step(mel, text, lens)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        loss = eng.train_step(mel, text, lens)
    e1.record()
    torch.cuda.synchronize()
    print(f"{name}: {e0.elapsed_time(e1) / steps:.2f} ms per training step, loss {float(loss):.4f}")
    if "--kineto" in sys.argv:  # in-situ kernel durations (CUPTI activity records: no replay, warm caches, real clocks)
        from torch.profiler import ProfilerActivity, profile
        with profile(activities=[ProfilerActivity.CUDA]) as prof:
            for _ in range(2):
                eng.train_step(mel, text, lens)
            torch.cuda.synchronize()
        tot, cnt = defaultdict(float), defaultdict(int)
        for ev in prof.events():
            if ev.device_type == torch.autograd.DeviceType.CUDA:
                tot[ev.name.split("(")[0]] += ev.device_time
                cnt[ev.name.split("(")[0]] += 1
        total = sum(tot.values())
        print(f"kineto: {total / 2e3:.2f} ms of kernel time per step")
        for k, v in sorted(tot.items(), key=lambda kv: -kv[1])[:24]:
            print(f"{v / 2e3:9.3f} ms {100 * v / total:5.1f} %  x{cnt[k] // 2:5d}  avg {v / cnt[k]:8.1f} us  {k[:90]}")
