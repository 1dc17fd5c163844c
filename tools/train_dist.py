"""Config-5 training step under torchrun (data parallel): ms per step with / without the bucketed all-reduce overlap.

  torchrun --nproc-per-node N --master-addr 127.0.0.1 tools/train_dist.py [--no-overlap] [--steps K]
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
import torch
import torch.distributed as dist
import weights as GW

from oron_tts_b200.f5tts import F5TTS
from oron_tts_b200.train import TrainEngine

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
steps = int(sys.argv[sys.argv.index("--steps") + 1]) if "--steps" in sys.argv else 5
torch.manual_seed(0)
with torch.device(dev):
    m = F5TTS.from_config(GW.CONFIGS["base"])
    for p in m.parameters():
        if float(p.detach().abs().max()) == 0.0:
            torch.nn.init.normal_(p, std=0.02)
eng = TrainEngine(m.train())
eng.reducer.overlap = "--no-overlap" not in sys.argv
B, Tn = 8, 1024
g = torch.Generator(device=dev).manual_seed(1 + rank)
mel = torch.randn(B, 100, Tn, device=dev, generator=g) * 1.5 - 3.0
text = torch.randint(4, 65, (B, Tn), device=dev, generator=g)
lens = torch.full((B,), Tn, device=dev, dtype=torch.long)
for _ in range(2):
    eng.train_step(mel, text, lens)
if world > 1:
    dist.barrier()
torch.cuda.synchronize()
e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
t_bwd = 0.0
e0.record()
for _ in range(steps):
    s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s0.record()
    eng.loss_and_grad(mel, text, lens)
    s1.record()
    eng.optimizer_step()
    torch.cuda.synchronize()
    t_bwd += s0.elapsed_time(s1)
e1.record()
torch.cuda.synchronize()
ms = torch.tensor([e0.elapsed_time(e1) / steps, t_bwd / steps], device=dev)
if world > 1:
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
if rank == 0:
    print(f"world {world} overlap={eng.reducer.overlap} NCCL_MAX_CTAS={os.environ.get('NCCL_MAX_CTAS')}: {float(ms[0]):.2f} ms per step "
          f"(forward + backward incl. issued all-reduces: {float(ms[1]):.2f} ms; reduce tail + optimizer: {float(ms[0] - ms[1]):.2f} ms)")
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
