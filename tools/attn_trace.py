"""Per-CTA clock64 timeline of the attention kernel at config-2 shapes."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from oron_tts_b200 import _lib as L
DEV = "cuda"
R, T = 2816, 1408
g = torch.Generator(device=DEV).manual_seed(1)
qkv = torch.randn(R, 3072, device=DEV, generator=g).bfloat16()
qkv[:, 2048:] = torch.randn(R, 1024, device=DEV, generator=g).half().view(torch.bfloat16)
o = torch.zeros(R, 1024, device=DEV, dtype=torch.bfloat16)
lens = torch.tensor([1406, 1406], device=DEV, dtype=torch.int32)
AWS = L.attention_workspace(2, T, 16, DEV)
fn = lambda: L.attention(qkv, o, nbatch=2, rows_per_batch=T, heads=16, seq_lens=lens, scale=0.125, workspace=AWS)
for _ in range(3): fn()
torch.cuda.synchronize()
dbg = torch.zeros(1024, 16, device=DEV, dtype=torch.int64)
L.lib().oron_debug_set_attention_stamps(dbg.data_ptr())
torch.cuda._sleep(200000)
fn(); torch.cuda.synchronize()
L.lib().oron_debug_set_attention_stamps(None)
d = dbg.cpu()
names = {1: "t2 s_full", 2: "t2 pass1", 3: "t2 o_wait", 4: "t2 pass2", 5: "t2 arrive", 7: "t3 s_full", 8: "t3 pass1", 9: "t3 o_wait", 10: "t3 pass2",
         11: "t3 arrive", 6: "t3 o_full(2) seen", 12: "mma: p_full(2)", 13: "mma: PV(2) issued", 14: "softmax end", 15: "cta end"}
starts = d[:, 0]
print("global start spread (cycles are per-SM clocks; only relative values inside a CTA are meaningful)")
for cta in (0, 100, 295, 296, 400, 575):
    base = int(d[cta, 0])
    print(f"  cta {cta}: " + ", ".join(f"{names[i]}={int(d[cta, i]) - base}" for i in sorted(names) if int(d[cta, i]) != 0))
d = d[:576]
dur = (d[:, 15] - d[:, 0]).float()
print("cta duration cycles: mean %.0f min %.0f max %.0f" % (dur.mean(), dur.min(), dur.max()))
