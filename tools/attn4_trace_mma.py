"""Per-CTA clock64 timeline of the attention kernel (attn_fwd4.cuh) at config-2 shapes."""
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
PER_ITEM = "--per-item" in sys.argv
AWS = None if PER_ITEM else L.attention_workspace(2, T, 16, DEV, seq_lens=lens)
fn = lambda: L.attention(qkv, o, nbatch=2, rows_per_batch=T, heads=16, seq_lens=lens, scale=0.125, workspace=AWS)
for _ in range(3): fn()
torch.cuda.synchronize()
dbg = torch.zeros(1024, 16, device=DEV, dtype=torch.int64)
for _ in range(200): fn()
L.lib().oron_debug_set_attention_stamps(dbg.data_ptr())
fn(); torch.cuda.synchronize()
L.lib().oron_debug_set_attention_stamps(None)
d = dbg.cpu()
names = {1: "mma:wait s_free", 2: "mma:wait q/k_full", 3: "mma:wait p_full", 4: "mma:wait v_full", 5: "mma:issue S (4)", 6: "mma:commits S", 7: "mma:issue PV (8)", 10: "mma:commits PV", 11: "mma:misc", 15: "cta end"}
NCTA = 352 if PER_ITEM else 296
d = d[:NCTA]
dur = (d[:, 15] - d[:, 0]).float()
print("cta duration cycles: mean %.0f min %.0f max %.0f" % (dur.mean(), dur.min(), dur.max()))
for i in (1, 2, 3, 4, 5, 6, 7, 10, 11):
    print("  mean %-20s %10.0f" % (names[i], d[:, i].float().mean()))
