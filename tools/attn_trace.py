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
WS = "--ws" in sys.argv  # key-split tail (needs the workspace); default = production path, one CTA per item
AWS = L.attention_workspace(2, T, 16, DEV, seq_lens=lens) if WS else None
fn = lambda: L.attention(qkv, o, nbatch=2, rows_per_batch=T, heads=16, seq_lens=lens, scale=0.125, workspace=AWS)
for _ in range(3): fn()
torch.cuda.synchronize()
dbg = torch.zeros(1024, 16, device=DEV, dtype=torch.int64)
L.lib().oron_debug_set_attention_stamps(dbg.data_ptr())
if "--hot" in sys.argv:  # sustained load first: the traced launch then runs at the clocks of a long run
    L.lib().oron_debug_set_attention_stamps(None)
    for _ in range(400): fn()
    L.lib().oron_debug_set_attention_stamps(dbg.data_ptr())
else:
    torch.cuda._sleep(200000)
fn(); torch.cuda.synchronize()
L.lib().oron_debug_set_attention_stamps(None)
d = dbg.cpu()
names = {1: "tj s_full", 2: "tj pass1", 3: "tj o_wait", 4: "tj pass2", 5: "tj arrive", 6: "t+1 s_full", 7: "t+2 s_full", 10: "fence+bar", 11: "count done", 12: "mma: p_full(2)", 13: "mma: PV(2) issued", 14: "softmax end", 15: "cta end"}
starts = d[:, 0]
print("global start spread (cycles are per-SM clocks; only relative values inside a CTA are meaningful)")
for cta in (0, 50, 100, 150, 200, 250, 295, 296, 320, 351):
    base = int(d[cta, 0])
    print(f"  cta {cta}: " + ", ".join(f"{names[i]}={int(d[cta, i]) - base}" for i in sorted(names) if int(d[cta, i]) != 0))
NCTA = 296 if WS else 352
d = d[:NCTA]
dur = (d[:, 15] - d[:, 0]).float()
print("cta duration cycles: mean %.0f min %.0f max %.0f" % (dur.mean(), dur.min(), dur.max()))

g0 = int(d[:, 8].min())
st = (d[:, 8] - g0).double() / 1e3
en = (d[:, 9] - g0).double() / 1e3
if not WS: print("globaltimer (us): full items start %.1f..%.1f end %.1f..%.1f | tail units start %.1f..%.1f end %.1f..%.1f" % (
    st[:296].min(), st[:296].max(), en[:296].min(), en[:296].max(), st[296:].min(), st[296:].max(), en[296:].min(), en[296:].max()))
if not WS: print("tail unit durations (us): mean %.2f max %.2f ; full item durations mean %.2f max %.2f" % (
    (en[296:] - st[296:]).mean(), (en[296:] - st[296:]).max(), (en[:296] - st[:296]).mean(), (en[:296] - st[:296]).max()))

print("kernel span (us): %.1f" % float(en.max() - st.min()))
print("cta spans (us): mean %.2f min %.2f max %.2f" % (float((en - st).mean()), float((en - st).min()), float((en - st).max())))
sys.exit(0)
dd = (en - st)
slow = torch.argsort(dd[296:], descending=True)[:3] + 296
for cta in slow.tolist() + [300]:
    base = int(d[cta, 0])
    print(f"  tail cta {cta} ({dd[cta]:.1f} us): " + ", ".join(f"{names[i]}={int(d[cta, i]) - base}" for i in sorted(names) if int(d[cta, i]) != 0))
