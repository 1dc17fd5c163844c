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
LITE = os.environ.get("ORON_ATT_TRACE", "") .startswith("l")
names = {13: "tiles", 1: "sm:wait s_full", 2: "sm:wait P free", 3: "sm:epilogues", 4: "sm:S load", 5: "sm:max", 6: "sm:exp", 7: "sm:P store", 10: "sm:loop misc", 11: "mma:wait p_full+v_full", 12: "sm:loop span", 14: "softmax end", 15: "cta end"}
NCTA = 352 if PER_ITEM else 296
for cta in (() if LITE else (0, 1, 50, 100, 150, 200, 250, 295)):
    base = int(d[cta, 0])
    print(f"  cta {cta}: " + ", ".join(f"{names[i]}={int(d[cta, i]) - (base if i in (14, 15) else 0)}" for i in sorted(names)))
d = d[:NCTA]
dur = (d[:, 15] - d[:, 0]).float()
print("cta duration cycles: mean %.0f min %.0f max %.0f" % (dur.mean(), dur.min(), dur.max()))
for i in (13, 10, 1, 4, 5, 6, 2, 7, 3, 12, 11):
    print("  mean %-20s %10.0f" % (names[i], d[:, i].float().mean()))
g0 = int(d[:, 8].min())
st = (d[:, 8] - g0).double() / 1e3
en = (d[:, 9] - g0).double() / 1e3
print("kernel span (us): %.1f; cta start %.1f..%.1f end %.1f..%.1f" % (float(en.max() - st.min()), st.min(), st.max(), en.min(), en.max()))
ghz = ((d[:, 15] - d[:, 0]).double() / (d[:, 9] - d[:, 8]).double())
print("SM clock during the traced launch (clock64 / globaltimer): mean %.3f GHz min %.3f max %.3f" % (float(ghz.mean()), float(ghz.min()), float(ghz.max())))
if LITE:
    import torch
    dur_us = (en - st)
    order = torch.argsort(dur_us)
    print("cta duration (us): min %.1f p25 %.1f median %.1f p75 %.1f max %.1f" % tuple(float(torch.quantile(dur_us, q)) for q in (0, .25, .5, .75, 1)))
    lo, hi = dur_us[:148], dur_us[148:296]
    print("mean duration of blocks 0..147: %.1f us, of blocks 148..295: %.1f us; correlation of the two halves (same SM if round-robin): %.2f" % (
        float(lo.mean()), float(hi.mean()), float(torch.corrcoef(torch.stack([lo, hi]))[0, 1])))
    print("slowest ctas:", [(int(i), round(float(dur_us[i]), 1)) for i in order[-12:].tolist()])
    print("fastest ctas:", [(int(i), round(float(dur_us[i]), 1)) for i in order[:12].tolist()])
