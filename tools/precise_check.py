"""fp32-mode accuracy check: PreciseDiT vs the live-reference fixture (tiny) and vs the CPU oracle (Small)."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
import weights as GW
from oracle import dit_oracle as DO
from oron_tts_b200.f5tts import F5TTS
from oron_tts_b200.precise import PreciseDiT
DEV = "cuda"
rel = lambda a, b: float((a.float().cpu() - b.float().cpu()).norm() / b.float().cpu().norm())
g = torch.load(os.path.join(ROOT, "tests", "golden", "dit_tiny.pt"), weights_only=False)
keys = torch.load(os.path.join(ROOT, "tests", "golden", "state_keys.pt"), weights_only=False)
def sd_of(name):
    sd = GW.fill_state_dict({k: torch.empty(s) for k, s in keys[name].items()}, GW.SEEDS[name])
    sd["cfm.backbone.rotary_embed.inv_freq"] = 1.0 / (10000 ** (torch.arange(0, 64, 2).float() / 64))
    return sd
m = F5TTS.from_config(GW.CONFIGS["tiny"]); m.load_state_dict(sd_of("tiny")); m = m.to(DEV).eval()
pd = PreciseDiT(m.cfm.backbone)
T_ = g["x"].shape[1]
mask = (torch.arange(T_)[None, :] < g["lens"][:, None]).to(DEV)
valid = torch.cat([mask, mask], 0).cpu()
out = pd.forward(g["x"].to(DEV), g["cond"].to(DEV), g["text"].to(DEV), g["time"].to(DEV), mask=mask, cfg_infer=True).cpu()
print("tiny cfg fwd vs live-reference golden:", rel(out[valid], g["fwd_cfg"][valid]))
ref = m.cfm.backbone(g["x"].to(DEV), g["cond"].to(DEV), g["text"].to(DEV), g["time"].to(DEV), mask=mask, cfg_infer=True).cpu()
print("tiny bf16 path:", rel(ref[valid], g["fwd_cfg"][valid]))
sd = sd_of("small")
m = F5TTS.from_config(GW.CONFIGS["small"]); m.load_state_dict(sd); m = m.to(DEV).eval()
pd = PreciseDiT(m.cfm.backbone)
gen = torch.Generator().manual_seed(3)
B, T_ = 2, 260
lens = torch.tensor([260, 181])
x = torch.randn(B, T_, 100, generator=gen); cond = torch.randn(B, T_, 100, generator=gen) * (torch.arange(T_)[None, :, None] < 70)
text = torch.randint(4, 65, (B, T_), generator=gen); text[1, 181:] = -1
time = torch.tensor([0.21, 0.67])
mask = torch.arange(T_)[None, :] < lens[:, None]
o = DO.dit_forward(sd, x, cond, text, time, mask, cfg_infer=True)
out = pd.forward(x.to(DEV), cond.to(DEV), text.to(DEV), time.to(DEV), mask=mask.to(DEV), cfg_infer=True).cpu()
valid = torch.cat([mask, mask], 0)
print("small cfg fwd vs oracle:", rel(out[valid], o[valid]))
