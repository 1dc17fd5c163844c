"""A few launches of the attention kernel at config-2 shapes for `ncu` (--ws: balanced schedule)."""
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
ws = L.attention_workspace(2, T, 16, DEV, seq_lens=lens) if "--ws" in sys.argv else None
for _ in range(3):
    L.attention(qkv, o, nbatch=2, rows_per_batch=T, heads=16, seq_lens=lens, scale=0.125, workspace=ws)
torch.cuda.synchronize()
print("ok", float(o.float().abs().mean()))
