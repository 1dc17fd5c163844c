"""Two launches each of dwconv7_ln (C = 512, config-4 size) and the iSTFT head for `ncu --set full`."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oron_tts_b200 import _lib as L  # noqa: E402

dev = torch.device("cuda", 0)
L.lib()
nb, T, D = 64, 2813, 512
R = nb * T
x = torch.randn(R, D, device=dev)
n = torch.empty(R, D, device=dev, dtype=torch.bfloat16)
w, wb = torch.randn(D, 7, device=dev), torch.randn(D, device=dev)
lw, lb = torch.randn(D, device=dev), torch.randn(D, device=dev)
for _ in range(3):
    L.dwconv7_ln(x, rows_per_batch=T, nbatch=nb, seq_lens=None, w=w, wb=wb, ln_w=lw, ln_b=lb, eps=1e-6, out=n)
torch.cuda.synchronize()
print("ok", float(n.float().abs().mean()))
