"""Two launches each of the log-mel and iSTFT-head kernels at config-4 size (64 x 30 s), for `ncu --set full`."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oron_tts_b200 import _lib as L  # noqa: E402
from oron_tts_b200.audio import AudioProcessor  # noqa: E402

dev = torch.device("cuda", 0)
L.lib()
nb, S = 64, 720000
wav = (torch.rand(nb, S, device=dev) * 2 - 1) * 0.3
ap = AudioProcessor()
for _ in range(2):
    mel = ap.mel_spectrogram(wav)
T = mel.shape[-1]
hs = torch.randn(nb * T, 1056, device=dev) * 0.5
wv = torch.empty(nb, (T - 1) * 256, device=dev)
win = torch.hann_window(1024, device=dev)
for _ in range(2):
    L.istft_head(hs, win, wv, rows_per_batch=T, nb=nb, n_frames=T, mode=0)
torch.cuda.synchronize()
print("ok", float(mel.mean()), float(wv.abs().mean()))
# dwconv7 + LayerNorm at the same size (C = 512)
R, D = nb * T, 512
x = torch.randn(R, D, device=dev)
n = torch.empty(R, D, device=dev, dtype=torch.bfloat16)
w, wb = torch.randn(D, 7, device=dev), torch.randn(D, device=dev)
lw, lb = torch.randn(D, device=dev), torch.randn(D, device=dev)
for _ in range(2):
    L.dwconv7_ln(x, rows_per_batch=T, nbatch=nb, seq_lens=None, w=w, wb=wb, ln_w=lw, ln_b=lb, eps=1e-6, out=n)
torch.cuda.synchronize()
print("ok dwconv", float(n.float().abs().mean()))
