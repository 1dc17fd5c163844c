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
