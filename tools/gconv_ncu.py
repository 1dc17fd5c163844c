"""One config-5-sized launch of the tensor-core grouped-conv weight gradient (and the CUDA-core kernel) for ncu:
  ncu --set full --clock-control none --import-source on -k regex:gconv -o gpurun_out/r02_gconv python tools/gconv_ncu.py"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch

from oron_tts_b200 import _lib_train as T

dev = "cuda"
g = torch.Generator(device=dev).manual_seed(0)
nb, rpb, C, cg, taps = 8, 1024, 1024, 64, 31
x = torch.randn(nb * rpb, C, device=dev, generator=g).bfloat16()
dy = (torch.randn(nb * rpb, C, device=dev, generator=g) * 0.1).bfloat16()
dw = torch.zeros(C, cg, taps, device=dev)
for _ in range(2):
    T.gconv_wgrad_tc(x, dy, rows_per_batch=rpb, nbatch=nb, cg=cg, taps=taps, dw=dw)
    T.gconv_wgrad(x, dy, rows_per_batch=rpb, nbatch=nb, cg=cg, taps=taps, seq_lens=None, dw=dw, db=None)
torch.cuda.synchronize()
e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
e[0].record()
for _ in range(20):
    T.gconv_wgrad_tc(x, dy, rows_per_batch=rpb, nbatch=nb, cg=cg, taps=taps, dw=dw)
e[1].record()
for _ in range(20):
    T.gconv_wgrad(x, dy, rows_per_batch=rpb, nbatch=nb, cg=cg, taps=taps, seq_lens=None, dw=dw, db=None)
e[2].record()
torch.cuda.synchronize()
fl = 2.0 * nb * rpb * C * cg * taps
print(f"gconv wgrad 8 x 1024 x 1024, 31 taps: tensor core {e[0].elapsed_time(e[1]) / 20 * 1e3:.1f} us "
      f"({fl / (e[0].elapsed_time(e[1]) / 20 * 1e-3) / 1e12:.1f} useful TFLOP/s), CUDA cores {e[1].elapsed_time(e[2]) / 20 * 1e3:.1f} us")
