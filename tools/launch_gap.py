"""Per-launch cost of back-to-back kernels inside a CUDA graph (config-2 shapes): isolates launch gaps /
prologue / teardown from in-kernel time."""
import math, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from oron_tts_b200 import _lib as L
DEV = "cuda"
R, T = 2816, 1408
g = torch.Generator(device=DEV).manual_seed(0)
rnd = lambda *s: torch.randn(*s, device=DEV, generator=g)
A1 = rnd(R, 1024).bfloat16(); A4 = rnd(R, 4096).bfloat16()
Wq = [(rnd(3072, 1024) / 32).bfloat16() for _ in range(4)]
W1 = [(rnd(4096, 1024) / 32).bfloat16() for _ in range(4)]
Wo = [(rnd(1024, 1024) / 32).bfloat16() for _ in range(4)]
W2 = [(rnd(1024, 4096) / 64).bfloat16() for _ in range(4)]
bq, b1, bo = rnd(3072), rnd(4096), rnd(1024)
cos, sin = rnd(T, 32), rnd(T, 32)
gate = rnd(1024) * 0.01
qkv = torch.empty(R, 3072, device=DEV, dtype=torch.bfloat16); hid = torch.empty(R, 4096, device=DEV, dtype=torch.bfloat16)
xres = rnd(R, 1024); nrm = torch.empty(R, 1024, device=DEV, dtype=torch.bfloat16)
tab = rnd(6 * 1024) * 0.1
def f_qkv(i): L.gemm(A1, Wq[i % 4], qkv, epilogue=L.EPI_QKV_ROPE, bias=bq, rows_per_batch=T, nbatch=2, block_n=256, rope_cos=cos, rope_sin=sin, rope_cols=2048, f16_from_col=2048, two_sm=True)
def f_ffn1(i): L.gemm(A1, W1[i % 4], hid, epilogue=L.EPI_BF16, bias=b1, act=L.ACT_GELU_TANH, rows_per_batch=T, nbatch=2, block_n=256, two_sm=True)
def f_out(i): L.gemm(A1, Wo[i % 4], xres, epilogue=L.EPI_GATE_RESID, bias=bo, gate=gate, rows_per_batch=T, nbatch=2, block_n=256, two_sm=True)
def f_ffn2(i): L.gemm(A4, W2[i % 4], xres, epilogue=L.EPI_GATE_RESID, bias=bo, gate=gate, rows_per_batch=T, nbatch=2, block_n=256, two_sm=True)
def f_ln(i): L.ln_modulate(xres, rows_per_batch=T, nbatch=2, eps=1e-6, scale=tab[1024:], shift=tab, add_one=True, out_bf16=nrm)
def f_mix(i): f_ln(i); f_qkv(i); f_out(i); f_ln(i); f_ffn1(i); f_ffn2(i)
def run(name, fn, n, per=1):
    for i in range(2): fn(i)
    torch.cuda.synchronize()
    gr = torch.cuda.CUDAGraph()
    with torch.cuda.graph(gr):
        for i in range(n): fn(i)
    gr.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5): gr.replay()
    e1.record(); torch.cuda.synchronize()
    print(f"{name:28s} {e0.elapsed_time(e1) * 1e3 / 5 / n / per:8.2f} us per launch  (PDL={os.environ.get('ORON_PDL', '1')})", flush=True)
run("qkv gemm x40 (4 weight sets)", f_qkv, 40)
run("ffn1 gemm x40", f_ffn1, 40)
run("outproj gemm x40", f_out, 40)
run("ffn2 gemm x40", f_ffn2, 40)
run("ln_modulate x40", f_ln, 40)
run("block mix (6 launches) x10", f_mix, 10, per=1)
