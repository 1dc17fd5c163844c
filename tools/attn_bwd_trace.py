"""Per-CTA clock64 timeline of the attention-backward kernels at config-5 shapes (8 x 1024 frames, 16 heads)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from oron_tts_b200 import _lib as L
from oron_tts_b200 import _lib_train as T
DEV = "cuda"
nb, rpb, H = 8, 1024, 16
HD, R = H * 64, nb * rpb
g = torch.Generator(device=DEV).manual_seed(1)
qkv = torch.randn(R, 3 * HD, device=DEV, generator=g).bfloat16()
qkv[:, 2 * HD:] = torch.randn(R, HD, device=DEV, generator=g).half().view(torch.bfloat16)
vb = torch.randn(R, HD, device=DEV, generator=g).bfloat16()
o = torch.zeros(R, HD, device=DEV, dtype=torch.bfloat16)
lse = torch.zeros(nb * H * rpb, device=DEV)
delta = torch.zeros_like(lse)
T.attention_fwd_lse(qkv, o, lse, nbatch=nb, rows_per_batch=rpb, heads=H, seq_lens=None, scale=0.125)
d_o = (torch.randn(R, HD, device=DEV, generator=g) * 0.1).bfloat16()
dqkv = torch.zeros(R, 3 * HD, device=DEV, dtype=torch.bfloat16)
inv = 1.0 / (10000 ** (torch.arange(0, 64, 2, device=DEV).float() / 64))
ang = torch.outer(torch.arange(rpb, device=DEV).float(), inv)
cos, sin = ang.cos().contiguous(), ang.sin().contiguous()
run = lambda: T.attention_bwd(qkv[:, :2 * HD], vb, o, d_o, dqkv, nbatch=nb, rows_per_batch=rpb, heads=H, seq_lens=None,
                              scale=0.125, rope_cos=cos, rope_sin=sin, lse=lse, delta=delta, have_lse=True)
for _ in range(3): run()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10): run()
e1.record(); torch.cuda.synchronize()
print(f"attention backward (dQ + dK/dV): {e0.elapsed_time(e1) / 10 * 1e3:.1f} us per call")
grid = (rpb // 128) * H * nb
dbg = torch.zeros(2 * grid, 16, device=DEV, dtype=torch.int64)
T.tlib().oron_debug_set_attention_bwd_stamps.argtypes = [__import__("ctypes").c_void_p]
T.tlib().oron_debug_set_attention_bwd_stamps(dbg.data_ptr())
run(); torch.cuda.synchronize()
T.tlib().oron_debug_set_attention_bwd_stamps(None)
d = dbg.cpu()
names = ["start", "stats done", "first S/dP ready", "second S/dP ready", "loop end", "acc ready", "stored"]
for mode in (0, 1):
    x = d[mode * grid:(mode + 1) * grid]
    print(f"mode {mode}: cycles since the CTA's first stamp (median over CTAs)")
    for k in range(1, 7):
        print(f"   {names[k]:22s} {int((x[:, k] - x[:, 0]).median())}")
    med = lambda a, b_: int((x[:, a] - x[:, b_]).median())
    print(f"   sub-iteration 6 (tile 3, first half), MMA thread: fetch/acc-wait {med(9, 8)}, issue S/dP(n+1) {med(10, 9)}, "
          f"wait bar_p {med(11, 10)}, issue acc {med(12, 11)}")
    print(f"   same, compute thread: wait bar_s {med(14, 13)}, tmem ld + math + stores + arrive {med(15, 14)}; "
          f"compute start relative to the MMA thread's loop top {med(13, 8)}")
    print(f"   per 128-column tile    {int(((x[:, 4] - x[:, 3]) / 7).median())}  (tiles 1..7)")
