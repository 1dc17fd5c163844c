"""Per-CTA clock64 timeline of the 2-SM GEMM (ORON_STAMP slots) at config-2 shapes, L2-warm."""
import math, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from oron_tts_b200 import _lib as L
DEV = "cuda"
R, T = 2816, 1408
g = torch.Generator(device=DEV).manual_seed(0)
rnd = lambda *s: torch.randn(*s, device=DEV, generator=g)
names = {0: "start", 1: "tma first issued", 2: "tma tile0 all issued", 3: "mma first full", 4: "mma tile0 issued", 5: "mma tile1 issued",
         6: "epi tile0 tfull", 7: "epi tile0 done", 8: "epi tile1 tfull", 9: "epi tile1 done", 10: "end",
         11: "c0 start", 12: "c0 tmem ld done", 13: "c0 staged", 14: "c0 bias loaded", 15: "c0 done"}
for (N, K, epi, bn) in ((3072, 1024, "bf16", 256), (1024, 1024, "gate", 256), (4096, 1024, "bf16", 256), (1024, 4096, "gate", 256)):
    A = rnd(R, K).bfloat16(); W = (rnd(N, K) / math.sqrt(K)).bfloat16(); bias = rnd(N)
    dbg = torch.zeros(148, 16, device=DEV, dtype=torch.int64)
    if epi == "bf16":
        out = torch.empty(R, N, device=DEV, dtype=torch.bfloat16)
        fn = lambda d=None: L.gemm(A, W, out, epilogue=L.EPI_BF16, bias=bias, rows_per_batch=T, nbatch=2, block_n=bn, two_sm=True, debug_stamps=d)
    else:
        x = rnd(R, N); gate = rnd(N)
        fn = lambda d=None: L.gemm(A, W, x, epilogue=L.EPI_GATE_RESID, bias=bias, gate=gate, rows_per_batch=T, nbatch=2, block_n=bn, two_sm=True, debug_stamps=d)
    for _ in range(3): fn()
    torch.cuda.synchronize()
    torch.cuda._sleep(200000)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); fn(dbg); e1.record(); torch.cuda.synchronize()
    d = dbg.cpu()
    print(f"== N={N} K={K} {epi} bn={bn}: {e0.elapsed_time(e1)*1e3:.1f} us (L2-warm)")
    for cta in (0, 1, 2, 147):
        base = int(d[cta, 0])
        print(f"  cta {cta}: " + ", ".join(f"{names[i]}={int(d[cta, i]) - base}" for i in range(1, 16) if int(d[cta, i]) != 0))
