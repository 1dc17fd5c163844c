"""Micro-benchmarks of the hot kernels at BASELINE config-2 shapes (CUDA events, L2-warm and L2-cold).
Usage: python tools/kernel_bench.py [gemm|attn|ln|conv|ffn|all] [--ncu]   (--ncu: one launch per case, for profiling)"""
import math
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch

from oron_tts_b200 import _lib as L

DEV = "cuda"
NCU = "--ncu" in sys.argv
what = next((a for a in sys.argv[1:] if not a.startswith("--")), "all")
R, D, T = 2816, 1024, 1408
if "--short" in sys.argv:  # a 1.5 s utterance with CFG: 2 x 256 padded rows
    R, T = 512, 256
flush = torch.empty(256 * 1024 * 1024, device=DEV, dtype=torch.uint8)


def timeit(fn, reps=20, cold=False):
    if NCU:
        fn()
        torch.cuda.synchronize()
        return float("nan")
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        if cold:
            flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    ts.sort()
    return ts[len(ts) // 2]


def gemm_cases():
    g = torch.Generator(device=DEV).manual_seed(0)
    rnd = lambda *s: torch.randn(*s, device=DEV, generator=g)
    bias = {n: rnd(n) for n in (1024, 3072, 4096)}
    A1 = rnd(R, 1024).bfloat16()
    A4 = rnd(R, 4096).bfloat16()
    W = {(n, k): (rnd(n, k) / math.sqrt(k)).bfloat16() for n, k in ((3072, 1024), (1024, 1024), (4096, 1024), (1024, 4096))}
    cos, sin = rnd(T, 32), rnd(T, 32)
    gate = rnd(6 * 1024)
    xres = rnd(R, 1024)
    out = {n: torch.empty(R, n, device=DEV, dtype=torch.bfloat16) for n in (3072, 4096, 1024)}
    cases = []
    for bn in (128, 256):
        cases.append((f"qkv_rope   N=3072 K=1024 bn={bn}", 2 * R * 3072 * 1024,
                      lambda bn=bn: L.gemm(A1, W[(3072, 1024)], out[3072], epilogue=L.EPI_QKV_ROPE, bias=bias[3072], rows_per_batch=T,
                                           nbatch=2, block_n=bn, rope_cos=cos, rope_sin=sin, rope_cols=2048)))
        cases.append((f"plain bf16 N=3072 K=1024 bn={bn}", 2 * R * 3072 * 1024,
                      lambda bn=bn: L.gemm(A1, W[(3072, 1024)], out[3072], epilogue=L.EPI_BF16, bias=bias[3072], rows_per_batch=T,
                                           nbatch=2, block_n=bn)))
        cases.append((f"ffn1 gelu  N=4096 K=1024 bn={bn}", 2 * R * 4096 * 1024,
                      lambda bn=bn: L.gemm(A1, W[(4096, 1024)], out[4096], epilogue=L.EPI_BF16, bias=bias[4096], act=L.ACT_GELU_TANH,
                                           rows_per_batch=T, nbatch=2, block_n=bn)))
        cases.append((f"outproj    N=1024 K=1024 bn={bn}", 2 * R * 1024 * 1024,
                      lambda bn=bn: L.gemm(A1, W[(1024, 1024)], xres, epilogue=L.EPI_GATE_RESID, bias=bias[1024], gate=gate,
                                           rows_per_batch=T, nbatch=2, block_n=bn)))
        cases.append((f"ffn2       N=1024 K=4096 bn={bn}", 2 * R * 1024 * 4096,
                      lambda bn=bn: L.gemm(A4, W[(1024, 4096)], xres, epilogue=L.EPI_GATE_RESID, bias=bias[1024], gate=gate,
                                           rows_per_batch=T, nbatch=2, block_n=bn)))
    for bn in (128, 256):
        cases.append((f"2SM qkv_rope   N=3072 K=1024 bn={bn}", 2 * R * 3072 * 1024,
                      lambda bn=bn: L.gemm(A1, W[(3072, 1024)], out[3072], epilogue=L.EPI_QKV_ROPE, bias=bias[3072], rows_per_batch=T,
                                           nbatch=2, block_n=bn, rope_cos=cos, rope_sin=sin, rope_cols=2048, two_sm=True)))
        cases.append((f"2SM ffn1 gelu  N=4096 K=1024 bn={bn}", 2 * R * 4096 * 1024,
                      lambda bn=bn: L.gemm(A1, W[(4096, 1024)], out[4096], epilogue=L.EPI_BF16, bias=bias[4096], act=L.ACT_GELU_TANH,
                                           rows_per_batch=T, nbatch=2, block_n=bn, two_sm=True)))
        cases.append((f"2SM outproj    N=1024 K=1024 bn={bn}", 2 * R * 1024 * 1024,
                      lambda bn=bn: L.gemm(A1, W[(1024, 1024)], xres, epilogue=L.EPI_GATE_RESID, bias=bias[1024], gate=gate,
                                           rows_per_batch=T, nbatch=2, block_n=bn, two_sm=True)))
        cases.append((f"2SM ffn2       N=1024 K=4096 bn={bn}", 2 * R * 1024 * 4096,
                      lambda bn=bn: L.gemm(A4, W[(1024, 4096)], xres, epilogue=L.EPI_GATE_RESID, bias=bias[1024], gate=gate,
                                           rows_per_batch=T, nbatch=2, block_n=bn, two_sm=True)))
    for bn in (128, 256):
        cases.append((f"2SM outproj SK N=1024 K=1024 bn={bn}", 2 * R * 1024 * 1024,
                      lambda bn=bn: L.gemm(A1, W[(1024, 1024)], xres, epilogue=L.EPI_GATE_RESID, bias=bias[1024], gate=gate,
                                           rows_per_batch=T, nbatch=2, block_n=bn, two_sm=True, stream_k=True)))
        cases.append((f"2SM ffn2    SK N=1024 K=4096 bn={bn}", 2 * R * 1024 * 4096,
                      lambda bn=bn: L.gemm(A4, W[(1024, 4096)], xres, epilogue=L.EPI_GATE_RESID, bias=bias[1024], gate=gate,
                                           rows_per_batch=T, nbatch=2, block_n=bn, two_sm=True, stream_k=True)))
    cases.append(("outproj    N=1024 K=1024 bn=64", 2 * R * 1024 * 1024,
                  lambda: L.gemm(A1, W[(1024, 1024)], xres, epilogue=L.EPI_GATE_RESID, bias=bias[1024], gate=gate,
                                 rows_per_batch=T, nbatch=2, block_n=64)))
    cases.append(("ffn2       N=1024 K=4096 bn=64", 2 * R * 1024 * 4096,
                  lambda: L.gemm(A4, W[(1024, 4096)], xres, epilogue=L.EPI_GATE_RESID, bias=bias[1024], gate=gate,
                                 rows_per_batch=T, nbatch=2, block_n=64)))
    return cases


def run(cases):
    for name, flops, fn in cases:
        w, c = timeit(fn), timeit(fn, cold=True)
        print(f"{name:40s} warm {w:8.1f} us {flops / w / 1e6:8.1f} TFLOP/s | cold {c:8.1f} us {flops / c / 1e6:8.1f} TFLOP/s", flush=True)


if what in ("gemm", "all"):
    run(gemm_cases())
if what in ("attn", "all"):
    g = torch.Generator(device=DEV).manual_seed(1)
    qkv = torch.randn(R, 3072, device=DEV, generator=g).bfloat16()
    qkv[:, 2048:] = torch.randn(R, 1024, device=DEV, generator=g).half().view(torch.bfloat16)
    o = torch.zeros(R, 1024, device=DEV, dtype=torch.bfloat16)
    lens = torch.tensor([1406, 1406], device=DEV, dtype=torch.int32)
    fl = 4 * 2 * 16 * 1406 * 1406 * 64
    for ver in (4, 3):
        L.lib().oron_debug_set_attention_version(ver)
        AWS = L.attention_workspace(2, T, 16, DEV, seq_lens=lens)
        run([(f"attention v{ver} T=1406 H=16 nb=2 (one CTA per item)", fl,
              lambda: L.attention(qkv, o, nbatch=2, rows_per_batch=T, heads=16, seq_lens=lens, scale=0.125)),
             (f"attention v{ver} T=1406 H=16 nb=2 (planned shares)", fl,
              lambda: L.attention(qkv, o, nbatch=2, rows_per_batch=T, heads=16, seq_lens=lens, scale=0.125, workspace=AWS))])
    # back-to-back launches inside one CUDA graph (what the ODE step sees: PDL overlap, warm L2), after 1.5 s of sustained
    # load: the SM clock ramps up from idle over many milliseconds, so short bursts are timed at a low clock
    def graph_time(ver):
        import time
        L.lib().oron_debug_set_attention_version(ver)
        AWS = L.attention_workspace(2, T, 16, DEV, seq_lens=lens)
        fn = lambda: L.attention(qkv, o, nbatch=2, rows_per_batch=T, heads=16, seq_lens=lens, scale=0.125, workspace=AWS)
        s_ = torch.cuda.Stream()
        with torch.cuda.stream(s_):
            for _ in range(3): fn()
            gph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(gph, stream=s_):
                for _ in range(22): fn()
            t_end = time.time() + 1.5
            while time.time() < t_end:
                for _ in range(20): gph.replay()
                torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(s_)
            for _ in range(50): gph.replay()
            e1.record(s_)
        torch.cuda.synchronize()
        us = e0.elapsed_time(e1) * 1e3 / (22 * 50)
        print(f"attention v{ver} planned, 22 launches per graph replay, sustained: {us:.1f} us per launch = {fl / us / 1e6:.1f} TFLOP/s", flush=True)
    if not NCU:
        for ver in ((4,) if os.environ.get("ORON_ATT_ABL") else (4, 3)):
            graph_time(ver)
    L.lib().oron_debug_set_attention_version(4)
if what in ("ln", "all"):
    g = torch.Generator(device=DEV).manual_seed(2)
    x = torch.randn(R, 1024, device=DEV, generator=g)
    tab = torch.randn(6 * 1024, device=DEV, generator=g)
    o = torch.empty(R, 1024, device=DEV, dtype=torch.bfloat16)
    fn = lambda: L.ln_modulate(x, rows_per_batch=T, nbatch=2, eps=1e-6, scale=tab[1024:], shift=tab, add_one=True, out_bf16=o)
    w, c = timeit(fn), timeit(fn, cold=True)
    byt = R * 1024 * 6
    print(f"ln_modulate R=2816 C=1024: warm {w:.1f} us ({byt / w / 1e3:.0f} GB/s) cold {c:.1f} us ({byt / c / 1e3:.0f} GB/s)")
if what in ("conv",):
    # ConvPositionEmbedding conv (k = 31, 64-channel groups) at config 2: resident-window kernel vs the generic per-tap fetch
    g = torch.Generator(device=DEV).manual_seed(5)
    x = torch.randn(R, D, device=DEV, generator=g).bfloat16()
    w = (torch.randn(D, 31 * 64, device=DEV, generator=g) / math.sqrt(31 * 64)).bfloat16()
    bias = torch.randn(D, device=DEV, generator=g) * 0.1
    lens = torch.tensor([1406, 1406], device=DEV, dtype=torch.int32)
    o = torch.empty(R, D, device=DEV, dtype=torch.bfloat16)
    fl = 2 * R * D * 64 * 31
    import time
    for mode, epi in (("1", L.EPI_MISH_MASK_BF16), ("0", L.EPI_MISH_MASK_BF16), ("1", L.EPI_BF16), ("0", L.EPI_BF16)):
        fn = lambda: L.gemm(x, w, o, epilogue=epi, bias=bias, rows_per_batch=T, nbatch=2, taps=31, cin_blocks=1, pad=15,
                            grouped=64, block_n=64, seq_lens=lens)
        os.environ["ORON_GCONV_RES"] = mode
        run([(f"grouped conv k=31 R={R} (ORON_GCONV_RES={mode}, epilogue {epi})", fl, fn)])
        if NCU:
            continue
        s_ = torch.cuda.Stream()  # 20 launches per CUDA-graph replay after 1.5 s of sustained load (no host time in the figure)
        with torch.cuda.stream(s_):
            for _ in range(3): fn()
            gph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(gph, stream=s_):
                for _ in range(20): fn()
            t_end = time.time() + 1.5
            while time.time() < t_end:
                for _ in range(20): gph.replay()
                torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(s_)
            for _ in range(50): gph.replay()
            e1.record(s_)
        torch.cuda.synchronize()
        us = e0.elapsed_time(e1) * 1e3 / (20 * 50)
        print(f"  in a CUDA graph, sustained: {us:.1f} us per launch = {fl / us / 1e6:.1f} TFLOP/s", flush=True)
    os.environ.pop("ORON_GCONV_RES")
if what in ("outproj",):
    # attention out-projection (N = K = 1024, gated residual) at config 2: tile widths of the 2-SM kernel, sustained graph replay
    import time
    g = torch.Generator(device=DEV).manual_seed(7)
    rnd = lambda *s: torch.randn(*s, device=DEV, generator=g)
    A = rnd(R, 1024).bfloat16()
    Ws = [(rnd(1024, 1024) / 32).bfloat16() for _ in range(22)]
    b, gate = rnd(1024), rnd(1024) * 0.01
    lens = torch.tensor([1406, 1406], device=DEV, dtype=torch.int32)
    fl = 2 * R * 1024 * 1024
    ref = None
    for bn, sk in ((256, False), (192, False), (128, False), (256, True)):
        xres = torch.zeros(R, 1024, device=DEV)
        fn = lambda i: L.gemm(A, Ws[i], xres, epilogue=L.EPI_GATE_RESID, bias=b, gate=gate, rows_per_batch=T, nbatch=2, block_n=bn,
                              two_sm=True, stream_k=sk, seq_lens=lens, mask_rows=True)
        fn(0)
        torch.cuda.synchronize()
        if ref is None:
            ref = xres.clone()
        print(f"bn={bn} sk={sk}: max |diff| vs bn=256 = {float((xres - ref).abs().max()):.3e}", flush=True)
        if NCU:
            continue
        s_ = torch.cuda.Stream()
        with torch.cuda.stream(s_):
            for i in range(3): fn(i)
            gph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(gph, stream=s_):
                for i in range(22): fn(i)
            t_end = time.time() + 1.5
            while time.time() < t_end:
                for _ in range(20): gph.replay()
                torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(s_)
            for _ in range(50): gph.replay()
            e1.record(s_)
        torch.cuda.synchronize()
        us = e0.elapsed_time(e1) * 1e3 / (22 * 50)
        print(f"  out-projection bn={bn} stream_k={sk}: {us:.2f} us per launch = {fl / us / 1e6:.1f} TFLOP/s", flush=True)
if what in ("ffn",):
    # FeedForward of one DiTBlock at config 2: two launches (up-projection + stream-K down-projection) against the fused launch,
    # 22 per CUDA-graph replay with 22 distinct weight sets (as in the ODE step: weights stream from HBM), sustained clocks
    import time
    g = torch.Generator(device=DEV).manual_seed(3)
    rnd = lambda *s: torch.randn(*s, device=DEV, generator=g)
    A1 = rnd(R, 1024).bfloat16()
    hid = torch.empty(R, 4096, device=DEV, dtype=torch.bfloat16)
    xres = rnd(R, 1024)
    gate = rnd(1024) * 0.01
    W1 = [(rnd(4096, 1024) / 32).bfloat16() for _ in range(22)]
    W2 = [(rnd(1024, 4096) / 64).bfloat16() for _ in range(22)]
    b1, b2 = rnd(4096), rnd(1024)
    fws = L.ffn_workspace(T, 2, 4096, DEV)
    stamps = None
    def layer(i, fused, sk=True):
        up = L.gemm(A1, W1[i], hid, epilogue=L.EPI_BF16, bias=b1, act=L.ACT_GELU_TANH, rows_per_batch=T, nbatch=2, block_n=256,
                    two_sm=True, desc_only=fused, debug_stamps=stamps if fused else None)
        dn = L.gemm(hid, W2[i], xres, epilogue=L.EPI_GATE_RESID, bias=b2, gate=gate, rows_per_batch=T, nbatch=2, block_n=256,
                    two_sm=True, stream_k=sk, desc_only=fused)
        if fused:
            L.ffn(up, dn, fws)
    fl = 2 * R * 1024 * 4096 * 2
    if NCU:
        for i in range(2):
            layer(i, False, True)
            layer(i, True, True)
        torch.cuda.synchronize()
        sys.exit(0)
    same = "--same-weights" in sys.argv  # every layer reads the same (L2-resident) weights: is the main loop slowed by HBM-cold operands?
    for name, fused, sk in (("two launches (stream-K down)", False, True), ("two launches (whole tiles)", False, False), ("fused launch", True, True)):
        s_ = torch.cuda.Stream()
        with torch.cuda.stream(s_):
            for i in range(3): layer(i, fused, sk)
            gph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(gph, stream=s_):
                for i in range(22): layer(0 if same else i, fused, sk)
            t_end = time.time() + 1.5
            while time.time() < t_end:
                for _ in range(20): gph.replay()
                torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(s_)
            for _ in range(50): gph.replay()
            e1.record(s_)
        torch.cuda.synchronize()
        us = e0.elapsed_time(e1) * 1e3 / (22 * 50)
        print(f"FFN {name}: {us:.1f} us per block = {fl / us / 1e6:.1f} TFLOP/s", flush=True)
    if "--trace" in sys.argv:
        # per-CTA stamps of three launches inside a sustained run (clock64 deltas per CTA, globaltimer across CTAs)
        stamps_all = torch.zeros(22, 148, 16, device=DEV, dtype=torch.int64)
        s_ = torch.cuda.Stream()
        with torch.cuda.stream(s_):
            gph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(gph, stream=s_):
                for i in range(22):
                    stamps = stamps_all[i]
                    layer(i, True)
            t_end = time.time() + 1.5
            while time.time() < t_end:
                for _ in range(20): gph.replay()
                torch.cuda.synchronize()
        st_all = stamps_all.cpu()
        for li in (10, 11):
            st = st_all[li]
            ns0 = int(st[:, 11].min())
            print(f"launch {li}: start spread {int(st[:,11].max())-ns0} ns, phase-2 first load {int(st[:,13].min())-ns0}..{int(st[:,13].max())-ns0} ns, "
                  f"end {int(st[:,12].min())-ns0}..{int(st[:,12].max())-ns0} ns; next launch starts {int(st_all[li+1][:,11].min())-ns0} ns")
            cyc = (st[:, 10] - st[:, 0]).float()
            print(f"  cycles per CTA start->end: mean {cyc.mean():.0f} max {cyc.max():.0f}; clock = {float(cyc.max()) / (int(st[:,12].max())-ns0) :.3f} GHz")
            for c in (0, 2, 54, 56, 100, 146):
                r = st[c]
                print(f"  cta {c:3d}: p2-first-load {int(r[2]-r[0]):6d} loads-done {int(r[3]-r[0]):6d} mma-p2-start {int(r[14]-r[0]):6d} mma-done {int(r[4]-r[0]):6d} end {int(r[10]-r[0]):6d} | "
                      f"producer waits: empty p1 {int(r[1]):6d} p2 {int(r[9]):6d} flags {int(r[8]):6d} | mma waits: full p1 {int(r[5]):6d} p2 {int(r[6]):6d} tempty {int(r[7]):6d}")
