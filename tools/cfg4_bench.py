"""Stage-by-stage timing of BASELINE config 4 (log-mel STFT + Vocos decode on 64 x 30 s clips) on one B200.

Usage: python tools/cfg4_bench.py [--nb 64] [--seconds 30] [--two-sm 0|1]
Each stage is timed alone with CUDA events (kernels here run 0.1-3 ms, so host launch latency is irrelevant);
algorithmic bytes / flops are printed next to the time so the HBM or tensor fraction can be read off directly.
"""

from __future__ import annotations

import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))


def timeit(fn, reps=5, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--nb", type=int, default=64)
    ap.add_argument("--seconds", type=float, default=30.0)
    args = ap.parse_args()
    import weights as GW

    from oron_tts_b200 import _lib as L
    from oron_tts_b200.audio import AudioProcessor
    from oron_tts_b200.vocos import Vocos

    dev = torch.device("cuda", 0)
    L.lib()
    peaks = {}
    pth = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(pth):
        peaks = json.load(open(pth))
    hbm = peaks.get("hbm_gbs", 6650.0)
    voc = Vocos()
    voc.load_state_dict(GW.fill_state_dict(voc.state_dict(), 4321), strict=True)
    voc = voc.to(dev).eval()
    apx = AudioProcessor()
    nb, S = args.nb, int(args.seconds * 24000)
    wav = (torch.rand(nb, S, device=dev) * 2 - 1) * 0.3
    out = {}
    with torch.inference_mode():
        mel = apx.mel_spectrogram(wav)
        T = mel.shape[-1]
        R = nb * T
        ms = timeit(lambda: apx.mel_spectrogram(wav))
        by = nb * S * 4 + R * 100 * 4
        out["logmel"] = dict(ms=round(ms, 3), gbs=round(by / ms / 1e6, 1), hbm_frac=round(by / ms / 1e6 / hbm, 4))
        ms = timeit(lambda: voc.decode(mel), reps=5, warm=3)
        out["vocos_decode"] = dict(ms=round(ms, 3), tflops=round(R * 27.0e6 / ms / 1e9, 1), rtf=round(ms / 1e3 / (nb * args.seconds), 8))

        # ---- the stages of Vocos.decode, one by one (same calls as vocos.py) ----
        pk = voc._pack()
        D, H = 512, 1536
        BF16, F32 = torch.bfloat16, torch.float32
        a0 = torch.zeros(R, pk["cin_pad"], device=dev, dtype=BF16)
        x = torch.randn(R, D, device=dev, dtype=F32)
        n = torch.empty(R, D, device=dev, dtype=BF16)
        h = torch.empty(R, H, device=dev, dtype=BF16)
        blk = pk["blocks"][0]
        melt = mel.transpose(1, 2).reshape(R, 100).float().contiguous()

        def stage(name, fn, by=None, fl=None):
            ms = timeit(fn)
            d = dict(ms=round(ms, 3))
            if by:
                d["gbs"] = round(by / ms / 1e6, 1)
                d["hbm_frac"] = round(by / ms / 1e6 / hbm, 3)
            if fl:
                d["tflops"] = round(fl / ms / 1e9, 1)
            out[name] = d

        stage("transpose+cast", lambda: L.cast_rows_bf16(mel.transpose(1, 2).reshape(R, 100).float().contiguous(), a0[:, :100]),
              by=R * 100 * 4 * 3 + R * 100 * 2)
        for two in (False, True):
            for bn in (128, 256):
                stage(f"embed_conv_bn{bn}_2sm{int(two)}", lambda: L.gemm(a0, pk["embed_w"], x, epilogue=L.EPI_F32, bias=pk["embed_b"], rows_per_batch=T,
                                                                          nbatch=nb, taps=7, cin_blocks=pk["cin_pad"] // 64, pad=3, block_n=bn, two_sm=two),
                      fl=2.0 * R * 7 * pk["cin_pad"] * D)
        stage("ln_f32", lambda: L.ln_modulate(x, rows_per_batch=T, nbatch=nb, eps=1e-6, scale=pk["norm_w"], shift=pk["norm_b"],
                                              add_one=False, out_f32=x), by=R * D * 8)
        stage("dwconv7_ln", lambda: L.dwconv7_ln(x, rows_per_batch=T, nbatch=nb, seq_lens=None, w=blk["dw_w"], wb=blk["dw_b"],
                                                 ln_w=blk["ln_w"], ln_b=blk["ln_b"], eps=1e-6, out=n), by=R * D * 6)
        for two in (False, True):
            for bn in (128, 256):
                stage(f"pw1_gelu_bn{bn}_2sm{int(two)}", lambda: L.gemm(n, blk["w1"], h, epilogue=L.EPI_BF16, bias=blk["b1"], act=L.ACT_GELU_ERF,
                                                                        rows_per_batch=T, nbatch=nb, block_n=bn, two_sm=two), fl=2.0 * R * D * H)
        for two in (False, True):
            for bn in (128, 256):
                stage(f"pw2_resid_bn{bn}_2sm{int(two)}", lambda: L.gemm(h, blk["w2"], x, epilogue=L.EPI_SCALE_RESID, bias=blk["b2"], rows_per_batch=T,
                                                                         nbatch=nb, addend=x, gate=blk["gamma"], block_n=bn, two_sm=two), fl=2.0 * R * D * H)
        nh = pk["head_w"].shape[0]
        ldh = (nh + 31) // 32 * 32
        hs = torch.empty(R, ldh, device=dev, dtype=F32)
        for two in (False, True):
            for bn in (128, 256):
                stage(f"head_gemm_bn{bn}_2sm{int(two)}", lambda: L.gemm(n, pk["head_w"], hs, epilogue=L.EPI_F32, bias=pk["head_b"], rows_per_batch=T,
                                                                         nbatch=nb, block_n=bn, n=nh, two_sm=two), fl=2.0 * R * D * nh)
        hs.normal_(0, 0.5)
        wv = torch.empty(nb, (T - 1) * 256, device=dev, dtype=F32)
        stage("istft_head", lambda: L.istft_head(hs, pk["window"], wv, rows_per_batch=T, nb=nb, n_frames=T, mode=0),
              by=R * nh * 4 + nb * (T - 1) * 256 * 4)
    for k, v in out.items():
        print(f"{k:28s} {json.dumps(v)}")


if __name__ == "__main__":
    main()
