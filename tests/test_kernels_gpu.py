"""Kernel-level numerics: every C-ABI entry point against the plain PyTorch op it replaces,
on the same seeded inputs (bf16 operands are rounded identically on both sides; the torch side
accumulates in fp32). These call through liboron_b200.so — no fallback exists."""

import math

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def L():
    from oron_tts_b200 import _lib

    _lib.lib()
    return _lib


def _rel(a: torch.Tensor, b: torch.Tensor) -> float:
    return float((a.float() - b.float()).norm() / (b.float().norm() + 1e-12))


def _bf(x):
    return x.to(torch.bfloat16)


DEV = "cuda"


# ------------------------------------------------------------------------------------------
# GEMM
# ------------------------------------------------------------------------------------------
@pytest.mark.parametrize(
    "M,N,K,bn",
    [
        (128, 128, 64, 128),
        (256, 256, 256, 128),
        (2816, 1024, 1024, 128),
        (2816, 3072, 1024, 256),
        (300, 1026, 512, 128),  # ragged M and N
        (32, 4096, 1024, 256),  # tiny M (modulation table shape)
        (1406, 1024, 4096, 64),
        (2816, 1024, 128, 128),
    ],
)
def test_gemm_bf16_plain(L, M, N, K, bn):
    g = torch.Generator(device=DEV).manual_seed(M * 7 + N * 3 + K)
    A = _bf(torch.randn(M, K, device=DEV, generator=g))
    W = _bf(torch.randn(N, K, device=DEV, generator=g) / math.sqrt(K))
    bias = torch.randn(N, device=DEV, generator=g)
    ldo = (N + 7) // 8 * 8
    out = torch.full((M, ldo), float("nan"), device=DEV, dtype=torch.bfloat16)
    L.gemm(A, W, out, epilogue=L.EPI_BF16, bias=bias, block_n=bn, n=N)
    ref = A.float() @ W.float().t() + bias
    torch.cuda.synchronize()
    assert torch.isfinite(out[:, :N].float()).all()
    assert _rel(out[:, :N], ref) < 6e-3
    assert float((out[:, :N].float() - ref).abs().max()) < 0.06 * float(ref.abs().max())


@pytest.mark.parametrize("act", ["gelu_tanh", "gelu_erf", "silu"])
def test_gemm_activation_epilogues(L, act):
    M, N, K = 512, 512, 256
    g = torch.Generator(device=DEV).manual_seed(11)
    A = _bf(torch.randn(M, K, device=DEV, generator=g))
    W = _bf(torch.randn(N, K, device=DEV, generator=g) / math.sqrt(K) * 2)
    bias = torch.randn(N, device=DEV, generator=g)
    out = torch.empty(M, N, device=DEV, dtype=torch.bfloat16)
    code = {"gelu_tanh": L.ACT_GELU_TANH, "gelu_erf": L.ACT_GELU_ERF, "silu": L.ACT_SILU}[act]
    L.gemm(A, W, out, epilogue=L.EPI_BF16, bias=bias, act=code)
    pre = A.float() @ W.float().t() + bias
    ref = {"gelu_tanh": lambda x: F.gelu(x, approximate="tanh"), "gelu_erf": F.gelu, "silu": F.silu}[act](pre)
    assert _rel(out, ref) < 6e-3


def test_gemm_f32_addend_and_tail(L):
    M, N, K = 1406, 100, 1024
    g = torch.Generator(device=DEV).manual_seed(5)
    A = _bf(torch.randn(M, K, device=DEV, generator=g))
    W = _bf(torch.randn(N, K, device=DEV, generator=g) / math.sqrt(K))
    bias = torch.randn(N, device=DEV, generator=g)
    out = torch.full((M, N), float("nan"), device=DEV)
    L.gemm(A, W, out, epilogue=L.EPI_F32, bias=bias)
    ref = A.float() @ W.float().t() + bias
    assert _rel(out, ref) < 1e-4
    add = torch.randn(M, 128, device=DEV, generator=g)
    out2 = torch.zeros(M, 128, device=DEV)
    L.gemm(A, W, out2, epilogue=L.EPI_F32, bias=bias, addend=add, n=N)
    assert _rel(out2[:, :N], ref + add[:, :N]) < 1e-4
    assert float(out2[:, N:].abs().max()) == 0.0


def test_gemm_qkv_rope(L):
    nb, T, D, H = 2, 384, 1024, 16
    M = nb * T
    g = torch.Generator(device=DEV).manual_seed(3)
    A = _bf(torch.randn(M, D, device=DEV, generator=g))
    W = _bf(torch.randn(3 * D, D, device=DEV, generator=g) / math.sqrt(D))
    bias = torch.randn(3 * D, device=DEV, generator=g) * 0.1
    inv_freq = 1.0 / (10000 ** (torch.arange(0, 64, 2, device=DEV).float() / 64))
    ang = torch.outer(torch.arange(T, device=DEV).float(), inv_freq)  # [T, 32]
    cos, sin = ang.cos().contiguous(), ang.sin().contiguous()
    for bn in (128, 256):
        out = torch.empty(M, 3 * D, device=DEV, dtype=torch.bfloat16)
        L.gemm(A, W, out, epilogue=L.EPI_QKV_ROPE, bias=bias, rows_per_batch=T, nbatch=nb, block_n=bn,
               rope_cos=cos, rope_sin=sin, rope_cols=2 * D, f16_from_col=2 * D)
        out = torch.cat([out[:, :2 * D].float(), out[:, 2 * D:].contiguous().view(torch.float16).float()], 1)
        pre = (A.float() @ W.float().t() + bias).view(nb, T, 3, H, 64)
        q, k, v = pre[:, :, 0], pre[:, :, 1], pre[:, :, 2]
        c = torch.cat([cos, cos], -1)[None, :, None, :]
        s = torch.cat([sin, sin], -1)[None, :, None, :]

        def rot(x):
            return torch.cat([-x[..., 32:], x[..., :32]], -1)

        ref = torch.stack([q * c + rot(q) * s, k * c + rot(k) * s, v], 2).reshape(M, 3 * D)
        assert _rel(out, ref) < 6e-3, bn


def test_gemm_gate_residual_masked(L):
    nb, T, D = 2, 256, 1024
    M = nb * T
    g = torch.Generator(device=DEV).manual_seed(9)
    A = _bf(torch.randn(M, D, device=DEV, generator=g))
    W = _bf(torch.randn(D, D, device=DEV, generator=g) / math.sqrt(D))
    bias = torch.randn(D, device=DEV, generator=g) * 0.1
    steps = 3
    table = torch.randn(steps, 1, 6 * D, device=DEV, generator=g)  # shared by all batch elements
    step = torch.tensor([2], device=DEV, dtype=torch.int32)
    lens = torch.tensor([256, 100], device=DEV, dtype=torch.int32)
    x0 = torch.randn(M, D, device=DEV, generator=g)
    x = x0.clone()
    gate_view = table.view(-1)[2 * D:]  # chunk 2 (gate_msa) of step 0
    L.gemm(A, W, x, epilogue=L.EPI_GATE_RESID, bias=bias, rows_per_batch=T, nbatch=nb, gate=gate_view, gate_ld=0,
           gate_nb=1, gate_step_stride=6 * D, step_ptr=step, seq_lens=lens, mask_rows=True)
    gate = table[2, 0, 2 * D:3 * D]
    upd = gate * (A.float() @ W.float().t() + bias)
    mask = (torch.arange(T, device=DEV)[None, :] < lens[:, None]).reshape(M, 1)
    ref = x0 + torch.where(mask, upd, torch.zeros_like(upd))
    assert _rel(x, ref) < 2e-3
    assert torch.equal(x[~mask.squeeze(1)], x0[~mask.squeeze(1)])


@pytest.mark.parametrize("N,K,T,lens", [(1024, 1024, 300, [300, 131]), (768, 768, 1408, [1406, 900]), (992, 256, 129, [129, 128]),
                                        (1024, 4096, 256, [256, 256])])
def test_gemm_gate_residual_tile_width_192(L, N, K, T, lens):
    """192-wide tiles of the 2-SM kernel (engine.resid_tile_width: out-projection / whole-tile down-projection, modules.py:279,
    299, 338, 343): same result as the 256-wide tiles bit for bit (same k order per output), incl. a last tile that is mostly
    (N = 1024: 64 of 192 columns) or raggedly (N = 992) out of range, and against an fp32 reference."""
    nb = len(lens)
    M = nb * T
    g = torch.Generator(device=DEV).manual_seed(19)
    A = _bf(torch.randn(M, K, device=DEV, generator=g))
    W = _bf(torch.randn(N, K, device=DEV, generator=g) / math.sqrt(K))
    bias = torch.randn(N, device=DEV, generator=g) * 0.1
    gate = torch.randn(N, device=DEV, generator=g)
    lens_t = torch.tensor(lens, device=DEV, dtype=torch.int32)
    x0 = torch.randn(M, N, device=DEV, generator=g)
    outs = {}
    for bn in (256, 192):
        x = x0.clone()
        L.gemm(A, W, x, epilogue=L.EPI_GATE_RESID, bias=bias, rows_per_batch=T, nbatch=nb, gate=gate, seq_lens=lens_t,
               mask_rows=True, block_n=bn, two_sm=True)
        outs[bn] = x
    assert torch.equal(outs[192], outs[256])
    mask = (torch.arange(T, device=DEV)[None, :] < lens_t[:, None]).reshape(M, 1)
    upd = gate * (A.float() @ W.float().t() + bias)
    ref = x0 + torch.where(mask, upd, torch.zeros_like(upd))
    assert _rel(outs[192], ref) < 2e-3
    assert torch.equal(outs[192][~mask.squeeze(1)], x0[~mask.squeeze(1)])


@pytest.mark.parametrize("N,K,T,lens,bn,sk", [(1024, 1024, 1408, [1406, 1406], 192, False), (1024, 4096, 1408, [1406, 1000], 256, True),
                                              (1024, 1024, 300, [300, 131, 7], 256, False), (768, 3072, 700, [700], 256, True),
                                              (768, 768, 129, [129, 77], 192, False), (1024, 4096, 1024, [1024] * 8, 256, True)])
def test_gemm_ln_tail(L, N, K, T, lens, bn, sk):
    """Gated-residual GEMM with the LayerNorm + modulation of the updated rows as the tail of the launch (oron_gemm_ln_bf16;
    modules.py:338-343 followed by :218 / :341 / :234) against the two separate launches: bit-identical without stream-K,
    within rounding of the sum order with it; odd m-tile counts (a phantom tile in the last pair), several batch elements,
    masked rows, a per-step modulation table, and three launches on ONE counter buffer (the launch re-arms it)."""
    nb = len(lens)
    M = nb * T
    g = torch.Generator(device=DEV).manual_seed(29)
    A = _bf(torch.randn(M, K, device=DEV, generator=g))
    W = _bf(torch.randn(N, K, device=DEV, generator=g) / math.sqrt(K))
    bias = torch.randn(N, device=DEV, generator=g) * 0.1
    steps = 3
    table = torch.randn(steps, 6 * N, device=DEV, generator=g) * 0.3
    tab = table.view(-1)
    step = torch.tensor([1], device=DEV, dtype=torch.int32)
    lens_t = torch.tensor(lens, device=DEV, dtype=torch.int32)
    x0 = torch.randn(M, N, device=DEV, generator=g)
    kw = dict(epilogue=L.EPI_GATE_RESID, bias=bias, rows_per_batch=T, nbatch=nb, gate=tab[2 * N:], gate_step_stride=6 * N,
              step_ptr=step, seq_lens=lens_t, mask_rows=True, block_n=bn, two_sm=True, stream_k=sk)
    lnkw = dict(eps=1e-6, scale=tab[4 * N:], shift=tab[3 * N:], step_stride=6 * N, add_one=True)
    x_ref = x0.clone()
    L.gemm(A, W, x_ref, **kw)
    n_ref = torch.empty(M, N, device=DEV, dtype=torch.bfloat16)
    L.ln_modulate(x_ref, rows_per_batch=T, nbatch=nb, step_ptr=step, out_bf16=n_ref, **lnkw)
    cnt = L.gemm_ln_counters(T, nb, DEV)
    for rep in range(3):
        x = x0.clone()
        n_out = torch.full((M, N), float("nan"), device=DEV, dtype=torch.bfloat16)
        d = L.gemm(A, W, x, desc_only=True, **kw)
        L.gemm_ln(d, cnt, out_bf16=n_out, **lnkw)
        torch.cuda.synchronize()
        assert int(cnt.abs().sum()) == 0, rep
        if sk:
            assert _rel(x, x_ref) < 1e-6
            assert _rel(n_out, n_ref) < 4e-3
            assert torch.isfinite(n_out.float()).all()
        else:
            assert torch.equal(x, x_ref)
            assert torch.equal(n_out, n_ref)


@pytest.mark.parametrize("N,K,T,nb", [(4096, 1024, 1408, 2), (3072, 768, 300, 3), (4000, 256, 129, 1)])
def test_gemm_tile_width_224(L, N, K, T, nb):
    """224-wide tiles of the 2-SM kernel (engine.up_tile_width: the FeedForward up-projection, modules.py:294-297; the tile's
    columns are drained as 128 + 96 by the two warps of a lane quarter): bit-identical to the 256-wide tiles (same k order per
    output) incl. a last tile that is mostly (N = 4096: 64 of 224 columns) or raggedly (N = 4000) out of range; fp32 reference."""
    M = nb * T
    g = torch.Generator(device=DEV).manual_seed(37)
    A = _bf(torch.randn(M, K, device=DEV, generator=g))
    W = _bf(torch.randn(N, K, device=DEV, generator=g) / math.sqrt(K))
    bias = torch.randn(N, device=DEV, generator=g) * 0.1
    outs = {}
    for bn in (256, 224):
        o = torch.full((M, N), float("nan"), device=DEV, dtype=torch.bfloat16)
        L.gemm(A, W, o, epilogue=L.EPI_BF16, bias=bias, act=L.ACT_GELU_TANH, rows_per_batch=T, nbatch=nb, block_n=bn, two_sm=True)
        outs[bn] = o
    assert torch.equal(outs[224], outs[256])
    ref = F.gelu(A.float() @ W.float().t() + bias, approximate="tanh")
    assert _rel(outs[224], ref) < 6e-3


def test_conv_gemm_grouped_k31_mish(L):
    """ConvPositionEmbedding conv (modules.py:120-141) as implicit GEMM, both epilogues."""
    nb, T, D, G, KS = 2, 256, 1024, 16, 31
    lens = torch.tensor([256, 131], device=DEV, dtype=torch.int32)
    g = torch.Generator(device=DEV).manual_seed(21)
    mask = (torch.arange(T, device=DEV)[None, :] < lens[:, None])  # [nb, T]
    x = torch.randn(nb, T, D, device=DEV, generator=g) * mask[..., None]
    w = torch.randn(D, D // G, KS, device=DEV, generator=g) / math.sqrt(D // G * KS)
    bias = torch.randn(D, device=DEV, generator=g) * 0.1
    xb = _bf(x).reshape(nb * T, D).contiguous()
    wb = _bf(w)
    # tap-major weight [D, 31*64]
    W2 = wb.permute(0, 2, 1).reshape(D, KS * (D // G)).contiguous()
    out = torch.empty(nb * T, D, device=DEV, dtype=torch.bfloat16)
    L.gemm(xb, W2, out, epilogue=L.EPI_MISH_MASK_BF16, bias=bias, rows_per_batch=T, nbatch=nb, taps=KS, cin_blocks=1,
           pad=KS // 2, grouped=64, block_n=64, seq_lens=lens)
    ref = F.conv1d(xb.float().view(nb, T, D).transpose(1, 2), wb.float(), bias, padding=KS // 2, groups=G)
    ref = F.mish(ref.masked_fill(~mask[:, None, :], 0.0)).masked_fill(~mask[:, None, :], 0.0)
    ref = ref.transpose(1, 2).reshape(nb * T, D)
    assert _rel(out, ref) < 8e-3
    add = torch.randn(nb * T, D, device=DEV, generator=g)
    out2 = torch.empty(nb * T, D, device=DEV)
    L.gemm(xb, W2, out2, epilogue=L.EPI_MISH_MASK_RESID, bias=bias, rows_per_batch=T, nbatch=nb, taps=KS,
           cin_blocks=1, pad=KS // 2, grouped=64, block_n=64, seq_lens=lens, addend=add)
    assert _rel(out2, ref + add) < 3e-3


@pytest.mark.parametrize("nb,T,lens", [(2, 256, [256, 131]), (3, 300, [300, 1, 177]), (1, 77, [77]), (2, 1408, [1406, 900])])
def test_conv_gemm_grouped_resident_window(L, monkeypatch, nb, T, lens):
    """The resident-window grouped conv (gconv_res_tcgen05.cuh: the window of 128 + taps - 1 rows is loaded once per tile and
    tap j reads it through a descriptor advanced by j rows) against F.conv1d and, bit for bit, against the generic kernel that
    fetches a shifted A tile per tap (same MMAs in the same order). modules.py:120-141."""
    D, G = 1024, 16
    lens_t = torch.tensor(lens, device=DEV, dtype=torch.int32)
    g = torch.Generator(device=DEV).manual_seed(23)
    mask = torch.arange(T, device=DEV)[None, :] < lens_t[:, None]
    x = torch.randn(nb, T, D, device=DEV, generator=g) * mask[..., None]
    bias = torch.randn(D, device=DEV, generator=g) * 0.1
    xb = _bf(x).reshape(nb * T, D).contiguous()
    for KS in (31, 7, 33):
        w = torch.randn(D, D // G, KS, device=DEV, generator=g) / math.sqrt(D // G * KS)
        wb = _bf(w)
        W2 = wb.permute(0, 2, 1).reshape(D, KS * (D // G)).contiguous()
        ref = F.conv1d(xb.float().view(nb, T, D).transpose(1, 2), wb.float(), bias, padding=KS // 2, groups=G)
        ref_m = F.mish(ref).masked_fill(~mask[:, None, :], 0.0).transpose(1, 2).reshape(nb * T, D)
        ref = ref.transpose(1, 2).reshape(nb * T, D)
        add = torch.randn(nb * T, D, device=DEV, generator=g)
        common = dict(bias=bias, rows_per_batch=T, nbatch=nb, taps=KS, cin_blocks=1, pad=KS // 2, grouped=64, block_n=64)
        outs = {}
        for mode in ("1", "0"):
            monkeypatch.setenv("ORON_GCONV_RES", mode)
            o_m = torch.empty(nb * T, D, device=DEV, dtype=torch.bfloat16)
            L.gemm(xb, W2, o_m, epilogue=L.EPI_MISH_MASK_BF16, seq_lens=lens_t, **common)
            o_r = torch.empty(nb * T, D, device=DEV)
            L.gemm(xb, W2, o_r, epilogue=L.EPI_MISH_MASK_RESID, seq_lens=lens_t, addend=add, **common)
            o_b = torch.empty(nb * T, D, device=DEV, dtype=torch.bfloat16)
            L.gemm(xb, W2, o_b, epilogue=L.EPI_BF16, **common)
            o_f = torch.empty(nb * T, D, device=DEV)
            L.gemm(xb, W2, o_f, epilogue=L.EPI_F32, **common)
            o_s = torch.empty(nb * T, D, device=DEV)
            L.gemm(xb, W2, o_s, epilogue=L.EPI_SCALE_RESID, seq_lens=lens_t, addend=add, **common)
            torch.cuda.synchronize()
            outs[mode] = (o_m, o_r, o_b, o_f, o_s)
        for a, b in zip(outs["1"], outs["0"]):
            assert torch.equal(a, b)
        o_m, o_r, o_b, o_f, o_s = outs["1"]
        assert _rel(o_m, ref_m) < 8e-3
        assert _rel(o_r, ref_m + add) < 3e-3
        assert _rel(o_b, ref) < 8e-3
        assert _rel(o_f, ref) < 1e-4
        ref_s = torch.where(mask.reshape(-1, 1), ref + add, torch.zeros_like(ref))
        assert _rel(o_s, ref_s) < 1e-4


def test_conv_gemm_dense_k7(L):
    """Vocos embed Conv1d(100 -> 512, k=7, pad=3) as a dense implicit GEMM (channels padded to 128)."""
    nb, T, Cin, Cout, KS = 2, 200, 100, 512, 7
    g = torch.Generator(device=DEV).manual_seed(22)
    x = torch.randn(nb, T, Cin, device=DEV, generator=g)
    w = torch.randn(Cout, Cin, KS, device=DEV, generator=g) / math.sqrt(Cin * KS)
    bias = torch.randn(Cout, device=DEV, generator=g) * 0.1
    xb = torch.zeros(nb * T, 128, device=DEV, dtype=torch.bfloat16)
    xb[:, :Cin] = _bf(x).reshape(nb * T, Cin)
    W2 = torch.zeros(Cout, KS, 128, device=DEV, dtype=torch.bfloat16)
    W2[:, :, :Cin] = _bf(w).permute(0, 2, 1)
    W2 = W2.reshape(Cout, KS * 128)
    out = torch.empty(nb * T, Cout, device=DEV)
    L.gemm(xb, W2, out, epilogue=L.EPI_F32, bias=bias, rows_per_batch=T, nbatch=nb, taps=KS, cin_blocks=2, pad=3,
           block_n=128)
    ref = F.conv1d(_bf(x).float().transpose(1, 2), _bf(w).float(), bias, padding=3).transpose(1, 2).reshape(nb * T, Cout)
    assert _rel(out, ref) < 1e-4


def test_gemm_embed_dual_and_scale_resid(L):
    nb, T, D, K = 2, 256, 1024, 128
    M = nb * T
    g = torch.Generator(device=DEV).manual_seed(31)
    A = _bf(torch.randn(M, K, device=DEV, generator=g))
    W = _bf(torch.randn(D, K, device=DEV, generator=g) / math.sqrt(K))
    add = torch.randn(M, D, device=DEV, generator=g)
    lens = torch.tensor([200, 256], device=DEV, dtype=torch.int32)
    o32 = torch.empty(M, D, device=DEV)
    o16 = torch.empty(M, D, device=DEV, dtype=torch.bfloat16)
    L.gemm(A, W, o32, epilogue=L.EPI_EMBED_DUAL, rows_per_batch=T, nbatch=nb, addend=add, seq_lens=lens, out2=o16)
    mask = (torch.arange(T, device=DEV)[None, :] < lens[:, None]).reshape(M, 1)
    ref = torch.where(mask, A.float() @ W.float().t() + add, torch.zeros(M, D, device=DEV))
    assert _rel(o32, ref) < 1e-4
    assert _rel(o16, ref) < 4e-3
    # SCALE_RESID with explicit row mask and per-column scale
    rv = (torch.rand(M, device=DEV, generator=g) > 0.3).to(torch.uint8)
    colscale = torch.randn(D, device=DEV, generator=g)
    bias = torch.randn(D, device=DEV, generator=g)
    o = torch.empty(M, D, device=DEV)
    L.gemm(A, W, o, epilogue=L.EPI_SCALE_RESID, bias=bias, rows_per_batch=T, nbatch=nb, addend=add, gate=colscale,
           row_valid=rv)
    ref2 = torch.where(rv.bool()[:, None], add + colscale * (A.float() @ W.float().t() + bias), torch.zeros_like(add))
    assert _rel(o, ref2) < 1e-4



# ------------------------------------------------------------------------------------------
# 2-SM (cta_group::2) GEMM: same contracts as the 1-SM kernel
# ------------------------------------------------------------------------------------------
@pytest.mark.parametrize(
    "nb,T,N,K,bn",
    [(1, 256, 256, 128, 256), (2, 1408, 3072, 1024, 256), (2, 1408, 1024, 4096, 128), (1, 384, 1024, 1024, 128),
     (3, 200, 1026, 512, 128), (1, 32, 4096, 1024, 256), (2, 1408, 4096, 1024, 256)],
)
def test_gemm_two_sm_plain(L, nb, T, N, K, bn):
    M = nb * T
    g = torch.Generator(device=DEV).manual_seed(M + N + K)
    A = _bf(torch.randn(M, K, device=DEV, generator=g))
    W = _bf(torch.randn(N, K, device=DEV, generator=g) / math.sqrt(K))
    bias = torch.randn(N, device=DEV, generator=g)
    ldo = (N + 7) // 8 * 8
    out = torch.full((M, ldo), float("nan"), device=DEV, dtype=torch.bfloat16)
    L.gemm(A, W, out, epilogue=L.EPI_BF16, bias=bias, block_n=bn, n=N, rows_per_batch=T, nbatch=nb, two_sm=True,
           act=L.ACT_GELU_TANH)
    ref = F.gelu(A.float() @ W.float().t() + bias, approximate="tanh")
    assert torch.isfinite(out[:, :N].float()).all()
    assert _rel(out[:, :N], ref) < 6e-3


def test_gemm_two_sm_fused_epilogues(L):
    nb, T, D, H = 2, 384, 1024, 16
    M = nb * T
    g = torch.Generator(device=DEV).manual_seed(13)
    A = _bf(torch.randn(M, D, device=DEV, generator=g))
    W = _bf(torch.randn(3 * D, D, device=DEV, generator=g) / math.sqrt(D))
    bias = torch.randn(3 * D, device=DEV, generator=g) * 0.1
    inv_freq = 1.0 / (10000 ** (torch.arange(0, 64, 2, device=DEV).float() / 64))
    ang = torch.outer(torch.arange(T, device=DEV).float(), inv_freq)
    cos, sin = ang.cos().contiguous(), ang.sin().contiguous()
    pre = (A.float() @ W.float().t() + bias).view(nb, T, 3, H, 64)
    c = torch.cat([cos, cos], -1)[None, :, None, :]
    s = torch.cat([sin, sin], -1)[None, :, None, :]
    rot = lambda x: torch.cat([-x[..., 32:], x[..., :32]], -1)
    q, k, v = pre[:, :, 0], pre[:, :, 1], pre[:, :, 2]
    ref = torch.stack([q * c + rot(q) * s, k * c + rot(k) * s, v], 2).reshape(M, 3 * D)
    for bn in (128, 256):
        out = torch.empty(M, 3 * D, device=DEV, dtype=torch.bfloat16)
        L.gemm(A, W, out, epilogue=L.EPI_QKV_ROPE, bias=bias, rows_per_batch=T, nbatch=nb, block_n=bn, rope_cos=cos,
               rope_sin=sin, rope_cols=2 * D, two_sm=True)
        assert _rel(out, ref) < 6e-3, bn
    # gated residual with row masking
    Wo = _bf(torch.randn(D, D, device=DEV, generator=g) / math.sqrt(D))
    bo = torch.randn(D, device=DEV, generator=g) * 0.1
    gate = torch.randn(D, device=DEV, generator=g)
    lens = torch.tensor([384, 100], device=DEV, dtype=torch.int32)
    x0 = torch.randn(M, D, device=DEV, generator=g)
    mask = (torch.arange(T, device=DEV)[None, :] < lens[:, None]).reshape(M, 1)
    upd = gate * (A.float() @ Wo.float().t() + bo)
    refx = x0 + torch.where(mask, upd, torch.zeros_like(upd))
    for bn in (128, 256):
        for sk in (False, True):  # stream-K: equal k-block shares per SM pair, partial sums land with f32 reductions
            x = x0.clone()
            L.gemm(A, Wo, x, epilogue=L.EPI_GATE_RESID, bias=bo, rows_per_batch=T, nbatch=nb, gate=gate, seq_lens=lens,
                   mask_rows=True, block_n=bn, two_sm=True, stream_k=sk)
            assert _rel(x, refx) < 2e-3, (bn, sk)
    # f32 with ragged N and addend
    Wp = _bf(torch.randn(100, D, device=DEV, generator=g) / math.sqrt(D))
    bp = torch.randn(100, device=DEV, generator=g)
    o = torch.full((M, 100), float("nan"), device=DEV)
    L.gemm(A, Wp, o, epilogue=L.EPI_F32, bias=bp, rows_per_batch=T, nbatch=nb, block_n=128, two_sm=True)
    assert _rel(o, A.float() @ Wp.float().t() + bp) < 1e-4

# ------------------------------------------------------------------------------------------
# attention
# ------------------------------------------------------------------------------------------
@pytest.mark.parametrize("version", [4, 3])
@pytest.mark.parametrize("nb,T,H,lens,qmul", [(1, 128, 2, None, 1.0), (2, 384, 4, [384, 130], 1.0), (2, 1408, 16, [1406, 1406], 1.0),
                                              (3, 200, 8, [200, 1, 77], 1.0), (2, 1408, 16, [1406, 300], 1.0), (5, 2816, 16, None, 1.0),
                                              (2, 1024, 4, [1000, 515], 6.0)])
def test_attention(L, nb, T, H, lens, qmul, version):
    """version 4 = attn_fwd4.cuh (production), 3 = the round-1 kernel kept for A/B. qmul > 1 spreads the scores so that the
    running maximum moves by more than 2^8 between key tiles (the lazy-rescale branch)."""
    g = torch.Generator(device=DEV).manual_seed(T + H)
    qkv = torch.randn(nb * T, 3 * H * 64, device=DEV, generator=g)
    qkv[:, : H * 64] *= qmul
    qkv = _bf(qkv)
    # the V third is IEEE f16 (bit-cast into the bf16-typed buffer), as the QKV GEMM epilogue writes it
    v16 = torch.randn(nb * T, H * 64, device=DEV, generator=g).half()
    qkv[:, 2 * H * 64:] = v16.view(torch.bfloat16)
    lens_t = torch.tensor(lens, device=DEV, dtype=torch.int32) if lens is not None else None
    x = qkv.float().view(nb, T, 3, H, 64)
    q, k = (x[:, :, i].transpose(1, 2) for i in range(2))
    v = v16.float().view(nb, T, H, 64).transpose(1, 2)
    ll = lens if lens is not None else [T] * nb
    mask = torch.arange(T, device=DEV)[None, :] < torch.tensor(ll, device=DEV)[:, None]
    ref = F.scaled_dot_product_attention(q, k, v, attn_mask=mask[:, None, None, :])
    ref = ref.transpose(1, 2).reshape(nb, T, H * 64)
    L.lib().oron_debug_set_attention_version(version)
    try:
        # schedules: -1 = planned workspace (equal shares of the (item, key tile) list when there are more items than
        # CTA slots, else one CTA per item), 1 = shares forced on the small shapes too (many parts per item, empty
        # shares; split items combined in-kernel), None = no workspace: one CTA per item. Each runs twice on the same
        # workspace (the arrival counters must come back to zero).
        for sched in (-1, 1, None):
            out = torch.zeros(nb * T, H * 64, device=DEV, dtype=torch.bfloat16)
            if sched is not None:
                L.lib().oron_debug_set_attention_schedule(sched)
                ws = L.attention_workspace(nb, T, H, DEV, seq_lens=lens_t)
            else:
                ws = None
            for _ in range(2):
                out.zero_()
                L.attention(qkv, out, nbatch=nb, rows_per_batch=T, heads=H, seq_lens=lens_t, scale=0.125, workspace=ws)
            o = out.view(nb, T, H * 64)
            for b in range(nb):
                assert _rel(o[b, : ll[b]], ref[b, : ll[b]]) < 1e-2, (sched, b)
    finally:
        L.lib().oron_debug_set_attention_schedule(-1)
        L.lib().oron_debug_set_attention_version(4)


# ------------------------------------------------------------------------------------------
# row-wise kernels
# ------------------------------------------------------------------------------------------
@pytest.mark.parametrize("C", [512, 1024])
def test_ln_modulate(L, C):
    nb, T = 2, 300
    g = torch.Generator(device=DEV).manual_seed(C)
    x = torch.randn(nb * T, C, device=DEV, generator=g) * 3 + 1
    table = torch.randn(4, nb, 6 * C, device=DEV, generator=g)
    step = torch.tensor([3], device=DEV, dtype=torch.int32)
    out = torch.empty(nb * T, C, device=DEV, dtype=torch.bfloat16)
    flat = table.view(-1)
    L.ln_modulate(x, rows_per_batch=T, nbatch=nb, eps=1e-6, scale=flat[C:], shift=flat, mod_ld=6 * C, mod_nb=nb,
                  step_stride=nb * 6 * C, step_ptr=step, add_one=True, out_bf16=out)
    shift, scale = table[3, :, :C], table[3, :, C:2 * C]
    ref = F.layer_norm(x.view(nb, T, C), (C,), eps=1e-6) * (1 + scale[:, None]) + shift[:, None]
    assert _rel(out.view(nb, T, C), ref) < 4e-3
    w, b = torch.randn(C, device=DEV, generator=g), torch.randn(C, device=DEV, generator=g)
    o32 = torch.empty(nb * T, C, device=DEV)
    L.ln_modulate(x, rows_per_batch=T, nbatch=nb, eps=1e-6, scale=w, shift=b, add_one=False, out_f32=o32)
    assert _rel(o32, F.layer_norm(x, (C,), w, b, eps=1e-6)) < 1e-5


def test_cfg_euler_step(L):
    nb, T, n_mels, steps = 2, 130, 100, 4
    rows = nb * T
    g = torch.Generator(device=DEV).manual_seed(1)
    x0 = torch.randn(rows, n_mels, device=DEV, generator=g)
    v = torch.randn(2 * rows, 128, device=DEV, generator=g)
    dt = torch.rand(steps, device=DEV, generator=g)
    step = torch.tensor([1], device=DEV, dtype=torch.int32)
    xb = torch.zeros(2 * rows, 128, device=DEV, dtype=torch.bfloat16)
    traj = torch.zeros(steps + 1, rows, n_mels, device=DEV)
    vout = torch.empty(rows, n_mels, device=DEV)
    x = x0.clone()
    L.cfg_euler_step(x, v, nb=nb, rows_per_batch=T, n_mels=n_mels, has_uncond=True, cfg=2.0, dt=dt, step_ptr=step,
                     xb=xb, traj=traj, v_out=vout)
    vc, vu = v[:rows, :n_mels], v[rows:, :n_mels]
    vg = vc + (vc - vu) * 2.0
    ref = x0 + vg * dt[1]
    assert torch.allclose(x, ref, atol=1e-6)
    assert torch.allclose(vout, vg, atol=1e-6)
    assert torch.equal(traj[2], x)
    assert int(step.item()) == 2
    assert torch.equal(xb[:rows, :n_mels], x.to(torch.bfloat16)) and torch.equal(xb[rows:, :n_mels], x.to(torch.bfloat16))
    assert float(xb[:, n_mels:].abs().max()) == 0.0


def test_time_sinusoid(L):
    t = torch.linspace(0, 1, 9, device=DEV)
    out = torch.empty(9, 256, device=DEV, dtype=torch.bfloat16)
    L.time_sinusoid(t, out)
    emb = torch.exp(torch.arange(128, device=DEV).float() * -(math.log(10000) / 127))
    e = 1000.0 * t[:, None] * emb[None]
    ref = torch.cat([e.sin(), e.cos()], -1)
    assert float((out.float() - ref).abs().max()) < 1e-2


def test_text_front_dwconv_grn(L):
    nb, T, C = 2, 140, 512
    g = torch.Generator(device=DEV).manual_seed(4)
    ids = torch.randint(1, 66, (nb, T), device=DEV, generator=g, dtype=torch.int32)
    ids[0, 100:] = 0
    ids[1, 10:20] = 0
    drop = torch.tensor([0, 1], device=DEV, dtype=torch.uint8)
    table = torch.randn(66, C, device=DEV, generator=g)
    pos = torch.randn(T, C, device=DEV, generator=g)
    x = torch.empty(nb * T, C, device=DEV)
    rv = torch.empty(nb * T, device=DEV, dtype=torch.uint8)
    L.text_embed_front(ids.view(-1), drop, table, pos, rows_per_batch=T, nb=nb, x=x, row_valid=rv)
    look = ids.clone().long()
    look[1] = 0
    ref = table[look] + pos[None]
    ref = ref.masked_fill((ids == 0)[..., None], 0.0)
    assert torch.equal(x.view(nb, T, C), ref)
    assert torch.equal(rv.view(nb, T).bool(), ids != 0)
    # dwconv7 + LN
    lens = torch.tensor([140, 90], device=DEV, dtype=torch.int32)
    w = torch.randn(C, 1, 7, device=DEV, generator=g) * 0.3
    wb, lw, lb = (torch.randn(C, device=DEV, generator=g) for _ in range(3))
    out = torch.empty(nb * T, C, device=DEV, dtype=torch.bfloat16)
    L.dwconv7_ln(x, rows_per_batch=T, nbatch=nb, seq_lens=lens, w=w.view(C, 7).contiguous(), wb=wb, ln_w=lw, ln_b=lb,
                 eps=1e-6, out=out)
    m = (torch.arange(T, device=DEV)[None, :] < lens[:, None])
    xin = x.view(nb, T, C) * m[..., None]
    y = F.conv1d(xin.transpose(1, 2), w, wb, padding=3, groups=C).transpose(1, 2)
    refln = F.layer_norm(y, (C,), lw, lb, eps=1e-6)
    o = out.view(nb, T, C).float()
    for b in range(nb):
        assert _rel(o[b, : lens[b]], refln[b, : lens[b]]) < 4e-3
    # GRN over valid frames only
    C2 = 1024
    h = _bf(torch.randn(nb * T, C2, device=DEV, generator=g))
    gamma, beta = torch.randn(C2, device=DEV, generator=g), torch.randn(C2, device=DEV, generator=g)
    gx2 = torch.empty(nb, C2, device=DEV)
    h2 = h.clone()
    L.grn(h2, rows_per_batch=T, nb=nb, seq_lens=lens, gamma=gamma, beta=beta, gx2=gx2)
    for b in range(nb):
        hb = h.view(nb, T, C2)[b, : lens[b]].float()[None]
        gx = torch.norm(hb, p=2, dim=1, keepdim=True)
        nx = gx / (gx.mean(dim=-1, keepdim=True) + 1e-6)
        refg = gamma * (hb * nx) + beta + hb
        assert _rel(h2.view(nb, T, C2)[b, : lens[b]], refg[0]) < 5e-3


@pytest.mark.parametrize("nb,T,C,aligned", [(1, 37, 512, True), (3, 301, 512, True), (2, 75, 256, True), (2, 75, 512, False),
                                            (4, 2813, 512, True)])
def test_dwconv7_ln_runs(L, nb, T, C, aligned):
    """Sliding-window kernel (aligned rows) and the warp-per-row fallback (rows offset by one float) against
    torch conv1d + layer_norm; odd run lengths, ragged seq_lens, several runs per sequence."""
    g = torch.Generator(device=DEV).manual_seed(40 + T)
    lens = torch.tensor([T if b % 2 == 0 else max(4, T - 11 * b) for b in range(nb)], device=DEV, dtype=torch.int32)
    ld = C + (0 if aligned else 4)
    buf = torch.randn(nb * T * ld + 4, device=DEV, generator=g)
    x = buf[(0 if aligned else 1):][: nb * T * ld].view(nb * T, ld)[:, :C]
    w = torch.randn(C, 1, 7, device=DEV, generator=g) * 0.3
    wb, lw, lb = (torch.randn(C, device=DEV, generator=g) for _ in range(3))
    out = torch.empty(nb * T, C, device=DEV, dtype=torch.bfloat16)
    L.dwconv7_ln(x, rows_per_batch=T, nbatch=nb, seq_lens=lens, w=w.view(C, 7).contiguous(), wb=wb, ln_w=lw, ln_b=lb,
                 eps=1e-6, out=out)
    m = torch.arange(T, device=DEV)[None, :] < lens[:, None]
    xin = x.reshape(nb, T, C) * m[..., None]
    y = F.conv1d(xin.transpose(1, 2), w, wb, padding=3, groups=C).transpose(1, 2)
    ref = F.layer_norm(y, (C,), lw, lb, eps=1e-6)
    o = out.view(nb, T, C).float()
    for b in range(nb):
        assert _rel(o[b, : lens[b]], ref[b, : lens[b]]) < 4e-3


# ------------------------------------------------------------------------------------------
# audio
# ------------------------------------------------------------------------------------------
def _mel_fb(n_freqs=513, n_mels=100, sr=24000):
    import torchaudio

    return torchaudio.functional.melscale_fbanks(n_freqs, 0.0, sr / 2, n_mels, sr, norm=None, mel_scale="htk")


@pytest.mark.parametrize("nb,S", [(1, 48000), (3, 120000), (2, 7777)])
def test_logmel(L, nb, S):
    import torchaudio

    g = torch.Generator(device=DEV).manual_seed(S)
    wav = (torch.rand(nb, S, device=DEV, generator=g) * 2 - 1) * 0.3
    wav[0] += 0.5 * torch.sin(torch.arange(S, device=DEV) * (2 * math.pi * 220 / 24000))
    window = torch.hann_window(1024, device=DEV)
    fb = _mel_fb().to(DEV).contiguous()
    T = 1 + S // 256
    out = torch.empty(nb, 100, T, device=DEV)
    L.logmel(wav, window, fb, out, clip=1e-5)
    tr = torchaudio.transforms.MelSpectrogram(sample_rate=24000, n_fft=1024, hop_length=256, win_length=1024,
                                              n_mels=100, center=True, power=1).to(DEV)
    ref = torch.log(torch.clamp(tr(wav), min=1e-5))
    assert out.shape == ref.shape
    assert float((out - ref).abs().max()) < 2e-3
    assert _rel(out, ref) < 1e-4


@pytest.mark.parametrize("nb,T,mode,pad", [(1, 37, 0, 0), (2, 300, 0, 0), (2, 64, 1, 0), (1, 2, 0, 0), (5, 33, 1, 7), (2400, 4, 0, 0),
                                           (3, 301, 0, 11), (64, 2813, 0, 0)])
def test_istft_head(L, nb, T, mode, pad):
    """Against torch.istft: odd / even frame counts, the two-frame minimum, more clips than warps in the launch (one run
    per clip), rows_per_batch > n_frames (padded activations), several runs per clip with their halo frames, and the
    config-4 size (64 x 2813 frames: 77-hop runs)."""
    g = torch.Generator(device=DEV).manual_seed(T)
    ld = 1056
    R = T + pad  # rows per clip in the activation buffer; only the first T are frames
    h = torch.randn(nb * R, ld, device=DEV, generator=g)
    window = torch.hann_window(1024, device=DEV)
    out = torch.empty(nb, (T - 1) * 256, device=DEV)
    L.istft_head(h, window, out, rows_per_batch=R, nb=nb, n_frames=T, mode=mode)
    # float64 reference: torch.istft in fp32 is itself off by 1e-2 for batches of several hundred clips (cuFFT plan choice)
    hv = h.view(nb, R, ld)[:, :T].double()
    w64 = window.double()
    if mode == 0:
        mag = torch.clip(torch.exp(hv[..., :513]), max=1e2)
        p = hv[..., 513:1026]
        spec = (mag * (torch.cos(p) + 1j * torch.sin(p))).transpose(1, 2)
        ref = torch.istft(spec, 1024, 256, 1024, w64, center=True, normalized=False)
    else:
        ri = hv[..., :1026].reshape(nb, T, 513, 2)
        spec = torch.complex(ri[..., 0], ri[..., 1]).transpose(1, 2)
        ref = torch.istft(spec, 1024, 256, 1024, w64, normalized=True, onesided=True)
    ref = ref.float()
    assert out.shape == ref.shape
    assert _rel(out, ref) < 1e-5
    assert float((out - ref).abs().max()) < 5e-5 * float(ref.abs().max())


def test_peak_normalize(L):
    g = torch.Generator(device=DEV).manual_seed(8)
    x = torch.randn(3, 50001, device=DEV, generator=g) * 0.2
    x[1] = 0.0
    out = torch.empty_like(x)
    scratch = torch.empty(3, device=DEV)
    L.peak_normalize(x, out, scratch)
    for b in range(3):
        mx = x[b].abs().max()
        ref = x[b] if mx < 1e-8 else torch.clamp(x[b] / (mx + 1e-7), -1.0, 1.0)
        assert torch.allclose(out[b], ref, atol=1e-7)


@pytest.mark.parametrize("nb,T,N,K,bn", [(2, 1408, 1024, 4096, 256), (2, 1408, 1024, 1024, 256), (3, 200, 512, 1536, 128),
                                         (1, 130, 1024, 1024, 256)])
def test_gemm_stream_k(L, nb, T, N, K, bn):
    """Stream-K gated residual at the config-2 out-proj / FFN-down shapes, a ragged case with phantom m-tiles and a
    tiny one (fewer tiles than SM pairs): x += gate * (A W^T + bias), rows beyond seq_len untouched."""
    M = nb * T
    g = torch.Generator(device=DEV).manual_seed(M + N + K + 1)
    A = _bf(torch.randn(M, K, device=DEV, generator=g))
    W = _bf(torch.randn(N, K, device=DEV, generator=g) / math.sqrt(K))
    bias = torch.randn(N, device=DEV, generator=g) * 0.1
    gate = torch.randn(N, device=DEV, generator=g)
    lens = torch.tensor([T if b % 2 == 0 else max(1, T - 37) for b in range(nb)], device=DEV, dtype=torch.int32)
    x0 = torch.randn(M, N, device=DEV, generator=g)
    mask = (torch.arange(T, device=DEV)[None, :] < lens[:, None]).reshape(M, 1)
    upd = gate * (A.float() @ W.float().t() + bias)
    ref = x0 + torch.where(mask, upd, torch.zeros_like(upd))
    for _ in range(2):
        x = x0.clone()
        L.gemm(A, W, x, epilogue=L.EPI_GATE_RESID, bias=bias, rows_per_batch=T, nbatch=nb, gate=gate, seq_lens=lens,
               mask_rows=True, block_n=bn, two_sm=True, stream_k=True)
        assert _rel(x, ref) < 2e-3
        assert torch.equal(x[~mask.squeeze(1)], x0[~mask.squeeze(1)])


@pytest.mark.parametrize("nb,T,D,F", [(2, 1408, 1024, 4096), (1, 130, 512, 1024), (3, 640, 1024, 2048), (1, 1408, 256, 512)])
def test_ffn_fused(L, nb, T, D, F):
    """FeedForward in one launch (oron_ffn_bf16) against fp32 torch and against the two separate launches: config-2 shape,
    an odd m-tile count (phantom tile), more phase-1 tiles than three waves, and fewer units than SM pairs. Replayed three
    times on one workspace: the kernel must leave its flags zeroed."""
    M = nb * T
    g = torch.Generator(device=DEV).manual_seed(M + D + F)
    A = _bf(torch.randn(M, D, device=DEV, generator=g))
    W1 = _bf(torch.randn(F, D, device=DEV, generator=g) / math.sqrt(D))
    W2 = _bf(torch.randn(D, F, device=DEV, generator=g) / math.sqrt(F))
    b1 = torch.randn(F, device=DEV, generator=g) * 0.1
    b2 = torch.randn(D, device=DEV, generator=g) * 0.1
    gate = torch.randn(D, device=DEV, generator=g)
    x0 = torch.randn(M, D, device=DEV, generator=g)
    h_ref = _bf(torch.nn.functional.gelu(A.float() @ W1.float().t() + b1, approximate="tanh"))
    ref = x0 + gate * (h_ref.float() @ W2.float().t() + b2)
    ws = L.ffn_workspace(T, nb, F, DEV)
    hid = torch.empty(M, F, device=DEV, dtype=torch.bfloat16)
    for _ in range(3):
        x = x0.clone()
        hid.zero_()
        up = L.gemm(A, W1, hid, epilogue=L.EPI_BF16, bias=b1, act=L.ACT_GELU_TANH, rows_per_batch=T, nbatch=nb, block_n=256,
                    two_sm=True, desc_only=True)
        dn = L.gemm(hid, W2, x, epilogue=L.EPI_GATE_RESID, bias=b2, gate=gate, rows_per_batch=T, nbatch=nb, block_n=256,
                    two_sm=True, desc_only=True)
        L.ffn(up, dn, ws)
        torch.cuda.synchronize()
        assert _rel(hid.float(), h_ref.float()) < 6e-3
        assert _rel(x, ref) < 3e-3
        assert int(ws.view(torch.int32).abs().sum()) == 0
    # the two-launch path computes the same thing
    x2, hid2 = x0.clone(), torch.empty_like(hid)
    L.gemm(A, W1, hid2, epilogue=L.EPI_BF16, bias=b1, act=L.ACT_GELU_TANH, rows_per_batch=T, nbatch=nb, block_n=256, two_sm=True)
    L.gemm(hid2, W2, x2, epilogue=L.EPI_GATE_RESID, bias=b2, gate=gate, rows_per_batch=T, nbatch=nb, block_n=256, two_sm=True)
    assert torch.equal(hid, hid2)
    assert _rel(x, x2) < 1e-5
