"""Kernel-level numerics of the training-step entry points (include/oron_b200_train.h) against torch autograd (fp32)
of the same op on the same inputs. Tolerances reflect bf16 operands / outputs where the ABI uses them."""

import math

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

from oron_tts_b200 import _lib as L  # noqa: E402
from oron_tts_b200 import _lib_train as T  # noqa: E402

DEV = "cuda"
BF16, F32 = torch.bfloat16, torch.float32


def _rel(a, b):
    a, b = a.detach().float().cpu(), b.detach().float().cpu()
    return float((a - b).norm() / (b.norm() + 1e-12))


def _mask(nb, rpb, lens):
    return (torch.arange(rpb, device=DEV)[None, :] < torch.tensor(lens, device=DEV)[:, None]).reshape(nb * rpb)


def test_transpose_mask_colsum():
    g = torch.Generator(device=DEV).manual_seed(0)
    nb, rpb, C = 2, 192, 200
    x = torch.randn(nb * rpb, 256, device=DEV, generator=g).to(BF16)[:, :C]
    lens = [192, 77]
    sl = torch.tensor(lens, device=DEV, dtype=torch.int32)
    out = torch.full((C, nb * rpb), 7.0, device=DEV, dtype=BF16)
    cs = torch.zeros(C, device=DEV)
    T.transpose(x, out, rows_per_batch=rpb, nbatch=nb, seq_lens=sl, colsum=cs)
    xm = x.float() * _mask(nb, rpb, lens)[:, None]
    assert torch.equal(out.float(), xm.t())
    assert _rel(cs, xm.sum(0)) < 1e-5
    out2 = torch.empty(C, nb * rpb, device=DEV, dtype=BF16)
    T.transpose(x, out2, rows_per_batch=rpb, nbatch=nb)
    assert torch.equal(out2.float(), x.float().t())


@pytest.mark.parametrize("C,affine", [(1024, False), (128, False), (512, True), (64, True)])
def test_ln_bwd(C, affine):
    g = torch.Generator(device=DEV).manual_seed(1)
    nb, rpb = 2, 128
    lens = [128, 70]
    sl = torch.tensor(lens, device=DEV, dtype=torch.int32)
    x = torch.randn(nb * rpb, C, device=DEV, generator=g) * 2 + 0.3
    dy = (torch.randn(nb * rpb, C, device=DEV, generator=g) * 0.1).to(BF16)
    m = _mask(nb, rpb, lens)
    if affine:
        scale = (1 + 0.1 * torch.randn(1, C, device=DEV, generator=g)).requires_grad_()
        shift = torch.zeros(1, C, device=DEV, requires_grad=True)
        mult = scale.expand(nb, C)
    else:
        scale = (0.2 * torch.randn(nb, C, device=DEV, generator=g)).requires_grad_()
        shift = torch.zeros(nb, C, device=DEV, requires_grad=True)
        mult = 1 + scale
    xr = x.clone().requires_grad_()
    y = F.layer_norm(xr, (C,), eps=1e-6).view(nb, rpb, C) * mult[:, None] + shift.expand(nb, C)[:, None]
    (y.reshape(-1, C) * dy.float() * m[:, None]).sum().backward()
    dx0 = torch.randn(nb * rpb, C, device=DEV, generator=g)
    dx = dx0.clone()
    ds = torch.zeros_like(scale.detach())
    dh = torch.zeros_like(shift.detach())
    T.ln_bwd(x, dy, rows_per_batch=rpb, nbatch=nb, eps=1e-6, scale=scale.detach(), mod_ld=0 if affine else C,
             add_one=not affine, seq_lens=sl, dx=dx, accumulate=True, dscale=ds, dshift=dh, dmod_ld=0 if affine else C)
    assert _rel(dx - dx0, xr.grad) < 1e-4
    assert _rel(ds, scale.grad) < 1e-4 and _rel(dh, shift.grad) < 1e-4
    dx2 = torch.full_like(dx0, 3.0)
    T.ln_bwd(x, dy, rows_per_batch=rpb, nbatch=nb, eps=1e-6, scale=scale.detach(), mod_ld=0 if affine else C,
             add_one=not affine, seq_lens=sl, dx=dx2, accumulate=False, dscale=None, dshift=None, dmod_ld=0)
    assert _rel(dx2, xr.grad) < 1e-4 and float(dx2[~m].abs().max()) == 0.0


@pytest.mark.parametrize("act,fn", [(L.ACT_GELU_TANH, lambda x: F.gelu(x, approximate="tanh")), (L.ACT_GELU_ERF, F.gelu),
                                    (L.ACT_SILU, F.silu), (T.ACT_MISH, F.mish)])
def test_act_fwd_bwd(act, fn):
    g = torch.Generator(device=DEV).manual_seed(2)
    x = (torch.randn(300, 256, device=DEV, generator=g) * 3)
    dy = torch.randn(300, 256, device=DEV, generator=g)
    xr = x.clone().requires_grad_()
    y = fn(xr)
    y.backward(dy)
    out = torch.empty_like(x)
    T.act_fwd(x, out, act)
    assert float((out - y).abs().max()) < 4e-3
    d = torch.empty_like(x)
    T.act_bwd(dy, x, d, act)
    assert _rel(d, xr.grad) < 1e-4
    xb, dyb = x.to(BF16), dy.to(BF16)
    ob = torch.empty_like(xb)
    T.act_fwd(xb, ob, act)
    assert _rel(ob, fn(xb.float())) < 6e-3
    db = torch.empty_like(xb)
    T.act_bwd(dyb, xb, db, act)
    xr2 = xb.float().requires_grad_()
    fn(xr2).backward(dyb.float())
    assert _rel(db, xr2.grad) < 6e-3
    # row mask (rows t >= seq_lens[b] are written as zeros)
    sl = torch.tensor([100, 150], device=DEV, dtype=torch.int32)
    m = _mask(2, 150, [100, 150])
    T.act_fwd(xb, ob, act, rows_per_batch=150, seq_lens=sl)
    T.act_bwd(dyb, xb, db, act, rows_per_batch=150, seq_lens=sl)
    assert float(ob[~m].float().abs().max()) == 0.0 and float(db[~m].float().abs().max()) == 0.0
    assert _rel(db, xr2.grad * m[:, None]) < 6e-3 and _rel(ob, fn(xb.float()) * m[:, None]) < 6e-3


def test_gate_resid_and_bwd():
    g = torch.Generator(device=DEV).manual_seed(3)
    nb, rpb, C = 2, 128, 256
    lens = [100, 128]
    sl = torch.tensor(lens, device=DEV, dtype=torch.int32)
    m = _mask(nb, rpb, lens)
    x = torch.randn(nb * rpb, C, device=DEV, generator=g)
    y = torch.randn(nb * rpb, C, device=DEV, generator=g).to(BF16)
    gate = torch.randn(nb, 3 * C, device=DEV, generator=g)[:, C:2 * C]
    x1 = x.clone()
    T.gate_resid(x1, y, rows_per_batch=rpb, nbatch=nb, gate=gate, gate_ld=3 * C, seq_lens=sl, mask_rows=True)
    ref = x + (gate[:, None, :] * y.float().view(nb, rpb, C)).reshape(-1, C) * m[:, None]
    assert _rel(x1, ref) < 1e-6
    x2 = torch.empty_like(x)
    T.gate_resid(x, y, rows_per_batch=rpb, nbatch=nb, gate=gate, gate_ld=3 * C, seq_lens=sl, mask_rows=True, out=x2)
    assert torch.equal(x2, x1)  # out of place (the training forward writes the next saved residual buffer)
    dx = torch.randn(nb * rpb, C, device=DEV, generator=g)
    dy = torch.empty(nb * rpb, C, device=DEV, dtype=BF16)
    dg = torch.zeros(nb, C, device=DEV)
    dbias = torch.zeros(C, device=DEV)
    T.gate_bwd(dx, y, rows_per_batch=rpb, nbatch=nb, gate=gate, gate_ld=3 * C, seq_lens=sl, dy=dy, dgate=dg, dgate_ld=C, dbias=dbias)
    dxm = dx * m[:, None]
    assert _rel(dbias, (gate[:, None, :] * dxm.view(nb, rpb, C)).sum((0, 1))) < 1e-5
    assert _rel(dy, (gate[:, None, :] * dxm.view(nb, rpb, C)).reshape(-1, C)) < 4e-3
    assert _rel(dg, (dxm * y.float()).view(nb, rpb, C).sum(1)) < 1e-5


def test_dwconv7_fwd_flip_wgrad():
    g = torch.Generator(device=DEV).manual_seed(4)
    nb, rpb, C = 2, 96, 64
    lens = [96, 50]
    sl = torch.tensor(lens, device=DEV, dtype=torch.int32)
    m = _mask(nb, rpb, lens).view(nb, rpb, 1)
    x = torch.randn(nb, rpb, C, device=DEV, generator=g) * m
    w = torch.randn(C, 1, 7, device=DEV, generator=g).requires_grad_()
    b = torch.randn(C, device=DEV, generator=g).requires_grad_()
    xr = x.clone().requires_grad_()
    # per-sequence zero padding == conv over the masked input restricted to valid rows
    y = torch.stack([F.pad(F.conv1d(xr[i, :lens[i]].t()[None], w, b, padding=3, groups=C)[0].t(), (0, 0, 0, rpb - lens[i]))
                     for i in range(nb)])
    dy = torch.randn(nb, rpb, C, device=DEV, generator=g) * m
    y.backward(dy)
    out = torch.empty(nb * rpb, C, device=DEV)
    T.dwconv7(x.view(-1, C), out, rows_per_batch=rpb, nbatch=nb, seq_lens=sl, w=w.detach().view(C, 7).contiguous(), bias=b.detach())
    assert _rel(out.view(nb, rpb, C) * m, y) < 1e-5
    dxo = torch.empty(nb * rpb, C, device=DEV)
    T.dwconv7(dy.view(-1, C).contiguous(), dxo, rows_per_batch=rpb, nbatch=nb, seq_lens=sl, w=w.detach().view(C, 7).contiguous(),
              bias=None, flip=True)
    assert _rel(dxo.view(nb, rpb, C) * m, xr.grad) < 1e-5
    dw = torch.zeros(C, 7, device=DEV)
    db = torch.zeros(C, device=DEV)
    T.dwconv7_wgrad(x.view(-1, C), dy.view(-1, C).contiguous(), rows_per_batch=rpb, nbatch=nb, seq_lens=sl, dw=dw, db=db)
    assert _rel(dw, w.grad.view(C, 7)) < 1e-5 and _rel(db, b.grad) < 1e-5


def test_grn_bwd():
    g = torch.Generator(device=DEV).manual_seed(5)
    nb, rpb, C = 2, 128, 128
    lens = [128, 128]
    sl = torch.tensor(lens, device=DEV, dtype=torch.int32)
    pre = (torch.randn(nb * rpb, C, device=DEV, generator=g)).to(BF16)
    gamma = (torch.randn(C, device=DEV, generator=g) * 0.5).requires_grad_()
    beta = (torch.randn(C, device=DEV, generator=g) * 0.5).requires_grad_()
    dy = (torch.randn(nb * rpb, C, device=DEV, generator=g) * 0.1).to(BF16)
    pr = pre.float().requires_grad_()
    h = F.gelu(pr).view(nb, rpb, C)
    gx = torch.linalg.vector_norm(h, ord=2, dim=1, keepdim=True)
    nx = gx / (gx.mean(dim=-1, keepdim=True) + 1e-6)
    y = gamma * (h * nx) + beta + h
    (y.reshape(-1, C) * dy.float()).sum().backward()
    # forward pieces through the product kernels: gx2 from oron_grn on bf16 h
    hb = torch.empty_like(pre)
    T.act_fwd(pre, hb, L.ACT_GELU_ERF)
    gx2 = torch.zeros(nb, C, device=DEV)
    L.grn(hb, rows_per_batch=rpb, nb=nb, seq_lens=sl, gamma=gamma.detach(), beta=beta.detach(), gx2=gx2)
    assert _rel(hb, y.reshape(-1, C)) < 1e-2
    A, nxo, coef = (torch.zeros(nb, C, device=DEV) for _ in range(3))
    dgam, dbet = torch.zeros(C, device=DEV), torch.zeros(C, device=DEV)
    dpre = torch.empty_like(pre)
    T.grn_bwd(dy, pre, dpre, rows_per_batch=rpb, nb=nb, seq_lens=sl, gamma=gamma.detach(), gx2=gx2, A=A, nx=nxo, coef=coef,
              dgamma=dgam, dbeta=dbet)
    assert _rel(dpre, pr.grad) < 1e-2
    assert _rel(dgam, gamma.grad) < 1e-2 and _rel(dbet, beta.grad) < 1e-4


def test_text_embed_bwd():
    g = torch.Generator(device=DEV).manual_seed(6)
    nb, rpb, C, V = 2, 64, 64, 66
    ids = torch.randint(0, V, (nb * rpb,), device=DEV, generator=g, dtype=torch.int32)
    drop = torch.tensor([0, 1], device=DEV, dtype=torch.uint8)
    dx = torch.randn(nb * rpb, C, device=DEV, generator=g)
    dt = torch.zeros(V, C, device=DEV)
    T.text_embed_bwd(ids, drop, dx, dt, rows_per_batch=rpb, nb=nb)
    ref = torch.zeros(V, C, device=DEV)
    eff = ids.long().clone()
    eff[rpb:] = 0
    keep = ids != 0
    ref.index_add_(0, eff[keep], dx[keep])
    assert _rel(dt, ref) < 1e-5


@pytest.mark.parametrize("nb,N,K", [(5, 2300, 384), (8, 22 * 6144 + 2048, 1024), (3, 700, 250), (12, 515, 512)])
def test_skinny_dgrad_wgrad(nb, N, K):
    """(5, 2300, 384) and the stacked AdaLN matrix of the Base model go through the 16-byte-load kernel, K % 8 != 0 and
    nb > 8 through the 4-byte one."""
    g = torch.Generator(device=DEV).manual_seed(7)
    dY = torch.randn(nb, N, device=DEV, generator=g)
    W = torch.randn(N, K, device=DEV, generator=g).to(BF16)
    X = torch.randn(nb, K, device=DEV, generator=g)
    dX = torch.zeros(nb, K, device=DEV)
    T.skinny_dgrad(dY, W, dX)
    assert _rel(dX, dY @ W.float()) < 1e-5
    dW = torch.ones(N, K, device=DEV)
    db = torch.ones(N, device=DEV)
    T.skinny_wgrad(dY, X, dW, db, accumulate=True)
    assert _rel(dW, 1 + dY.t() @ X) < 1e-5 and _rel(db, 1 + dY.sum(0)) < 1e-5
    T.skinny_wgrad(dY, X, dW, None, accumulate=False)
    assert _rel(dW, dY.t() @ X) < 1e-5


@pytest.mark.parametrize("C,cg", [(128, 8), (1024, 64)])
def test_gconv_wgrad(C, cg):
    g = torch.Generator(device=DEV).manual_seed(8)
    nb, rpb, taps = 2, 128, 31
    lens = [128, 60]
    sl = torch.tensor(lens, device=DEV, dtype=torch.int32)
    m = _mask(nb, rpb, lens).view(nb, rpb, 1)
    x = (torch.randn(nb, rpb, C, device=DEV, generator=g) * m).to(BF16)
    dy = (torch.randn(nb, rpb, C, device=DEV, generator=g) * 0.1 * m).to(BF16)
    w = torch.zeros(C, cg, taps, device=DEV, requires_grad=True)
    b = torch.zeros(C, device=DEV, requires_grad=True)
    y = F.conv1d(x.float().transpose(1, 2), w, b, padding=taps // 2, groups=C // cg).transpose(1, 2)
    (y * dy.float()).sum().backward()
    dw = torch.zeros(C, cg, taps, device=DEV)
    db = torch.zeros(C, device=DEV)
    T.gconv_wgrad(x.view(-1, C), dy.view(-1, C), rows_per_batch=rpb, nbatch=nb, cg=cg, taps=taps, seq_lens=sl, dw=dw, db=db)
    assert _rel(dw, w.grad) < 1e-4 and _rel(db, b.grad) < 1e-4


@pytest.mark.parametrize("C,cg,nb,rpb", [(128, 8, 2, 128), (1024, 64, 2, 128), (512, 32, 3, 320), (1024, 64, 8, 1024)])
def test_gconv_wgrad_tensor_core(C, cg, nb, rpb):
    """tcgen05 weight gradient of the grouped k = 31 conv against autograd (fp32 conv on the bf16-rounded operands) and the
    CUDA-core kernel; ragged lengths (rows beyond the length are zero in both operands, as the training step guarantees),
    accumulation into a non-zero dw, and the config-5 size."""
    g = torch.Generator(device=DEV).manual_seed(C + nb)
    taps = 31
    lens = [rpb if b % 2 == 0 else max(1, rpb - 68) for b in range(nb)]
    m = _mask(nb, rpb, lens).view(nb, rpb, 1)
    x = (torch.randn(nb, rpb, C, device=DEV, generator=g) * m).to(BF16)
    dy = (torch.randn(nb, rpb, C, device=DEV, generator=g) * 0.1 * m).to(BF16)
    w = torch.zeros(C, cg, taps, device=DEV, requires_grad=True)
    y = F.conv1d(x.float().transpose(1, 2), w, None, padding=taps // 2, groups=C // cg).transpose(1, 2)
    (y * dy.float()).sum().backward()
    dw = torch.full((C, cg, taps), 0.25, device=DEV)
    T.gconv_wgrad_tc(x.view(-1, C), dy.view(-1, C), rows_per_batch=rpb, nbatch=nb, cg=cg, taps=taps, dw=dw)
    assert _rel(dw - 0.25, w.grad) < 1e-4
    dw2 = torch.zeros(C, cg, taps, device=DEV)
    T.gconv_wgrad(x.view(-1, C), dy.view(-1, C), rows_per_batch=rpb, nbatch=nb, cg=cg, taps=taps,
                  seq_lens=torch.tensor(lens, device=DEV, dtype=torch.int32), dw=dw2, db=None)
    assert _rel(dw - 0.25, dw2) < 1e-4


@pytest.mark.parametrize("rows,C,ld,off", [(8192, 1024, 1024, 0), (1000, 1024, 3072, 1024), (333, 200, 200, 0), (70, 64, 72, 8), (300, 72, 80, 4)])
def test_colsum_bf16(rows, C, ld, off):
    """Bias gradients: out += column sums, through the 16-byte-load kernel (aligned views) and the 4-byte fallback."""
    g = torch.Generator(device=DEV).manual_seed(rows + C)
    buf = torch.randn(rows, ld, device=DEV, generator=g).to(BF16)
    x = buf[:, off:off + C]
    out = torch.full((C,), 0.5, device=DEV)
    T.colsum(x, out)
    ref = 0.5 + x.float().sum(0)
    assert float((out - ref).abs().max()) < 1e-3 * float(ref.abs().max() + 1)


@pytest.mark.parametrize("rows,C,ld,off", [(8192, 1024, 3072, 2048), (100, 66, 70, 2)])
def test_f16_to_bf16(rows, C, ld, off):
    """V of the forward (IEEE f16 bit patterns inside the fused qkv buffer) -> bf16 operand of the backward."""
    g = torch.Generator(device=DEV).manual_seed(rows)
    buf = torch.randn(rows, ld, device=DEV, generator=g).half()
    x = buf[:, off:off + C]
    out = torch.zeros(rows, C, device=DEV, dtype=BF16)
    T.f16_to_bf16(x.view(torch.bfloat16) if False else x, out)
    assert torch.equal(out, x.float().to(BF16))


def test_cfm_loss():
    g = torch.Generator(device=DEV).manual_seed(9)
    rows, M = 512, 100
    pred = torch.randn(rows, 128, device=DEV, generator=g)[:, :M]
    flow = torch.randn(rows, M, device=DEV, generator=g)
    span = (torch.rand(rows, device=DEV, generator=g) < 0.6).to(torch.uint8)
    cnt = span.sum().to(torch.int32).reshape(1)
    ls = torch.zeros(1, device=DEV)
    dp = torch.full((rows, 128), 5.0, device=DEV, dtype=BF16)
    T.cfm_loss(pred, flow, span, cnt, ls, dp, n_mels=M)
    pr = pred.clone().requires_grad_()
    loss = F.mse_loss(pr, flow, reduction="none")[span.bool()].mean()
    loss.backward()
    assert abs(float(ls) / (int(cnt) * M) - float(loss)) < 1e-5 * float(loss)
    assert _rel(dp[:, :M], pr.grad) < 4e-3 and float(dp[:, M:].abs().max()) == 0.0


def test_sumsq_adamw_clip_vs_torch():
    g = torch.Generator(device=DEV).manual_seed(10)
    n = 100003
    p0 = torch.randn(n, device=DEV, generator=g)
    ref = torch.nn.Parameter(p0.clone())
    opt = torch.optim.AdamW([ref], lr=3e-4, betas=(0.9, 0.999), weight_decay=0.01)
    p, m, v = p0.clone(), torch.zeros(n, device=DEV), torch.zeros(n, device=DEV)
    pb = torch.empty(n, device=DEV, dtype=BF16)
    for step in range(1, 4):
        grad = torch.randn(n, device=DEV, generator=g) * (0.01 if step == 2 else 1.0)
        ref.grad = grad.clone()
        torch.nn.utils.clip_grad_norm_([ref], 1.0)
        opt.step()
        ss = torch.zeros(1, device=DEV)
        T.sumsq(grad, ss)
        assert abs(float(ss) - float((grad.double() ** 2).sum())) < 1e-4 * float(ss)
        T.adamw_clip(p, grad, m, v, pb, ss, grad_scale=1.0, max_norm=1.0, lr=3e-4, beta1=0.9, beta2=0.999, eps=1e-8, wd=0.01,
                     step=step)
        assert float((p - ref.detach()).abs().max()) < 2e-6
    assert torch.equal(pb, p.to(BF16))
    sk = torch.zeros(1, device=DEV, dtype=torch.int32)
    bad = torch.full((1,), float("inf"), device=DEV)
    before = p.clone()
    T.adamw_clip(p, grad, m, v, pb, bad, grad_scale=1.0, max_norm=1.0, lr=3e-4, beta1=0.9, beta2=0.999, eps=1e-8, wd=0.01,
                 step=4, skipped=sk)
    assert int(sk) == 1 and torch.equal(p, before)


def _rope_tables(T_, dev):
    inv = 1.0 / (10000 ** (torch.arange(0, 64, 2, device=dev).float() / 64))
    ang = torch.outer(torch.arange(T_, device=dev).float(), inv)
    return ang.cos().contiguous(), ang.sin().contiguous()


def _rope(x, cos, sin):  # x [B, H, T, 64]; cos/sin [T, 32]
    c, s = torch.cat([cos, cos], -1), torch.cat([sin, sin], -1)
    rot = torch.cat([-x[..., 32:], x[..., :32]], dim=-1)
    return x * c + rot * s


@pytest.mark.parametrize("nb,rpb,H,lens", [(1, 128, 1, [128]), (2, 256, 2, [256, 150]), (2, 384, 3, [300, 384])])
def test_attention_bwd_vs_torch(nb, rpb, H, lens):
    g = torch.Generator(device=DEV).manual_seed(11)
    HD = H * 64
    R = nb * rpb
    sl = torch.tensor(lens, device=DEV, dtype=torch.int32)
    m = _mask(nb, rpb, lens)
    cos, sin = _rope_tables(rpb, DEV)
    qkv_pre = (torch.randn(R, 3 * HD, device=DEV, generator=g)).to(BF16).float().requires_grad_()
    q, k, v = (t.view(nb, rpb, H, 64).transpose(1, 2) for t in qkv_pre.split(HD, dim=1))
    qr, kr = _rope(q, cos, sin), _rope(k, cos, sin)
    # the kernels see bf16 post-RoPE q / k (as the QKV GEMM epilogue writes them)
    qb, kb = qr.detach().to(BF16), kr.detach().to(BF16)
    qr2 = qr + (qb.float() - qr).detach()
    kr2 = kr + (kb.float() - kr).detach()
    key_mask = m.view(nb, 1, 1, rpb)
    s = (qr2 @ kr2.transpose(-1, -2)) / 8.0
    s = s.masked_fill(~key_mask, float("-inf"))
    o = (s.softmax(-1) @ v).transpose(1, 2).reshape(R, HD)
    d_o = (torch.randn(R, HD, device=DEV, generator=g) * 0.1 * m[:, None]).to(BF16)
    (o * d_o.float()).sum().backward()
    qk = torch.cat([qb.transpose(1, 2).reshape(R, HD), kb.transpose(1, 2).reshape(R, HD)], dim=1).contiguous()
    vb = qkv_pre.detach()[:, 2 * HD:].to(BF16).contiguous()
    dqkv = torch.full((R, 3 * HD), 9.0, device=DEV, dtype=BF16)
    lse = torch.zeros(nb * H * rpb, device=DEV)
    delta = torch.zeros(nb * H * rpb, device=DEV)
    T.attention_bwd(qk, vb, o.detach().to(BF16), d_o, dqkv, nbatch=nb, rows_per_batch=rpb, heads=H, seq_lens=sl, scale=0.125,
                    rope_cos=cos, rope_sin=sin, lse=lse, delta=delta)
    torch.cuda.synchronize()
    ref = qkv_pre.grad * m[:, None]
    for name, lo in (("dq", 0), ("dk", HD), ("dv", 2 * HD)):
        assert _rel(dqkv[:, lo:lo + HD], ref[:, lo:lo + HD]) < 2e-2, name
    assert float(dqkv[~m].float().abs().max()) == 0.0 if (~m).any() else True
    # log-sum-exp from the forward kernel (the training path): same gradients without the backward's pre-pass
    qkv16 = torch.cat([qk, vb.float().to(torch.float16).view(torch.bfloat16)], dim=1).contiguous()
    out = torch.zeros(R, HD, device=DEV, dtype=BF16)
    lse_f = torch.zeros(nb * H * rpb, device=DEV)
    T.attention_fwd_lse(qkv16, out, lse_f, nbatch=nb, rows_per_batch=rpb, heads=H, seq_lens=sl, scale=0.125)
    assert _rel(out[m], o.detach()[m]) < 2e-2
    valid = m.view(nb, 1, rpb).expand(nb, H, rpb).reshape(-1)
    assert float((lse_f - lse)[valid].abs().max()) < 2e-2
    dqkv2 = torch.full((R, 3 * HD), 9.0, device=DEV, dtype=BF16)
    T.attention_bwd(qk, vb, out, d_o, dqkv2, nbatch=nb, rows_per_batch=rpb, heads=H, seq_lens=sl, scale=0.125, rope_cos=cos,
                    rope_sin=sin, lse=lse_f, delta=delta, have_lse=True)
    for name, lo in (("dq", 0), ("dk", HD), ("dv", 2 * HD)):
        assert _rel(dqkv2[:, lo:lo + HD], ref[:, lo:lo + HD]) < 2e-2, name


@pytest.mark.parametrize("R,n_out,k_in,bn", [(2048, 1024, 512, 256), (1024, 256, 4096, 256), (1024, 100, 1024, 256),
                                             (2048, 1024, 612, 128)])
def test_gemm_mn_major_operands(R, n_out, k_in, bn):
    """Backward-pass GEMMs straight from the row-major activations (no transposed copies): dW = dY^T X with both
    operands MN-major, dX = dY W with an MN-major B."""
    g = torch.Generator(device=DEV).manual_seed(12)
    ko = (k_in + 7) // 8 * 8
    dY = (torch.randn(R, 1024 if n_out <= 1024 else n_out, device=DEV, generator=g) * 0.1).to(BF16)[:, :n_out]
    X = torch.randn(R, ko, device=DEV, generator=g).to(BF16)[:, :k_in]
    dW = torch.full((n_out, ko), 7.0, device=DEV)[:, :k_in]
    L.gemm(dY, X, dW, epilogue=L.EPI_F32, a_mn=True, b_mn=True, two_sm=True, block_n=bn)
    ref = dY.float().t() @ X.float()
    assert _rel(dW, ref) < 1e-5 + 2e-3 * 0, _rel(dW, ref)
    dW2 = dW.clone()
    L.gemm(dY, X, dW2, epilogue=L.EPI_F32, a_mn=True, b_mn=True, two_sm=True, block_n=bn, addend=dW2)
    assert _rel(dW2, 2 * ref) < 1e-5
    if k_in % 4 == 0:  # stream-K: equal (tile, k-block) shares per SM pair, partial sums added into the output
        dW3 = dW.clone()
        L.gemm(dY, X, dW3, epilogue=L.EPI_F32, a_mn=True, b_mn=True, two_sm=True, block_n=bn, stream_k=True)
        assert _rel(dW3, 2 * ref) < 1e-5
    if n_out % 64 == 0 and k_in % 8 == 0:
        W = (torch.randn(n_out, k_in, device=DEV, generator=g) * 0.05).to(BF16)
        dX = torch.zeros(R, k_in, device=DEV, dtype=BF16)
        L.gemm(dY, W, dX, epilogue=L.EPI_BF16, b_mn=True, two_sm=True, block_n=bn, rows_per_batch=R // 2, nbatch=2)
        assert _rel(dX, dY.float() @ W.float()) < 5e-3


def test_dropout_masks():
    """nn.Dropout inside the activation / gated-residual kernels: keep rate, 1/(1-p) scaling, and the backward kernels
    recompute exactly the forward's mask from (seed, element index)."""
    g = torch.Generator(device=DEV).manual_seed(13)
    rows, C, p = 512, 1024, 0.3
    x = (torch.randn(rows, C, device=DEV, generator=g) + 2.0).to(BF16)  # gelu(x) != 0 almost everywhere
    ref = F.gelu(x.float(), approximate="tanh")
    out = torch.empty_like(x)
    T.act_fwd(x, out, L.ACT_GELU_TANH, dropout_p=p, dropout_seed=123)
    kept = out.float() != 0
    live = ref.abs() > 1e-3
    rate = float(kept[live].float().mean())
    assert abs(rate - (1 - p)) < 5e-3, rate
    assert _rel(out.float()[kept], (ref / (1 - p))[kept]) < 6e-3
    out2 = torch.empty_like(x)
    T.act_fwd(x, out2, L.ACT_GELU_TANH, dropout_p=p, dropout_seed=123)
    assert torch.equal(out, out2)
    T.act_fwd(x, out2, L.ACT_GELU_TANH, dropout_p=p, dropout_seed=124)
    assert float(((out2.float() != 0) != kept)[live].float().mean()) > 0.3  # another seed, another mask
    dy = torch.ones(rows, C, device=DEV, dtype=BF16)
    dpre = torch.empty_like(x)
    T.act_bwd(dy, x, dpre, L.ACT_GELU_TANH, dropout_p=p, dropout_seed=123)
    assert torch.equal((dpre.float() != 0)[live], kept[live])
    # gated residual: x += gate * dropout(y); backward dy = gate * dx * mask / (1 - p), dgate = sum dx * dropout(y)
    nb, rpb, Cg = 2, 128, 256
    y = (torch.randn(nb * rpb, Cg, device=DEV, generator=g) + 3.0).to(BF16)
    gate = torch.randn(nb, Cg, device=DEV, generator=g)
    x0 = torch.zeros(nb * rpb, Cg, device=DEV)
    T.gate_resid(x0, y, rows_per_batch=rpb, nbatch=nb, gate=gate, gate_ld=Cg, seq_lens=None, mask_rows=False, dropout_p=p,
                 dropout_seed=7)
    yd = x0.view(nb, rpb, Cg) / gate[:, None, :]  # = dropout(y)
    k2 = yd.abs() > 1e-6
    assert abs(float(k2.float().mean()) - (1 - p)) < 1.5e-2
    assert _rel(yd[k2], (y.float().view(nb, rpb, Cg) / (1 - p))[k2]) < 1e-3
    dx = torch.randn(nb * rpb, Cg, device=DEV, generator=g)
    dyo = torch.empty(nb * rpb, Cg, device=DEV, dtype=BF16)
    dg = torch.zeros(nb, Cg, device=DEV)
    T.gate_bwd(dx, y, rows_per_batch=rpb, nbatch=nb, gate=gate, gate_ld=Cg, seq_lens=None, dy=dyo, dgate=dg, dgate_ld=Cg,
               dropout_p=p, dropout_seed=7)
    mask = k2.float() / (1 - p)
    assert _rel(dyo, (gate[:, None, :] * dx.view(nb, rpb, Cg) * mask).reshape(-1, Cg)) < 5e-3
    assert _rel(dg, (dx.view(nb, rpb, Cg) * y.float().view(nb, rpb, Cg) * mask).sum(1)) < 1e-4


def test_gemm_fused_gelu_dropout_epilogues():
    """FeedForward up-projection with the training epilogues: forward keeps the pre-activation and writes
    dropout(gelu(pre)); the data-gradient GEMM applies gelu'(pre) and the same mask. Checked against the un-fused
    kernels (act_fwd / act_bwd), which share the mask definition."""
    g = torch.Generator(device=DEV).manual_seed(14)
    R, D, H, p, seed = 1024, 256, 1024, 0.25, 99
    x = torch.randn(R, D, device=DEV, generator=g).to(BF16)
    w1 = (torch.randn(H, D, device=DEV, generator=g) / 16).to(BF16)
    b1 = torch.randn(H, device=DEV, generator=g) * 0.1
    hid, pre = torch.empty(R, H, device=DEV, dtype=BF16), torch.empty(R, H, device=DEV, dtype=BF16)
    L.gemm(x, w1, hid, epilogue=L.EPI_GELU_DROP_DUAL, bias=b1, out2=pre, block_n=256, two_sm=True, dropout_p=p, dropout_seed=seed,
           rows_per_batch=R // 2, nbatch=2)
    pre_ref = x.float() @ w1.float().t() + b1
    assert _rel(pre, pre_ref) < 5e-3
    hid_ref = torch.empty_like(hid)
    T.act_fwd(pre, hid_ref, L.ACT_GELU_TANH, dropout_p=p, dropout_seed=seed)
    assert torch.equal(hid.float() != 0, hid_ref.float() != 0) or float(((hid.float() != 0) != (hid_ref.float() != 0)).float().mean()) < 1e-3
    assert _rel(hid, hid_ref) < 1e-2
    assert abs(float((hid.float() != 0).float().mean()) - (1 - p)) < 2e-2
    # backward: dhpre = (dy @ w2) * gelu'(pre) * mask
    w2 = (torch.randn(D, H, device=DEV, generator=g) / 32).to(BF16)
    dy = (torch.randn(R, D, device=DEV, generator=g) * 0.1).to(BF16)
    dpre = torch.empty(R, H, device=DEV, dtype=BF16)
    L.gemm(dy, w2, dpre, epilogue=L.EPI_GELU_DROP_BWD, out2=pre, b_mn=True, two_sm=True, block_n=256, dropout_p=p, dropout_seed=seed,
           rows_per_batch=R // 2, nbatch=2)
    dh = torch.empty(R, H, device=DEV, dtype=BF16)
    L.gemm(dy, w2, dh, epilogue=L.EPI_BF16, b_mn=True, two_sm=True, block_n=256, rows_per_batch=R // 2, nbatch=2)
    ref = torch.empty_like(dh)
    T.act_bwd(dh, pre, ref, L.ACT_GELU_TANH, dropout_p=p, dropout_seed=seed)
    assert _rel(dpre, ref) < 1e-2


def test_gemm_gate_resid_dual_epilogue():
    """Out-projection / FFN down-projection of the training forward: y = A W^T + b kept as bf16, and
    out = addend + gate[b] * dropout(mask(y)) written to a separate residual buffer -- against GEMM + oron_gate_resid."""
    g = torch.Generator(device=DEV).manual_seed(15)
    nb, rpb, K, N, p, seed = 2, 384, 512, 256, 0.2, 4242
    R = nb * rpb
    lens = [384, 200]
    sl = torch.tensor(lens, device=DEV, dtype=torch.int32)
    a = torch.randn(R, K, device=DEV, generator=g).to(BF16)
    w = (torch.randn(N, K, device=DEV, generator=g) / 22).to(BF16)
    b = torch.randn(N, device=DEV, generator=g) * 0.1
    x_in = torch.randn(R, N, device=DEV, generator=g)
    gate = torch.randn(nb, 3 * N, device=DEV, generator=g)[:, N:2 * N]
    for mask_rows, pp in ((True, p), (False, 0.0)):
        y = torch.empty(R, N, device=DEV, dtype=BF16)
        out = torch.full((R, N), 7.0, device=DEV)
        L.gemm(a, w, out, epilogue=L.EPI_GATE_RESID_DUAL, bias=b, out2=y, addend=x_in, gate=gate, gate_ld=3 * N, gate_nb=nb,
               seq_lens=sl, mask_rows=mask_rows, block_n=256, two_sm=True, dropout_p=pp, dropout_seed=seed, rows_per_batch=rpb,
               nbatch=nb)
        y_ref = torch.empty(R, N, device=DEV, dtype=BF16)
        L.gemm(a, w, y_ref, epilogue=L.EPI_BF16, bias=b, block_n=256, two_sm=True, rows_per_batch=rpb, nbatch=nb)
        assert torch.equal(y, y_ref)
        ref = torch.empty_like(x_in)
        T.gate_resid(x_in, y_ref, rows_per_batch=rpb, nbatch=nb, gate=gate, gate_ld=3 * N, seq_lens=sl, mask_rows=mask_rows,
                     dropout_p=pp, dropout_seed=seed, out=ref)
        # the fused epilogue gates the fp32 accumulator, the reference path the bf16-rounded y
        assert _rel(out, ref) < 2e-3, (mask_rows, _rel(out, ref))
        if mask_rows:
            m = _mask(nb, rpb, lens)
            assert torch.equal(out[~m], x_in[~m])
