"""Backward kernels of the attention layers (csrc/grad_ops.cu, csrc/attn_bwd.cu) through the C ABI.

Each kernel is checked against fp64 autograd of the same maths on the GPU; the assembled block backward
(mha_block / ffn_block, dropout 0.1, train mode) against the older recompute-with-torch-ops backward
(EDA_BACKWARD=torch), which regenerates the very same dropout masks.

Tolerances: tf32 tensor-core operands (10-bit mantissa, round-to-nearest) with fp32 accumulation: a gradient GEMM over
R rows has relative error ~ 2^-11 * sqrt(R) / sqrt(R) per element, i.e. ~1e-3 of the gradient's scale; checked as
relative Frobenius error <= 5e-3 per kernel, <= 2e-2 for an assembled block (two chained tf32 stages + exp).
"""
import math
import os
import sys

import pytest
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))
import attn_cases as ac  # noqa: E402

pytestmark = pytest.mark.gpu


def rel(a, b):
    return ((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-30)).item()


def _round_tf32(x):
    return (x.view(torch.int32) + 0x1000 & ~0x1FFF).view(torch.float32) if x.dtype == torch.float32 else x


@pytest.mark.parametrize("R,N,K,nprob", [(700, 288, 288, 1), (8192, 288, 256, 1), (100, 64, 8, 1), (5000, 128, 144, 3),
                                         (31, 256, 288, 5), (200000, 64, 16, 1)])
def test_wgrad_kernel(R, N, K, nprob):
    from eda_b200 import attn_ops as ops

    g = torch.Generator().manual_seed(R + N + K)
    probs, refs = [], []
    for i in range(nprob):
        Ri = max(1, R - 37 * i)
        dy = torch.randn(Ri, N, generator=g).cuda()
        x = torch.randn(Ri, K, generator=g).cuda()
        dw = torch.zeros(N + 3, K, device="cuda")[1:N + 1]  # a row-offset view, like a slice of in_proj_weight's gradient
        db = torch.zeros(N, device="cuda") if i % 2 == 0 else None
        probs.append(dict(dy=dy, x=x, dw=dw, db=db))
        refs.append((dy.double().t() @ x.double(), dy.double().sum(0)))
    ops.wgrad(probs, N, K)
    for pr, (rw, rb) in zip(probs, refs):
        assert rel(pr["dw"], rw) <= 5e-3
        if pr["db"] is not None:
            assert rel(pr["db"], rb) <= 1e-4
    # accumulation: a second call doubles the result
    ops.wgrad(probs[:1], N, K)
    assert rel(probs[0]["dw"], 2 * refs[0][0]) <= 5e-3


@pytest.mark.parametrize("R,N,p", [(640, 288, 0.0), (8192, 288, 0.1), (5, 256, 0.0), (1000, 64, 0.3)])
def test_layernorm_backward_kernel(R, N, p):
    from eda_b200 import attn_ops as ops

    g = torch.Generator().manual_seed(R + N)
    u = (torch.randn(R, N, generator=g) * 2 + 0.5).cuda()
    dy = torch.randn(R, N, generator=g).cuda()
    gam, bet = (1 + 0.2 * torch.randn(N, generator=g)).cuda(), (0.1 * torch.randn(N, generator=g)).cuda()
    dgam, dbet = torch.zeros(N, device="cuda"), torch.zeros(N, device="cuda")
    seed = 4242
    du, dproj = ops.layernorm_backward(dy, u, gam, 1e-5, dgam, dbet, dropout=(p, seed) if p > 0 else None)
    ud = u.double().requires_grad_(True)
    gd, bd = gam.double().requires_grad_(True), bet.double().requires_grad_(True)
    y = torch.nn.functional.layer_norm(ud, (N,), gd, bd, 1e-5)
    y.backward(dy.double())
    assert rel(du, ud.grad) <= 1e-4 and rel(dgam, gd.grad) <= 1e-4 and rel(dbet, bd.grad) <= 1e-4
    if p > 0:
        keep = ops.dropout_mask(seed, p, R, N, 3, 0, u.device)
        torch.testing.assert_close(dproj, du * keep / (1 - p), rtol=1e-6, atol=1e-7)
    else:
        assert dproj is du


@pytest.mark.parametrize("B,Nq,Nk,masked,p", [(2, 80, 80, True, 0.0), (2, 256, 132, True, 0.1), (2, 200, 1024, False, 0.0),
                                              (1, 1024, 1024, False, 0.1), (3, 7, 5, True, 0.0), (2, 129, 257, True, 0.1)])
def test_attention_backward_kernel(B, Nq, Nk, masked, p):
    from eda_b200 import attn_ops as ops

    H, D = 8, 36
    E = H * D
    g = torch.Generator().manual_seed(B * 1000 + Nq + Nk)
    q = (torch.randn(B, Nq, E, generator=g) * 1.5).cuda()
    k, v = (_round_tf32(torch.randn(B, Nk, E, generator=g)).cuda() for _ in range(2))
    dctx = torch.randn(B, Nq, E, generator=g).cuda()
    mask = ac.ragged_mask(B, Nk, max(1, Nk // 3), g).cuda() if masked else None
    seed = 99
    drop = (p, seed) if p > 0 else None
    ld = (Nk + 3) & ~3
    vt = torch.zeros(B, E, ld, device="cuda")
    vt[:, :, :Nk] = v.transpose(1, 2)
    lse = torch.empty(B, H, Nq, device="cuda")
    c = ops.attention_raw(q.view(-1, E), k.view(-1, E), vt, mask, B, Nq, Nk, H, dropout=drop, lse=lse)
    dq, dk, dv = ops.attention_backward_raw(q.view(-1, E), k.view(-1, E), vt, dctx.view(-1, E), c, lse, mask, B, Nq, Nk, H,
                                            dropout=drop)
    # fp64 autograd of the same maths with the same keep-mask
    qd, kd, vd = (t.double().requires_grad_(True) for t in (q, k, v))
    s = (qd.view(B, Nq, H, D).transpose(1, 2) / math.sqrt(D)) @ kd.view(B, Nk, H, D).transpose(1, 2).transpose(-1, -2)
    if masked:
        s = s.masked_fill(mask.view(B, 1, 1, Nk), float("-inf"))
    pr = torch.softmax(s, -1)
    lse_ref = torch.logsumexp(s, -1)
    if p > 0:
        keep = ops.dropout_mask(seed, p, B * H * Nq, Nk, 1, 0, q.device).view(B, H, Nq, Nk)
        pr = pr * keep.double() / (1 - p)
    ref = (pr @ vd.view(B, Nk, H, D).transpose(1, 2)).transpose(1, 2).reshape(B, Nq, E)
    ref.backward(dctx.double())
    torch.testing.assert_close(lse.double(), lse_ref, rtol=2e-3, atol=2e-3)
    assert rel(dq.view(B, Nq, E), qd.grad) <= 5e-3
    assert rel(dk.view(B, Nk, E), kd.grad) <= 5e-3
    assert rel(dv.view(B, Nk, E), vd.grad) <= 5e-3
    if masked:  # ignored keys receive exactly zero gradient
        assert dk.view(B, Nk, E)[mask].abs().max() == 0 and dv.view(B, Nk, E)[mask].abs().max() == 0


def _block_grads(mode, fn):
    old = os.environ.get("EDA_BACKWARD")
    os.environ["EDA_BACKWARD"] = mode
    try:
        return fn()
    finally:
        if old is None:
            del os.environ["EDA_BACKWARD"]
        else:
            os.environ["EDA_BACKWARD"] = old


@pytest.mark.parametrize("train", [False, True])
@pytest.mark.parametrize("Nq,Nk,self_attn", [(256, 132, False), (300, 300, True), (80, 1024, False)])
def test_mha_block_backward_cuda_vs_torch(train, Nq, Nk, self_attn):
    """LayerNorm(residual + dropout(MHA(q + qpos, k + kpos, v))): CUDA backward vs the torch recompute backward."""
    from eda_b200 import attn_ops as ops

    E, H, B = 288, 8, 2
    g = torch.Generator().manual_seed(Nq * 7 + Nk)
    mha = torch.nn.MultiheadAttention(E, H, dropout=0.1)
    norm = torch.nn.LayerNorm(E)
    dropm = torch.nn.Dropout(0.1)
    with torch.no_grad():
        for prm in list(mha.parameters()) + list(norm.parameters()):
            prm.copy_(torch.randn(prm.shape, generator=g) * (0.06 if prm.dim() > 1 else 0.3) + (1.0 if prm.dim() == 1 else 0.0))
    mods = torch.nn.ModuleList([mha, norm, dropm]).cuda()
    mods.train(train)
    x = torch.randn(B, Nq, E, generator=g).cuda()
    mem = x if self_attn else torch.randn(B, Nk, E, generator=g).cuda()
    qpos = torch.randn(B, Nq, E, generator=g).cuda() * 0.5
    kpos = qpos if self_attn else None
    mask = ac.ragged_mask(B, mem.size(1), 20, g).cuda()
    w = torch.randn(B, Nq, E, generator=g).cuda()

    def run():
        torch.manual_seed(7)
        xs = x.clone().requires_grad_(True)
        ms = xs if self_attn else mem.clone().requires_grad_(True)
        qp = qpos.clone().requires_grad_(True)
        for prm in mods.parameters():
            prm.grad = None
        y = ops.mha_block(mha, xs, ms, ms, q_pos=qp, k_pos=(qp if self_attn else kpos), key_padding_mask=mask,
                          residual=xs, norm=norm, out_dropout=dropm)
        (y * w).sum().backward()
        out = {"x": xs.grad.clone(), "qpos": qp.grad.clone(), "y": y.detach().clone()}
        if not self_attn:
            out["mem"] = ms.grad.clone()
        for n, prm in mods.named_parameters():
            out[n] = prm.grad.clone()
        return out

    got = _block_grads("cuda", run)
    want = _block_grads("torch", run)
    assert torch.equal(got["y"], want["y"])
    for n in want:
        assert rel(got[n], want[n]) <= 2e-2, (n, rel(got[n], want[n]))


@pytest.mark.parametrize("train", [False, True])
def test_ffn_block_backward_cuda_vs_torch(train):
    from eda_b200 import attn_ops as ops

    E, Fh, R = 288, 256, 700
    g = torch.Generator().manual_seed(11)
    ffn = torch.nn.Sequential(torch.nn.Linear(E, Fh), torch.nn.ReLU(), torch.nn.Dropout(0.1), torch.nn.Linear(Fh, E),
                              torch.nn.Dropout(0.1))
    norm = torch.nn.LayerNorm(E)
    mods = torch.nn.ModuleList([ffn, norm]).cuda()
    mods.train(train)
    x = torch.randn(2, R // 2, E, generator=g).cuda()
    w = torch.randn(2, R // 2, E, generator=g).cuda()

    def run():
        torch.manual_seed(3)
        xs = x.clone().requires_grad_(True)
        for prm in mods.parameters():
            prm.grad = None
        y = ops.ffn_block(ffn, xs, norm)
        (y * w).sum().backward()
        out = {"x": xs.grad.clone(), "y": y.detach().clone()}
        for n, prm in mods.named_parameters():
            out[n] = prm.grad.clone()
        return out

    got = _block_grads("cuda", run)
    want = _block_grads("torch", run)
    assert torch.equal(got["y"], want["y"])
    for n in want:
        assert rel(got[n], want[n]) <= 2e-2, (n, rel(got[n], want[n]))
