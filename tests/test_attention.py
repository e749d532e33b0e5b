"""Cross-modal attention layers.

CPU suite : the oracle (oracle/attention_oracle.py) against the golden outputs of the REFERENCE's own
            models/encoder_decoder_layers.py (tests/golden/attn_*.npz, made by make_golden_attention.py),
            state-dict key/shape identity of this repo's mirror classes, loud failure on CPU tensors.
GPU suite : the CUDA layers (eda_b200/encoder_decoder_layers.py -> eda_linear_forward /
            eda_attention_forward through the C ABI) against the same fixtures and against the oracle at the
            BASELINE sizes (V=1024, L=80, D=132, K=256).

Tolerance for the CUDA path (tf32 tensor-core operands, fp32 accumulation, fp32 softmax / LayerNorm;
SURVEY.md 8c): one fused block (a GEMM, an attention core) rtol = atol = 2e-3; a whole layer = 6-7 chained
blocks with LayerNorm in between, measured per-block max error 1.5e-3 / rms 2.5e-4 on O(1) outputs:
max error <= 5e-3 and rms error <= 1e-3; the 3-layer encoder: 1e-2 / 2e-3.  Oracle vs reference
fixture: 2e-5 (same fp32 maths, different association).
"""
import math
import os
import sys

import numpy as np
import pytest
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))
import attn_cases as ac  # noqa: E402

from eda_b200 import encoder_decoder_layers as edl  # noqa: E402
from oracle import attention_oracle as ao  # noqa: E402

TOL = dict(rtol=2e-3, atol=2e-3)


def build(kind):
    if kind == "bi_encoder_layer":
        return edl.BiEncoderLayer(ac.E, dropout=0.1, activation="relu", n_heads=ac.HEADS, dim_feedforward=ac.FF,
                                  self_attend_lang=True, self_attend_vis=True, use_butd_enc_attn=True)
    if kind == "bi_encoder":
        return edl.BiEncoder(build("bi_encoder_layer"), 3)
    return edl.BiDecoderLayer(ac.E, ac.HEADS, ac.FF, 0.1, "relu", self_position_embedding="loc_learned", butd=True)


def oracle_run(kind, sd, inp):
    if kind == "bi_encoder_layer":
        v, t = ao.bi_encoder_layer(sd, "", inp["vis"], inp["pos"], None, inp["text"], inp["text_mask"], inp["det"],
                                   inp["det_mask"])
        return dict(vis=v, text=t)
    if kind == "bi_encoder":
        v, t = ao.bi_encoder(sd, "", 3, inp["vis"], inp["pos"], None, inp["text"], inp["text_mask"], inp["det"],
                             inp["det_mask"])
        return dict(vis=v, text=t)
    q = ao.bi_decoder_layer(sd, "", inp["query"], inp["vis"], inp["text"], inp["query_pos"], None, inp["text_mask"],
                            inp["det"], inp["det_mask"])
    return dict(query=q)


def module_run(kind, m, inp):
    if kind in ("bi_encoder_layer", "bi_encoder"):
        v, t = m(inp["vis"], inp["pos"], None, inp["text"], inp["text_mask"], {}, detected_feats=inp["det"],
                 detected_mask=inp["det_mask"])
        return dict(vis=v, text=t)
    q = m(inp["query"], inp["vis"], inp["text"], inp["query_pos"], None, inp["text_mask"], detected_feats=inp["det"],
          detected_mask=inp["det_mask"])
    return dict(query=q)


def assert_layer_close(got, want, depth=1):
    """max / rms error bounds for `depth` chained layers (see the module docstring)."""
    err = (got.double() - want.double()).abs()
    assert torch.isfinite(got).all()
    lim_max, lim_rms = (5e-3, 1e-3) if depth == 1 else (1e-2, 2e-3)
    rms = err.pow(2).mean().sqrt().item()
    assert err.max().item() <= lim_max and rms <= lim_rms, f"max err {err.max().item():.3e}, rms err {rms:.3e}"


# ------------------------------------------------------------------------------------------------
# CPU suite
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name", list(ac.CASES))
def test_oracle_matches_reference_fixture(name):
    kind = ac.CASES[name][0]
    fx = np.load(os.path.join(HERE, "golden", f"attn_{name}.npz"))
    m = ac.fill_params(build(kind), seed=100 + len(name)).eval()
    sd = m.state_dict()
    # the mirror has exactly the reference's state-dict keys and shapes (checkpoints load)
    assert sorted(sd.keys()) == list(fx["keys"])
    assert [str(tuple(sd[k].shape)) for k in sorted(sd.keys())] == list(fx["shapes"])
    with torch.no_grad():
        out = oracle_run(kind, sd, ac.make_inputs(name))
    for k, v in out.items():
        torch.testing.assert_close(v, torch.from_numpy(fx[k]), rtol=2e-5, atol=2e-5)


GRAD_CASES = ["enc_layer", "dec_layer"]


def _grad_w(shape, seed):
    return torch.randn(shape, generator=torch.Generator().manual_seed(seed))  # make_golden_attention.grad_weights


def _check_against_grad_fixture(fx, input_grads, param_grads, tol):
    """fx: tests/golden/attn_grad_<case>.npz — gradients of the REFERENCE module (autograd on CPU): inputs and 1-D
    parameters in full, matrices as row sums / column sums / norm (+ one full matrix)."""
    def rel(a, b):
        return ((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-20)).item()

    checked, failures = 0, []
    for key in fx.files:
        kind_, name = key.split(".", 1)
        want = torch.from_numpy(np.asarray(fx[key]))
        if kind_ == "input":
            got = input_grads[name]
        else:
            g = param_grads[name]
            g2 = g.reshape(g.shape[0], -1) if g.dim() > 1 else g
            got = {"param": g, "rowsum": g2.sum(1) if g.dim() > 1 else g, "colsum": g2.sum(0) if g.dim() > 1 else g,
                   "norm": g2.norm()}[kind_]
        # absolute floor of 1e-4 per element: some gradients are zero up to rounding (the key-projection bias — softmax
        # is shift-invariant — and a bias in front of a BatchNorm), a relative comparison of noise means nothing
        # (the floor follows the tolerance: with tf32 operands such a "zero" is a sum of ~1e-3-relative rounding errors)
        floor = max(1e-4, 0.06 * tol) * math.sqrt(want.numel())
        if kind_ in ("rowsum", "colsum"):
            # a sum over a matrix whose entries are right to `tol` can be off by tol * ||matrix||: the column sums of a
            # weight gradient in front of a LayerNorm vanish by exact cancellation (sum_n du[r, n] = 0), which tf32
            # rounding of du does not preserve — the matrix itself (param.* / norm.* / rowsum.*) is what is checked
            floor += tol * float(fx["norm." + name])
        if kind_ == "norm":
            ok = abs(got.item() - want.item()) <= tol * want.item() + floor
            err = abs(got.item() - want.item())
        else:
            err = (got.cpu().double() - want.double()).norm().item()
            ok = err <= tol * want.double().norm().item() + floor
        if not ok:
            failures.append((key, round(err, 6), round(float(want.double().norm()), 6)))
        checked += 1
    assert not failures, failures
    assert checked >= 50


@pytest.mark.parametrize("name", GRAD_CASES)
def test_oracle_gradients_match_reference_fixture(name):
    """Backward parity is pinned by the reference too: autograd through the oracle restatement reproduces the gradients
    the reference's own modules give on the same inputs (fixtures made by make_golden_attention.py)."""
    kind = ac.CASES[name][0]
    fx = np.load(os.path.join(HERE, "golden", f"attn_grad_{name}.npz"))
    m = ac.fill_params(build(kind), seed=100 + len(name)).eval()
    sd = {k: v.detach().clone().requires_grad_(v.is_floating_point()) for k, v in m.state_dict().items()}
    inp = ac.make_inputs(name)
    leaves = {k: inp[k].clone().requires_grad_(True) for k in (("vis", "text") if kind == "bi_encoder_layer" else ("query",))}
    out = oracle_run(kind, sd, {**inp, **leaves})
    loss = sum((v * _grad_w(v.shape, 1234 + i)).sum() for i, v in enumerate(out.values()))
    loss.backward()
    _check_against_grad_fixture(fx, {k: v.grad for k, v in leaves.items()},
                                {k: v.grad for k, v in sd.items() if v.grad is not None}, tol=1e-3)


def test_layers_refuse_cpu_tensors():
    m = build("bi_decoder_layer").eval()
    inp = ac.make_inputs("dec_layer")
    with pytest.raises(RuntimeError):
        module_run("bi_decoder_layer", m, inp)


def test_oracle_masked_keys_have_no_influence():
    m = ac.fill_params(build("bi_encoder_layer"), seed=3).eval()
    sd = m.state_dict()
    inp = ac.make_inputs("enc_layer")
    with torch.no_grad():
        a = oracle_run("bi_encoder_layer", sd, inp)
        inp2 = dict(inp)
        inp2["det"] = inp["det"].clone()
        inp2["det"][inp["det_mask"]] = 1e3  # garbage in padded boxes
        b = oracle_run("bi_encoder_layer", sd, inp2)
    torch.testing.assert_close(a["vis"], b["vis"], rtol=0, atol=0)


# ------------------------------------------------------------------------------------------------
# GPU suite
# ------------------------------------------------------------------------------------------------
def _cuda(d):
    return {k: v.cuda() for k, v in d.items()}


def _round_tf32(t):
    """cvt.rna.tf32.f32 (round to nearest, ties away) emulated on the fp32 bit pattern: what eda_linear_forward's
    round_tf32 epilogue does to the K / V projections before eda_attention_forward consumes them."""
    i = t.contiguous().view(torch.int32)
    return ((i + 0x1000) & ~0x1FFF).view(torch.float32)


def _channel_major(v):
    """(B, Nk, E) -> (B, E, ld) with ld = Nk rounded up to 4: the layout eda_attention_forward takes for V
    (the V-projection GEMM writes it directly; here built with torch for the kernel-level tests)."""
    B, Nk, E = v.shape
    ld = (Nk + 3) & ~3
    out = torch.full((B, E, ld), float("nan"), device=v.device)  # padding must never be read
    out[:, :, :Nk] = v.transpose(1, 2)
    return out


@pytest.mark.gpu
@pytest.mark.parametrize("R,K,N,relu,ln,pos", [
    (300, 288, 288, False, False, True), (128, 288, 256, True, False, False), (1000, 256, 288, False, True, False),
    (77, 6, 288, True, False, False), (513, 3, 288, False, False, False), (2048, 288, 288, False, True, True),
    (5, 288, 64, False, False, False), (260, 40, 16, True, False, True),
])
def test_linear_kernel(R, K, N, relu, ln, pos):
    from eda_b200 import attn_ops as ops

    g = torch.Generator().manual_seed(R + K + N)
    x = torch.randn(R, K, generator=g).cuda()
    p = torch.randn(R, K, generator=g).cuda() if pos else None
    W = (torch.randn(N, K, generator=g) / math.sqrt(K)).cuda()
    b = torch.randn(N, generator=g).cuda()
    res = torch.randn(R, N, generator=g).cuda() if ln else None
    gam, bet = (1 + 0.1 * torch.randn(N, generator=g)).cuda(), (0.1 * torch.randn(N, generator=g)).cuda()
    packed = ops.pack_weight(W)
    (y,) = ops.linear_raw([dict(x=x, pos=p, w_packed=packed, bias=b, residual=res)], K, N, relu=relu,
                          ln=(gam, bet, 1e-5) if ln else None)
    xin = (x + p) if pos else x
    ref = (xin.double() @ W.double().t() + b.double())
    if relu:
        ref = ref.clamp_min(0)
    if ln:
        ref = torch.nn.functional.layer_norm(ref + res.double(), (N,), gam.double(), bet.double(), 1e-5)
        torch.testing.assert_close(y.double(), ref, **TOL)
    else:
        # tf32 operands: each product carries <= 2 * 2^-11 relative error -> bound on sum |x||w|
        bound = xin.abs().double() @ W.abs().double().t()
        assert ((y.double() - ref).abs() <= 1.1e-3 * bound + 1e-5).all()


@pytest.mark.gpu
def test_linear_kernel_three_problems_one_launch():
    from eda_b200 import attn_ops as ops

    g = torch.Generator().manual_seed(5)
    xs = [torch.randn(r, 288, generator=g).cuda() for r in (130, 80, 1024)]
    Ws = [(torch.randn(288, 288, generator=g) / 17).cuda() for _ in range(3)]
    bs = [torch.randn(288, generator=g).cuda() for _ in range(3)]
    ys = ops.linear_raw([dict(x=x, w_packed=ops.pack_weight(W), bias=b) for x, W, b in zip(xs, Ws, bs)], 288, 288)
    for x, W, b, y in zip(xs, Ws, bs, ys):
        torch.testing.assert_close(y.double(), x.double() @ W.double().t() + b.double(), **TOL)


@pytest.mark.gpu
@pytest.mark.parametrize("B,Nq,Nk,masked", [(2, 80, 80, True), (2, 256, 132, True), (2, 200, 1024, False),
                                            (1, 1024, 1024, False), (3, 7, 5, True), (2, 129, 257, True)])
def test_attention_kernel(B, Nq, Nk, masked):
    from eda_b200 import attn_ops as ops

    H, D = 8, 36
    g = torch.Generator().manual_seed(B * 1000 + Nq + Nk)
    q, k, v = (torch.randn(B, n, H * D, generator=g).cuda() for n in (Nq, Nk, Nk))
    q = q * 2.0  # sharper softmax
    k, v = _round_tf32(k), _round_tf32(v)  # the kernel's contract: K / V hold tf32-representable values
    mask = ac.ragged_mask(B, Nk, max(1, Nk // 3), g).cuda() if masked else None
    ctx = ops.attention_raw(q.view(-1, H * D), k.view(-1, H * D), _channel_major(v), mask, B, Nq, Nk, H).view(B, Nq, H * D)
    qd = q.double().view(B, Nq, H, D).transpose(1, 2) / 6.0
    kd = k.double().view(B, Nk, H, D).transpose(1, 2)
    vd = v.double().view(B, Nk, H, D).transpose(1, 2)
    s = qd @ kd.transpose(-1, -2)
    if masked:
        s = s.masked_fill(mask.view(B, 1, 1, Nk), float("-inf"))
    ref = (torch.softmax(s, -1) @ vd).transpose(1, 2).reshape(B, Nq, H * D)
    torch.testing.assert_close(ctx.double(), ref, **TOL)


@pytest.mark.gpu
def test_attention_fully_masked_row_is_nan_like_reference():
    from eda_b200 import attn_ops as ops

    H, D, B, Nq, Nk = 8, 36, 2, 40, 50
    g = torch.Generator().manual_seed(0)
    q, k, v = (_round_tf32(torch.randn(B, n, H * D, generator=g)).cuda() for n in (Nq, Nk, Nk))
    mask = torch.zeros(B, Nk, dtype=torch.bool).cuda()
    mask[1] = True
    ctx = ops.attention_raw(q.view(-1, H * D), k.view(-1, H * D), _channel_major(v), mask, B, Nq, Nk, H).view(B, Nq, H * D)
    assert torch.isfinite(ctx[0]).all() and torch.isnan(ctx[1]).all()


@pytest.mark.gpu
@pytest.mark.parametrize("name", list(ac.CASES))
def test_cuda_layers_match_reference_fixture(name):
    kind = ac.CASES[name][0]
    fx = np.load(os.path.join(HERE, "golden", f"attn_{name}.npz"))
    m = ac.fill_params(build(kind), seed=100 + len(name)).eval().cuda()
    with torch.no_grad():
        out = module_run(kind, m, _cuda(ac.make_inputs(name)))
    for k, v in out.items():
        assert_layer_close(v.cpu(), torch.from_numpy(fx[k]), depth=3 if kind == "bi_encoder" else 1)


def _full_inputs(B, V, L, D, K, seed):
    g = torch.Generator().manual_seed(seed)
    r = lambda *s: torch.randn(*s, generator=g)
    return dict(vis=r(B, V, ac.E), pos=0.5 * r(B, V, ac.E), text=r(B, L, ac.E), text_mask=ac.ragged_mask(B, L, 20, g),
                det=r(B, D, ac.E), det_mask=ac.ragged_mask(B, D, 20, g), query=r(B, K, ac.E),
                query_pos=torch.cat([4 * torch.rand(B, K, 3, generator=g) - 2, torch.rand(B, K, 3, generator=g) + .2], -1))


@pytest.mark.gpu
@pytest.mark.parametrize("kind", ["bi_encoder_layer", "bi_encoder", "bi_decoder_layer"])
def test_cuda_layers_match_oracle_at_baseline_sizes(kind):
    """V=1024 seeds, L=80 tokens (ragged), D=132 boxes (ragged), K=256 queries (BASELINE.json configs[2])."""
    B = 2
    inp = _full_inputs(B, 1024, 80, 132, 256, seed=11)
    m = ac.fill_params(build(kind), seed=21).eval()
    with torch.no_grad():
        want = oracle_run(kind, m.state_dict(), inp)
        got = module_run(kind, m.cuda(), _cuda(inp))
    for k in want:
        assert_layer_close(got[k].cpu(), want[k], depth=3 if kind == "bi_encoder" else 1)


@pytest.mark.gpu
def test_cuda_masked_keys_have_no_influence():
    m = ac.fill_params(build("bi_decoder_layer"), seed=4).eval().cuda()
    inp = _cuda(ac.make_inputs("dec_layer"))
    with torch.no_grad():
        a = module_run("bi_decoder_layer", m, inp)["query"]
        inp["det"] = inp["det"].clone()
        inp["det"][inp["det_mask"]] = 1e3
        inp["text"] = inp["text"].clone()
        inp["text"][inp["text_mask"]] = -1e3
        b = module_run("bi_decoder_layer", m, inp)["query"]
    assert torch.equal(a, b)


@pytest.mark.gpu
def test_seq_first_wrappers_match_reference_layout():
    m = ac.fill_params(edl.PosTransformerEncoderLayerNoFFN(ac.E, ac.HEADS, 0.1), seed=9).eval()
    g = torch.Generator().manual_seed(1)
    src, pos = torch.randn(50, 2, ac.E, generator=g), torch.randn(50, 2, ac.E, generator=g)
    with torch.no_grad():
        want = ao.self_attention(m.state_dict(), "", src.transpose(0, 1), pos.transpose(0, 1)).transpose(0, 1)
        got = m.cuda()(src.cuda(), pos.cuda())
    assert got.shape == (50, 2, ac.E)
    torch.testing.assert_close(got.cpu(), want, **TOL)


@pytest.mark.gpu
def test_backward_matches_torch_autograd():
    """dropout = 0 training step through one decoder layer: grads of the CUDA path (recompute backward)
    against autograd through the oracle maths run on the GPU in fp32."""
    torch.backends.cuda.matmul.allow_tf32 = False
    m = edl.BiDecoderLayer(ac.E, ac.HEADS, ac.FF, 0.0, "relu", self_position_embedding="loc_learned", butd=True)
    ac.fill_params(m, seed=5).cuda().eval()  # eval: BatchNorm1d of the pos-embed uses running stats in both paths
    inp = _cuda(ac.make_inputs("dec_layer"))
    q = inp["query"].clone().requires_grad_(True)
    out = m(q, inp["vis"], inp["text"], inp["query_pos"], None, inp["text_mask"], detected_feats=inp["det"],
            detected_mask=inp["det_mask"])
    w = torch.randn(out.shape, generator=torch.Generator().manual_seed(0)).to(out.device)
    (out * w).sum().backward()
    got = {n: p.grad.clone() for n, p in m.named_parameters() if p.grad is not None}
    got_q = q.grad.clone()

    sd = {k: v.detach().clone().requires_grad_(v.is_floating_point()) for k, v in m.state_dict().items()}
    q2 = inp["query"].clone().requires_grad_(True)
    ref = ao.bi_decoder_layer(sd, "", q2, inp["vis"], inp["text"], inp["query_pos"], None, inp["text_mask"], inp["det"],
                              inp["det_mask"])
    (ref * w).sum().backward()

    # Gradients are piecewise: a ReLU unit of the FFN / pos-embed whose pre-activation sits within the forward
    # tolerance of zero flips between the two paths and changes one row of the gradient by O(0.1) (CPU check:
    # 1e-3 input noise moves fp64 gradients by up to 0.3).  So: relative Frobenius error, not element-wise.
    def rel(a, b):
        return ((a - b).norm() / b.norm().clamp_min(1e-12)).item()

    assert rel(got_q, q2.grad) <= 5e-2
    checked = 0
    for n, gval in got.items():
        assert rel(gval, sd[n].grad) <= 5e-2, n
        checked += 1
    assert checked > 20


# ------------------------------------------------------------------------------------------------
# train-mode dropout (in-kernel counter-based masks; the backward pass regenerates them)
# ------------------------------------------------------------------------------------------------
@pytest.mark.gpu
def test_linear_dropout_matches_regenerated_mask():
    from eda_b200 import attn_ops as ops

    g = torch.Generator().manual_seed(3)
    R, K, N, p, seed = 700, 288, 288, 0.1, 12345
    x = torch.randn(R, K, generator=g).cuda()
    W = (torch.randn(N, K, generator=g) / math.sqrt(K)).cuda()
    b = torch.randn(N, generator=g).cuda()
    res = torch.randn(R, N, generator=g).cuda()
    gam, bet = (1 + 0.1 * torch.randn(N, generator=g)).cuda(), (0.1 * torch.randn(N, generator=g)).cuda()
    (y,) = ops.linear_raw([dict(x=x, w_packed=ops.pack_weight(W), bias=b, residual=res)], K, N, relu=True,
                          ln=(gam, bet, 1e-5), dropout=(p, seed))
    keep = ops.dropout_mask(seed, p, R, N, 3, 0, x.device)
    assert abs(keep.mean().item() - (1 - p)) < 5e-3  # 200k Bernoulli(0.9) draws
    assert not torch.equal(keep, ops.dropout_mask(seed + 1, p, R, N, 3, 0, x.device))
    ref = (x.double() @ W.double().t() + b.double()).clamp_min(0) * keep.double() / (1 - p)
    ref = torch.nn.functional.layer_norm(ref + res.double(), (N,), gam.double(), bet.double(), 1e-5)
    torch.testing.assert_close(y.double(), ref, **TOL)


@pytest.mark.gpu
def test_attention_dropout_matches_regenerated_mask():
    from eda_b200 import attn_ops as ops

    H, D, B, Nq, Nk, p, seed = 8, 36, 2, 200, 300, 0.1, 777
    g = torch.Generator().manual_seed(9)
    q = torch.randn(B, Nq, H * D, generator=g).cuda()
    k, v = (_round_tf32(torch.randn(B, Nk, H * D, generator=g)).cuda() for _ in range(2))
    mask = ac.ragged_mask(B, Nk, 100, g).cuda()
    ctx = ops.attention_raw(q.view(-1, H * D), k.view(-1, H * D), _channel_major(v), mask, B, Nq, Nk, H,
                            dropout=(p, seed)).view(B, Nq, H * D)
    keep = ops.dropout_mask(seed, p, B * H * Nq, Nk, 1, 0, q.device).view(B, H, Nq, Nk)
    assert abs(keep.mean().item() - (1 - p)) < 5e-3
    qd = q.double().view(B, Nq, H, D).transpose(1, 2) / 6.0
    kd = k.double().view(B, Nk, H, D).transpose(1, 2)
    vd = v.double().view(B, Nk, H, D).transpose(1, 2)
    s_ = (qd @ kd.transpose(-1, -2)).masked_fill(mask.view(B, 1, 1, Nk), float("-inf"))
    pr = torch.softmax(s_, -1) * keep.double() / (1 - p)
    ref = (pr @ vd).transpose(1, 2).reshape(B, Nq, H * D)
    torch.testing.assert_close(ctx.double(), ref, **TOL)


@pytest.mark.gpu
def test_train_mode_dropout_layer_forward_backward():
    """The reference's default configuration (dropout 0.1, train mode) runs on the CUDA path: repeatable under
    torch.manual_seed, different from eval mode, and its backward uses exactly the masks the forward applied
    (a second backward with the same seeds gives identical gradients; with p -> eval the eval output returns)."""
    m = ac.fill_params(build("bi_encoder_layer"), seed=8).cuda().train()
    inp = _cuda(ac.make_inputs("enc_layer"))

    def run():
        torch.manual_seed(1234)
        vis = inp["vis"].clone().requires_grad_(True)
        v, t = m(vis, inp["pos"], None, inp["text"], inp["text_mask"], {}, detected_feats=inp["det"],
                 detected_mask=inp["det_mask"])
        (v.pow(2).mean() + t.pow(2).mean()).backward()
        return v.detach(), t.detach(), vis.grad.clone()

    v1, t1, g1 = run()
    for prm in m.parameters():
        prm.grad = None
    v2, t2, g2 = run()
    assert torch.equal(v1, v2) and torch.equal(t1, t2) and torch.allclose(g1, g2, rtol=1e-4, atol=1e-6)
    assert torch.isfinite(g1).all() and g1.abs().sum() > 0
    with torch.no_grad():
        ve, te = m.eval()(inp["vis"], inp["pos"], None, inp["text"], inp["text_mask"], {}, detected_feats=inp["det"],
                          detected_mask=inp["det_mask"])
    assert (v1 - ve).abs().max() > 1e-2  # dropout really happened


@pytest.mark.gpu
@pytest.mark.parametrize("name", GRAD_CASES)
def test_cuda_backward_matches_reference_gradient_fixture(name):
    """The CUDA backward (attention / LayerNorm / weight-gradient kernels) against the gradients of the REFERENCE's own
    modules (fixture): relative Frobenius error <= 5e-2 per tensor (tf32 operands; ReLU units within the forward
    tolerance of zero flip between the two paths)."""
    kind = ac.CASES[name][0]
    fx = np.load(os.path.join(HERE, "golden", f"attn_grad_{name}.npz"))
    m = ac.fill_params(build(kind), seed=100 + len(name)).eval().cuda()
    inp = _cuda(ac.make_inputs(name))
    leaves = {k: inp[k].clone().requires_grad_(True) for k in (("vis", "text") if kind == "bi_encoder_layer" else ("query",))}
    out = module_run(kind, m, {**inp, **leaves})
    loss = sum((v * _grad_w(v.shape, 1234 + i).cuda()).sum() for i, v in enumerate(out.values()))
    loss.backward()
    _check_against_grad_fixture(fx, {k: v.grad for k, v in leaves.items()},
                                {k: p.grad for k, p in m.named_parameters() if p.grad is not None}, tol=5e-2)
