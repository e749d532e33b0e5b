"""Module-level parity: PointnetSAModuleVotes / PointnetFPModule.

CPU suite : the oracle restatement (oracle/modules_oracle.py) == golden vectors produced by the
            reference's own Python modules (tests/golden/make_golden_modules.py); state-dict keys of the
            host mirror == the reference's.
GPU suite : the CUDA product (fused tcgen05 path) vs the same golden vectors.  Index outputs exact;
            new_features within the TF32 tolerance stated below (the fused MLP multiplies in TF32 with
            fp32 accumulation: inputs rounded to 10 mantissa bits, 3 chained layers).
"""
import glob
import os

import numpy as np
import pytest
import torch

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
SA_FILES = sorted(glob.glob(os.path.join(GOLDEN, "sa_*_*.npz")))
FP_FILES = sorted(glob.glob(os.path.join(GOLDEN, "fp_*_*.npz")))
ids = lambda fs: [os.path.basename(f)[:-4] for f in fs]  # noqa: E731

# Stated tolerance of the fused TF32 forward against the reference's fp32 forward: operands are
# rounded to TF32 (10 mantissa bits, relative 2^-11) before each of the 3 chained contractions and
# accumulated in fp32, so the error scales with the activation magnitude, not with the element:
#     |got - ref| <= TF32_RTOL * |ref| + TF32_ATOL_REL * max|ref|
# (measured on B200: 0.4e-3 .. 0.9e-3 * max|ref| for one module, 2.3e-3 * max|ref| after two chained
# train-mode modules, where BatchNorm renormalises the first module's error).
TF32_RTOL, TF32_ATOL_REL = 1e-2, 2.5e-3


def assert_close_tf32(got, want, atol_rel=TF32_ATOL_REL):
    torch.testing.assert_close(got, want, rtol=TF32_RTOL, atol=atol_rel * want.abs().max().item())


def _t(a):
    return torch.from_numpy(np.ascontiguousarray(a))


def _sd(z):
    return {k[3:]: _t(z[k]) for k in z.files if k.startswith("sd.")}


def test_fixtures_present():
    assert len(SA_FILES) == 8 and len(FP_FILES) == 2


@pytest.mark.parametrize("path", SA_FILES, ids=ids(SA_FILES))
def test_oracle_sa_module_matches_reference_python(oracle, path):
    from oracle import modules_oracle as mo

    z = np.load(path)
    sd = _sd(z)
    training = path.endswith("_train.npz")
    layers = mo.layers_from_state_dict(sd, "mlp_module.", 3)
    new_xyz, feats, inds, _ = mo.sa_module_forward(_t(z["xyz"]), _t(z["features"]), int(z["npoint"]),
                                                   float(z["radius"]), int(z["nsample"]), layers, True, training,
                                                   update_running=training)
    assert torch.equal(inds, _t(z["inds"]))
    assert torch.equal(new_xyz, _t(z["new_xyz"]))
    torch.testing.assert_close(feats, _t(z["new_features"]), rtol=1e-5, atol=1e-5)
    if training:
        for k in z.files:
            if k.startswith("after."):
                torch.testing.assert_close(sd[k[6:]], _t(z[k]), rtol=1e-5, atol=1e-6)


@pytest.mark.parametrize("path", FP_FILES, ids=ids(FP_FILES))
def test_oracle_fp_module_matches_reference_python(oracle, path):
    from oracle import modules_oracle as mo

    z = np.load(path)
    layers = mo.layers_from_state_dict(_sd(z), "mlp.", 2)
    out = mo.fp_module_forward(_t(z["unknown"]), _t(z["known"]), _t(z["unknow_feats"]), _t(z["known_feats"]), layers,
                               path.endswith("_train.npz"))
    torch.testing.assert_close(out, _t(z["out"]), rtol=1e-5, atol=1e-5)


def test_host_mirror_has_reference_state_dict_keys_and_ctor_semantics():
    from eda_b200.pointnet2.pointnet2_modules import PointnetFPModule, PointnetSAModuleVotes

    z = np.load(SA_FILES[0])
    mlp = [int(v) for v in z["mlp"]]
    m = PointnetSAModuleVotes(npoint=int(z["npoint"]), radius=float(z["radius"]), nsample=int(z["nsample"]), mlp=mlp,
                              use_xyz=True, normalize_xyz=True)
    assert mlp[0] == int(z["mlp"][0]) + 3  # the caller's list is widened in place, like the reference (:204-206)
    m.load_state_dict(_sd(z), strict=True)
    z = np.load(FP_FILES[0])
    PointnetFPModule(mlp=[80, 64, 96]).load_state_dict(_sd(z), strict=True)
    with pytest.raises(RuntimeError, match="CPU not supported"):
        m(torch.rand(1, 64, 3), torch.rand(1, 3, 64))


# ---------------------------------------------------------------------------------------------
def _build_sa(z, fuse=True):
    from eda_b200.pointnet2.pointnet2_modules import PointnetSAModuleVotes

    m = PointnetSAModuleVotes(npoint=int(z["npoint"]), radius=float(z["radius"]), nsample=int(z["nsample"]),
                              mlp=[int(v) for v in z["mlp"]], use_xyz=True, normalize_xyz=True)
    m.load_state_dict(_sd(z), strict=True)
    m.fuse = fuse
    return m.cuda()


@pytest.mark.gpu
@pytest.mark.parametrize("fuse", [True, False], ids=["fused", "unfused"])
@pytest.mark.parametrize("path", SA_FILES, ids=ids(SA_FILES))
def test_cuda_sa_module_matches_reference_python(path, fuse):
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    z = np.load(path)
    training = path.endswith("_train.npz")
    m = _build_sa(z, fuse)
    m.train(training)
    if fuse:
        assert m._fusable(_t(z["features"])) is not None, "configuration should take the fused path"
    feats = _t(z["features"]).cuda().requires_grad_(True)
    new_xyz, out, inds = m(_t(z["xyz"]).cuda(), feats)
    assert torch.equal(inds.cpu(), _t(z["inds"]))
    assert torch.equal(new_xyz.cpu(), _t(z["new_xyz"]))
    assert out.is_contiguous() and out.shape == z["new_features"].shape
    want = _t(z["new_features"])
    err = (out.detach().cpu() - want).abs()
    print(f"{os.path.basename(path)} fuse={fuse}: max abs err {err.max():.3e}, max |ref| {want.abs().max():.3e}")
    if fuse:
        assert_close_tf32(out.detach().cpu(), want)
    else:
        torch.testing.assert_close(out.detach().cpu(), want, rtol=1e-4, atol=1e-4)
    # backward.  Unfused: the reference's op sequence in fp32 -> element-wise agreement.  Fused: the CUDA backward
    # (csrc/sa_bwd.cu) recomputes the layers with tf32 GEMM operands like the forward kernel, so a max-pool arg-max or
    # ReLU gate whose candidates are within ~1e-3 of each other may resolve differently from the fp32 reference and
    # move a whole gradient entry (tests/test_backward.py pins every kernel of the chain exactly on shared
    # pre-activations) -> relative Frobenius error per tensor, bound = the measured flip noise.
    out.backward(_t(z["grad_out"]).cuda())
    if fuse:
        def rel(a, b):
            return ((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-20)).item()

        assert rel(feats.grad.cpu(), _t(z["grad_features"])) <= 0.1
        for k, p in m.named_parameters():
            assert rel(p.grad.cpu(), _t(z["grad." + k])) <= 0.1, k
    else:
        torch.testing.assert_close(feats.grad.cpu(), _t(z["grad_features"]), rtol=2e-3, atol=2e-4)
        for k, p in m.named_parameters():
            torch.testing.assert_close(p.grad.cpu(), _t(z["grad." + k]), rtol=2e-3, atol=2e-3, msg=lambda s: f"{k}: {s}")
    if training:
        sd = m.state_dict()
        for k in z.files:
            if k.startswith("after."):
                torch.testing.assert_close(sd[k[6:]].cpu(), _t(z[k]), rtol=5e-3, atol=5e-4)


@pytest.mark.gpu
@pytest.mark.parametrize("fuse", [True, False], ids=["own_kernels", "unfused"])
@pytest.mark.parametrize("path", FP_FILES, ids=ids(FP_FILES))
def test_cuda_fp_module_matches_reference_python(path, fuse):
    """PointnetFPModule vs the reference's own Python module (fixtures).  fuse=True is the product path: 3-NN ->
    eda_fp_gather_rows -> tcgen05 GEMMs + BatchNorm kernels, forward and backward, no torch layer (tf32 operands: the
    stated TF32 tolerance).  fuse=False is the reference's op-by-op composition on the CUDA ops + torch fp32 layers."""
    from eda_b200.pointnet2.pointnet2_modules import PointnetFPModule

    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    z = np.load(path)
    m = PointnetFPModule(mlp=[80, 64, 96])
    m.load_state_dict(_sd(z), strict=True)
    training = path.endswith("_train.npz")
    m = m.cuda().train(training)
    m.fuse = fuse
    if fuse:
        assert m._rows_layers() is not None, "configuration should take the own-kernel path"
    kf = _t(z["known_feats"]).cuda().requires_grad_(True)
    uf = _t(z["unknow_feats"]).cuda().requires_grad_(True)
    out = m(_t(z["unknown"]).cuda(), _t(z["known"]).cuda(), uf, kf)
    assert out.is_contiguous() and out.shape == z["out"].shape
    out.backward(_t(z["grad_out"]).cuda())
    if fuse:
        assert_close_tf32(out.detach().cpu(), _t(z["out"]))

        def rel(a, b):
            return ((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-20)).item()

        # two tf32 GEMMs + ReLU gates within the forward tolerance of zero: relative Frobenius bound
        assert rel(kf.grad.cpu(), _t(z["grad_known_feats"])) <= 2e-2
        assert rel(uf.grad.cpu(), _t(z["grad_unknow_feats"])) <= 2e-2
        for k, p in m.named_parameters():
            if ("grad." + k) in z.files:
                assert rel(p.grad.cpu(), _t(z["grad." + k])) <= 2e-2, k
        if training:
            sd = m.state_dict()
            for k in z.files:
                if k.startswith("after."):
                    torch.testing.assert_close(sd[k[6:]].cpu(), _t(z[k]), rtol=5e-3, atol=5e-4)
    else:
        torch.testing.assert_close(out.detach().cpu(), _t(z["out"]), rtol=1e-4, atol=1e-4)
        torch.testing.assert_close(kf.grad.cpu(), _t(z["grad_known_feats"]), rtol=1e-3, atol=1e-4)
        torch.testing.assert_close(uf.grad.cpu(), _t(z["grad_unknow_feats"]), rtol=1e-3, atol=1e-4)


@pytest.mark.gpu
@pytest.mark.parametrize("training", [False, True], ids=["eval", "train"])
def test_cuda_sa_fused_vs_unfused_at_backbone_shapes(training):
    """SA1 and SA2 of the backbone (models/backbone_module.py:44-60) at B=2, N=50 000: the fused kernel against
    the unfused composition of the same module (fp32 cuDNN), chained so SA2 consumes the point-major copy."""
    from eda_b200 import synthetic
    from eda_b200.pointnet2.pointnet2_modules import PointnetSAModuleVotes

    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.manual_seed(0)
    pc = synthetic.point_clouds(2, 50000, "surface").cuda()
    xyz = pc[..., :3].contiguous()
    feats = pc[..., 3:].transpose(1, 2).contiguous()
    sa1 = PointnetSAModuleVotes(npoint=2048, radius=0.2, nsample=64, mlp=[3, 64, 64, 128], use_xyz=True,
                                normalize_xyz=True).cuda().train(training)
    sa2 = PointnetSAModuleVotes(npoint=1024, radius=0.4, nsample=32, mlp=[128, 128, 128, 256], use_xyz=True,
                                normalize_xyz=True).cuda().train(training)
    import copy

    ref1, ref2 = copy.deepcopy(sa1), copy.deepcopy(sa2)
    ref1.fuse = ref2.fuse = False
    x1, f1, i1 = sa1(xyz, feats)
    x2, f2, i2 = sa2(x1, f1)
    rx1, rf1, ri1 = ref1(xyz, feats)
    rx2, rf2, ri2 = ref2(rx1, rf1)
    assert torch.equal(i1, ri1) and torch.equal(i2, ri2) and torch.equal(x2, rx2)
    for name, a, b in (("sa1", f1, rf1), ("sa2", f2, rf2)):
        err = (a - b).abs().max().item()
        print(f"{name} training={training}: max abs err {err:.3e} (max |ref| {b.abs().max().item():.3e})")
        assert_close_tf32(a, b, atol_rel=TF32_ATOL_REL if name == "sa1" else 2 * TF32_ATOL_REL)
    if training:
        for (k, v), (_, rv) in zip(sa2.state_dict().items(), ref2.state_dict().items()):
            if "running" in k:
                torch.testing.assert_close(v, rv, rtol=5e-3, atol=5e-4)
