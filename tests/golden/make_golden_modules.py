"""Golden vectors for the MODULE-level path, produced by the reference's own Python
(/root/reference/pointnet2/pointnet2_modules.py, imported unmodified) in THIS container on CPU.
The reference's native `_ext` is CUDA-only, so `pointnet2._ext` is bound to the C oracle, which is
itself pinned bit-exactly against the reference's compiled `_ext` (tests/golden/ext_*.npz).

    python tests/golden/make_golden_modules.py      # needs /root/reference; writes tests/golden/{sa,fp}_*.npz
"""
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
REF = "/root/reference"

from eda_b200 import synthetic  # noqa: E402
from oracle import pointnet2_oracle as orc  # noqa: E402


def import_reference_modules():
    pkg = types.ModuleType("pointnet2")
    pkg.__path__ = []
    ext = types.ModuleType("pointnet2._ext")
    for name in ("furthest_point_sampling", "gather_points", "gather_points_grad", "ball_query", "group_points",
                 "group_points_grad", "three_nn", "three_interpolate", "three_interpolate_grad"):
        setattr(ext, name, getattr(orc, name))
    pkg._ext = ext
    sys.modules["pointnet2"] = pkg
    sys.modules["pointnet2._ext"] = ext
    sys.path.insert(0, os.path.join(REF, "pointnet2"))
    import pointnet2_modules  # the reference's

    assert pointnet2_modules.__file__.startswith(REF)
    return pointnet2_modules


def randomise_bn(module, g):
    """Non-trivial BatchNorm parameters / running stats so eval-mode folding is actually tested."""
    for m in module.modules():
        if isinstance(m, torch.nn.BatchNorm2d):
            m.weight.data = 0.5 + torch.rand(m.num_features, generator=g)
            m.bias.data = 0.2 * torch.randn(m.num_features, generator=g)
            m.running_mean.data = 0.1 * torch.randn(m.num_features, generator=g)
            m.running_var.data = 0.5 + torch.rand(m.num_features, generator=g)


def sd_np(module):
    return {"sd." + k: v.detach().numpy().copy() for k, v in module.state_dict().items()}


def main():
    ref = import_reference_modules()
    torch.manual_seed(0)
    cases = [
        # name, B, N, C, npoint, radius, nsample, mlp, family
        ("sa_a", 2, 2048, 3, 256, 0.3, 16, [3, 32, 32, 64], "surface"),
        ("sa_b", 2, 1024, 16, 128, 0.4, 32, [16, 64, 64, 128], "uniform"),
        ("sa_c", 1, 4096, 3, 64, 0.25, 64, [3, 64, 64, 128], "surface"),
        ("sa_d", 1, 512, 128, 64, 0.8, 16, [128, 128, 128, 256], "dup"),
    ]
    for name, B, N, C, npoint, radius, nsample, mlp, family in cases:
        g = torch.Generator().manual_seed(hash(name) % 1000 if False else sum(map(ord, name)))
        xyz = synthetic.point_clouds(B, N, family, seed=11 + N, channels=0)
        feats = torch.randn(B, C, N, generator=g)
        for mode in ("eval", "train"):
            torch.manual_seed(1)
            m = ref.PointnetSAModuleVotes(npoint=npoint, radius=radius, nsample=nsample, mlp=list(mlp), use_xyz=True,
                                          normalize_xyz=True)
            randomise_bn(m, torch.Generator().manual_seed(7))
            sd0 = sd_np(m)
            m.train(mode == "train")
            f = feats.clone().requires_grad_(True)
            new_xyz, new_feat, inds = m(xyz, f)
            go = torch.randn(new_feat.shape, generator=torch.Generator().manual_seed(3))
            new_feat.backward(go)
            out = dict(xyz=xyz.numpy(), features=feats.numpy(), new_xyz=new_xyz.detach().numpy(),
                       new_features=new_feat.detach().numpy(), inds=inds.numpy(), grad_out=go.numpy(),
                       grad_features=f.grad.numpy(), npoint=npoint, radius=np.float32(radius), nsample=nsample,
                       mlp=np.array(mlp), **sd0)
            for k, p in m.named_parameters():
                out["grad." + k] = p.grad.numpy()
            if mode == "train":
                for k, v in m.state_dict().items():
                    if "running" in k:
                        out["after." + k] = v.numpy().copy()
            np.savez_compressed(os.path.join(HERE, f"{name}_{mode}.npz"), **out)
            print("wrote", name, mode, tuple(new_feat.shape))

    # feature propagation (FP1-like, reduced)
    g = torch.Generator().manual_seed(5)
    unknown = synthetic.point_clouds(2, 256, "surface", seed=3, channels=0)
    known = unknown[:, :96].contiguous()
    uf = torch.randn(2, 32, 256, generator=g)
    kf = torch.randn(2, 48, 96, generator=g)
    for mode in ("eval", "train"):
        torch.manual_seed(2)
        m = ref.PointnetFPModule(mlp=[80, 64, 96])
        randomise_bn(m, torch.Generator().manual_seed(8))
        sd0 = sd_np(m)
        m.train(mode == "train")
        k = kf.clone().requires_grad_(True)
        u = uf.clone().requires_grad_(True)
        y = m(unknown, known, u, k)
        go = torch.randn(y.shape, generator=torch.Generator().manual_seed(4))
        y.backward(go)
        out = dict(unknown=unknown.numpy(), known=known.numpy(), unknow_feats=uf.numpy(), known_feats=kf.numpy(),
                   out=y.detach().numpy(), grad_out=go.numpy(), grad_known_feats=k.grad.numpy(),
                   grad_unknow_feats=u.grad.numpy(), **sd0)
        for kk, p in m.named_parameters():
            out["grad." + kk] = p.grad.numpy()
        np.savez_compressed(os.path.join(HERE, f"fp_a_{mode}.npz"), **out)
        print("wrote fp_a", mode, tuple(y.shape))


if __name__ == "__main__":
    main()
