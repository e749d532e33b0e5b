"""PointNet++ set-abstraction and feature-propagation modules with the constructor / forward
signatures and state-dict keys of the reference's pointnet2/pointnet2_modules.py
(PointnetSAModuleVotes 164-272, PointnetFPModule 356-416), running on the sm_100a kernels of
libeda_b200.so.

PointnetSAModuleVotes.forward takes the FUSED path — FPS -> gather -> ball query -> one kernel for
grouping + 3-layer MLP + max-pool (tcgen05) — whenever the configuration is the one the backbone
uses (ball-query grouping, max pooling, a 3-layer conv/BN/ReLU SharedMLP within the kernel's
width limits).  Other configurations (avg / rbf pooling, GroupAll, uniform resampling, other
depths) run the same maths through the unfused CUDA ops + torch layers, like the reference.
"""
from typing import List

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import fused
from . import pointnet2_utils
from . import pytorch_utils as pt_utils


class PointnetSAModuleVotes(nn.Module):
    def __init__(self, *, mlp: List[int], npoint: int = None, radius: float = None, nsample: int = None,
                 bn: bool = True, use_xyz: bool = True, pooling: str = "max", sigma: float = None,
                 normalize_xyz: bool = False, sample_uniformly: bool = False, ret_unique_cnt: bool = False):
        super().__init__()
        self.npoint = npoint
        self.radius = radius
        self.nsample = nsample
        self.pooling = pooling
        self.use_xyz = use_xyz
        self.sigma = sigma if sigma is not None else (self.radius / 2 if self.radius is not None else None)
        self.normalize_xyz = normalize_xyz
        self.ret_unique_cnt = ret_unique_cnt
        if npoint is not None:
            self.grouper = pointnet2_utils.QueryAndGroup(
                radius, nsample, use_xyz=use_xyz, ret_grouped_xyz=True, normalize_xyz=normalize_xyz,
                sample_uniformly=sample_uniformly, ret_unique_cnt=ret_unique_cnt)
        else:
            self.grouper = pointnet2_utils.GroupAll(use_xyz, ret_grouped_xyz=True)
        mlp_spec = mlp  # the reference widens the CALLER's list in place (pointnet2_modules.py:204-206); so do we
        if use_xyz and len(mlp_spec) > 0:
            mlp_spec[0] += 3
        self.mlp_module = pt_utils.SharedMLP(mlp_spec, bn=bn)
        self.fuse = True  # set False to force the unfused composition (tests compare the two)

    # ------------------------------------------------------------------------------------------
    def _fusable(self, features):
        if not self.fuse or self.npoint is None or self.pooling != "max" or not self.use_xyz:
            return None
        if self.grouper.sample_uniformly or self.ret_unique_cnt:
            return None
        layers = self.mlp_module.fusable_layers()
        if layers is None or len(layers) != 3:
            return None
        C = 0 if features is None else features.size(1)
        if layers[0][0].in_channels != C + 3:
            return None
        if not fused.fusable(C, [conv.out_channels for conv, _ in layers], self.nsample):
            return None
        return layers

    def forward(self, xyz: torch.Tensor, features: torch.Tensor = None, inds: torch.Tensor = None):
        """xyz (B,N,3), features (B,C,N), inds (B,npoint) optional ->
        new_xyz (B,npoint,3), new_features (B,mlp[-1],npoint), inds (B,npoint) [, unique_cnt]."""
        xyz_flipped = xyz.transpose(1, 2).contiguous()
        if inds is None:
            inds = pointnet2_utils.furthest_point_sample(xyz, self.npoint)
        else:
            assert inds.shape[1] == self.npoint
        new_xyz = (pointnet2_utils.gather_operation(xyz_flipped, inds).transpose(1, 2).contiguous()
                   if self.npoint is not None else None)

        layers = self._fusable(features) if xyz.is_cuda else None
        if layers is not None:
            idx = pointnet2_utils.ball_query(self.radius, self.nsample, xyz, new_xyz)
            new_features, feat_pm = fused.FusedSAFunction.apply(self, xyz, new_xyz, features, idx,
                                                                *fused.sa_params(layers))
            fused.attach_point_major(new_features, feat_pm)  # lets the next fused layer skip a transpose
            return new_xyz, new_features, inds

        unique_cnt = None
        if not self.ret_unique_cnt:
            grouped_features, grouped_xyz = self.grouper(xyz, new_xyz, features)
        else:
            grouped_features, grouped_xyz, unique_cnt = self.grouper(xyz, new_xyz, features)
        new_features = self.mlp_module(grouped_features)  # (B, mlp[-1], npoint, nsample)
        if self.pooling == "max":
            new_features = F.max_pool2d(new_features, kernel_size=[1, new_features.size(3)])
        elif self.pooling == "avg":
            new_features = F.avg_pool2d(new_features, kernel_size=[1, new_features.size(3)])
        elif self.pooling == "rbf":
            rbf = torch.exp(-1 * grouped_xyz.pow(2).sum(1, keepdim=False) / (self.sigma ** 2) / 2)
            new_features = torch.sum(new_features * rbf.unsqueeze(1), -1, keepdim=True) / float(self.nsample)
        new_features = new_features.squeeze(-1)
        if not self.ret_unique_cnt:
            return new_xyz, new_features, inds
        return new_xyz, new_features, inds, unique_cnt


class PointnetFPModule(nn.Module):
    """Feature propagation: 3-NN inverse-distance interpolation + skip concat + SharedMLP
    (pointnet2/pointnet2_modules.py:356-416).

    On CUDA the whole module runs on this package's kernels: eda_three_nn -> eda_fp_gather_rows (weights +
    interpolation + concat in one pass, row-major) -> per layer tcgen05 GEMM + BatchNorm statistics / apply kernels
    (eda_b200/rows_mlp.py), forward and backward; the result returns to the reference's (B, C, n) layout through the
    transpose kernel and carries its point-major copy along for the next consumer."""

    def __init__(self, *, mlp: List[int], bn: bool = True):
        super().__init__()
        self.mlp = pt_utils.SharedMLP(mlp, bn=bn)
        self.fuse = True  # False: the reference's op-by-op composition on torch layers (tests compare the two)

    def _rows_layers(self):
        from .. import rows_mlp

        layers = self.mlp.fusable_layers()
        if layers is None or any(bn is None for _, bn in layers):
            return None
        if any(conv.out_channels % 16 or conv.out_channels > 320 or conv.in_channels % 16 for conv, _ in layers):
            return None
        return [rows_mlp.Layer(conv.weight, conv.bias, bn, True, conv, "w") for conv, bn in layers]

    def forward(self, unknown: torch.Tensor, known: torch.Tensor, unknow_feats: torch.Tensor,
                known_feats: torch.Tensor) -> torch.Tensor:
        """unknown (B,n,3), known (B,m,3), unknow_feats (B,C1,n), known_feats (B,C2,m) -> (B,mlp[-1],n)."""
        layers = self._rows_layers() if (self.fuse and known is not None and known_feats.is_cuda) else None
        if layers is not None:
            from .. import rows_mlp

            B, n = unknown.size(0), unknown.size(1)
            dist2, idx = pointnet2_utils.three_nn_squared(unknown, known)
            x0 = rows_mlp.fp_rows(dist2, idx, known_feats, unknow_feats)
            out_pm = rows_mlp.rows_mlp(x0, layers).view(B, n, -1)
            new_features = rows_mlp.transpose_last2(out_pm)
            fused.attach_point_major(new_features, out_pm.detach())
            return new_features
        if known is not None:
            dist, idx = pointnet2_utils.three_nn(unknown, known)
            dist_recip = 1.0 / (dist + 1e-8)
            norm = torch.sum(dist_recip, dim=2, keepdim=True)
            weight = dist_recip / norm
            interpolated = pointnet2_utils.three_interpolate(known_feats, idx, weight)
        else:
            interpolated = known_feats.expand(*known_feats.size()[0:2], unknown.size(1))
        new_features = torch.cat([interpolated, unknow_feats], dim=1) if unknow_feats is not None else interpolated
        new_features = self.mlp(new_features.unsqueeze(-1))
        return new_features.squeeze(-1)
