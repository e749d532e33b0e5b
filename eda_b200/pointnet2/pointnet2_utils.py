"""Autograd wrappers and grouping modules with the names and call signatures of the reference's
pointnet2/pointnet2_utils.py (FurthestPointSampling 51-80, GatherOperation 83-117, ThreeNN 120-149,
ThreeInterpolate 152-206, GroupingOperation 209-257, BallQuery 260-291, QueryAndGroup 294-376,
GroupAll 379-426), bound to the CUDA ops of libeda_b200.so through `_ext`."""
import torch
import torch.nn as nn
from torch.autograd import Function

from . import _ext


class FurthestPointSampling(Function):
    @staticmethod
    def forward(ctx, xyz, npoint):
        """xyz (B,N,3) f32 -> (B,npoint) i32 indices, bit-exact with the reference kernel."""
        inds = _ext.furthest_point_sampling(xyz, npoint)
        ctx.mark_non_differentiable(inds)
        return inds

    @staticmethod
    def backward(ctx, grad=None):
        return None, None


furthest_point_sample = FurthestPointSampling.apply


class GatherOperation(Function):
    @staticmethod
    def forward(ctx, features, idx):
        """features (B,C,N), idx (B,npoint) i32 -> (B,C,npoint)."""
        ctx.save_for_backward(idx)
        ctx.n = features.size(2)
        return _ext.gather_points(features, idx)

    @staticmethod
    def backward(ctx, grad_out):
        (idx,) = ctx.saved_tensors
        return _ext.gather_points_grad(grad_out.contiguous(), idx, ctx.n), None


gather_operation = GatherOperation.apply


class ThreeNN(Function):
    @staticmethod
    def forward(ctx, unknown, known):
        """unknown (B,n,3), known (B,m,3) -> dist (B,n,3) (Euclidean, i.e. sqrt of the op's dist2), idx (B,n,3)."""
        dist2, idx = _ext.three_nn(unknown, known)
        dist = torch.sqrt(dist2)
        ctx.mark_non_differentiable(dist, idx)
        return dist, idx

    @staticmethod
    def backward(ctx, a=None, b=None):
        return None, None


three_nn = ThreeNN.apply


def three_nn_squared(unknown, known):
    """The native op's own outputs (SQUARED distances, indices), non-differentiable like ThreeNN: what the fused
    feature-propagation path consumes (the sqrt and the weight arithmetic happen inside eda_fp_gather_rows)."""
    with torch.no_grad():
        return _ext.three_nn(unknown.detach().contiguous(), known.detach().contiguous())


class ThreeInterpolate(Function):
    @staticmethod
    def forward(ctx, features, idx, weight):
        """features (B,c,m), idx (B,n,3) i32, weight (B,n,3) -> (B,c,n)."""
        ctx.save_for_backward(idx, weight)
        ctx.m = features.size(2)
        return _ext.three_interpolate(features, idx, weight)

    @staticmethod
    def backward(ctx, grad_out):
        idx, weight = ctx.saved_tensors
        return _ext.three_interpolate_grad(grad_out.contiguous(), idx, weight, ctx.m), None, None


three_interpolate = ThreeInterpolate.apply


class GroupingOperation(Function):
    @staticmethod
    def forward(ctx, features, idx):
        """features (B,C,N), idx (B,npoint,nsample) i32 -> (B,C,npoint,nsample)."""
        ctx.save_for_backward(idx)
        ctx.n = features.size(2)
        return _ext.group_points(features, idx)

    @staticmethod
    def backward(ctx, grad_out):
        (idx,) = ctx.saved_tensors
        return _ext.group_points_grad(grad_out.contiguous(), idx, ctx.n), None


grouping_operation = GroupingOperation.apply


class BallQuery(Function):
    @staticmethod
    def forward(ctx, radius, nsample, xyz, new_xyz):
        """xyz (B,N,3), new_xyz (B,npoint,3) -> (B,npoint,nsample) i32 (note the argument order)."""
        inds = _ext.ball_query(new_xyz, xyz, radius, nsample)
        ctx.mark_non_differentiable(inds)
        return inds

    @staticmethod
    def backward(ctx, a=None):
        return None, None, None, None


ball_query = BallQuery.apply


class QueryAndGroup(nn.Module):
    """Ball query + grouping (unfused composition; PointnetSAModuleVotes fuses this with its MLP when it can)."""

    def __init__(self, radius, nsample, use_xyz=True, ret_grouped_xyz=False, normalize_xyz=False,
                 sample_uniformly=False, ret_unique_cnt=False):
        super().__init__()
        self.radius, self.nsample, self.use_xyz = radius, nsample, use_xyz
        self.ret_grouped_xyz = ret_grouped_xyz
        self.normalize_xyz = normalize_xyz
        self.sample_uniformly = sample_uniformly
        self.ret_unique_cnt = ret_unique_cnt
        if self.ret_unique_cnt:
            assert self.sample_uniformly

    def query(self, xyz, new_xyz):
        """The neighbour lists alone (plus unique counts when sampling uniformly)."""
        idx = ball_query(self.radius, self.nsample, xyz, new_xyz)
        unique_cnt = None
        if self.sample_uniformly:
            # resample the padded tail uniformly from the distinct hits (pointnet2_utils.py:334-343);
            # host-side and RNG-dependent in the reference as well
            unique_cnt = torch.zeros((idx.shape[0], idx.shape[1]))
            for b in range(idx.shape[0]):
                for r in range(idx.shape[1]):
                    uniq = torch.unique(idx[b, r, :])
                    k = uniq.shape[0]
                    unique_cnt[b, r] = k
                    pick = torch.randint(0, k, (self.nsample - k,), dtype=torch.long)
                    idx[b, r, :] = torch.cat((uniq, uniq[pick]))
        return idx, unique_cnt

    def group(self, xyz, new_xyz, features, idx):
        grouped_xyz = grouping_operation(xyz.transpose(1, 2).contiguous(), idx)  # (B,3,npoint,nsample)
        grouped_xyz = grouped_xyz - new_xyz.transpose(1, 2).unsqueeze(-1)
        if self.normalize_xyz:
            grouped_xyz = grouped_xyz / self.radius
        if features is not None:
            grouped_features = grouping_operation(features, idx)
            new_features = torch.cat([grouped_xyz, grouped_features], dim=1) if self.use_xyz else grouped_features
        else:
            assert self.use_xyz, "Cannot have not features and not use xyz as a feature!"
            new_features = grouped_xyz
        return new_features, grouped_xyz

    def forward(self, xyz, new_xyz, features=None):
        idx, unique_cnt = self.query(xyz, new_xyz)
        new_features, grouped_xyz = self.group(xyz, new_xyz, features, idx)
        ret = [new_features]
        if self.ret_grouped_xyz:
            ret.append(grouped_xyz)
        if self.ret_unique_cnt:
            ret.append(unique_cnt)
        return ret[0] if len(ret) == 1 else tuple(ret)


class GroupAll(nn.Module):
    """Treats the whole cloud as one group: (B, 3 + C, 1, N)."""

    def __init__(self, use_xyz=True, ret_grouped_xyz=False):
        super().__init__()
        self.use_xyz = use_xyz
        self.ret_grouped_xyz = ret_grouped_xyz

    def forward(self, xyz, new_xyz, features=None):
        grouped_xyz = xyz.transpose(1, 2).unsqueeze(2)
        if features is not None:
            grouped_features = features.unsqueeze(2)
            new_features = torch.cat([grouped_xyz, grouped_features], dim=1) if self.use_xyz else grouped_features
        else:
            new_features = grouped_xyz
        return (new_features, grouped_xyz) if self.ret_grouped_xyz else new_features
