"""Query-generation and prediction-head modules with the class names, constructor / forward signatures and state-dict
keys of the reference's models/modules.py (PointsObjClsModule 19-49, PositionEmbeddingLearned 52-69,
GeneralSamplingModule 72-91, ThreeLayerMLP 94-114, ClsAgnosticPredictHead 117-178) — SURVEY.md 8(f) rank 2, the layer of
tiny Conv1d + BatchNorm1d + Dropout kernels that `BeaUTyDETR.forward` interleaves with every decoder layer
(models/bdetr.py:187-205,266-337).

Every 1x1 Conv1d is a row-major GEMM on the tcgen05 kernel, BatchNorm1d (batch statistics in training mode,
synchronised over ranks for SyncBatchNorm; running statistics folded into the GEMM in inference), ReLU and Dropout run
on this package's kernels through eda_b200/rows_mlp.py, forward and backward; the three heads of a
ClsAgnosticPredictHead that share their input start from one transposed copy of it.  Parameters live in the same
torch containers the reference builds, so initialisation, `state_dict()` and checkpoint loading are unchanged.
There is no CPU path: CPU tensors raise RuntimeError like everywhere else in the package.
"""
import numpy as np
import torch.nn as nn

from . import attn_ops as ops
from . import rows_mlp
from .encoder_decoder_layers import PositionEmbeddingLearned  # same module, same keys (models/modules.py:52-69)
from .pointnet2.pointnet2_utils import gather_operation

__all__ = ["PointsObjClsModule", "PositionEmbeddingLearned", "GeneralSamplingModule", "ThreeLayerMLP",
           "ClsAgnosticPredictHead"]


def _rows(x):
    """(B, C, N) -> rows (B*N, C) on the transpose kernel (differentiable)."""
    B, C, N = x.shape
    return rows_mlp.transpose_last2(x).reshape(B * N, C), B, N


def _channels_first(rows, B, N):
    """rows (B*N, C) -> (B, C, N)."""
    return rows_mlp.transpose_last2(rows.view(B, N, -1))


class PointsObjClsModule(nn.Module):
    """Object candidate point prediction from seed point features."""

    def __init__(self, seed_feature_dim):
        super().__init__()
        self.in_dim = seed_feature_dim
        self.conv1 = nn.Conv1d(self.in_dim, self.in_dim, 1)
        self.bn1 = nn.BatchNorm1d(self.in_dim)
        self.conv2 = nn.Conv1d(self.in_dim, self.in_dim, 1)
        self.bn2 = nn.BatchNorm1d(self.in_dim)
        self.conv3 = nn.Conv1d(self.in_dim, 1, 1)

    def forward(self, seed_features):
        """seed_features (B, C, num_seed) -> logits (B, 1, num_seed)."""
        ops._require_cuda(seed_features, "seed_features")
        x, B, N = _rows(seed_features)
        layers = [rows_mlp.Layer(self.conv1.weight, self.conv1.bias, self.bn1, True, self.conv1, "w"),
                  rows_mlp.Layer(self.conv2.weight, self.conv2.bias, self.bn2, True, self.conv2, "w"),
                  rows_mlp.Layer(self.conv3.weight, self.conv3.bias, None, False, self.conv3, "w")]
        return _channels_first(rows_mlp.rows_mlp(x, layers), B, N)


class GeneralSamplingModule(nn.Module):
    def __init__(self):
        super().__init__()

    def forward(self, xyz, features, sample_inds):
        """xyz (B,K,3), features (B,C,K), sample_inds (B,M) int32 -> (B,M,3), (B,C,M), sample_inds."""
        xyz_flipped = xyz.transpose(1, 2).contiguous()
        new_xyz = gather_operation(xyz_flipped, sample_inds).transpose(1, 2).contiguous()
        new_features = gather_operation(features, sample_inds).contiguous()
        return new_xyz, new_features, sample_inds


class ThreeLayerMLP(nn.Module):
    """A 3-layer MLP with normalization and dropout."""

    def __init__(self, dim, out_dim):
        super().__init__()
        self.net = nn.Sequential(
            nn.Conv1d(dim, dim, 1, bias=False),
            nn.BatchNorm1d(dim),
            nn.ReLU(),
            nn.Dropout(0.3),
            nn.Conv1d(dim, dim, 1, bias=False),
            nn.BatchNorm1d(dim),
            nn.ReLU(),
            nn.Dropout(0.3),
            nn.Conv1d(dim, out_dim, 1)
        )

    def rows_layers(self):
        n = self.net
        return [rows_mlp.Layer(n[0].weight, None, n[1], True, n[0], "w", dropout=n[3]),
                rows_mlp.Layer(n[4].weight, None, n[5], True, n[4], "w", dropout=n[7]),
                rows_mlp.Layer(n[8].weight, n[8].bias, None, False, n[8], "w")]

    def forward_rows(self, x_rows):
        """x_rows (R, dim) -> (R, out_dim)."""
        return rows_mlp.rows_mlp(x_rows, self.rows_layers())

    def forward(self, x):
        """Forward pass, x can be (B, dim, N)."""
        ops._require_cuda(x)
        rows, B, N = _rows(x)
        return _channels_first(self.forward_rows(rows), B, N)


class ClsAgnosticPredictHead(nn.Module):
    def __init__(self, num_class, num_heading_bin, num_proposal, seed_feat_dim=256, objectness=True, heading=False,
                 compute_sem_scores=True):
        super().__init__()
        self.num_class = num_class
        self.num_heading_bin = num_heading_bin
        self.num_proposal = num_proposal
        self.seed_feat_dim = seed_feat_dim
        self.objectness = objectness
        self.heading = heading
        self.compute_sem_scores = compute_sem_scores
        if objectness:
            self.objectness_scores_head = ThreeLayerMLP(seed_feat_dim, 1)
        self.center_residual_head = ThreeLayerMLP(seed_feat_dim, 3)
        if heading:
            self.heading_class_head = nn.Conv1d(seed_feat_dim, num_heading_bin, 1)
            self.heading_residual_head = nn.Conv1d(seed_feat_dim, num_heading_bin, 1)
        self.size_pred_head = ThreeLayerMLP(seed_feat_dim, 3)
        if compute_sem_scores:
            self.sem_cls_scores_head = ThreeLayerMLP(seed_feat_dim, self.num_class)

    def forward(self, features, base_xyz, end_points, prefix=''):
        """features (B, C, num_proposal), base_xyz (B, num_proposal, 3) -> center, pred_size; fills end_points."""
        batch_size = features.shape[0]
        num_proposal = features.shape[-1]
        ops._require_cuda(features, "features")
        # the heads consume (proposal, channel) rows and the reference transposes every head's output back to
        # (B, num_proposal, .) anyway: one transposed copy of the input, no transpose on the way out
        rows, B, N = _rows(features)
        head = lambda m: m.forward_rows(rows).view(B, N, -1)  # noqa: E731
        conv = lambda c: rows_mlp.rows_mlp(rows, [rows_mlp.Layer(c.weight, c.bias, None, False, c, "w")]).view(B, N, -1)  # noqa: E731

        if self.objectness:
            objectness_scores = head(self.objectness_scores_head)  # (batch_size, num_proposal, 1)
            end_points[f'{prefix}objectness_scores'] = objectness_scores.squeeze(-1)

        center_residual = head(self.center_residual_head)  # (B, num_proposal, 3)
        center = base_xyz + center_residual

        if self.heading:
            heading_scores = conv(self.heading_class_head)
            heading_residuals_normalized = conv(self.heading_residual_head)
            heading_residuals = heading_residuals_normalized * (np.pi / self.num_heading_bin)
            end_points[f'{prefix}heading_scores'] = heading_scores
            end_points[f'{prefix}heading_residuals_normalized'] = heading_residuals_normalized
            end_points[f'{prefix}heading_residuals'] = heading_residuals

        pred_size = head(self.size_pred_head).reshape([batch_size, num_proposal, 3])

        if self.compute_sem_scores:
            sem_cls_scores = head(self.sem_cls_scores_head)  # (B, num_proposal, num_class)

        end_points[f'{prefix}base_xyz'] = base_xyz
        end_points[f'{prefix}center'] = center
        end_points[f'{prefix}pred_size'] = pred_size
        if self.compute_sem_scores:
            end_points[f'{prefix}sem_cls_scores'] = sem_cls_scores
        return center, pred_size
