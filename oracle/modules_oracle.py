"""CPU restatement of the reference's set-abstraction / feature-propagation MODULE maths
(TEST INFRASTRUCTURE).  Plain torch fp32 on CPU tensors + the C oracle for the index ops.

Follows, with file:line of the reference:
  sa_module_forward   pointnet2/pointnet2_modules.py:210-272 (PointnetSAModuleVotes.forward, max pooling)
                      pointnet2/pointnet2_utils.py:317-376   (QueryAndGroup.forward)
                      pointnet2/pytorch_utils.py:11-36,67-120 (SharedMLP = [1x1 conv, BN2d, ReLU] x L)
  fp_module_forward   pointnet2/pointnet2_modules.py:371-416 (PointnetFPModule.forward)
                      pointnet2/pointnet2_utils.py:120-206   (three_nn returns sqrt(dist2); three_interpolate)

Pinned against the reference's own Python modules (imported from /root/reference with `_ext` bound to
the C oracle) by tests/golden/make_golden_modules.py -> tests/golden/sa_*.npz, fp_*.npz.
Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may import this file.
"""
import torch
import torch.nn.functional as F

from . import pointnet2_oracle as ops


def shared_mlp(x, layers, training, eps=1e-5, momentum=0.1, update_running=False):
    """x (B,C,H,W).  layers: list of dicts {weight (Cout,Cin,1,1), bias|None, bn: None | dict(weight,bias,
    running_mean, running_var)}.  Train mode normalises with batch statistics (biased variance)."""
    for L in layers:
        x = F.conv2d(x, L["weight"], L.get("bias"))
        bn = L.get("bn")
        if bn is not None:
            if training:
                rm = bn["running_mean"] if update_running else None
                rv = bn["running_var"] if update_running else None
                x = F.batch_norm(x, rm, rv, bn["weight"], bn["bias"], True, momentum, eps)
            else:
                x = F.batch_norm(x, bn["running_mean"], bn["running_var"], bn["weight"], bn["bias"], False, momentum, eps)
        x = F.relu(x)
    return x


def query_and_group(xyz, new_xyz, features, radius, nsample, normalize_xyz, use_xyz=True):
    idx = ops.ball_query(new_xyz, xyz, radius, nsample)
    grouped_xyz = ops.group_points(xyz.transpose(1, 2).contiguous(), idx)
    grouped_xyz = grouped_xyz - new_xyz.transpose(1, 2).unsqueeze(-1)
    if normalize_xyz:
        grouped_xyz = grouped_xyz / radius
    if features is None:
        return grouped_xyz, idx
    grouped = ops.group_points(features.contiguous(), idx)
    return (torch.cat([grouped_xyz, grouped], dim=1) if use_xyz else grouped), idx


def sa_module_forward(xyz, features, npoint, radius, nsample, layers, normalize_xyz, training, inds=None,
                      update_running=False):
    """-> new_xyz (B,npoint,3), new_features (B,C_out,npoint), inds (B,npoint) int32, idx (B,npoint,nsample)."""
    if inds is None:
        inds = ops.furthest_point_sampling(xyz.contiguous(), npoint)
    new_xyz = ops.gather_points(xyz.transpose(1, 2).contiguous(), inds).transpose(1, 2).contiguous()
    grouped, idx = query_and_group(xyz.contiguous(), new_xyz, features, radius, nsample, normalize_xyz)
    y = shared_mlp(grouped, layers, training, update_running=update_running)
    y = F.max_pool2d(y, kernel_size=[1, y.size(3)]).squeeze(-1)
    return new_xyz, y, inds, idx


def fp_module_forward(unknown, known, unknow_feats, known_feats, layers, training, update_running=False):
    dist2, idx = ops.three_nn(unknown.contiguous(), known.contiguous())
    dist = torch.sqrt(dist2)
    dist_recip = 1.0 / (dist + 1e-8)
    weight = dist_recip / torch.sum(dist_recip, dim=2, keepdim=True)
    interpolated = ops.three_interpolate(known_feats.contiguous(), idx, weight.contiguous())
    x = torch.cat([interpolated, unknow_feats], dim=1) if unknow_feats is not None else interpolated
    return shared_mlp(x.unsqueeze(-1), layers, training, update_running=update_running).squeeze(-1)


def layers_from_state_dict(sd, prefix, nlayers):
    """Builds the `layers` list from reference-named state-dict entries (`<prefix>layer{i}.conv.weight`, ...)."""
    out = []
    for i in range(nlayers):
        base = f"{prefix}layer{i}."
        L = {"weight": sd[base + "conv.weight"], "bias": sd.get(base + "conv.bias"), "bn": None}
        if base + "bn.bn.weight" in sd:
            L["bn"] = {k: sd[base + "bn.bn." + k] for k in ("weight", "bias", "running_mean", "running_var")}
        out.append(L)
    return out


BACKBONE_SA = [  # models/backbone_module.py:44-78: (npoint, radius, nsample)
    (2048, 0.2, 64), (1024, 0.4, 32), (512, 0.8, 16), (256, 1.2, 16)]


def backbone_forward(sd, pointcloud, training=False):
    """Pointnet2Backbone.forward (models/backbone_module.py:92-144) over a reference-named state dict
    (`sa{1..4}.mlp_module.layer{i}.*`, `fp{1,2}.mlp.layer{i}.*`).  Returns the end_points dict."""
    xyz = pointcloud[..., 0:3].contiguous()
    features = pointcloud[..., 3:].transpose(1, 2).contiguous() if pointcloud.size(-1) > 3 else None
    ep = {}
    for i, (npoint, radius, nsample) in enumerate(BACKBONE_SA, start=1):
        layers = layers_from_state_dict(sd, f"sa{i}.mlp_module.", 3)
        xyz, features, inds, _ = sa_module_forward(xyz, features, npoint, radius, nsample, layers, True, training)
        ep[f"sa{i}_xyz"], ep[f"sa{i}_features"], ep[f"sa{i}_inds"] = xyz, features, inds
    f = fp_module_forward(ep["sa3_xyz"], ep["sa4_xyz"], ep["sa3_features"], ep["sa4_features"],
                          layers_from_state_dict(sd, "fp1.mlp.", 2), training)
    f = fp_module_forward(ep["sa2_xyz"], ep["sa3_xyz"], ep["sa2_features"], f,
                          layers_from_state_dict(sd, "fp2.mlp.", 2), training)
    ep["fp2_features"], ep["fp2_xyz"] = f, ep["sa2_xyz"]
    ep["fp2_inds"] = ep["sa1_inds"][:, :ep["fp2_xyz"].shape[1]]
    return ep
