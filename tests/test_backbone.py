"""Pointnet2Backbone mirror (models/backbone_module.py:26-144)."""
import pytest
import torch

from eda_b200 import synthetic
from eda_b200.backbone_module import Pointnet2Backbone

REF_KEYS_SAMPLE = [  # names a reference checkpoint carries
    "sa1.mlp_module.layer0.conv.weight", "sa1.mlp_module.layer0.bn.bn.running_mean",
    "sa4.mlp_module.layer2.bn.bn.num_batches_tracked", "fp1.mlp.layer0.conv.weight", "fp2.mlp.layer1.bn.bn.weight"]


def _randomise_bn(m, seed):
    g = torch.Generator().manual_seed(seed)
    for mod in m.modules():
        if isinstance(mod, torch.nn.BatchNorm2d):
            mod.running_mean.copy_(0.1 * torch.randn(mod.num_features, generator=g))
            mod.running_var.copy_(0.5 + torch.rand(mod.num_features, generator=g))
            mod.weight.data.copy_(1 + 0.1 * torch.randn(mod.num_features, generator=g))
            mod.bias.data.copy_(0.1 * torch.randn(mod.num_features, generator=g))


def test_backbone_state_dict_keys_and_shapes():
    m = Pointnet2Backbone(input_feature_dim=3, width=1)
    sd = m.state_dict()
    for k in REF_KEYS_SAMPLE:
        assert k in sd, k
    assert sd["sa1.mlp_module.layer0.conv.weight"].shape == (64, 6, 1, 1)
    assert sd["sa2.mlp_module.layer0.conv.weight"].shape == (128, 131, 1, 1)
    assert sd["fp2.mlp.layer1.conv.weight"].shape == (288, 256, 1, 1)
    assert len(sd) == 16 * 6  # 16 conv+BN layers x (conv.weight + 5 BN entries)


@pytest.mark.gpu
def test_backbone_matches_oracle_and_overlap_is_invisible():
    from oracle import modules_oracle as mo

    torch.manual_seed(0)
    m = Pointnet2Backbone(input_feature_dim=3, width=1).eval()
    _randomise_bn(m, 1)
    pc = synthetic.point_clouds(2, 6000, "surface")
    with torch.no_grad():
        want = mo.backbone_forward(m.state_dict(), pc)
        md = m.cuda()
        got = md(pc.cuda())
        md.overlap_fps = False
        plain = md(pc.cuda())
    for k in ("sa1_inds", "sa2_inds", "fp2_inds"):
        assert torch.equal(got[k].cpu(), want[k]), k
    for k in ("sa1_xyz", "sa2_xyz", "sa3_xyz", "sa4_xyz", "fp2_xyz"):
        assert torch.equal(got[k].cpu(), want[k]), k
    for k in ("sa1_features", "sa2_features", "sa3_features", "sa4_features", "fp2_features"):
        # tf32 tensor-core MLP vs fp32 oracle; error grows mildly through the 4 chained stages (3 layers each)
        err = (got[k].cpu() - want[k]).abs()
        assert err.max() <= 1e-2 * max(1.0, want[k].abs().max().item()), (k, err.max().item())
        assert err.pow(2).mean().sqrt() <= 2e-3, (k, err.pow(2).mean().sqrt().item())
        assert torch.equal(got[k], plain[k]), k  # side-stream FPS chain changes nothing
    assert got["fp2_features"].shape == (2, 288, 1024)


@pytest.mark.gpu
@pytest.mark.parametrize("every", [512, 300, 2048])
def test_pipelined_sa1_equals_plain(every):
    """SA1 fed chunk by chunk from the in-flight sampler (progress milestones + stream-ordered waits) gives
    exactly what the plain path gives: same indices, same neighbour lists, same features."""
    from eda_b200.backbone_module import fps_chain
    from eda_b200.pointnet2 import fused
    from eda_b200.pointnet2.pointnet2_modules import PointnetSAModuleVotes

    torch.manual_seed(0)
    sa = PointnetSAModuleVotes(npoint=2048, radius=0.2, nsample=64, mlp=[3, 64, 64, 128], use_xyz=True,
                               normalize_xyz=True).cuda().eval()
    _randomise_bn(sa, 2)
    pc = synthetic.point_clouds(3, 20000, "surface").cuda()
    xyz, feats = pc[..., :3].contiguous(), pc[..., 3:].transpose(1, 2).contiguous()
    side = torch.cuda.Stream()
    with torch.no_grad():
        want_xyz, want_f, want_inds = sa(xyz, feats)
        for _ in range(2):  # twice: the progress counter is never reset, the second call starts from a non-zero base
            (handle, _ev), = fps_chain(xyz, [2048], side, pipeline_every=every)
            got_xyz, got_f, _, got_inds = fused.sa_forward_pipelined(sa, xyz, feats, handle)
            torch.cuda.synchronize()
            assert torch.equal(got_inds, want_inds)
            assert torch.equal(got_xyz, want_xyz)
            assert torch.equal(got_f, want_f)


@pytest.mark.gpu
def test_pipelined_sa1_many_scenes_in_waves():
    """B = 40 scenes x 8-CTA clusters do not fit the GPU at once, so the sampler's clusters run in waves: the scenes of
    the first wave publish ALL their milestones before later scenes start.  Each milestone has its own counter word,
    so a chunk is only released when every scene reached it (a single summed word released chunks early
    and the consumers read unwritten centre indices)."""
    from eda_b200.backbone_module import fps_chain
    from eda_b200.pointnet2 import fused
    from eda_b200.pointnet2.pointnet2_modules import PointnetSAModuleVotes

    torch.manual_seed(0)
    sa = PointnetSAModuleVotes(npoint=1024, radius=0.2, nsample=32, mlp=[3, 64, 64, 128], use_xyz=True,
                               normalize_xyz=True).cuda().eval()
    _randomise_bn(sa, 2)
    pc = synthetic.point_clouds(40, 50000, "surface").cuda()
    xyz, feats = pc[..., :3].contiguous(), pc[..., 3:].transpose(1, 2).contiguous()
    side = torch.cuda.Stream()
    with torch.no_grad():
        want_xyz, want_f, want_inds = sa(xyz, feats)
        for _ in range(2):
            (handle, _ev), = fps_chain(xyz, [1024], side, pipeline_every=256)
            got_xyz, got_f, _, got_inds = fused.sa_forward_pipelined(sa, xyz, feats, handle)
            torch.cuda.synchronize()
            assert torch.equal(got_inds, want_inds)
            assert torch.equal(got_xyz, want_xyz)
            assert torch.equal(got_f, want_f)


@pytest.mark.gpu
def test_pipelined_chunks_with_identity_flags_mixed():
    """Identity-verified scenes publish all milestones at once while the others are still sampling: with several
    chunks the consumer must still wait for the slow scenes (one counter word per milestone)."""
    from eda_b200.pointnet2 import _ext, fused
    from eda_b200.pointnet2.pointnet2_modules import PointnetSAModuleVotes

    torch.manual_seed(0)
    B, n, m = 6, 8192, 2048
    base = synthetic.point_clouds(B, 30000, "surface", channels=0).cuda().contiguous()
    inds1 = _ext.furthest_point_sampling(base, n)
    cur = torch.gather(base, 1, inds1.long().unsqueeze(-1).expand(-1, -1, 3)).contiguous()  # FPS-ordered sets
    perm = torch.randperm(n, generator=torch.Generator().manual_seed(0)).cuda()
    cur[1::2] = cur[1::2][:, perm]  # odd scenes are shuffled: they need the real, serial sampler
    flags = _ext.fps_identity_flags(cur, m)
    assert flags.tolist() == [0, 1, 0, 1, 0, 1]
    sa = PointnetSAModuleVotes(npoint=m, radius=0.3, nsample=16, mlp=[3, 64, 64, 128], use_xyz=True,
                               normalize_xyz=True).cuda().eval()
    _randomise_bn(sa, 3)
    feats = torch.rand(B, 3, n, generator=torch.Generator().manual_seed(1)).cuda()
    side = torch.cuda.Stream()
    with torch.no_grad():
        want_xyz, want_f, want_inds = sa(cur, feats)
        side.wait_stream(torch.cuda.current_stream())
        handle = fused.launch_pipelined_fps(cur, m, 256, side, not_identity=flags)
        got_xyz, got_f, _, got_inds = fused.sa_forward_pipelined(sa, cur, feats, handle)
        torch.cuda.synchronize()
    assert torch.equal(got_inds, want_inds)
    assert torch.equal(got_xyz, want_xyz)
    assert torch.equal(got_f, want_f)


@pytest.mark.gpu
def test_backbone_xyz_only_input_eval_and_train():
    """input_feature_dim = 0 (the class default; the reference without --use_color): SA1 has no input features.
    ADVICE r1: the eval / no_grad forward dereferenced features.device on None."""
    from oracle import modules_oracle as mo

    torch.manual_seed(0)
    m = Pointnet2Backbone(input_feature_dim=0, width=1).eval()
    _randomise_bn(m, 4)
    pc = synthetic.point_clouds(2, 6000, "surface", channels=0)
    with torch.no_grad():
        want = mo.backbone_forward(m.state_dict(), pc)
        got = m.cuda()(pc.cuda())
    assert torch.equal(got["sa1_inds"].cpu(), want["sa1_inds"])
    err = (got["fp2_features"].cpu() - want["fp2_features"]).abs()
    assert err.max() <= 1e-2 * max(1.0, want["fp2_features"].abs().max().item())
    m.train()
    out = m(pc.cuda())
    out["fp2_features"].sum().backward()
    assert m.sa1.mlp_module.layer0.conv.weight.grad is not None


@pytest.mark.gpu
@pytest.mark.parametrize("family", ["surface", "uniform", "dup", "lattice", "origin"])
def test_fps_identity_shortcut_is_bit_exact(family):
    """Stages 2-4 of the backbone sample from FPS-ordered sets.  The parallel identity check +
    flag-consulting sampler must return exactly what the full serial algorithm (and hence the reference) returns:
    the identity where every step has a strict maximiser, the full result where ties / duplicates decide."""
    from eda_b200.pointnet2 import _ext
    from oracle import pointnet2_oracle as orc

    B, N = 3, 20000
    xyz = synthetic.point_clouds(B, N, family, channels=0).cuda().contiguous()
    inds1 = _ext.furthest_point_sampling(xyz, 2048)
    sub = torch.gather(xyz, 1, inds1.long().unsqueeze(-1).expand(-1, -1, 3)).contiguous()
    for n, m in ((2048, 1024), (1024, 512), (512, 256)):
        cur = sub[:, :n].contiguous()
        full = _ext.furthest_point_sampling(cur, m)
        flags = _ext.fps_identity_flags(cur, m)
        short = _ext.furthest_point_sampling(cur, m, not_identity=flags)
        assert torch.equal(short, full), (family, n, m)
        assert torch.equal(full.cpu(), orc.furthest_point_sampling(cur.cpu(), m))  # and both equal the oracle
        ident = torch.arange(m, dtype=torch.int32, device="cuda")
        for b in range(B):
            if flags[b].item() == 0:
                assert torch.equal(full[b], ident)       # verified scenes really are the identity
        if family in ("surface", "uniform"):
            assert flags.sum().item() == 0                # continuous data: no ties, the shortcut is taken
        if family == "lattice":
            assert flags.sum().item() == B                # every step is a tie: always the full algorithm
    # a set that is NOT in FPS order is never reported as verified
    perm = torch.randperm(2048, generator=torch.Generator().manual_seed(0)).cuda()
    shuffled = sub[:, perm].contiguous()
    flags = _ext.fps_identity_flags(shuffled, 1024)
    assert flags.sum().item() == B
    assert torch.equal(_ext.furthest_point_sampling(shuffled, 1024, not_identity=flags),
                       _ext.furthest_point_sampling(shuffled, 1024))
    # the sampler obeys the flags (forced zeros -> identity), i.e. the flags are what decides
    forced = _ext.furthest_point_sampling(shuffled, 1024, not_identity=torch.zeros(B, dtype=torch.int32, device="cuda"))
    assert torch.equal(forced, torch.arange(1024, dtype=torch.int32, device="cuda").expand(B, -1))
