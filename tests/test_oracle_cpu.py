"""CPU tests of the oracle (oracle/pointnet2_oracle.c) against independent statements of what the
reference kernels compute.  No GPU needed.  The oracle's pin against the reference's own compiled
`_ext` is tests/test_golden.py (fixtures produced on the B200 box by tests/golden/make_golden.py).
"""
import numpy as np
import pytest
import torch

from eda_b200 import synthetic


def _bitrev(v, bits):
    r = 0
    for _ in range(bits):
        r = (r << 1) | (v & 1)
        v >>= 1
    return r


def fps_order_model(xyz, m, bs):
    """FPS restated through the total order of SURVEY.md A.2 (no thread layout, no tree):
    winner = max over points of (d2, -bitrev(k mod BS), -(k div BS)).  Distances are computed in
    float64 from float32 inputs, so only use on data where every d2 is exactly representable."""
    n = xyz.shape[0]
    p = xyz.astype(np.float64)
    lb = int(np.log2(bs))
    k = np.arange(n)
    code = np.array([_bitrev(int(i) % bs, lb) for i in k], dtype=np.int64) * (1 << 32) + k // bs
    skip = (p * p).sum(1) <= 1e-3
    temp = np.full(n, 1e10)
    out = np.zeros(m, dtype=np.int32)
    old = 0
    for j in range(1, m):
        d = ((p - p[old]) ** 2).sum(1)
        temp = np.where(skip, temp, np.minimum(temp, d))
        cand = np.where(skip, -1.0, temp)
        best = cand.max()
        if best < 0:  # every point skipped: besti stays 0 (sampling_gpu.cu:95)
            old = 0
        else:
            tie = np.flatnonzero(cand == best)
            old = int(tie[np.argmin(code[tie])])
        out[j] = old
    return out


@pytest.mark.parametrize("n,m", [(4096, 256), (1000, 128), (512, 64), (300, 50), (37, 20)])
def test_fps_tiebreak_matches_order_model_on_lattice(oracle, n, m):
    # lattice coordinates are multiples of 0.25: all squared distances are exact in fp32 and fp64,
    # and almost every arg-max is a tie -> this pins the tree's tie-break, not the arithmetic
    xyz = synthetic.scene_xyz(n, "lattice", seed=5)
    got = oracle.furthest_point_sampling(xyz[None], m)[0].numpy()
    bs = oracle.opt_n_threads(n)
    want = fps_order_model(xyz.numpy(), m, bs)
    np.testing.assert_array_equal(got, want)


def test_opt_n_threads(oracle):
    # include/cuda_utils.h:18-22
    assert oracle.opt_n_threads(50000) == 512
    assert oracle.opt_n_threads(2048) == 512
    assert oracle.opt_n_threads(512) == 512
    assert oracle.opt_n_threads(511) == 256
    assert oracle.opt_n_threads(256) == 256
    assert oracle.opt_n_threads(37) == 32
    assert oracle.opt_n_threads(1) == 1


@pytest.mark.parametrize("family", ["uniform", "surface", "origin"])
def test_fps_is_furthest_point_sampling(oracle, family):
    n, m = 4096, 128
    xyz = synthetic.scene_xyz(n, family, seed=11)
    inds = oracle.furthest_point_sampling(xyz[None], m)[0].numpy()
    p = xyz.numpy().astype(np.float64)
    skip = (p * p).sum(1) <= 1e-3
    assert inds[0] == 0
    temp = np.full(n, np.inf)
    for j in range(1, m):
        temp = np.minimum(temp, ((p - p[inds[j - 1]]) ** 2).sum(1))
        cand = np.where(skip, -1.0, temp)
        assert not skip[inds[j]]
        assert cand[inds[j]] >= cand.max() * (1 - 1e-5)  # the pick is a furthest point up to fp32 rounding
    assert len(set(inds.tolist())) == m


def test_fps_all_points_skipped_returns_zeros(oracle):
    xyz = (torch.rand(2, 64, 3) - 0.5) * 0.01  # |p|^2 << 1e-3 everywhere
    inds = oracle.furthest_point_sampling(xyz.contiguous(), 16)
    assert torch.equal(inds, torch.zeros(2, 16, dtype=torch.int32))


def ball_query_bruteforce(new_xyz, xyz, r, ns):
    q = new_xyz.astype(np.float64)
    p = xyz.astype(np.float64)
    r2 = float(np.float32(r) * np.float32(r))
    out = np.zeros((q.shape[0], ns), dtype=np.int32)
    for j in range(q.shape[0]):
        hits = np.flatnonzero(((p - q[j]) ** 2).sum(1) < r2)[:ns]
        if len(hits):
            out[j, :] = hits[0]
            out[j, : len(hits)] = hits
    return out


@pytest.mark.parametrize("r,ns", [(0.5, 8), (0.75, 32), (0.2, 4)])
def test_ball_query_matches_definition_on_lattice(oracle, r, ns):
    # exact arithmetic again (0.5^2, 0.75^2 and all d2 exactly representable); r = 0.2 gives
    # self-only balls: every slot is the centre's own (first) index
    xyz = synthetic.scene_xyz(2048, "lattice", seed=3)
    new_xyz = xyz[:256].clone()
    got = oracle.ball_query(new_xyz[None], xyz[None], r, ns)[0].numpy()
    want = ball_query_bruteforce(new_xyz.numpy(), xyz.numpy(), r, ns)
    np.testing.assert_array_equal(got, want)


def test_ball_query_config1_properties(oracle):
    # BASELINE.json configs[0]: B=2 N=4096 r=0.2 nsample=32
    pc = synthetic.point_clouds(2, 4096, "surface", channels=0)
    inds = oracle.furthest_point_sampling(pc, 512)
    new_xyz = torch.stack([pc[b, inds[b].long()] for b in range(2)])
    idx = oracle.ball_query(new_xyz, pc, 0.2, 32)
    assert idx.shape == (2, 512, 32) and idx.dtype == torch.int32
    for b in range(2):
        p = pc[b].numpy().astype(np.float64)
        q = new_xyz[b].numpy().astype(np.float64)
        I = idx[b].numpy()
        d2 = ((p[I] - q[:, None, :]) ** 2).sum(-1)
        assert (d2 < 0.2 ** 2 * (1 + 1e-5)).all()            # every listed neighbour is inside the ball
        first = I[:, :1]
        uniq = np.where(I == first, -1, I)                   # padding repeats the first hit
        for j in range(I.shape[0]):
            row = uniq[j][uniq[j] >= 0]
            assert (np.diff(row) > 0).all()                  # ascending index order
        assert (I[:, 0] <= inds[b].numpy()).all()            # the centre itself is in its ball


def test_ball_query_empty_ball_is_zero(oracle):
    xyz = torch.rand(1, 128, 3) + 10.0
    new_xyz = torch.zeros(1, 4, 3)
    idx = oracle.ball_query(new_xyz, xyz.contiguous(), 0.1, 8)
    assert torch.equal(idx, torch.zeros(1, 4, 8, dtype=torch.int32))


def test_gather_group_match_torch_indexing(oracle):
    g = torch.Generator().manual_seed(0)
    pts = torch.randn(2, 5, 100, generator=g)
    idx = torch.randint(0, 100, (2, 7), generator=g, dtype=torch.int32)
    out = oracle.gather_points(pts, idx)
    want = torch.gather(pts, 2, idx.long()[:, None, :].expand(-1, 5, -1))
    assert torch.equal(out, want)
    gidx = torch.randint(0, 100, (2, 7, 4), generator=g, dtype=torch.int32)
    out = oracle.group_points(pts, gidx)
    want = torch.gather(pts, 2, gidx.long().reshape(2, 1, 28).expand(-1, 5, -1)).reshape(2, 5, 7, 4)
    assert torch.equal(out, want)
    # gradients = transposes of the gathers
    go = torch.randn(2, 5, 7, 4, generator=g)
    gp = oracle.group_points_grad(go, gidx, 100)
    want = torch.zeros(2, 5, 100).scatter_add_(2, gidx.long().reshape(2, 1, 28).expand(-1, 5, -1), go.reshape(2, 5, 28))
    torch.testing.assert_close(gp, want, rtol=1e-6, atol=1e-6)
    go = torch.randn(2, 5, 7, generator=g)
    gp = oracle.gather_points_grad(go, idx, 100)
    want = torch.zeros(2, 5, 100).scatter_add_(2, idx.long()[:, None, :].expand(-1, 5, -1), go)
    torch.testing.assert_close(gp, want, rtol=1e-6, atol=1e-6)


def test_three_nn_and_interpolate(oracle):
    g = torch.Generator().manual_seed(1)
    unknown = torch.rand(2, 50, 3, generator=g)
    known = torch.rand(2, 20, 3, generator=g)
    dist2, idx = oracle.three_nn(unknown, known)
    d = ((unknown[:, :, None, :].double() - known[:, None, :, :].double()) ** 2).sum(-1)
    want_d, want_i = torch.topk(d, 3, dim=2, largest=False)
    assert torch.equal(idx.long(), want_i)
    torch.testing.assert_close(dist2.double(), want_d, rtol=1e-5, atol=1e-7)
    # fewer than 3 known points: unfilled slots keep (float)1e40 = inf and index 0 (interpolate_gpu.cu:30-34)
    dist2, idx = oracle.three_nn(unknown, known[:, :2].contiguous())
    assert torch.isinf(dist2[..., 2]).all() and (idx[..., 2] == 0).all()
    # interpolation and its transpose
    feats = torch.randn(2, 6, 20, generator=g)
    w = torch.rand(2, 50, 3, generator=g)
    _, idx = oracle.three_nn(unknown, known)
    out = oracle.three_interpolate(feats, idx, w)
    gathered = torch.gather(feats[:, :, None, :].expand(-1, -1, 50, -1), 3, idx.long()[:, None].expand(-1, 6, -1, -1))
    want = (gathered * w[:, None]).sum(-1)
    torch.testing.assert_close(out, want, rtol=1e-5, atol=1e-6)
    go = torch.randn(2, 6, 50, generator=g)
    gp = oracle.three_interpolate_grad(go, idx, w, 20)
    want = torch.zeros(2, 6, 20)
    contrib = go[:, :, :, None] * w[:, None]
    want.scatter_add_(2, idx.long().reshape(2, 1, 150).expand(-1, 6, -1), contrib.reshape(2, 6, 150))
    torch.testing.assert_close(gp, want, rtol=1e-5, atol=1e-6)


@pytest.mark.parametrize("family", ["surface", "uniform", "dup", "lattice", "origin"])
def test_fps_of_an_fps_ordered_set_is_the_identity_when_no_step_is_a_tie(oracle, family):
    """SURVEY.md A.4 / the product's FPS shortcut for the backbone's stages 2-4, on the oracle alone: whenever the
    strict-maximiser criterion holds, the reference algorithm (literal emulation, tie-break included) returns
    0..m-1; continuous data always satisfies it, tie-heavy data never does, and a shuffled set is never verified."""
    B, N = 2, 6000
    xyz = synthetic.point_clouds(B, N, family, channels=0).contiguous()
    inds1 = oracle.furthest_point_sampling(xyz, 512)
    sub = torch.gather(xyz, 1, inds1.long().unsqueeze(-1).expand(-1, -1, 3)).contiguous()
    for n, m in ((512, 256), (256, 128)):
        cur = sub[:, :n].contiguous()
        ver = oracle.fps_identity_verified(cur, m)
        full = oracle.furthest_point_sampling(cur, m)
        for b in range(B):
            if ver[b].item() == 1:
                assert torch.equal(full[b], torch.arange(m, dtype=torch.int32)), (family, n, m, b)
        if family in ("surface", "uniform"):
            assert ver.sum().item() == B
        if family == "lattice":
            assert ver.sum().item() == 0
    perm = torch.randperm(512, generator=torch.Generator().manual_seed(0))
    assert oracle.fps_identity_verified(sub[:, perm].contiguous(), 256).sum().item() == 0
