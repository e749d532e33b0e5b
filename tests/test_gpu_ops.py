"""GPU parity tests proper: the CUDA ops (through the C ABI) against the CPU oracle on the same
seeded inputs, and against the reference's own compiled `_ext` when oracle/_ref travelled.
Index outputs and pure copies must be BIT-EXACT; scatter-add gradients (atomics in the reference
too) are compared with rtol 1e-5."""
import numpy as np
import pytest
import torch

from eda_b200 import synthetic

pytestmark = pytest.mark.gpu

DEV = "cuda:0"


def _cuda(t):
    return t.to(DEV).contiguous()


# ---- FPS -------------------------------------------------------------------------------------
FPS_CASES = [
    # (family, B, N, m) — config 1 and the shapes of SA1..SA4, odd sizes, N < 512 (BS < 512), ties
    ("surface", 2, 4096, 512), ("uniform", 2, 4096, 512), ("dup", 2, 4096, 512), ("lattice", 2, 4096, 512),
    ("origin", 2, 4096, 512), ("surface", 2, 2048, 1024), ("surface", 2, 1024, 512), ("surface", 2, 512, 256),
    ("lattice", 2, 2048, 1024), ("lattice", 1, 1000, 333), ("dup", 3, 777, 100), ("lattice", 2, 300, 64),
    ("uniform", 2, 37, 20), ("surface", 1, 20000, 512), ("lattice", 1, 20000, 256), ("dup", 1, 9999, 300),
]


@pytest.mark.parametrize("family,B,N,m", FPS_CASES)
def test_fps_bit_exact_vs_oracle(ext, oracle, family, B, N, m):
    xyz = synthetic.point_clouds(B, N, family, seed=100 + N, channels=0)
    want = oracle.furthest_point_sampling(xyz, m)
    got = ext.furthest_point_sampling(_cuda(xyz), m).cpu()
    assert got.dtype == torch.int32 and got.shape == (B, m)
    assert torch.equal(got, want), f"first mismatch at {torch.nonzero(got != want)[0].tolist()}"


@pytest.mark.parametrize("threads", [256, 512])
@pytest.mark.parametrize("cl", [1, 2, 4, 8, 16])
def test_fps_every_cluster_size_bit_exact(ext, oracle, cl, threads, monkeypatch):
    # EDA_FPS_CLUSTER / EDA_FPS_THREADS force the decomposition: the result may not depend on it
    monkeypatch.setenv("EDA_FPS_CLUSTER", str(cl))
    monkeypatch.setenv("EDA_FPS_THREADS", str(threads))
    for family, N, m in [("lattice", 8192, 300), ("dup", 5000, 300), ("lattice", 1000, 200), ("dup", 300, 100)]:
        xyz = synthetic.point_clouds(2, N, family, seed=cl, channels=0)
        want = oracle.furthest_point_sampling(xyz, m)
        got = ext.furthest_point_sampling(_cuda(xyz), m).cpu()
        assert torch.equal(got, want)


def test_fps_full_size_sa1(ext, oracle):
    # BASELINE config: N = 50 000 -> 2048 (one scene through the oracle keeps the CPU side to seconds)
    xyz = synthetic.point_clouds(2, 50000, "surface", channels=0)
    got = ext.furthest_point_sampling(_cuda(xyz), 2048).cpu()
    want = oracle.furthest_point_sampling(xyz[:1].contiguous(), 2048)
    assert torch.equal(got[:1], want)
    # size-independent properties on the rest: indices unique, start at 0, coverage radius shrinks
    for b in range(2):
        assert got[b, 0] == 0 and len(set(got[b].tolist())) == 2048


def test_fps_large_n_and_global_fallback(ext, oracle):
    xyz = synthetic.point_clouds(1, 200000, "surface", seed=9, channels=0)
    got = ext.furthest_point_sampling(_cuda(xyz), 256).cpu()
    want = oracle.furthest_point_sampling(xyz, 256)
    assert torch.equal(got, want)
    xyz = synthetic.point_clouds(1, 300000, "uniform", seed=10, channels=0)  # beyond the register variant
    got = ext.furthest_point_sampling(_cuda(xyz), 64).cpu()
    want = oracle.furthest_point_sampling(xyz, 64)
    assert torch.equal(got, want)


def test_fps_all_skipped_and_m_equals_1(ext):
    xyz = ((torch.rand(2, 64, 3) - 0.5) * 0.01).contiguous()
    got = ext.furthest_point_sampling(_cuda(xyz), 16).cpu()
    assert torch.equal(got, torch.zeros(2, 16, dtype=torch.int32))
    got = ext.furthest_point_sampling(_cuda(torch.rand(3, 100, 3)), 1).cpu()
    assert torch.equal(got, torch.zeros(3, 1, dtype=torch.int32))


def test_fps_of_fps_ordered_set_is_identity_without_ties(ext):
    # models/backbone_module.py:142 relies on this
    xyz = _cuda(synthetic.point_clouds(2, 20000, "uniform", seed=3, channels=0))
    inds = ext.furthest_point_sampling(xyz, 2048)
    sub = torch.gather(xyz, 1, inds.long()[..., None].expand(-1, -1, 3)).contiguous()
    again = ext.furthest_point_sampling(sub, 1024).cpu()
    assert torch.equal(again, torch.arange(1024, dtype=torch.int32).expand(2, -1))


# ---- ball query ------------------------------------------------------------------------------
BQ_CASES = [
    # (family, B, N, M, r, ns)
    ("surface", 2, 4096, 512, 0.2, 32), ("uniform", 2, 4096, 512, 0.2, 32), ("dup", 2, 4096, 512, 0.2, 32),
    ("lattice", 2, 4096, 512, 0.5, 32), ("surface", 2, 2048, 1024, 0.4, 32), ("surface", 2, 1024, 512, 0.8, 16),
    ("surface", 2, 512, 256, 1.2, 16), ("surface", 1, 4099, 100, 0.3, 64), ("lattice", 2, 1001, 77, 0.75, 7),
    ("uniform", 1, 50, 50, 10.0, 64), ("surface", 1, 20000, 2048, 0.2, 64), ("surface", 1, 6001, 33, 0.2, 1),
]


@pytest.mark.parametrize("family,B,N,M,r,ns", BQ_CASES)
def test_ball_query_bit_exact_vs_oracle(ext, oracle, family, B, N, M, r, ns):
    xyz = synthetic.point_clouds(B, N, family, seed=7 + N, channels=0)
    inds = oracle.furthest_point_sampling(xyz, M)
    new_xyz = torch.gather(xyz, 1, inds.long()[..., None].expand(-1, -1, 3)).contiguous()
    want = oracle.ball_query(new_xyz, xyz, r, ns)
    got = ext.ball_query(_cuda(new_xyz), _cuda(xyz), r, ns).cpu()
    assert got.dtype == torch.int32 and got.shape == (B, M, ns)
    assert torch.equal(got, want)


def test_ball_query_empty_balls_and_unaligned_views(ext, oracle):
    xyz = (torch.rand(2, 1000, 3) + 10.0).contiguous()
    new_xyz = torch.zeros(2, 9, 3)
    got = ext.ball_query(_cuda(new_xyz), _cuda(xyz), 0.1, 8).cpu()
    assert torch.equal(got, torch.zeros(2, 9, 8, dtype=torch.int32))
    # a scene that starts at a 4-byte-aligned (not 16-byte-aligned) address: the TMA path must not be taken blindly
    buf = torch.rand(1 + 2 * 1000 * 3).to(DEV)
    xyz_d = buf[1:].view(2, 1000, 3)
    assert xyz_d.is_contiguous() and xyz_d.data_ptr() % 16 != 0
    q = xyz_d[:, :64].contiguous()
    got = ext.ball_query(q, xyz_d, 0.3, 16).cpu()
    want = oracle.ball_query(q.cpu(), xyz_d.cpu().contiguous(), 0.3, 16)
    assert torch.equal(got, want)


def test_ball_query_full_size_sa1(ext, oracle):
    xyz = synthetic.point_clouds(2, 50000, "surface", channels=0)
    xyz_d = _cuda(xyz)
    inds = ext.furthest_point_sampling(xyz_d, 2048)
    new_xyz = torch.gather(xyz_d, 1, inds.long()[..., None].expand(-1, -1, 3)).contiguous()
    got = ext.ball_query(new_xyz, xyz_d, 0.2, 64).cpu()
    want = oracle.ball_query(new_xyz.cpu(), xyz, 0.2, 64)
    assert torch.equal(got, want)


# ---- gathers / grouping ----------------------------------------------------------------------
@pytest.mark.parametrize("B,C,N,M,S", [(2, 3, 4096, 512, 32), (2, 131, 2048, 64, 32), (1, 259, 512, 256, 16),
                                       (3, 1, 17, 5, 3), (2, 288, 1024, 256, 1)])
def test_gather_group_forward_exact_and_grads(ext, oracle, B, C, N, M, S):
    g = torch.Generator().manual_seed(B * 1000 + C)
    pts = torch.randn(B, C, N, generator=g)
    idx = torch.randint(0, N, (B, M), generator=g, dtype=torch.int32)
    gidx = torch.randint(0, N, (B, M, S), generator=g, dtype=torch.int32)
    gidx[:, :, -1] = gidx[:, :, 0]  # repeated indices, as ball-query padding produces
    assert torch.equal(ext.gather_points(_cuda(pts), _cuda(idx)).cpu(), oracle.gather_points(pts, idx))
    assert torch.equal(ext.group_points(_cuda(pts), _cuda(gidx)).cpu(), oracle.group_points(pts, gidx))
    go = torch.randn(B, C, M, generator=g)
    torch.testing.assert_close(ext.gather_points_grad(_cuda(go), _cuda(idx), N).cpu(),
                               oracle.gather_points_grad(go, idx, N), rtol=1e-5, atol=1e-5)
    go = torch.randn(B, C, M, S, generator=g)
    torch.testing.assert_close(ext.group_points_grad(_cuda(go), _cuda(gidx), N).cpu(),
                               oracle.group_points_grad(go, gidx, N), rtol=1e-5, atol=1e-5)


# ---- 3-NN / interpolation --------------------------------------------------------------------
@pytest.mark.parametrize("family,B,n,m,C", [("surface", 2, 512, 256, 256), ("surface", 2, 1024, 512, 256),
                                            ("lattice", 2, 700, 300, 5), ("uniform", 1, 33, 2, 4),
                                            ("dup", 2, 3000, 1500, 3)])
def test_three_nn_and_interpolate(ext, oracle, family, B, n, m, C):
    unknown = synthetic.point_clouds(B, n, family, seed=21, channels=0)
    known = unknown[:, :m].contiguous() if family != "uniform" else synthetic.point_clouds(B, m, family, seed=22, channels=0)
    want_d, want_i = oracle.three_nn(unknown, known)
    got_d, got_i = ext.three_nn(_cuda(unknown), _cuda(known))
    assert torch.equal(got_i.cpu(), want_i)
    assert torch.equal(got_d.cpu(), want_d)  # same fp32 expression -> same bits (inf for unfilled slots)
    g = torch.Generator().manual_seed(5)
    feats = torch.randn(B, C, m, generator=g)
    w = torch.rand(B, n, 3, generator=g)
    assert torch.equal(ext.three_interpolate(_cuda(feats), _cuda(want_i), _cuda(w)).cpu(),
                       oracle.three_interpolate(feats, want_i, w))
    go = torch.randn(B, C, n, generator=g)
    torch.testing.assert_close(ext.three_interpolate_grad(_cuda(go), _cuda(want_i), _cuda(w), m).cpu(),
                               oracle.three_interpolate_grad(go, want_i, w, m), rtol=1e-5, atol=1e-5)


# ---- the reference's own compiled _ext, when it travelled --------------------------------------
def test_reference_ext_agrees_on_all_index_ops(ext, ref_ext, oracle):
    """Three-way: reference `_ext` (unmodified sources, sm_100a) == CUDA product == CPU oracle."""
    for family in synthetic.FAMILIES:
        xyz = synthetic.point_clouds(2, 4096, family, seed=42, channels=0)
        xyz_d = _cuda(xyz)
        r_inds = ref_ext.furthest_point_sampling(xyz_d, 512)
        assert torch.equal(ext.furthest_point_sampling(xyz_d, 512), r_inds), family
        assert torch.equal(oracle.furthest_point_sampling(xyz, 512), r_inds.cpu()), family
        new_xyz = torch.gather(xyz_d, 1, r_inds.long()[..., None].expand(-1, -1, 3)).contiguous()
        r = 0.5 if family == "lattice" else 0.2
        r_idx = ref_ext.ball_query(new_xyz, xyz_d, r, 32)
        assert torch.equal(ext.ball_query(new_xyz, xyz_d, r, 32), r_idx), family
        assert torch.equal(oracle.ball_query(new_xyz.cpu(), xyz, r, 32), r_idx.cpu()), family
        rd, ri = ref_ext.three_nn(xyz_d[:, :700].contiguous(), new_xyz)
        gd, gi = ext.three_nn(xyz_d[:, :700].contiguous(), new_xyz)
        assert torch.equal(gi, ri) and torch.equal(gd, rd), family
        feats = torch.randn(2, 16, 512, device=DEV)
        w = torch.rand(2, 700, 3, device=DEV)
        assert torch.equal(ext.three_interpolate(feats, ri, w), ref_ext.three_interpolate(feats, ri, w))
        pts = torch.randn(2, 9, 4096, device=DEV)
        assert torch.equal(ext.group_points(pts, r_idx), ref_ext.group_points(pts, r_idx))
        assert torch.equal(ext.gather_points(pts, r_inds), ref_ext.gather_points(pts, r_inds))
        go = torch.randn(2, 9, 512, 32, device=DEV)
        torch.testing.assert_close(ext.group_points_grad(go, r_idx, 4096), ref_ext.group_points_grad(go, r_idx, 4096),
                                   rtol=1e-5, atol=1e-5)


def test_reference_ext_agrees_at_full_size(ext, ref_ext):
    xyz_d = _cuda(synthetic.point_clouds(2, 50000, "dup", seed=77, channels=0))
    r_inds = ref_ext.furthest_point_sampling(xyz_d, 2048)
    assert torch.equal(ext.furthest_point_sampling(xyz_d, 2048), r_inds)
    new_xyz = torch.gather(xyz_d, 1, r_inds.long()[..., None].expand(-1, -1, 3)).contiguous()
    assert torch.equal(ext.ball_query(new_xyz, xyz_d, 0.2, 64), ref_ext.ball_query(new_xyz, xyz_d, 0.2, 64))
