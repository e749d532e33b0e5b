"""Pins the CPU oracle against outputs of the REFERENCE's own compiled `_ext` (unmodified sources,
sm_100a, run on a B200 by tests/golden/make_golden.py).  Runs on CPU.  The GPU-marked twin checks the
CUDA product against the same fixtures."""
import glob
import hashlib
import os

import numpy as np
import pytest
import torch

from eda_b200 import synthetic

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
SMALL = sorted(p for p in glob.glob(os.path.join(GOLDEN, "ext_*.npz")) if "N50000" not in p)


def _t(a):
    return torch.from_numpy(np.ascontiguousarray(a))


def test_fixtures_present():
    assert len(SMALL) >= 6 and os.path.exists(os.path.join(GOLDEN, "ext_surface_N50000_m2048.npz"))


def _check_small(ops, path, to_dev=lambda t: t, to_cpu=lambda t: t):
    z = np.load(path)
    xyz = _t(z["xyz"])
    m = z["fps_inds"].shape[1]
    inds = to_cpu(ops.furthest_point_sampling(to_dev(xyz), m))
    assert torch.equal(inds, _t(z["fps_inds"])), "FPS indices differ from the reference _ext"
    new_xyz = torch.gather(xyz, 1, inds.long()[..., None].expand(-1, -1, 3)).contiguous()
    idx = to_cpu(ops.ball_query(to_dev(new_xyz), to_dev(xyz), float(z["radius"]), int(z["nsample"])))
    assert torch.equal(idx, _t(z["ball_idx"])), "ball-query lists differ from the reference _ext"
    unknown = xyz[:, : z["nn_idx"].shape[1]].contiguous()
    d2, nn_idx = ops.three_nn(to_dev(unknown), to_dev(new_xyz))
    assert torch.equal(to_cpu(nn_idx), _t(z["nn_idx"]))
    assert torch.equal(to_cpu(d2), _t(z["nn_dist2"]))
    out = ops.three_interpolate(to_dev(_t(z["feats"])), to_dev(_t(z["nn_idx"])), to_dev(_t(z["weight"])))
    assert torch.equal(to_cpu(out), _t(z["interp"]))


@pytest.mark.parametrize("path", SMALL, ids=[os.path.basename(p) for p in SMALL])
def test_oracle_matches_reference_ext_golden(oracle, path):
    _check_small(oracle, path)


def test_synthetic_generator_is_stable():
    # the full-size fixture stores only a hash of its input: the generator must reproduce it
    z = np.load(os.path.join(GOLDEN, "ext_surface_N50000_m2048.npz"))
    xyz = synthetic.point_clouds(1, 50000, "surface", seed=synthetic.SEED, channels=0)
    assert hashlib.sha256(xyz.numpy().tobytes()).digest() == z["xyz_sha"].tobytes()


def _check_full(ops, to_dev=lambda t: t, to_cpu=lambda t: t):
    z = np.load(os.path.join(GOLDEN, "ext_surface_N50000_m2048.npz"))
    xyz = synthetic.point_clouds(1, 50000, "surface", seed=synthetic.SEED, channels=0)
    if hashlib.sha256(xyz.numpy().tobytes()).digest() != z["xyz_sha"].tobytes():
        pytest.fail("synthetic generator drifted from the one that made the fixture")
    inds = to_cpu(ops.furthest_point_sampling(to_dev(xyz), 2048))
    assert torch.equal(inds, _t(z["fps_inds"]))
    new_xyz = torch.gather(xyz, 1, inds.long()[..., None].expand(-1, -1, 3)).contiguous()
    idx = to_cpu(ops.ball_query(to_dev(new_xyz), to_dev(xyz), 0.2, 64))
    assert torch.equal(idx[:, :, :8], _t(z["ball_idx_first8"]))
    assert torch.equal(idx.long().sum(-1), _t(z["ball_idx_sum"]))


def test_oracle_matches_reference_ext_full_size(oracle):
    _check_full(oracle)


@pytest.mark.gpu
@pytest.mark.parametrize("path", SMALL, ids=[os.path.basename(p) for p in SMALL])
def test_cuda_matches_reference_ext_golden(ext, path):
    _check_small(ext, path, to_dev=lambda t: t.cuda(), to_cpu=lambda t: t.cpu())


@pytest.mark.gpu
def test_cuda_matches_reference_ext_full_size(ext):
    _check_full(ext, to_dev=lambda t: t.cuda(), to_cpu=lambda t: t.cpu())
