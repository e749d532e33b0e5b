"""Generate golden vectors from the REFERENCE's own compiled `_ext` (run on the B200 box):

    gpurun -- python tests/golden/make_golden.py        # writes gpurun_out/golden/*.npz
    cp gpurun_out/golden/*.npz tests/golden/            # commit them

The reference ships no golden vectors for this path (its only test is a gradcheck,
pointnet2/pointnet2_test.py:18-30), so the CPU oracle is pinned against outputs of the unmodified
reference sources compiled for sm_100a (oracle/build_ref.py -> oracle/_ref/pointnet2/_ext*.so),
on this repo's seeded synthetic inputs.  Inputs are stored with the outputs so the fixtures are
self-contained; BASELINE.json configs[0] (B=2, N=4096, r=0.2, nsample=32) is the main case.
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from eda_b200 import synthetic  # noqa: E402
from oracle import ref_loader  # noqa: E402

OUT = os.path.join(ROOT, "gpurun_out", "golden")


def main():
    ref = ref_loader.load_reference_ext()
    assert ref is not None, "oracle/_ref/pointnet2/_ext*.so missing: run oracle/build_ref.py where /root/reference exists"
    os.makedirs(OUT, exist_ok=True)
    dev = "cuda:0"
    cases = [("surface", 2, 4096, 512, 0.2, 32), ("uniform", 1, 4096, 512, 0.2, 32), ("dup", 1, 4096, 512, 0.2, 32),
             ("lattice", 1, 4096, 512, 0.5, 32), ("origin", 1, 4096, 512, 0.2, 32), ("lattice", 1, 300, 64, 0.75, 8)]
    for family, B, N, m, r, ns in cases:
        xyz = synthetic.point_clouds(B, N, family, seed=synthetic.SEED, channels=0)
        xd = xyz.to(dev)
        inds = ref.furthest_point_sampling(xd, m)
        new_xyz = torch.gather(xd, 1, inds.long()[..., None].expand(-1, -1, 3)).contiguous()
        idx = ref.ball_query(new_xyz, xd, r, ns)
        unknown = xd[:, : min(N, 600)].contiguous()
        d2, nn_idx = ref.three_nn(unknown, new_xyz)
        g = torch.Generator().manual_seed(3)
        feats = torch.randn(B, 4, m, generator=g).to(dev)
        w = torch.rand(B, unknown.size(1), 3, generator=g).to(dev)
        interp = ref.three_interpolate(feats, nn_idx, w)
        np.savez_compressed(
            os.path.join(OUT, f"ext_{family}_N{N}_m{m}.npz"), xyz=xyz.numpy(), radius=np.float32(r),
            nsample=np.int32(ns), fps_inds=inds.cpu().numpy(), ball_idx=idx.cpu().numpy(),
            nn_dist2=d2.cpu().numpy(), nn_idx=nn_idx.cpu().numpy(), feats=feats.cpu().numpy(), weight=w.cpu().numpy(),
            interp=interp.cpu().numpy())
        print("wrote", family, N, m, flush=True)
    # full-size SA1 case: only a checksum-sized record (inputs are regenerated from the seed)
    xyz = synthetic.point_clouds(1, 50000, "surface", seed=synthetic.SEED, channels=0)
    xd = xyz.to(dev)
    inds = ref.furthest_point_sampling(xd, 2048)
    new_xyz = torch.gather(xd, 1, inds.long()[..., None].expand(-1, -1, 3)).contiguous()
    idx = ref.ball_query(new_xyz, xd, 0.2, 64)
    np.savez_compressed(os.path.join(OUT, "ext_surface_N50000_m2048.npz"), xyz_sha=np.frombuffer(
        __import__("hashlib").sha256(xyz.numpy().tobytes()).digest(), dtype=np.uint8), fps_inds=inds.cpu().numpy(),
        ball_idx_first8=idx[:, :, :8].cpu().numpy(), ball_idx_sum=idx.long().sum(-1).cpu().numpy())
    print("wrote full-size record")


if __name__ == "__main__":
    main()
