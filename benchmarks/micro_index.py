"""Micro-benchmark of the index kernels (FPS, ball query) at the BASELINE shapes, next to the
reference's own compiled `_ext` on the same GPU.  Development tool; bench.py is the contract.

    gpurun -- python benchmarks/micro_index.py > gpurun_out/micro_index.json
"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from eda_b200 import synthetic  # noqa: E402
from eda_b200.pointnet2 import _ext  # noqa: E402


def timeit(fn, warm=3, iters=10):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); b.synchronize()
        ts.append(a.elapsed_time(b))
    ts.sort()
    return {"min_ms": ts[0], "med_ms": ts[len(ts) // 2]}


def main():
    dev = "cuda:0"
    B = int(os.environ.get("B", 8))
    res = {"B": B, "gpu": torch.cuda.get_device_name(0)}
    ref = None
    try:
        from oracle import ref_loader
        ref = ref_loader.load_reference_ext()
    except Exception as e:  # noqa: BLE001
        res["ref_error"] = repr(e)
    stages = [(50000, 2048, 0.2, 64), (2048, 1024, 0.4, 32), (1024, 512, 0.8, 16), (512, 256, 1.2, 16)]
    xyz = synthetic.point_clouds(B, 50000, "surface", channels=0).to(dev)
    for N, m, r, ns in stages:
        key = f"N{N}_m{m}"
        res[f"fps_{key}"] = timeit(lambda: _ext.furthest_point_sampling(xyz, m))
        combos = [(8, 256), (8, 512), (16, 256)] if N == 50000 else [(1, 256), (1, 512), (1, 1024), (2, 256)]
        for cl, thr in combos:
            os.environ["EDA_FPS_CLUSTER"] = str(cl)
            os.environ["EDA_FPS_THREADS"] = str(thr)
            try:
                res[f"fps_{key}_cl{cl}_t{thr}"] = timeit(lambda: _ext.furthest_point_sampling(xyz, m))
            except RuntimeError as e:
                res[f"fps_{key}_cl{cl}_t{thr}"] = repr(e)
        os.environ.pop("EDA_FPS_CLUSTER")
        os.environ.pop("EDA_FPS_THREADS")
        inds = _ext.furthest_point_sampling(xyz, m)
        new_xyz = torch.gather(xyz, 1, inds.long()[..., None].expand(-1, -1, 3)).contiguous()
        res[f"bq_{key}_ns{ns}"] = timeit(lambda: _ext.ball_query(new_xyz, xyz, r, ns))
        if ref is not None:
            res[f"ref_fps_{key}"] = timeit(lambda: ref.furthest_point_sampling(xyz, m), warm=1, iters=3)
            res[f"ref_bq_{key}_ns{ns}"] = timeit(lambda: ref.ball_query(new_xyz, xyz, r, ns), warm=1, iters=3)
            idx = ref.ball_query(new_xyz, xyz, r, ns)
            feats = torch.randn(B, 128, N, device=dev)
            res[f"ref_group_{key}_C128"] = timeit(lambda: ref.group_points(feats, idx), warm=1, iters=3)
            res[f"group_{key}_C128"] = timeit(lambda: _ext.group_points(feats, idx))
        xyz = new_xyz
    # stress sweep (BASELINE configs[4]), one scene batch of B
    for N in (20000, 100000, 200000):
        x = synthetic.point_clouds(B, N, "surface", seed=5, channels=0).to(dev)
        res[f"fps_N{N}_m2048"] = timeit(lambda: _ext.furthest_point_sampling(x, 2048), warm=1, iters=3)
        inds = _ext.furthest_point_sampling(x, 2048)
        q = torch.gather(x, 1, inds.long()[..., None].expand(-1, -1, 3)).contiguous()
        for ns in (16, 32, 64):
            res[f"bq_N{N}_m2048_ns{ns}"] = timeit(lambda: _ext.ball_query(q, x, 0.2, ns), warm=1, iters=3)
    print(json.dumps(res, indent=1))


if __name__ == "__main__":
    main()
