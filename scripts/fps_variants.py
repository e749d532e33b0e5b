"""FPS decomposition variants at the SA1 shape (B=8, N=50000 -> 2048) and the sweep sizes: ms per launch for forced
(cluster size, threads per CTA) plans, next to the default plan.  Results are bit-identical by construction
(tests/test_gpu_ops.py sweeps the same environment variables)."""
import ctypes, json, os, sys
import torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from benchmarks import kernels as kn
from eda_b200 import synthetic
from eda_b200.pointnet2 import _ext

dev = torch.device("cuda", 0)
out = {}
for B, N in ((8, 50000), (8, 20000), (8, 100000), (16, 50000)):
    xyz = synthetic.point_clouds(B, N, "surface", channels=0).to(dev).contiguous()
    ref = None
    for cl, th in ((0, 0), (4, 256), (4, 512), (8, 256), (8, 512), (8, 1024), (16, 256), (16, 512)):
        if cl:
            os.environ["EDA_FPS_CLUSTER"], os.environ["EDA_FPS_THREADS"] = str(cl), str(th)
        else:
            os.environ.pop("EDA_FPS_CLUSTER", None); os.environ.pop("EDA_FPS_THREADS", None)
        try:
            inds = _ext.furthest_point_sampling(xyz, 2048)
            if ref is None:
                ref = inds.clone()
            same = bool(torch.equal(inds, ref))
            ms = kn.time_ms(lambda: _ext.furthest_point_sampling(xyz, 2048), 2, 5)
            out[f"B{B}_N{N}_cl{cl}_t{th}"] = {"ms": ms, "same": same}
        except Exception as e:  # noqa: BLE001
            out[f"B{B}_N{N}_cl{cl}_t{th}"] = {"error": repr(e)[:100]}
print(json.dumps(out, indent=1))
