"""Weight gradient: device time of eda_wgrad on the step's shapes (CUDA-graph replay).  Run twice to compare kernels:
EDA_WGRAD_TC=0 python scripts/wgrad_compare.py   (warp-level mma.sync kernel only)
python scripts/wgrad_compare.py                  (tcgen05 kernel where eligible)"""
import os, sys, torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from eda_b200 import attn_ops as ops
from benchmarks.kernels import time_ms
dev = torch.device("cuda", 0)
g = torch.Generator().manual_seed(0)
for R, N, K, nprob in ((8192, 288, 288, 1), (8192, 288, 288, 3), (8192, 256, 288, 1), (8192, 288, 256, 1), (2048, 288, 288, 1),
                       (2048, 288, 288, 3), (640, 288, 288, 1), (640, 288, 768, 1)):
    probs = []
    for i in range(nprob):
        probs.append(dict(dy=torch.randn(R, N, generator=g).to(dev), x=torch.randn(R, K, generator=g).to(dev),
                          dw=torch.zeros(N, K, device=dev), db=torch.zeros(N, device=dev)))
    ms = time_ms(lambda: ops.wgrad(probs, N, K), 3, 20, graph=True)
    fl = 2.0 * R * N * K * nprob
    print(f"R {R:5d} N {N:3d} K {K:3d} x{nprob}: {ms * 1e3:7.1f} us  {fl / ms / 1e9:6.1f} TFLOP/s   (EDA_WGRAD_TC={os.environ.get('EDA_WGRAD_TC', '1')})")
