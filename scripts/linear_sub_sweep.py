"""linear_kernel device time (CUDA-graph replay) for the step's row counts; run with EDA_LINEAR_SUB=1|2|3 to force the
ring-stage width (32 / 64 / 96 K-columns)."""
import os, sys, torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from eda_b200 import attn_ops as ops
from benchmarks.kernels import time_ms
dev = torch.device("cuda", 0)
g = torch.Generator().manual_seed(0)
for R, K, N, ln, nprob in ((8192, 288, 288, False, 1), (8192, 288, 288, False, 3), (8192, 288, 288, True, 1), (2048, 288, 288, False, 1),
                           (2048, 288, 288, False, 3), (2048, 288, 288, True, 1), (640, 288, 288, False, 1), (2048, 288, 256, False, 1), (2048, 256, 288, True, 1)):
    x = torch.randn(R, K, generator=g).to(dev)
    ws = [ops.pack_weight((torch.randn(N, K, generator=g) / K ** 0.5).to(dev)) for _ in range(nprob)]
    b = torch.randn(N, generator=g).to(dev)
    res = torch.randn(R, N, generator=g).to(dev) if ln else None
    lnp = (torch.ones(N, device=dev), torch.zeros(N, device=dev), 1e-5) if ln else None
    ms = time_ms(lambda: ops.linear_raw([dict(x=x, w_packed=w, bias=b, residual=res) for w in ws], K, N, ln=lnp), 3, 30, graph=True)
    print(f"R {R:5d} K {K} N {N} ln {int(ln)} x{nprob}: {ms * 1e3:6.2f} us   SUB={os.environ.get('EDA_LINEAR_SUB', 'auto')}")
