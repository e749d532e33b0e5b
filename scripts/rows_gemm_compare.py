"""Row-streaming GEMMs of the SA stage at their step shapes: device time and achieved HBM bandwidth against the compulsory
4 (K + N) bytes per row.  EDA_ROWS_GEMM_CTAS=1 python scripts/rows_gemm_compare.py compares with one CTA per SM."""
import os, sys, torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from eda_b200 import attn_ops as ops
from benchmarks.kernels import time_ms
dev = torch.device("cuda", 0)
g = torch.Generator().manual_seed(0)
shapes = [(1048576, 8, 64, 0, 1), (1048576, 64, 64, 1, 1), (1048576, 64, 128, 1, 1), (1048576, 128, 64, 0, 0), (1048576, 64, 64, 0, 0),
          (262144, 136, 128, 0, 1), (262144, 128, 128, 1, 1), (262144, 128, 256, 1, 1), (262144, 256, 128, 0, 0), (262144, 128, 128, 0, 0)]
for R, K, N, pro, st in shapes:
    x = torch.randn(R, K, generator=g).to(dev)
    W = (torch.randn(N, K, generator=g) / K ** 0.5).to(dev)
    sc = torch.rand(K, generator=g).to(dev) + 0.5 if pro else None
    sh = torch.randn(K, generator=g).to(dev) if pro else None
    stats = torch.zeros(2 * N, dtype=torch.float64, device=dev) if st else None
    ms = time_ms(lambda: ops.rows_gemm(x, W, in_scale=sc, in_shift=sh, stats=stats), 2, 5)
    gb = 4.0 * R * (K + N) / 1e9
    print(f"R {R:8d} K {K:3d} N {N:3d} prologue {pro} stats {st}: {ms * 1e3:7.1f} us  {gb / ms:6.2f} TB/s (compulsory {gb * 1e3:.0f} MB)"
          f"   CTAS={os.environ.get('EDA_ROWS_GEMM_CTAS', '2')}")
