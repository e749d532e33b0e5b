"""LayerNorm backward device time (graph replay) at the step's row counts; EDA_LN_BWD_ROWS / EDA_LN_BWD_CAP vary the grid."""
import os, sys, torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from eda_b200 import attn_ops as ops
from benchmarks.kernels import time_ms
dev = torch.device("cuda", 0)
for R in (640, 2048, 8192):
    N = 288
    dy, u = torch.randn(R, N, device=dev), torch.randn(R, N, device=dev)
    g = torch.randn(N, device=dev); dg, db = torch.zeros(N, device=dev), torch.zeros(N, device=dev)
    ms = time_ms(lambda: ops.layernorm_backward(dy, u, g, 1e-5, dg, db), 3, 30, graph=True)
    print(f"rows {R:5d}: {ms * 1e3:6.2f} us   ROWS={os.environ.get('EDA_LN_BWD_ROWS', '8')} CAP={os.environ.get('EDA_LN_BWD_CAP', '1')}")
