"""Phase timeline (cycles) of CTA 0 of the tcgen05 weight-gradient kernel."""
import ctypes, os, sys, torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from eda_b200 import _lib, attn_ops as ops
lib = _lib.load()
dev = torch.device("cuda", 0)
for R, N, K in ((640, 288, 288), (2048, 288, 288), (8192, 288, 288)):
    dy, x = torch.randn(R, N, device=dev), torch.randn(R, K, device=dev)
    dw, db = torch.zeros(N, K, device=dev), torch.zeros(N, device=dev)
    for _ in range(3):
        ops.wgrad([dict(dy=dy, x=x, dw=dw, db=db)], N, K)
    torch.cuda.synchronize()
    ts = (ctypes.c_longlong * 16)()
    lib.eda_debug_timestamps(ts, -16)
    t = list(ts)
    print(f"R={R}: setup {t[1]-t[0]}  first stage landed +{t[2]-t[1]}  rounded +{t[3]-t[2]}  accumulators complete +{t[4]-t[3]}  "
          f"bias sums +{t[5]-t[4]}  epilogue +{t[6]-t[5]}  teardown +{t[7]-t[6]}  total {t[7]-t[0]} cycles")
