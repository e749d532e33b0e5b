"""Attention backward: warp-level mma.sync kernel vs the tcgen05 kernel on the step's four attention shapes (device time,
CUDA-graph replay of back-to-back calls; the tc path's transposes / channel-major copies are inside its time)."""
import os, sys, torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from eda_b200 import attn_ops as ops
from benchmarks.kernels import time_ms
dev = torch.device("cuda", 0)
B, E, H = 8, 288, 8
g = torch.Generator().manual_seed(0)
r = lambda *s: torch.randn(*s, generator=g).to(dev)
for name, Nq, Nk in (("vis_self", 1024, 1024), ("cross_v", 256, 1024), ("cross_vl", 1024, 80), ("dec_self", 256, 256)):
    ld = (Nk + 3) & ~3
    q, k = r(B * Nq, E), r(B * Nk, E)
    vt = r(B, E, ld); dctx = r(B * Nq, E)
    lse = torch.empty(B, H, Nq, device=dev)
    c = ops.attention_raw(q, k, vt, None, B, Nq, Nk, H, lse=lse)
    fwd = time_ms(lambda: ops.attention_raw(q, k, vt, None, B, Nq, Nk, H, lse=lse), 3, 10, graph=True)
    out = {}
    for impl in ("mma", "tc"):
        out[impl] = time_ms(lambda: ops.attention_backward_raw(q, k, vt, dctx, c, lse, None, B, Nq, Nk, H, impl=impl), 3, 10, graph=True)
    a = ops.attention_backward_raw(q, k, vt, dctx, c, lse, None, B, Nq, Nk, H, impl="mma")
    b = ops.attention_backward_raw(q, k, vt, dctx, c, lse, None, B, Nq, Nk, H, impl="tc")
    err = max(float((x - y).abs().max() / x.abs().max()) for x, y in zip(a, b))
    print(f"{name:9s} Nq {Nq:4d} Nk {Nk:4d}: forward {fwd * 1e3:7.1f} us   backward mma {out['mma'] * 1e3:7.1f} us   tc {out['tc'] * 1e3:7.1f} us   max rel diff {err:.2e}")
