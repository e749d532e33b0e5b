"""Per-kernel timings of the backward building blocks at the shapes of the attention stack and the SA stages
(B = 8 scenes, E = 288, H = 8): CUDA events, median of 20 after 5 warm-ups, L2 not flushed (operands of these
kernels are produced by the preceding kernel in the real step, i.e. L2-warm there too).  Each sample brackets ONE call of
the Python wrapper, so entries below ~100 us are dominated by host time (allocation + ctypes + launch), not by the kernel:
device times per kernel are in profiles/r1_bwd_kernel_metrics.json (ncu) — e.g. the 1024 x 1024 attention backward is
74 + 90 us on the device, 369 us through the wrapper.

  python benchmarks/micro_bwd.py [out.json]
"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from eda_b200 import attn_ops as ops  # noqa: E402


def timeit(fn, warm=5, iters=20):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b) * 1e3)
    ts.sort()
    return round(ts[len(ts) // 2], 1)


def main():
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    B, E, H = 8, 288, 8
    g = torch.Generator().manual_seed(0)
    r = lambda *s: torch.randn(*s, generator=g).to(dev)
    res = {"unit": "us", "attention": {}, "gemm": {}, "sa": {}}
    W = r(E, E) / 17
    wt = ops.pack_weight_t(W)
    for Nq, Nk in [(80, 80), (1024, 1024), (80, 1024), (1024, 80), (1024, 132), (256, 256), (256, 80), (256, 132),
                   (256, 1024)]:
        q, k = r(B * Nq, E), r(B * Nk, E)
        ld = (Nk + 3) & ~3
        vt = r(B, E, ld)
        dctx = r(B * Nq, E)
        lse = torch.empty(B, H, Nq, device=dev)
        c = ops.attention_raw(q, k, vt, None, B, Nq, Nk, H, lse=lse)
        e = {}
        e["attn_fwd"] = timeit(lambda: ops.attention_raw(q, k, vt, None, B, Nq, Nk, H, lse=lse))
        flops = 4.0 * Nq * Nk * 36 * H * B
        for impl in ("tc", "mma"):
            key = f"attn_bwd_{impl} (2 launches + layout copies)"
            e[key] = timeit(lambda: ops.attention_backward_raw(q, k, vt, dctx, c, lse, None, B, Nq, Nk, H, impl=impl))
            e[f"attn_bwd_{impl}_TFLOPs(7 products)"] = round(3.5 * flops / (e[key] * 1e-6) / 1e12, 1)
        res["attention"][f"{Nq}x{Nk}"] = e
    for R in (640, 1056, 2048, 8192):
        x, dy = r(R, E), r(R, E)
        dw = torch.zeros(3 * E, E, device=dev)
        db = torch.zeros(3 * E, device=dev)
        e = {}
        e["dgrad_1"] = timeit(lambda: ops.linear_raw([dict(x=dy, w_packed=wt)], E, E))
        e["dgrad_3"] = timeit(lambda: ops.linear_raw([dict(x=dy, w_packed=wt)] * 3, E, E))
        e["wgrad_1"] = timeit(lambda: ops.wgrad([dict(dy=dy, x=x, dw=dw[:E], db=db[:E])], E, E))
        e["wgrad_5"] = timeit(lambda: ops.wgrad([dict(dy=dy, x=x, dw=dw[i * E:(i + 1) * E]) for i in (0, 1, 2, 0, 1)], E, E))
        u = r(R, E)
        gam = torch.ones(E, device=dev)
        dg, dbb = torch.zeros(E, device=dev), torch.zeros(E, device=dev)
        e["ln_bwd"] = timeit(lambda: ops.layernorm_backward(dy, u, gam, 1e-5, dg, dbb))
        e["pack_t"] = timeit(lambda: ops.pack_weight_t(W))
        e["zeros_flat"] = timeit(lambda: torch.zeros(4 * E * E + 6 * E, device=dev))
        res["gemm"][f"R={R}"] = e
    # SA-stage GEMMs (rows = B * npoint * nsample)
    for name, R, dims in (("sa1", 8 * 2048 * 64, (16, 64, 64, 128)), ("sa2", 8 * 1024 * 32, (144, 128, 128, 256)),
                          ("sa3", 8 * 512 * 16, (272, 128, 128, 256)), ("sa4", 8 * 256 * 16, (272, 128, 128, 256))):
        e = {}
        for l in range(3):
            K, N = dims[l], dims[l + 1]
            x, dy = r(R, K), r(R, N)
            Wl = r(N, K) / 10
            wp, wpt = ops.pack_weight(Wl), ops.pack_weight_t(Wl)
            dw = torch.zeros(N, K, device=dev)
            e[f"L{l + 1} fwd {K}->{N}"] = timeit(lambda: ops.linear_raw([dict(x=x, w_packed=wp)], K, N), 2, 8)
            e[f"L{l + 1} dgrad {N}->{K}"] = timeit(lambda: ops.linear_raw([dict(x=dy, w_packed=wpt)], N, K), 2, 8)
            e[f"L{l + 1} wgrad"] = timeit(lambda: ops.wgrad([dict(dy=dy, x=x, dw=dw)], N, K), 2, 8)
            del x, dy
        res["sa"][name] = e
    print(json.dumps(res, indent=1))
    if len(sys.argv) > 1:
        with open(sys.argv[1], "w") as f:
            json.dump(res, f, indent=1)


if __name__ == "__main__":
    main()
