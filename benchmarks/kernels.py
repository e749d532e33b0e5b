"""Stand-alone kernel timings behind bench.py's `roofline`, `rooflines` and `sweep` objects: each kernel of the hot
path timed ALONE with CUDA events on the launching stream (after warm-up, synchronised on both sides), against the
roofline that bounds it.  Algorithmic bytes / flops per launch are SURVEY.md 8(d)'s (restated in DESIGN.md section 3).

Nothing here is on the product path; nothing here touches oracle/.
"""
import ctypes
import json
import math
import os

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def peaks():
    """(HBM GB/s, tf32 TFLOP/s burst, tf32 sustained, SM MHz max, source string).  kind::tf32 runs at half the bf16 rate."""
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        d = json.load(open(path))
        return (float(d["hbm_gbs"]), float(d["bf16_tflops"]) / 2, float(d.get("bf16_tflops_sustained", d["bf16_tflops"])) / 2,
                float(d.get("sm_max_mhz", 1965.0)), "measured (MEASURED_PEAKS.json: hbm_gbs; bf16_tflops / 2 for kind::tf32)")
    except Exception:  # noqa: BLE001
        return 6650.0, 2250.0 / 2 * 0.75, 2250.0 / 2 * 0.62, 1965.0, "fallback (B200_PROFILING.md)"


def time_ms(fn, warmup=3, iters=20, graph=False):
    """Average device time of fn() in ms: `iters` back-to-back launches between two events on the current stream.
    graph=True captures the `iters` launches into one CUDA graph first and times its replay: for kernels of a few tens of
    microseconds the Python / ctypes / tensor-map-encode time of a launch (10 - 15 us) would otherwise be what is measured
    (the training step replays a graph, so this is also how the kernel runs in the product)."""
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    if graph:
        g = torch.cuda.CUDAGraph()
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            with torch.cuda.graph(g, stream=side):
                for _ in range(iters):
                    fn()
        torch.cuda.synchronize()
        g.replay()  # warm replay
        torch.cuda.synchronize()
        reps = 5
        a.record()
        for _ in range(reps):
            g.replay()
        b.record()
        torch.cuda.synchronize()
        ms = a.elapsed_time(b) / (iters * reps)
        del g
        return ms
    a.record()
    for _ in range(iters):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / iters


def traffic_table():
    """ncu DRAM bytes per launch of the kernels below (from one `ncu --set full` capture, profiles/roofline_traffic.json)."""
    try:
        return json.load(open(os.path.join(ROOT, "profiles", "roofline_traffic.json")))
    except Exception:  # noqa: BLE001
        return {}


def fps_point(B, N, m, dev, sm_mhz, with_floor=True):
    """FPS N -> m on B scenes: ms, algorithmic GB/s ((m-1) * N * 20 B per scene), cycles per iteration and the measured
    latency floor of the reduction / cluster-exchange chain at the same decomposition (eda_selftest_fps_exchange)."""
    from eda_b200 import _lib, synthetic
    from eda_b200.pointnet2 import _ext

    lib = _lib.load()
    xyz = synthetic.point_clouds(B, N, "surface", channels=0).to(dev).contiguous()
    ms = time_ms(lambda: _ext.furthest_point_sampling(xyz, m), 2, 5)
    alg = (m - 1) * N * 20 * B
    cl, th, ppt = ctypes.c_int(), ctypes.c_int(), ctypes.c_int()
    lib.eda_fps_plan(B, N, m, ctypes.byref(cl), ctypes.byref(th), ctypes.byref(ppt))
    out = {"B": B, "N": N, "m": m, "ms": ms, "algorithmic_bytes": alg, "algorithmic_GBps": alg / (ms * 1e-3) / 1e9,
           "compulsory_bytes": B * (12 * N + 4 * m), "cluster": cl.value, "threads": th.value,
           "points_per_thread": ppt.value, "cycles_per_iteration": ms * 1e-3 * sm_mhz * 1e6 / max(m - 1, 1)}
    if with_floor and cl.value >= 1 and th.value in (256, 512):
        sink = torch.zeros(B, dtype=torch.int32, device=dev)
        st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
        rc = lib.eda_selftest_fps_exchange(B, cl.value, th.value, m, ctypes.c_void_p(sink.data_ptr()), st)
        if rc == 0:
            fl = time_ms(lambda: lib.eda_selftest_fps_exchange(B, cl.value, th.value, m, ctypes.c_void_p(sink.data_ptr()), st), 2, 5)
            out["exchange_floor_ms"] = fl
            out["exchange_floor_cycles_per_iteration"] = fl * 1e-3 * sm_mhz * 1e6 / max(m - 1, 1)
            out["frac_of_latency_floor"] = fl / ms
    return out


def ball_query_point(B, N, m, radius, nsample, dev):
    """Ball query of m FPS centres against N points: ms and algorithmic GB/s (m*N*12 + m*ns*4 bytes per scene)."""
    from eda_b200 import synthetic
    from eda_b200.pointnet2 import _ext

    xyz = synthetic.point_clouds(B, N, "surface", channels=0).to(dev).contiguous()
    inds = _ext.furthest_point_sampling(xyz, m)
    new_xyz = torch.gather(xyz, 1, inds.long().unsqueeze(-1).expand(-1, -1, 3)).contiguous()
    ms = time_ms(lambda: _ext.ball_query(new_xyz, xyz, radius, nsample), 2, 10)
    alg = (m * N * 12 + m * nsample * 4) * B
    return {"B": B, "N": N, "m": m, "radius": radius, "nsample": nsample, "ms": ms, "algorithmic_bytes": alg,
            "algorithmic_GBps": alg / (ms * 1e-3) / 1e9, "compulsory_bytes": B * (12 * N + 12 * m + 4 * m * nsample)}


def attention_point(B, Nq, Nk, dev, H=8, E=288):
    """Attention core (QK^T, softmax, PV) on projected inputs: ms, TFLOP/s over the 4*Nq*Nk*E contraction flops."""
    from eda_b200 import attn_ops as ops

    g = torch.Generator().manual_seed(0)
    q = torch.randn(B * Nq, E, generator=g).to(dev)
    k = torch.randn(B * Nk, E, generator=g).to(dev)
    ld = (Nk + 3) & ~3
    vt = torch.randn(B, E, ld, generator=g).to(dev)
    ms = time_ms(lambda: ops.attention_raw(q, k, vt, None, B, Nq, Nk, H), 3, 20, graph=True)
    flops = 4.0 * Nq * Nk * E * B
    return {"B": B, "Nq": Nq, "Nk": Nk, "ms": ms, "flops": flops, "TFLOPs": flops / (ms * 1e-3) / 1e12}


def linear_point(R, K, N, dev, ln=False):
    """eda_linear_forward on one (R,K) x (N,K)^T problem (bias; optionally residual + LayerNorm): ms and TFLOP/s."""
    from eda_b200 import attn_ops as ops

    g = torch.Generator().manual_seed(0)
    x = torch.randn(R, K, generator=g).to(dev)
    w = (torch.randn(N, K, generator=g) / math.sqrt(K)).to(dev)
    b = torch.randn(N, generator=g).to(dev)
    wp = ops.pack_weight(w)
    res = torch.randn(R, N, generator=g).to(dev) if ln else None
    lnp = (torch.ones(N, device=dev), torch.zeros(N, device=dev), 1e-5) if ln else None
    # L2-warm: the same input every launch (how the kernel meets its input in the step: written by the kernel before it)
    warm = time_ms(lambda: ops.linear_raw([dict(x=x, w_packed=wp, bias=b, residual=res)], K, N, ln=lnp), 3, 30, graph=True)
    # L2-cold: inputs / outputs rotate through more buffers than the 126 MB L2 holds
    nbuf = max(2, int(160e6 // (4.0 * R * (K + N))) + 1)
    xs = [x.clone() for _ in range(nbuf)]
    ys = [torch.empty(R, N, device=dev) for _ in range(nbuf)]
    state = {"i": 0}

    def cold():
        i = state["i"] = (state["i"] + 1) % nbuf
        ops.linear_raw([dict(x=xs[i], w_packed=wp, bias=b, residual=res, y_into=ys[i])], K, N, ln=lnp)

    ms = time_ms(cold, 3, 2 * nbuf, graph=True)
    flops = 2.0 * R * K * N
    return {"R": R, "K": K, "N": N, "ln_epilogue": ln, "ms": ms, "ms_l2_warm": warm, "flops": flops,
            "TFLOPs": flops / (ms * 1e-3) / 1e12, "TFLOPs_l2_warm": flops / (warm * 1e-3) / 1e12,
            "timing": f"CUDA-graph replay of back-to-back launches; inputs and outputs rotate through {nbuf} buffers "
                      f"({nbuf * 4.0 * R * (K + N) / 1e6:.0f} MB > L2)",
            "min_bytes": 4.0 * (R * K + R * N * (2 if ln else 1) + N * K)}


def wgrad_point(R, N, K, nprob, dev):
    """eda_wgrad (dW += dY^T X, db += column sums) on `nprob` problems of one shape in one launch: ms and TFLOP/s."""
    from eda_b200 import attn_ops as ops

    g = torch.Generator().manual_seed(0)
    probs = [dict(dy=torch.randn(R, N, generator=g).to(dev), x=torch.randn(R, K, generator=g).to(dev),
                  dw=torch.zeros(N, K, device=dev), db=torch.zeros(N, device=dev)) for _ in range(nprob)]
    ms = time_ms(lambda: ops.wgrad(probs, N, K), 3, 20, graph=True)
    flops = 2.0 * R * N * K * nprob
    return {"R": R, "N": N, "K": K, "problems": nprob, "ms": ms, "flops": flops, "TFLOPs": flops / (ms * 1e-3) / 1e12,
            "min_bytes": 4.0 * nprob * (R * (N + K) + N * K)}


def sa_mlp_point(B, N, M, S, C, widths, radius, dev):
    """Fused group + 3-layer MLP + max-pool (eval-mode weights): ms and TFLOP/s over 2*M*S*sum(Cin*Cout) per scene."""
    from eda_b200 import synthetic
    from eda_b200.pointnet2 import _ext
    from eda_b200.pointnet2.pointnet2_modules import PointnetSAModuleVotes

    torch.manual_seed(0)
    sa = PointnetSAModuleVotes(npoint=M, radius=radius, nsample=S, mlp=[C] + list(widths), use_xyz=True,
                               normalize_xyz=True).to(dev).eval()
    pc = synthetic.point_clouds(B, N, "surface", channels=3).to(dev)
    xyz = pc[..., :3].contiguous()
    feats = torch.randn(B, C, N, device=dev) if C != 3 else pc[..., 3:].transpose(1, 2).contiguous()
    inds = _ext.furthest_point_sampling(xyz, M)
    new_xyz = torch.gather(xyz, 1, inds.long().unsqueeze(-1).expand(-1, -1, 3)).contiguous()
    idx = _ext.ball_query(new_xyz, xyz, radius, S)
    from eda_b200.pointnet2 import fused

    layers = sa.mlp_module.fusable_layers()
    feat_pm = fused.point_major(feats)
    with torch.no_grad():
        ms = time_ms(lambda: fused.sa_forward_raw(xyz, new_xyz, feat_pm, idx, layers, radius, True, False), 3, 20)
    dims = [C + 3] + list(widths)
    flops = 2.0 * B * M * S * sum(a * b for a, b in zip(dims[:-1], dims[1:]))
    return {"B": B, "N": N, "M": M, "S": S, "ms": ms, "flops": flops, "TFLOPs": flops / (ms * 1e-3) / 1e12}


def sweep(dev, B=8, quick=False):
    """BASELINE.json configs[4]: N in {20k, 50k, 100k, 200k} x nsample in {16, 32, 64} (npoint 2048, r = 0.2) for FPS /
    ball query (GB/s against the HBM peak, cycles per iteration against the exchange floor), plus the attention core at
    the encoder's vis-self and the decoder's cross_v shapes (TFLOP/s against the tf32 peak).  ncu counters of the same
    points (DRAM / L2 bytes, tensor-pipe %) are merged in from profiles/r2_sweep_ncu.json when present."""
    hbm, tf32, _, sm_mhz, src = peaks()
    Ns = (20000, 50000) if quick else (20000, 50000, 100000, 200000)
    out = {"peaks": {"hbm_GBps": hbm, "tf32_TFLOPs": tf32, "source": src}, "fps": [], "ball_query": [], "attention": []}
    for N in Ns:
        p = fps_point(B, N, 2048, dev, sm_mhz)
        p["frac_of_hbm_peak_algorithmic"] = p["algorithmic_GBps"] / hbm
        out["fps"].append(p)
        for ns in (16, 32, 64):
            q = ball_query_point(B, N, 2048, 0.2, ns, dev)
            q["frac_of_hbm_peak_algorithmic"] = q["algorithmic_GBps"] / hbm
            out["ball_query"].append(q)
    for name, Nq, Nk in (("vis_self", 1024, 1024), ("cross_v", 256, 1024), ("cross_vl", 1024, 80), ("dec_self", 256, 256)):
        a = attention_point(B, Nq, Nk, dev)
        a["name"] = name
        a["frac_of_tf32_peak"] = a["TFLOPs"] / tf32
        out["attention"].append(a)
    try:
        ncu = json.load(open(os.path.join(ROOT, "profiles", "r2_sweep_ncu.json")))
        out["ncu_counters"] = ncu
    except Exception:  # noqa: BLE001
        out["ncu_counters"] = None
    return out


if __name__ == "__main__":
    import sys

    sys.path.insert(0, ROOT)
    d = torch.device("cuda", 0)
    print(json.dumps(sweep(d, quick="--quick" in sys.argv), indent=1))
