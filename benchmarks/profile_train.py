"""Where the training step's time goes: torch.profiler (CUPTI) kernel table + named ranges around every
autograd.Function backward of this package.  Diagnostic only (numbers under a profiler are never bench values).

  python benchmarks/profile_train.py [out.json]
"""
import json
import os
import sys

import torch
from torch.profiler import ProfilerActivity, profile, record_function

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "benchmarks"))
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
import attn_cases as ac  # noqa: E402
import micro_train as mt  # noqa: E402

from eda_b200 import attn_ops, ddp, synthetic  # noqa: E402
from eda_b200.pointnet2 import fused  # noqa: E402


def wrap_backward(cls, name):
    orig = cls.backward

    def bw(ctx, *g):
        with record_function(name):
            return orig(ctx, *g)

    cls.backward = staticmethod(bw)


def main():
    out_path = sys.argv[1] if len(sys.argv) > 1 else None
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    for cls, name in ((fused.FusedSAFunction, "bw_sa"), (attn_ops._MHABlockFn, "bw_mha"),
                      (attn_ops._FFNBlockFn, "bw_ffn"), (attn_ops._LinearFn, "bw_linear")):
        if hasattr(cls, "backward"):
            wrap_backward(cls, name)
    B, N, L, D, K = 8, 50000, 80, 132, 256
    g = torch.Generator().manual_seed(100)
    r = lambda *s: torch.randn(*s, generator=g).to(dev)
    pc = synthetic.point_clouds(B, N, "surface", seed=synthetic.SEED).to(dev)
    text, det, query, pos = r(B, L, ac.E), r(B, D, ac.E), r(B, K, ac.E), 0.5 * r(B, 1024, ac.E)
    text_mask = ac.ragged_mask(B, L, 20, g).to(dev)
    det_mask = ac.ragged_mask(B, D, 20, g).to(dev)
    qpos = torch.cat([4 * torch.rand(B, K, 3, generator=g) - 2, torch.rand(B, K, 3, generator=g) + .2], -1).to(dev)
    args = (pc, pos, text, text_mask, det, det_mask, query, qpos)
    torch.manual_seed(0)
    model = mt.HotPath()
    ac.fill_params(model.encoder, 1)
    for i, d in enumerate(model.decoder):
        ac.fill_params(d, 10 + i)
    model = model.to(dev).train()
    fg = ddp.FlatGradients(model)
    torch.backends.cuda.matmul.allow_tf32 = False

    def step():
        fg.zero()
        with record_function("fwd"):
            loss = mt.loss_of(model(*args))
        with record_function("bwd"):
            loss.backward()
            fg.sync()
        return loss.detach()

    for _ in range(3):
        step()
    torch.cuda.synchronize()
    # wall-clock split with events (no profiler)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    fg.zero()
    ev[0].record()
    loss = mt.loss_of(model(*args))
    ev[1].record()
    loss.backward()
    fg.sync()
    ev[2].record()
    torch.cuda.synchronize()
    del loss
    res = {"fwd_ms": ev[0].elapsed_time(ev[1]), "bwd_ms": ev[1].elapsed_time(ev[2])}
    with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
        step()
        torch.cuda.synchronize()
    ka = prof.key_averages()
    rows = []
    for e in ka:
        dt = getattr(e, "device_time_total", None)
        if dt is None:
            dt = getattr(e, "cuda_time_total", 0)
        sdt = getattr(e, "self_device_time_total", None)
        if sdt is None:
            sdt = getattr(e, "self_cuda_time_total", 0)
        rows.append({"name": e.key[:110], "count": e.count, "device_us": dt, "self_device_us": sdt,
                     "cpu_us": e.cpu_time_total})
    ranges = [r_ for r_ in rows if r_["name"] in ("fwd", "bwd", "bw_sa", "bw_mha", "bw_ffn", "bw_linear")]
    kernels = sorted(rows, key=lambda r_: -r_["self_device_us"])[:45]
    from torch.autograd import DeviceType
    res["device_busy_ms_total"] = sum(e.self_device_time_total for e in ka
                                      if getattr(e, "device_type", None) == DeviceType.CUDA) / 1000.0
    res["ranges"] = ranges
    res["top_self_device"] = kernels
    # ---- the same step as one CUDA graph: per-kernel device time of a replay (no host gaps) ----
    try:
        from eda_b200.graphs import GraphedTrainStep

        gstep = GraphedTrainStep(model, mt.loss_of, list(args), fg)
        for _ in range(3):
            gstep(*args)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        gstep(*args)
        e1.record()
        torch.cuda.synchronize()
        res["graphed_ms"] = e0.elapsed_time(e1)
        with profile(activities=[ProfilerActivity.CUDA]) as prof2:
            gstep(*args)
            torch.cuda.synchronize()
        agg = {}
        for ev in prof2.events():
            if ev.device_type == DeviceType.CUDA:
                a = agg.setdefault(ev.name[:90], [0, 0.0])
                a[0] += 1
                a[1] += ev.device_time
        res["graphed_kernel_sum_ms"] = sum(v[1] for v in agg.values()) / 1000.0
        res["graphed_kernels"] = sorted(([k, v[0], round(v[1] / 1000.0, 3)] for k, v in agg.items()), key=lambda r_: -r_[2])[:40]
    except Exception as e:  # noqa: BLE001
        res["graphed_error"] = repr(e)[:400]
    print(json.dumps(res, indent=1))
    if out_path:
        with open(out_path, "w") as f:
            json.dump(res, f, indent=1)


if __name__ == "__main__":
    main()
