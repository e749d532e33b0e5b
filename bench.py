#!/usr/bin/env python
"""bench.py — the measurement contract.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

Workload = BASELINE.json's metric, on the configuration it is quoted on (configs[3]): ONE TRAINING STEP (forward +
backward + DDP gradient all-reduce) of the hot path — Pointnet2Backbone (models/backbone_module.py:26-144) ->
3 x BiEncoderLayer -> 6 x BiDecoderLayer (models/encoder_decoder_layers.py:189-407) — on B = 8 scenes per GPU of
N = 50 000 points (xyz + rgb), L = 80 text tokens, D = 132 detected boxes, K = 256 queries, d_model 288, 8 heads,
train-mode BatchNorm (SyncBatchNorm semantics on > 1 GPU like main_utils.py:335-338), dropout 0 (the parity
configuration), synthetic quadratic loss (the reference's loss / heads / text tower are outside the path),
weak scaling over GPUs.  Metric: scenes/s.

One JSON line on stdout (rank 0):
  value         K steps with all inputs resident in HBM (CUDA events per step, L2 flushed between steps), max over ranks
  e2e           the same step through the public API with HOST inputs: every step copies all eight input tensors from
                pinned host memory and reads the loss back; copies inside the timed region
  roofline      the dominant kernel family of the step (the tcgen05 linear GEMM) timed alone against the measured tf32 peak
  cpu_baseline  the reference's own Python modules on CPU (+ the C port of `_ext`, which the reference does not have on
                CPU), same step, bounded sample
  extras        ref_cuda (the reference CUDA build timed on this box: the denominator of the >= 10x north-star target),
                eval_forward (configs[2] shapes), sa_forward (configs[1]), rooflines (FPS / ball query / attention / SA
                MLP), sweep (configs[4]), dropout_0.1
`--impl reference`: the same training step on the host CPU (reference Python modules + C port of `_ext`), bounded sample.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

B_PER_GPU = 8
N_POINTS = 50000
L_TEXT, D_BOXES, K_QUERIES = 80, 132, 256
METRIC = "scenes/sec fwd+bwd (N=50000 pts, 256 queries, L=80)"
WORKLOAD = ("configs[3]: training step (forward + backward + DDP gradient all-reduce) of the hot path "
            "Pointnet2Backbone -> 3 BiEncoderLayer -> 6 BiDecoderLayer; B=8 scenes/GPU, N=50000 points (xyz+rgb), L=80 text "
            "tokens, D=132 boxes, K=256 queries, d_model 288, 8 heads, train-mode BatchNorm, dropout 0, synthetic "
            "quadratic loss")
CPU_SAMPLE_SCENES = 2  # scenes per step of the CPU legs (bounded sample of the same workload)


def config():
    """Identical in both arms (the driver compares them)."""
    return {"workload": WORKLOAD, "batch_per_gpu": B_PER_GPU, "n_points": N_POINTS, "text_tokens": L_TEXT,
            "boxes": D_BOXES, "queries": K_QUERIES, "d_model": 288, "loss": "synthetic quadratic (hotpath.quadratic_loss)",
            "batchnorm": "train mode (batch statistics; synchronised across ranks when n_gpus > 1)", "dropout": 0.0,
            "sharding": "batch only; one gradient all-reduce per step",
            "l2": "512 MB memset between timed steps (a step streams ~2 GB of activations, far beyond the 126 MB L2)"}


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="eda_b200", choices=["eda_b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true", help="skip the cpu_baseline leg (development)")
    ap.add_argument("--no-extras", action="store_true", help="skip ref_cuda / eval / sweep / roofline extras (development)")
    ap.add_argument("--quick-sweep", action="store_true", help="sweep only N in {20k, 50k} (development)")
    return ap.parse_args()


# ------------------------------------------------------------------------------------------------
# CPU legs: the reference's own Python modules + the C port of `_ext` (oracle/) — reported baseline and reference arm
# ------------------------------------------------------------------------------------------------
def run_cpu(steps, warmup, scenes):
    """Times `steps` training steps of the hot path over `scenes` scenes on the host CPU.  Returns (scenes/s, ms/step,
    cores, kind)."""
    import torch

    from eda_b200 import hotpath
    from oracle import pointnet2_oracle, ref_hotpath, ref_model

    pointnet2_oracle.build()
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    torch.manual_seed(0)
    model = ref_hotpath.build(ref_model.oracle_ext(), dropout=0.0).train()
    inputs = hotpath.synthetic_inputs(scenes, N_POINTS, L_TEXT, D_BOXES, K_QUERIES, seed=100)

    def step():
        for p in model.parameters():
            p.grad = None
        loss = hotpath.quadratic_loss(model(*inputs))
        loss.backward()
        return float(loss.detach())

    for _ in range(warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    dt = time.perf_counter() - t0
    return scenes * steps / dt, dt / steps * 1e3, cores


def cpu_sample_text(scenes, cores, ms):
    return (f"{scenes} of the {B_PER_GPU} scenes of one batch per step (reference Python modules unmodified: "
            f"Pointnet2Backbone, BiEncoder, BiDecoderLayer under torch CPU autograd with {cores} intra-op threads + the C "
            f"port of the nine `_ext` ops, which the reference only has for CUDA; one thread per scene), "
            f"{ms:.0f} ms per step; scenes/s = scenes * steps / time")


def reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return  # rank 0 alone runs the CPU reference arm
    warm = max(min(args.warmup, 2), 1)
    sps, ms, cores = run_cpu(args.steps, warm, CPU_SAMPLE_SCENES)
    line = {
        "impl": "reference", "metric": METRIC, "value": sps, "unit": "scenes/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": warm, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": config(),
        "cpu_baseline": {"value": sps, "unit": "scenes/s", "cores": cores, "kind": "port",
                         "sample": cpu_sample_text(CPU_SAMPLE_SCENES, cores, ms)},
        "e2e": {"value": sps, "unit": "scenes/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
# clocks sampler
# ------------------------------------------------------------------------------------------------
class ClockSampler:
    FIELDS = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits", "-lms", "100", "-i",
                 str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:  # noqa: BLE001
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), line.strip()))

    def stop(self, t0, t1):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()  # the exact PID we started
        try:
            self.proc.wait(timeout=2)
        except Exception:  # noqa: BLE001
            self.proc.kill()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        sm, mx, reasons = [], None, set()
        for ts, line in self.rows:
            parts = [p.strip() for p in line.split(",")]
            if len(parts) < 6:
                continue
            try:
                clk = float(parts[0])
                mx = float(parts[1])
            except ValueError:
                continue
            if t0 <= ts <= t1 + 0.2:
                sm.append(clk)
                for n, v in zip(names, parts[2:6]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
        if not sm:  # region shorter than the sampling period: take everything we saw
            for ts, line in self.rows:
                try:
                    sm.append(float(line.split(",")[0]))
                except ValueError:
                    pass
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm)}


# ------------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------------
def build_train_step(dev, world, rank, dropout, inputs):
    """Model + flat gradient bucket + the step recorded as one CUDA graph; returns (step callable, state dict)."""
    import torch

    from eda_b200 import _lib, ddp, hotpath
    from eda_b200.graphs import GraphedTrainStep

    lib = _lib.load()
    torch.manual_seed(0)
    model = hotpath.HotPath(dropout=dropout).to(dev).train()
    ddp.broadcast_parameters(model)
    sync_bn, peer = False, False
    if world > 1:
        from eda_b200 import syncbn

        ddp.convert_sync_batchnorm(model)  # what main_utils.py:335-338 does when more than one GPU is used
        sync_bn = True
        # statistics exchange over NVLink peer memory, fused with the BatchNorm finalisation (falls back to NCCL)
        peer = os.environ.get("EDA_PEER_REDUCE", "1") != "0" and syncbn.enable_peer_reduce(dev)
    fg = ddp.FlatGradients(model)
    if world > 1 and os.environ.get("EDA_OVERLAP_ALLREDUCE", "1") != "0":
        fg.enable_overlap()  # bucketed all-reduce issued during the backward pass (captured inside the graph)
    l0 = lib.eda_launch_count()
    warm = 3
    gstep = GraphedTrainStep(model, hotpath.quadratic_loss, inputs, fg, warmup=warm)
    kernels_per_step = (lib.eda_launch_count() - l0) // (warm + 1)  # warm-up steps + the captured one launch the same list
    overlapped = bool(getattr(gstep, "overlapped_allreduce", False))

    def step():
        loss = gstep(*inputs)
        if not overlapped:
            fg.all_reduce_mean()
        return loss

    return step, dict(model=model, fg=fg, gstep=gstep, kernels_per_step=int(kernels_per_step), sync_bn=sync_bn,
                      overlapped=overlapped, peer=bool(peer))


def gpu_arm(args):
    import torch
    import torch.distributed as dist

    from eda_b200 import _lib, hotpath

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device: the product path has no CPU fallback")
    # the contract is ONE JSON line on stdout: libraries (NCCL's version banner, symmetric-memory setup) write to fd 1 at
    # will, so fd 1 is pointed at stderr for the duration of the run and the line goes out through a saved descriptor
    sys.stdout.flush()
    json_fd = os.dup(1)
    os.dup2(2, 1)
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        # the contract is ONE JSON line on stdout: NCCL prints "NCCL version ..." there at the VERSION level
        if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION"):
            os.environ["NCCL_DEBUG"] = "WARN"
        dist.init_process_group("nccl", device_id=dev)
    lib = _lib.load()  # fails loudly when the CUDA library is missing
    K = args.steps
    W = max(args.warmup, 3)

    host = [t.pin_memory() for t in hotpath.synthetic_inputs(B_PER_GPU, N_POINTS, L_TEXT, D_BOXES, K_QUERIES, seed=100 + rank)]
    inputs = [t.to(dev) for t in host]
    step, st = build_train_step(dev, world, rank, 0.0, inputs)
    flush = torch.empty(512 * 1024 * 1024 // 4, dtype=torch.float32, device=dev)  # > 126 MB L2

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(W):
        step()
    barrier()
    # ---------------- value: inputs resident in HBM ----------------
    sampler = ClockSampler(local)
    sampler.start()
    time.sleep(0.25)
    evs = []
    t0 = time.perf_counter()
    for _ in range(K):
        flush.zero_()  # evict L2 between timed iterations (not timed)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        loss = step()
        b.record()
        evs.append((a, b))
    barrier()
    t1 = time.perf_counter()
    clocks = sampler.stop(t0, t1)
    total_ms = sum(a.elapsed_time(b) for a, b in evs)
    loss_value = float(loss.item())

    # ---------------- e2e: host inputs -> device -> step -> loss back on the host, every step ----------------
    loss_host = torch.zeros(K, dtype=torch.float32).pin_memory()

    def e2e_step(i):
        for dst, src in zip(inputs, host):
            dst.copy_(src, non_blocking=True)
        out = step()
        loss_host[i:i + 1].copy_(out.reshape(1), non_blocking=True)

    for i in range(2):
        e2e_step(i)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(K):
        e2e_step(i)
    e1.record()
    barrier()
    e2e_ms = e0.elapsed_time(e1)
    h2d = sum(t.numel() * t.element_size() for t in host)
    d2h = 4

    times = torch.tensor([total_ms, e2e_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(times, op=dist.ReduceOp.MAX)
    total_ms, e2e_ms = times.tolist()
    scenes = B_PER_GPU * world * K
    line = {
        "metric": METRIC, "value": scenes / (total_ms * 1e-3), "unit": "scenes/s", "n_gpus": world, "steps": K,
        "warmup": W, "ms_per_step": total_ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "tf32", "data": "synthetic", "config": config(),
        "e2e": {"value": scenes / (e2e_ms * 1e-3), "unit": "scenes/s", "h2d_bytes_per_step": h2d,
                "d2h_bytes_per_step": d2h, "ms_per_step": e2e_ms / K},
        "gpu_launches": st["kernels_per_step"] * K,
        "clocks": clocks,
        "execution": {"step": "one CUDA graph per step (forward, backward, weight packing%s)" %
                              (", bucketed gradient all-reduce overlapped with backward" if st["overlapped"] else
                               "); NCCL all-reduce of the flat fp32 gradient bucket after it"),
                      "kernels_per_step": st["kernels_per_step"], "gradient_floats": int(st["fg"].flat.numel()),
                      "sync_batchnorm": st["sync_bn"], "loss": loss_value,
                      "batchnorm_statistics_exchange": ("NVLink peer-memory kernel fused with the finalisation" if st["peer"]
                                                        else "NCCL all-reduce") if st["sync_bn"] else None,
                      "gradient_regions": st["fg"].regions,
                      "index_paths": "fp32, bit-exact", "contractions": "tcgen05 kind::tf32, fp32 accumulate"},
    }
    if True:
        extras = {}
        if not args.no_extras:
            try:
                extras = extra_legs(args, dev, world, rank, inputs, host, total_ms / K)
            except Exception as e:  # noqa: BLE001  (never lose the main line to an extra leg)
                extras = {"extras_error": repr(e)[:400]}
        line.update(extras)
    if "roofline" not in line:
        line["roofline"] = {"bound": "tensor", "achieved": None, "peak": None, "unit": "TFLOP/s", "frac": None,
                            "traffic": None, "note": "extras skipped"}
    if rank == 0:
        if not args.no_cpu_baseline and world == 1:
            try:
                sps, ms, cores = run_cpu(2, 1, CPU_SAMPLE_SCENES)
                line["cpu_baseline"] = {"value": sps, "unit": "scenes/s", "cores": cores, "kind": "port",
                                        "sample": "2 timed steps; " + cpu_sample_text(CPU_SAMPLE_SCENES, cores, ms)}
            except Exception as e:  # noqa: BLE001
                line["cpu_baseline"] = {"value": None, "unit": "scenes/s", "cores": os.cpu_count(), "kind": "port",
                                        "sample": "failed: " + repr(e)[:200]}
        sys.stdout.flush()
        os.write(json_fd, (json.dumps(line) + "\n").encode())
    if world > 1:
        # CUDA graphs that captured NCCL collectives (synchronised BatchNorm, overlapped gradient all-reduce) must be
        # released before the communicator goes away; the process then leaves without the NCCL teardown, which can
        # block on captured work
        st.clear()
        del step
        import gc
        gc.collect()
        torch.cuda.synchronize()
        dist.barrier()
        sys.stdout.flush()
        sys.stderr.flush()
        os._exit(0)


def extra_legs(args, dev, world, rank, inputs, host, step_ms):
    """Everything besides the headline: kernel rooflines, the reference CUDA build on this box, eval forward, SA
    forward, dropout 0.1, the configs[4] sweep.  Multi-GPU runs only keep the collective-bearing legs."""
    import torch
    import torch.distributed as dist

    from benchmarks import kernels as kn
    from eda_b200 import hotpath
    from eda_b200.graphs import GraphedCallable

    out = {}
    hbm, tf32, tf32_sus, sm_mhz, src = kn.peaks()
    traffic = kn.traffic_table()

    def maxed(ms):
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return t.item()

    # ---- the step with the reference's default dropout 0.1 (masks from the in-kernel hash) ----
    try:
        step_d, st_d = build_train_step(dev, world, rank, 0.1, inputs)
        ms = maxed(kn.time_ms(step_d, 3, max(5, args.steps // 2)))
        out["dropout_0.1"] = {"ms_per_step": ms, "value": B_PER_GPU * world / (ms * 1e-3), "unit": "scenes/s"}
        del step_d, st_d
    except Exception as e:  # noqa: BLE001
        out["dropout_0.1"] = {"error": repr(e)[:200]}
    # ---- the reference's published operating point is batch 12 per GPU (scripts/train_scanrefer.sh:7): the same step
    # at B = 12 with its peak memory (the SA stage keeps its pre-activations for the backward pass: 1.9 GB at B = 8) ----
    if world == 1:
        try:
            torch.cuda.reset_peak_memory_stats(dev)
            in12 = [t.to(dev) for t in hotpath.synthetic_inputs(12, N_POINTS, L_TEXT, D_BOXES, K_QUERIES, seed=100)]
            step12, st12 = build_train_step(dev, world, rank, 0.0, in12)
            ms = kn.time_ms(step12, 3, max(5, args.steps // 2))
            out["batch_12"] = {"ms_per_step": ms, "value": 12 / (ms * 1e-3), "unit": "scenes/s",
                               "peak_memory_GB": torch.cuda.max_memory_allocated(dev) / 1e9}
            st12.clear()
            del step12, st12, in12
        except Exception as e:  # noqa: BLE001
            out["batch_12"] = {"error": repr(e)[:200]}
    if world > 1 and rank != 0:
        return out

    # ---- roofline of the dominant kernel family: the tcgen05 linear GEMM (q/k/v/out projections, FFNs, activation
    # gradients: ~220 launches per step), heaviest shape = the 8192-row visual stream, 288 x 288 weights ----
    lin = kn.linear_point(B_PER_GPU * 1024, 288, 288, dev)
    out["roofline"] = {
        "kernel": "linear_kernel (eda_linear_forward) R=8192 K=288 N=288, bias epilogue; timed alone: " + lin["timing"],
        "bound": "tensor", "achieved": lin["TFLOPs"], "peak": tf32, "unit": "TFLOP/s", "frac": lin["TFLOPs"] / tf32,
        "traffic": traffic.get("linear_kernel_8192x288x288_dram_bytes_per_launch"), "peak_source": src + " burst",
        "kernel_ms": lin["ms"], "kernel_ms_l2_warm": lin["ms_l2_warm"], "achieved_l2_warm": lin["TFLOPs_l2_warm"],
        "algorithmic_flops_per_launch": lin["flops"], "min_bytes_per_launch": lin["min_bytes"],
        "hbm_frac_at_min_bytes": lin["min_bytes"] / (lin["ms"] * 1e-3) / 1e9 / hbm,
        "share_of_step": traffic.get("linear_kernel_share_of_step"),
    }
    if world > 1:
        return out  # the remaining legs are single-GPU measurements (BENCH), not part of the scaling run
    rl = {}
    fps = kn.fps_point(B_PER_GPU, N_POINTS, 2048, dev, sm_mhz)
    fps.update({"bound": "latency (serial arg-max chain); HBM by SURVEY 8d convention", "hbm_peak_GBps": hbm,
                "frac_of_hbm_peak_algorithmic": fps["algorithmic_GBps"] / hbm,
                "dram_bytes_per_launch_ncu": traffic.get("fps_cluster_kernel_dram_bytes_per_launch"),
                "note": "points and running minima are register-resident: DRAM traffic is the compulsory read only, so "
                        "algorithmic GB/s says nothing about efficiency; the honest bound is cycles_per_iteration vs "
                        "exchange_floor_cycles_per_iteration (measured here by eda_selftest_fps_exchange)"})
    rl["fps_sa1"] = fps
    bq = kn.ball_query_point(B_PER_GPU, N_POINTS, 2048, 0.2, 64, dev)
    bq.update({"hbm_peak_GBps": hbm, "frac_of_hbm_peak_algorithmic": bq["algorithmic_GBps"] / hbm,
               "dram_bytes_per_launch_ncu": traffic.get("ball_query_kernel_dram_bytes_per_launch")})
    rl["ball_query_sa1"] = bq
    for name, Nq, Nk in (("attention_vis_self", 1024, 1024), ("attention_cross_v", 256, 1024)):
        a = kn.attention_point(B_PER_GPU, Nq, Nk, dev)
        a.update({"tf32_peak_TFLOPs": tf32, "frac_of_tf32_peak": a["TFLOPs"] / tf32})
        rl[name] = a
    wg = kn.wgrad_point(B_PER_GPU * 1024, 288, 288, 3, dev)
    wg.update({"tf32_peak_TFLOPs": tf32, "frac_of_tf32_peak": wg["TFLOPs"] / tf32,
               "kernel": "wgrad_tc_kernel (eda_wgrad): the q/k/v in-projection weight gradients of one attention block"})
    rl["wgrad_qkv"] = wg
    sa = kn.sa_mlp_point(B_PER_GPU, N_POINTS, 2048, 64, 3, [64, 64, 128], 0.2, dev)
    sa.update({"tf32_peak_TFLOPs": tf32, "frac_of_tf32_peak": sa["TFLOPs"] / tf32})
    rl["sa_mlp_sa1"] = sa
    out["rooflines"] = rl

    # ---- eval forward of the hot path (configs[2] shapes), ours graphed ----
    from eda_b200 import hotpath as hp

    torch.manual_seed(0)
    model = hp.HotPath(dropout=0.0).to(dev).eval()
    g = GraphedCallable(lambda *a: model(*a)[0], inputs)
    ms = kn.time_ms(lambda: g(*inputs), 3, 20)
    out["eval_forward"] = {"ms": ms, "value": B_PER_GPU / (ms * 1e-3), "unit": "scenes/s",
                           "workload": "configs[2] shapes: hot-path forward, eval-mode BN, one CUDA graph"}

    # ---- the reference CUDA build on this box (R-GPU): its compiled `_ext` + its own Python
    # modules, same parameters, same inputs; TEST/BASELINE infrastructure, never on the product path ----
    try:
        out["ref_cuda"] = ref_cuda_leg(dev, model, inputs, out["eval_forward"]["ms"], step_ms)
    except Exception as e:  # noqa: BLE001
        out["ref_cuda"] = {"unavailable": repr(e)[:300]}
    del model, g

    # ---- configs[1]: SA1 + SA2 forward through the module API (last round's headline) ----
    try:
        out["sa_forward"] = sa_forward_leg(dev, host[0])
    except Exception as e:  # noqa: BLE001
        out["sa_forward"] = {"error": repr(e)[:200]}
    # ---- configs[4] sweep ----
    try:
        out["sweep"] = kn.sweep(dev, B_PER_GPU, quick=args.quick_sweep)
    except Exception as e:  # noqa: BLE001
        out["sweep"] = {"error": repr(e)[:200]}
    return out


def ref_cuda_leg(dev, ours_eval, inputs, ours_fwd_ms, ours_step_ms):
    import torch

    from benchmarks import kernels as kn
    from eda_b200 import hotpath
    from oracle import ref_hotpath, ref_loader

    ext = ref_loader.load_reference_ext()
    if ext is None:
        return {"unavailable": "oracle/_ref/pointnet2/_ext*.so did not travel"}
    ref = ref_hotpath.build(ext, dropout=0.0).to(dev)
    ref.load_state_dict(ours_eval.state_dict(), strict=True)
    res = {"what": "unmodified reference Python modules + the reference's own `_ext` compiled for sm_100a (oracle/_ref), "
                   "eager PyTorch 2.11, same parameters and inputs"}
    saved = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
    try:
        for label, tf32 in (("fp32", False), ("tf32_allowed", True)):
            torch.backends.cudnn.allow_tf32 = tf32
            torch.backends.cuda.matmul.allow_tf32 = tf32
            ref.eval()
            with torch.no_grad():
                fwd = kn.time_ms(lambda: ref(*inputs), 2, 5)
            ref.train()

            def train_step():
                for p in ref.parameters():
                    p.grad = None
                hotpath.quadratic_loss(ref(*inputs)).backward()

            trn = kn.time_ms(train_step, 2, 5)
            res[label] = {"forward_ms": fwd, "forward_scenes_per_s": B_PER_GPU / (fwd * 1e-3), "fwd_bwd_ms": trn,
                          "fwd_bwd_scenes_per_s": B_PER_GPU / (trn * 1e-3),
                          "speedup_forward": fwd / ours_fwd_ms, "speedup_fwd_bwd": trn / ours_step_ms}
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = saved
    return res


def sa_forward_leg(dev, pc_host):
    """configs[1]: SA1 -> SA2 forward, eval-mode BN, FPS chain on a side stream, chunk-pipelined."""
    import torch

    from benchmarks import kernels as kn
    from eda_b200.backbone_module import fps_chain
    from eda_b200.pointnet2 import fused
    from eda_b200.pointnet2.pointnet2_modules import PointnetSAModuleVotes

    torch.manual_seed(0)
    sa1 = PointnetSAModuleVotes(use_xyz=True, normalize_xyz=True, npoint=2048, radius=0.2, nsample=64,
                                mlp=[3, 64, 64, 128]).to(dev).eval()
    sa2 = PointnetSAModuleVotes(use_xyz=True, normalize_xyz=True, npoint=1024, radius=0.4, nsample=32,
                                mlp=[128, 128, 128, 256]).to(dev).eval()
    pc = pc_host.to(dev)
    side = torch.cuda.Stream(device=dev)

    def step():
        xyz = pc[..., :3].contiguous()
        feats = pc[..., 3:].transpose(1, 2).contiguous()
        (fps1, _), (fps2, _) = fps_chain(xyz, [2048, 1024], side, pipeline_every=(512, 1024))
        x1, f1, _, _ = fused.sa_forward_pipelined(sa1, xyz, feats, fps1)
        return fused.sa_forward_pipelined(sa2, x1, f1, fps2)

    with torch.no_grad():
        ms = kn.time_ms(step, 3, 20)
    return {"ms": ms, "value": B_PER_GPU / (ms * 1e-3), "unit": "scenes/s",
            "workload": "configs[1]: SA1(2048,r.2,ns64,[6,64,64,128]) -> SA2(1024,r.4,ns32,[131,128,128,256]) forward, "
                        "B=8, N=50000, eval-mode BN, device-resident; launched eagerly from Python (chunk-pipelined sampler, "
                        "not a graph): host-launch-bound, varies with the host's load"}


def main():
    args = parse()
    if args.impl == "reference":
        reference_arm(args)
    else:
        gpu_arm(args)


if __name__ == "__main__":
    main()
