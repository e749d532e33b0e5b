"""Multi-GPU checks of the data-parallel training path, run under torchrun by tests/test_multi_gpu.py:

    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 --master-port P tests/mp_train_check.py

Every rank holds its shard of ONE global batch.  With synchronised BatchNorm (main_utils.py:335-338) and the averaged
gradient all-reduce (main_utils.py:343-346), W ranks x B scenes must reproduce a single process on the concatenated
batch of W*B scenes: same outputs for the own scenes, same (averaged) gradients, same BatchNorm running statistics.
Checked for both statistic reducers (NCCL all-reduce; this package's NVLink peer-memory kernel), eagerly and as one
captured CUDA graph with the bucketed all-reduce overlapped with the backward pass.  Prints one JSON line on rank 0.
"""
import json
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from eda_b200 import ddp, hotpath, syncbn  # noqa: E402
from eda_b200.graphs import GraphedTrainStep  # noqa: E402


def rel(a, b):
    return ((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-30)).item()


def log(msg):
    print(f"[mpcheck rank {os.environ.get('RANK')}] {msg}", file=sys.stderr, flush=True)


def main():
    import faulthandler

    # a stuck collective must become a traceback and a non-zero exit, never a hung GPU box
    faulthandler.dump_traceback_later(int(os.environ.get("EDA_MPCHECK_TIMEOUT", "150")), exit=True)
    world = int(os.environ["WORLD_SIZE"])
    rank = int(os.environ["RANK"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    B, N = 2, 8192
    full = hotpath.synthetic_inputs(B * world, N, seed=7)
    mine = [t[rank * B:(rank + 1) * B].contiguous().to(dev) for t in full]
    res = {"world": world}

    def fresh():
        torch.manual_seed(0)
        return hotpath.HotPath(dropout=0.0).to(dev).train()

    # ---- single process on the concatenated batch (every rank computes it: the reference result) ----
    ref = fresh()
    out_ref = ref(*[t.to(dev) for t in full])
    hotpath.quadratic_loss(out_ref).backward()
    g_ref = torch.cat([p.grad.flatten() for p in ref.parameters()])
    bufs_ref = {k: v.clone() for k, v in ref.state_dict().items() if "running" in k}
    loss_full = hotpath.quadratic_loss(out_ref).item()

    def check(tag, model, fg, out, loss):
        res[tag + "_out_rel"] = max(rel(o, r[rank * B:(rank + 1) * B]) for o, r in zip(out, out_ref))
        res[tag + "_grad_rel"] = rel(fg.flat, g_ref)
        sd = model.state_dict()
        res[tag + "_running_rel"] = max(rel(sd[k], v) for k, v in bufs_ref.items())
        losses = torch.tensor([loss], device=dev)
        dist.all_reduce(losses)
        res[tag + "_loss_rel"] = abs(losses.item() / world - loss_full) / abs(loss_full)
        gsum = fg.flat.clone()
        dist.all_reduce(gsum, op=dist.ReduceOp.MAX)
        res[tag + "_grads_identical_across_ranks"] = bool(torch.equal(gsum, fg.flat))

    log("single-process reference done")
    # ---- 1. eager, NCCL reducer for the statistics, one flat all-reduce ----
    m = ddp.convert_sync_batchnorm(fresh())
    fg = ddp.FlatGradients(m)
    out = m(*mine)
    loss = hotpath.quadratic_loss(out)
    loss.backward()
    fg.all_reduce_mean()
    check("eager_nccl", m, fg, out, loss.item())

    log("eager nccl done: " + json.dumps(res))
    # ---- 2. eager, peer-memory reducer, overlapped bucketed all-reduce ----
    peer_ok = syncbn.enable_peer_reduce(dev)
    res["peer_reduce_available"] = bool(peer_ok)
    log(f"peer reduce available: {peer_ok}")
    if not peer_ok:
        res["peer_error"] = getattr(syncbn.default_reducer(), "peer_error", None)
    m = ddp.convert_sync_batchnorm(fresh())
    fg = ddp.FlatGradients(m).enable_overlap()
    res["regions"] = fg.regions
    fg.zero()
    out = m(*mine)
    loss = hotpath.quadratic_loss(out)
    loss.backward()
    res["regions_launched_in_backward"] = list(fg._launched)
    fg.all_reduce_mean()
    check("eager_peer_overlap", m, fg, out, loss.item())
    if peer_ok:
        res["peer_error_word"] = syncbn.default_reducer().peer.error_word()

    log("eager peer + overlap done: " + json.dumps(res))
    # ---- 3. the same as ONE CUDA graph (statistics exchange and gradient all-reduce captured) ----
    m = ddp.convert_sync_batchnorm(fresh())
    fg = ddp.FlatGradients(m).enable_overlap()
    step = GraphedTrainStep(m, hotpath.quadratic_loss, mine, fg)
    log("graph captured")
    loss = step(*mine)
    torch.cuda.synchronize()
    with torch.no_grad():
        m.eval()
        # running statistics after exactly one (replayed) step must equal the single-process ones
        m.train()
    res["graph_grad_rel"] = rel(fg.flat, g_ref)
    res["graph_running_rel"] = max(rel(m.state_dict()[k], v) for k, v in bufs_ref.items())
    losses = loss.clone().reshape(1)
    dist.all_reduce(losses)
    res["graph_loss_rel"] = abs(losses.item() / world - loss_full) / abs(loss_full)
    # timing: graphed step, SyncBN on, overlapped all-reduce
    for _ in range(3):
        step(*mine)
    torch.cuda.synchronize()
    dist.barrier()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(10):
        step(*mine)
    b.record()
    torch.cuda.synchronize()
    res["graph_ms_per_step_small"] = a.elapsed_time(b) / 10
    if peer_ok:
        res["peer_error_word_after_graph"] = syncbn.default_reducer().peer.error_word()

    # outputs: tf32-level agreement (per-rank vs whole-batch BatchNorm sums differ in summation order only, but the
    # tf32 GEMM operands amplify last-bit differences of the statistics); gradients of the AVERAGED loss; running stats
    ok = all(res[k] <= 5e-3 for k in res if k.endswith("_out_rel")) and \
        all(res[k] <= 2e-2 for k in res if k.endswith("_grad_rel")) and \
        all(res[k] <= 1e-3 for k in res if k.endswith("_running_rel")) and \
        all(res[k] <= 1e-4 for k in res if k.endswith("_loss_rel")) and \
        all(res[k] for k in res if k.endswith("identical_across_ranks"))
    res["ok"] = bool(ok)
    flag = torch.tensor([1 if ok else 0], device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        res["all_ranks_ok"] = bool(flag.item())
        print("MPCHECK " + json.dumps(res), flush=True)
    log("done")
    ok_all = bool(flag.item())
    # CUDA graphs that captured NCCL collectives must be gone before the communicator is torn down
    del step
    import gc
    gc.collect()
    torch.cuda.synchronize()
    dist.barrier()
    faulthandler.cancel_dump_traceback_later()
    sys.stdout.flush()
    os._exit(0 if ok_all else 1)  # skip the communicator teardown: nothing to save, and it may wait on captured work


if __name__ == "__main__":
    main()
