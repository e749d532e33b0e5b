"""Multi-process host logic of the data-parallel path on CPU: world_size 2, gloo backend."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from eda_b200 import ddp


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _model():
    torch.manual_seed(0)
    return torch.nn.Sequential(torch.nn.Linear(6, 16), torch.nn.ReLU(), torch.nn.Linear(16, 3))


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        m = _model()
        if rank != 0:
            for p in m.parameters():  # diverge, then prove the broadcast repairs it
                p.data.add_(1.0)
        ddp.broadcast_parameters(m, src=0)
        ref = _model()
        same = all(torch.equal(a, b) for a, b in zip(m.parameters(), ref.parameters()))

        g = torch.Generator().manual_seed(42)
        x, y = torch.randn(8, 6, generator=g), torch.randn(8, 3, generator=g)
        idx = list(ddp.shard_range(8, rank, world))
        fg = ddp.FlatGradients(m)
        loss = ((m(x[idx]) - y[idx]) ** 2).mean()
        loss.backward()
        views_ok = fg.check_views()
        fg.all_reduce_mean()
        # single-process gradient of the mean over the full batch == mean over ranks of the shard means
        full = _model()
        ((full(x) - y) ** 2).mean().backward()
        err = max((a.grad - b.grad).abs().max().item() for a, b in zip(m.parameters(), full.parameters()))
        h = fg.all_reduce_mean(async_op=True)  # async handle path (mean of identical grads = same grads)
        h.wait()
        err2 = max((a.grad - b.grad).abs().max().item() for a, b in zip(m.parameters(), full.parameters()))
        out.put((rank, same, views_ok, err, err2, idx))
    finally:
        dist.destroy_process_group()


def test_shard_range_partitions():
    for n in (0, 1, 7, 8, 64):
        for world in (1, 2, 3, 8):
            parts = [list(ddp.shard_range(n, r, world)) for r in range(world)]
            assert sum(parts, []) == list(range(n))
            assert max(len(p) for p in parts) - min(len(p) for p in parts) <= 1


def test_flat_gradients_single_process():
    m = _model()
    fg = ddp.FlatGradients(m)
    (m(torch.ones(2, 6)).sum()).backward()
    assert fg.check_views() and fg.flat.abs().sum() > 0
    assert fg.nbytes == 4 * sum(p.numel() for p in m.parameters())
    assert fg.all_reduce_mean() is None  # no process group: a no-op, not an error
    fg.zero()
    assert all(p.grad.abs().sum() == 0 for p in m.parameters())


@pytest.mark.timeout(120)
def test_two_ranks_gloo_broadcast_and_gradient_mean():
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    res = [out.get(timeout=100) for _ in procs]
    for p in procs:
        p.join(timeout=30)
        assert p.exitcode == 0
    shards = []
    for rank, same, views_ok, err, err2, idx in sorted(res):
        assert same, "broadcast_parameters did not equalise the ranks"
        assert views_ok
        assert err < 1e-6 and err2 < 1e-6
        shards += idx
    assert shards == list(range(8))


class _TwoBlocks(torch.nn.Module):
    """Two top-level children + a BatchNorm: stands in for backbone / decoder regions of the hot path."""

    def __init__(self):
        super().__init__()
        torch.manual_seed(0)
        self.front = torch.nn.Sequential(torch.nn.Linear(6, 16), torch.nn.BatchNorm1d(16), torch.nn.ReLU())
        self.back = torch.nn.Sequential(torch.nn.Linear(16, 8), torch.nn.ReLU(), torch.nn.Linear(8, 3))

    def forward(self, x):
        return self.back(self.front(x))


def _worker_overlap(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from eda_b200 import syncbn

        g = torch.Generator().manual_seed(7)
        x, y = torch.randn(8, 6, generator=g), torch.randn(8, 3, generator=g)
        idx = list(ddp.shard_range(8, rank, world))
        # ---- overlapped, bucketed all-reduce: regions in backward order, launched from the gradient hooks ----
        m = _TwoBlocks()
        fg = ddp.FlatGradients(m).enable_overlap()
        regions = list(fg.regions)
        errs = []
        for _ in range(2):  # twice: the per-step bookkeeping must re-arm
            fg.zero()
            ((m(x[idx]) - y[idx]) ** 2).mean().backward()
            launched = list(fg._launched)       # every region went out during / at the end of backward
            fg.all_reduce_mean()                # = finish()
            ref = _TwoBlocks()
            ga = []
            for r in range(world):              # mean over ranks of the per-shard gradients, computed locally
                ref.zero_grad()
                sh = list(ddp.shard_range(8, r, world))
                ((ref(x[sh]) - y[sh]) ** 2).mean().backward()
                ga.append(torch.cat([p.grad.flatten() for p in ref.parameters()]))
            want = torch.stack(ga).mean(0)
            errs.append((fg.flat - want).abs().max().item())
        # ---- synchronised BatchNorm host logic ----
        bn = m.front[1]
        none_before = syncbn.reducer_of(bn) is None          # plain BatchNorm: not synchronised
        ddp.convert_sync_batchnorm(m)
        red = syncbn.reducer_of(bn)
        t = torch.full((4,), float(rank + 1), dtype=torch.float64)
        red.all_reduce_sum_(t)
        total = red.total_count(10.0)
        bn.eval()
        none_eval = syncbn.reducer_of(bn) is None            # eval mode uses running statistics: nothing to sync
        converted = torch.nn.SyncBatchNorm.convert_sync_batchnorm(_TwoBlocks())
        sbn = converted.front[1]
        is_sync = isinstance(sbn, torch.nn.SyncBatchNorm) and syncbn.reducer_of(sbn) is not None
        unequal = False
        try:
            red.total_count(30.0 + rank)                      # ranks disagree on the row count -> refused
        except RuntimeError:
            unequal = True
        out.put((rank, regions, launched, errs, none_before, t.tolist(), total, none_eval, is_sync, unequal))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(120)
def test_two_ranks_gloo_overlapped_regions_and_syncbn_host_logic():
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker_overlap, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    res = [out.get(timeout=100) for _ in procs]
    for p in procs:
        p.join(timeout=30)
        assert p.exitcode == 0
    for rank, regions, launched, errs, none_before, t, total, none_eval, is_sync, unequal in sorted(res):
        assert regions == [(4, 8), (0, 4)]           # back (4 tensors) first, then front (linear w, b, bn w, b)
        assert launched == [True, True]
        assert max(errs) < 1e-6
        assert none_before and none_eval and is_sync and unequal
        assert t == [3.0] * 4 and total == 20.0


def test_overlap_is_a_noop_without_a_process_group():
    m = _TwoBlocks()
    fg = ddp.FlatGradients(m).enable_overlap()
    assert fg.regions is None
    m(torch.randn(4, 6)).sum().backward()
    fg.finish()
    assert fg.all_reduce_mean() is None
