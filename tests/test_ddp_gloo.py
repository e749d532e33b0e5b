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
