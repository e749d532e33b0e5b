"""csrc/peer_reduce.cu on ONE GPU: `world` ranks are emulated by `world` exchange buffers and `world` CUDA streams of
the same device, each stream playing one rank (the kernels of the different "ranks" run concurrently and exchange
through the same store / flag / spin protocol they use across NVLink; only the symmetric-memory plumbing is absent — that
part is covered by tests/test_multi_gpu.py on 2 GPUs)."""
import ctypes

import pytest
import torch

pytestmark = pytest.mark.gpu


def _p(t):
    return ctypes.c_void_p(t.data_ptr())


def _setup(world, max_elems):
    from eda_b200 import _lib

    lib = _lib.load()
    nbytes = lib.eda_peer_buffer_bytes(world, max_elems)
    assert nbytes > 0
    bufs = [torch.zeros(nbytes, dtype=torch.uint8, device="cuda") for _ in range(world)]
    ptrs = torch.tensor([b.data_ptr() for b in bufs], dtype=torch.int64).cuda()
    streams = [torch.cuda.Stream() for _ in range(world)]
    torch.cuda.synchronize()
    return lib, bufs, ptrs, streams


@pytest.mark.parametrize("world", [1, 2, 4, 8])
@pytest.mark.parametrize("dtype", [torch.float32, torch.float64])
def test_peer_allreduce_emulated_ranks(world, dtype):
    max_elems = 1024
    lib, bufs, ptrs, streams = _setup(world, max_elems)
    g = torch.Generator().manual_seed(world)
    for n in (1, 7, 576, 1024, 576):  # repeated calls: epochs advance, both parities are used
        data = [torch.randn(n, generator=g, dtype=torch.float64).to(dtype).cuda() for _ in range(world)]
        want = torch.stack([d.double() for d in data]).sum(0)
        torch.cuda.synchronize()
        for r in range(world):
            rc = lib.eda_peer_allreduce(_p(ptrs), world, r, max_elems, _p(data[r]), n, 1 if dtype == torch.float64 else 0,
                                        ctypes.c_void_p(streams[r].cuda_stream))
            assert rc == 0
        torch.cuda.synchronize()
        for r in range(world):
            assert torch.equal(data[r], data[0]), "every rank must hold bit-identical sums"
            torch.testing.assert_close(data[r].double(), want, rtol=1e-6 if dtype == torch.float32 else 1e-14, atol=1e-6)
    for b in bufs:
        assert int(b[4:8].view(torch.int32).item()) == 0  # no peer ever timed out


def test_bn_finalize_peer_matches_finalize_of_summed_statistics():
    from eda_b200 import _lib

    world, C, max_elems = 4, 288, 1024
    lib, bufs, ptrs, streams = _setup(world, max_elems)
    g = torch.Generator().manual_seed(3)
    rows = 1000.0
    mean = torch.randn(C, generator=g, dtype=torch.float64)
    stats = []
    for r in range(world):  # per-rank [sum z, sum z^2] of `rows` rows each
        m = mean + 0.1 * torch.randn(C, generator=g, dtype=torch.float64)
        v = 0.5 + torch.rand(C, generator=g, dtype=torch.float64)
        stats.append(torch.cat([m * rows, (v + m * m) * rows]).cuda())
    gamma = (1 + 0.1 * torch.randn(C, generator=g)).cuda()
    beta = (0.1 * torch.randn(C, generator=g)).cuda()
    rm0, rv0 = torch.randn(C, generator=g).cuda(), (0.5 + torch.rand(C, generator=g)).cuda()
    st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    # reference: the plain kernel on the summed statistics
    tot = torch.stack(stats).sum(0)
    rm, rv = rm0.clone(), rv0.clone()
    sc, sh, mu, inv = (torch.empty(C, device="cuda") for _ in range(4))
    assert lib.eda_bn_finalize(_p(tot), rows * world, _p(gamma), _p(beta), 1e-5, 0.1, _p(rm), _p(rv), 1, C, _p(sc), _p(sh),
                               _p(mu), _p(inv), st) == 0
    outs = []
    torch.cuda.synchronize()
    for r in range(world):
        o = dict(rm=rm0.clone(), rv=rv0.clone(), **{k: torch.empty(C, device="cuda") for k in ("sc", "sh", "mu", "inv")})
        outs.append(o)
    torch.cuda.synchronize()
    for r in range(world):
        o = outs[r]
        rc = lib.eda_bn_finalize_peer(_p(ptrs), world, r, max_elems, _p(stats[r]), rows * world, _p(gamma), _p(beta), 1e-5,
                                      0.1, _p(o["rm"]), _p(o["rv"]), 1, C, _p(o["sc"]), _p(o["sh"]), _p(o["mu"]),
                                      _p(o["inv"]), ctypes.c_void_p(streams[r].cuda_stream))
        assert rc == 0
    torch.cuda.synchronize()
    for o in outs:
        for k, want in (("sc", sc), ("sh", sh), ("mu", mu), ("inv", inv), ("rm", rm), ("rv", rv)):
            assert torch.equal(o[k], outs[0][k])                       # identical on every rank
            torch.testing.assert_close(o[k], want, rtol=1e-6, atol=1e-6)  # and equal to the single-process result


def test_peer_timeout_sets_the_error_word_instead_of_hanging():
    """Only one of two "ranks" shows up: after the bounded wait its kernel finishes and reports through the error word."""
    world, max_elems = 2, 64
    lib, bufs, ptrs, streams = _setup(world, max_elems)
    data = torch.ones(8, device="cuda")
    rc = lib.eda_peer_allreduce(_p(ptrs), world, 0, max_elems, _p(data), 8, 0, ctypes.c_void_p(streams[0].cuda_stream))
    assert rc == 0
    torch.cuda.synchronize()  # returns after ~2 s
    assert int(bufs[0][4:8].view(torch.int32).item()) != 0
