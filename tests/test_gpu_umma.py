"""Hardware check of the tcgen05 conventions in eda_b200/csrc/umma.cuh (descriptor encodings,
chunk-major K-major smem layout, TMEM addressing, A-from-TMEM) through eda_selftest_umma."""
import ctypes

import pytest
import torch

pytestmark = pytest.mark.gpu


def _tf32_trunc(t):
    return (t.view(torch.int32) & ~0x1FFF).view(torch.float32)


@pytest.mark.parametrize("mode", [0, 1])
@pytest.mark.parametrize("N,K", [(16, 16), (64, 16), (128, 64), (256, 128), (48, 32), (208, 96)])
def test_umma_tf32_matches_cpu(mode, N, K):
    from eda_b200 import _lib

    lib = _lib.load()
    g = torch.Generator().manual_seed(N * 1000 + K + mode)
    A = torch.randn(128, K, generator=g)
    W = torch.randn(N, K, generator=g)
    # make rows/columns distinguishable so a transposed / permuted layout cannot pass by accident
    A += torch.arange(128)[:, None] * 0.01
    W += torch.arange(N)[:, None] * 0.02
    Ad, Wd = A.cuda(), W.cuda()
    D = torch.full((128, N), float("nan"), device="cuda")
    rc = lib.eda_selftest_umma(ctypes.c_void_p(Ad.data_ptr()), ctypes.c_void_p(Wd.data_ptr()), N, K, mode,
                               ctypes.c_void_p(D.data_ptr()), ctypes.c_void_p(torch.cuda.current_stream().cuda_stream))
    _lib.check(rc, "selftest_umma")
    torch.cuda.synchronize()
    D = D.cpu().double()
    exact = A.double() @ W.double().T
    bound = (A.abs().double() @ W.abs().double().T)  # sum |a||w|
    err = (D - exact).abs()
    assert torch.isfinite(D).all()
    assert (err <= 2.0 ** -9 * bound + 1e-6).all(), f"max err/bound = {(err / bound).max().item():.3e}"
    trunc = _tf32_trunc(A).double() @ _tf32_trunc(W).double().T
    print(f"mode={mode} N={N} K={K}: max|D-exact|/bound={(err / bound).max():.2e} "
          f"max|D-trunc|/bound={((D - trunc).abs() / bound).max():.2e}")
