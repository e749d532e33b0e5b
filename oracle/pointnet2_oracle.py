"""ctypes front-end of oracle/pointnet2_oracle.c exposing the reference's `pointnet2._ext`
op surface (pointnet2/_ext_src/src/bindings.cpp:11-24) on CPU torch tensors.

TEST INFRASTRUCTURE: used as the checker in tests/ and as the CPU baseline in bench.py.
The reference itself has no CPU path for these ops ("CPU not supported",
pointnet2/_ext_src/src/sampling.cpp:87), so this is a *port*, pinned against the compiled
reference `_ext` on the GPU box (see tests/golden/README.md).
"""
import ctypes
import os
import subprocess
from concurrent.futures import ThreadPoolExecutor

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "liboracle.so")


def build(force=False):
    src = os.path.join(_HERE, "pointnet2_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.run(["make", "-s", "-C", _HERE, "-B", "liboracle.so"], check=True)
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = ctypes.CDLL(build())
    return _lib


def _f(t):
    assert t.dtype == torch.float32 and t.is_contiguous() and t.device.type == "cpu", "oracle: float32 contiguous CPU tensor expected"
    return ctypes.c_void_p(t.data_ptr())


def _i(t):
    assert t.dtype == torch.int32 and t.is_contiguous() and t.device.type == "cpu", "oracle: int32 contiguous CPU tensor expected"
    return ctypes.c_void_p(t.data_ptr())


def _over_batch(B, fn):
    """Run fn(b) for every scene; ctypes drops the GIL so host threads run scenes in parallel
    (the image's gcc has no libgomp, so the C file is single-threaded per call)."""
    nthr = min(B, os.cpu_count() or 1)
    if nthr <= 1:
        for b in range(B):
            fn(b)
        return
    with ThreadPoolExecutor(max_workers=nthr) as ex:
        list(ex.map(fn, range(B)))


def opt_n_threads(n):
    return lib().oracle_opt_n_threads(int(n))


def furthest_point_sampling(points, nsamples):
    B, N, _ = points.shape
    out = torch.zeros(B, nsamples, dtype=torch.int32)
    L = lib()
    _over_batch(B, lambda b: L.oracle_furthest_point_sampling(1, N, int(nsamples), _f(points[b]), _i(out[b])))
    return out


def fps_identity_verified(points, nsamples):
    """(B,n,3) -> (B,) int32: 1 where FPS(points[b], nsamples) is provably 0..nsamples-1 with no tie-break involved
    (the criterion behind eda_fps_identity_check; a property of the reference algorithm, see the C file)."""
    B, n, _ = points.shape
    L = lib()
    L.oracle_fps_identity_verified.restype = ctypes.c_int
    out = torch.zeros(B, dtype=torch.int32)
    for b in range(B):
        out[b] = L.oracle_fps_identity_verified(int(n), int(nsamples), _f(points[b].contiguous()))
    return out


def ball_query(new_xyz, xyz, radius, nsample):
    B, M, _ = new_xyz.shape
    N = xyz.shape[1]
    idx = torch.zeros(B, M, nsample, dtype=torch.int32)
    L = lib()
    _over_batch(B, lambda b: L.oracle_ball_query(1, N, M, ctypes.c_float(radius), int(nsample), _f(new_xyz[b]), _f(xyz[b]), _i(idx[b])))
    return idx


def group_points(points, idx):
    B, C, N = points.shape
    _, M, S = idx.shape
    out = torch.zeros(B, C, M, S, dtype=torch.float32)
    lib().oracle_group_points(B, C, N, M, S, _f(points), _i(idx), _f(out))
    return out


def group_points_grad(grad_out, idx, n):
    B, C, M, S = grad_out.shape
    out = torch.zeros(B, C, n, dtype=torch.float32)
    lib().oracle_group_points_grad(B, C, int(n), M, S, _f(grad_out), _i(idx), _f(out))
    return out


def gather_points(points, idx):
    B, C, N = points.shape
    M = idx.shape[1]
    out = torch.zeros(B, C, M, dtype=torch.float32)
    lib().oracle_gather_points(B, C, N, M, _f(points), _i(idx), _f(out))
    return out


def gather_points_grad(grad_out, idx, n):
    B, C, M = grad_out.shape
    out = torch.zeros(B, C, n, dtype=torch.float32)
    lib().oracle_gather_points_grad(B, C, int(n), M, _f(grad_out), _i(idx), _f(out))
    return out


def three_nn(unknown, known):
    B, N, _ = unknown.shape
    M = known.shape[1]
    dist2 = torch.zeros(B, N, 3, dtype=torch.float32)
    idx = torch.zeros(B, N, 3, dtype=torch.int32)
    lib().oracle_three_nn(B, N, M, _f(unknown), _f(known), _f(dist2), _i(idx))
    return [dist2, idx]


def three_interpolate(points, idx, weight):
    B, C, M = points.shape
    N = idx.shape[1]
    out = torch.zeros(B, C, N, dtype=torch.float32)
    lib().oracle_three_interpolate(B, C, M, N, _f(points), _i(idx), _f(weight), _f(out))
    return out


def three_interpolate_grad(grad_out, idx, weight, m):
    B, C, N = grad_out.shape
    out = torch.zeros(B, C, m, dtype=torch.float32)
    lib().oracle_three_interpolate_grad(B, C, N, int(m), _f(grad_out), _i(idx), _f(weight), _f(out))
    return out
