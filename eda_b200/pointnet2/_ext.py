"""`pointnet2._ext` op surface on top of libeda_b200.so.

Mirrors the 9 pybind functions of the reference (pointnet2/_ext_src/src/bindings.cpp:11-24):
same names, positional arguments, dtypes, shapes and error behaviour
(pointnet2/_ext_src/include/utils.h:10-30: non-contiguous / wrong dtype / CPU tensor -> RuntimeError),
so `pointnet2_utils.py` can bind to it unchanged.

Like the reference, every op launches on the CURRENT CUDA stream of the input's device, is
asynchronous with respect to the host, and returns freshly allocated outputs.  Unlike the
reference there is no hidden (B,N) scratch allocation in FPS (running minima live in registers)
and a failed launch raises RuntimeError instead of exit(-1) (cuda_utils.h:35-44).
"""
import ctypes

import torch

from .. import _lib

__all__ = [
    "furthest_point_sampling", "gather_points", "gather_points_grad", "ball_query", "group_points",
    "group_points_grad", "three_nn", "three_interpolate", "three_interpolate_grad",
]


def _contig(t, name):
    if not t.is_contiguous():
        raise RuntimeError(f"{name} must be a contiguous tensor")


def _is_float(t, name):
    if t.dtype != torch.float32:
        raise RuntimeError(f"{name} must be a float tensor")


def _is_int(t, name):
    if t.dtype != torch.int32:
        raise RuntimeError(f"{name} must be an int tensor")


def _cuda(t, name):
    if not t.is_cuda:
        raise RuntimeError(f"{name} must be a CUDA tensor")


def _require_cuda_primary(t):
    if not t.is_cuda:
        raise RuntimeError("CPU not supported")  # sampling.cpp:39,65,87 etc.


def _p(t):
    return ctypes.c_void_p(t.data_ptr())


def _stream(t):
    idx = t.device.index
    return ctypes.c_void_p(torch._C._cuda_getCurrentRawStream(idx if idx is not None else torch.cuda.current_device()))


def fps_identity_flags(points, nsamples):
    """(B,n,3) f32 -> (B,) i32 device flags: 0 where FPS(points[b], nsamples) is verified to be 0..nsamples-1 (the input
    is in FPS order and no step is decided by a tie), 1 where the full algorithm must run (eda_fps_identity_check)."""
    _contig(points, "points")
    _is_float(points, "points")
    _require_cuda_primary(points)
    lib = _lib.load()
    B, n = points.size(0), points.size(1)
    nsamples = int(nsamples)
    flags = torch.empty((B,), dtype=torch.int32, device=points.device)
    dsel = torch.empty((B, nsamples), dtype=torch.float32, device=points.device)
    with torch.cuda.device(points.device):
        rc = lib.eda_fps_identity_check(_p(points), B, n, nsamples, _p(dsel), _p(flags), _stream(points))
    _lib.check(rc, "fps_identity_check")
    return flags


def furthest_point_sampling(points, nsamples, not_identity=None):
    """(B,N,3) f32 -> (B,nsamples) i32.  sampling.cpp:70-91.  `not_identity`: optional flags from fps_identity_flags
    for the same (points, nsamples): verified scenes get 0..nsamples-1 without running the serial algorithm."""
    _contig(points, "points")
    _is_float(points, "points")
    _require_cuda_primary(points)
    lib = _lib.load()
    B, N = points.size(0), points.size(1)
    nsamples = int(nsamples)
    out = torch.empty((B, nsamples), dtype=torch.int32, device=points.device)
    nbytes = lib.eda_fps_scratch_bytes(B, N, nsamples)
    scratch = torch.empty((nbytes,), dtype=torch.uint8, device=points.device) if nbytes else None
    with torch.cuda.device(points.device):
        if not_identity is None:
            rc = lib.eda_furthest_point_sampling(_p(points), B, N, nsamples, _p(scratch) if nbytes else None, _p(out),
                                                 _stream(points))
        else:
            rc = lib.eda_furthest_point_sampling_ex(_p(points), B, N, nsamples, _p(scratch) if nbytes else None, _p(out),
                                                    None, 0, _p(not_identity), _stream(points))
    _lib.check(rc, "furthest_point_sampling")
    return out


def gather_points(points, idx):
    """(B,C,N) f32, (B,M) i32 -> (B,C,M).  sampling.cpp:20-44."""
    _contig(points, "points"); _contig(idx, "idx"); _is_float(points, "points"); _is_int(idx, "idx")
    if points.is_cuda:
        _cuda(idx, "idx")
    _require_cuda_primary(points)
    lib = _lib.load()
    B, C, N = points.shape
    M = idx.size(1)
    out = torch.empty((B, C, M), dtype=torch.float32, device=points.device)
    with torch.cuda.device(points.device):
        rc = lib.eda_gather_points(_p(points), _p(idx), B, C, N, M, _p(out), _stream(points))
    _lib.check(rc, "gather_points")
    return out


def gather_points_grad(grad_out, idx, n):
    """(B,C,M) f32, (B,M) i32 -> (B,C,n).  sampling.cpp:45-69."""
    _contig(grad_out, "grad_out"); _contig(idx, "idx"); _is_float(grad_out, "grad_out"); _is_int(idx, "idx")
    if grad_out.is_cuda:
        _cuda(idx, "idx")
    _require_cuda_primary(grad_out)
    lib = _lib.load()
    B, C, M = grad_out.shape
    out = torch.empty((B, C, int(n)), dtype=torch.float32, device=grad_out.device)
    with torch.cuda.device(grad_out.device):
        rc = lib.eda_gather_points_grad(_p(grad_out), _p(idx), B, C, int(n), M, _p(out), _stream(grad_out))
    _lib.check(rc, "gather_points_grad")
    return out


def ball_query(new_xyz, xyz, radius, nsample):
    """(B,M,3), (B,N,3) f32 -> (B,M,nsample) i32.  ball_query.cpp:13-37."""
    _contig(new_xyz, "new_xyz"); _contig(xyz, "xyz"); _is_float(new_xyz, "new_xyz"); _is_float(xyz, "xyz")
    if new_xyz.is_cuda:
        _cuda(xyz, "xyz")
    _require_cuda_primary(new_xyz)
    lib = _lib.load()
    B, M = new_xyz.size(0), new_xyz.size(1)
    N = xyz.size(1)
    nsample = int(nsample)
    idx = torch.empty((B, M, nsample), dtype=torch.int32, device=new_xyz.device)
    with torch.cuda.device(new_xyz.device):
        rc = lib.eda_ball_query(_p(new_xyz), _p(xyz), B, N, M, float(radius), nsample, _p(idx), _stream(new_xyz))
    _lib.check(rc, "ball_query")
    return idx


def group_points(points, idx):
    """(B,C,N) f32, (B,M,S) i32 -> (B,C,M,S).  group_points.cpp:17-41."""
    _contig(points, "points"); _contig(idx, "idx"); _is_float(points, "points"); _is_int(idx, "idx")
    if points.is_cuda:
        _cuda(idx, "idx")
    _require_cuda_primary(points)
    lib = _lib.load()
    B, C, N = points.shape
    M, S = idx.size(1), idx.size(2)
    out = torch.empty((B, C, M, S), dtype=torch.float32, device=points.device)
    with torch.cuda.device(points.device):
        rc = lib.eda_group_points(_p(points), _p(idx), B, C, N, M, S, _p(out), _stream(points))
    _lib.check(rc, "group_points")
    return out


def group_points_grad(grad_out, idx, n):
    """(B,C,M,S) f32, (B,M,S) i32 -> (B,C,n).  group_points.cpp:43-65."""
    _contig(grad_out, "grad_out"); _contig(idx, "idx"); _is_float(grad_out, "grad_out"); _is_int(idx, "idx")
    if grad_out.is_cuda:
        _cuda(idx, "idx")
    _require_cuda_primary(grad_out)
    lib = _lib.load()
    B, C, M, S = grad_out.shape
    out = torch.empty((B, C, int(n)), dtype=torch.float32, device=grad_out.device)
    with torch.cuda.device(grad_out.device):
        rc = lib.eda_group_points_grad(_p(grad_out), _p(idx), B, C, int(n), M, S, _p(out), _stream(grad_out))
    _lib.check(rc, "group_points_grad")
    return out


def three_nn(unknowns, knows):
    """(B,n,3), (B,m,3) f32 -> [dist2 (B,n,3) f32, idx (B,n,3) i32].  interpolate.cpp:19-46."""
    _contig(unknowns, "unknowns"); _contig(knows, "knows"); _is_float(unknowns, "unknowns"); _is_float(knows, "knows")
    if unknowns.is_cuda:
        _cuda(knows, "knows")
    _require_cuda_primary(unknowns)
    lib = _lib.load()
    B, n = unknowns.size(0), unknowns.size(1)
    m = knows.size(1)
    dist2 = torch.empty((B, n, 3), dtype=torch.float32, device=unknowns.device)
    idx = torch.empty((B, n, 3), dtype=torch.int32, device=unknowns.device)
    with torch.cuda.device(unknowns.device):
        rc = lib.eda_three_nn(_p(unknowns), _p(knows), B, n, m, _p(dist2), _p(idx), _stream(unknowns))
    _lib.check(rc, "three_nn")
    return [dist2, idx]


def three_interpolate(points, idx, weight):
    """(B,C,m) f32, (B,n,3) i32, (B,n,3) f32 -> (B,C,n).  interpolate.cpp:48-76."""
    _contig(points, "points"); _contig(idx, "idx"); _contig(weight, "weight")
    _is_float(points, "points"); _is_int(idx, "idx"); _is_float(weight, "weight")
    if points.is_cuda:
        _cuda(idx, "idx"); _cuda(weight, "weight")
    _require_cuda_primary(points)
    lib = _lib.load()
    B, C, m = points.shape
    n = idx.size(1)
    out = torch.empty((B, C, n), dtype=torch.float32, device=points.device)
    with torch.cuda.device(points.device):
        rc = lib.eda_three_interpolate(_p(points), _p(idx), _p(weight), B, C, m, n, _p(out), _stream(points))
    _lib.check(rc, "three_interpolate")
    return out


def three_interpolate_grad(grad_out, idx, weight, m):
    """(B,C,n) f32, (B,n,3) i32, (B,n,3) f32 -> (B,C,m).  interpolate.cpp:77-104."""
    _contig(grad_out, "grad_out"); _contig(idx, "idx"); _contig(weight, "weight")
    _is_float(grad_out, "grad_out"); _is_int(idx, "idx"); _is_float(weight, "weight")
    if grad_out.is_cuda:
        _cuda(idx, "idx"); _cuda(weight, "weight")
    _require_cuda_primary(grad_out)
    lib = _lib.load()
    B, C, n = grad_out.shape
    out = torch.empty((B, C, int(m)), dtype=torch.float32, device=grad_out.device)
    with torch.cuda.device(grad_out.device):
        rc = lib.eda_three_interpolate_grad(_p(grad_out), _p(idx), _p(weight), B, C, n, int(m), _p(out),
                                            _stream(grad_out))
    _lib.check(rc, "three_interpolate_grad")
    return out
