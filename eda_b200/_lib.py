"""ctypes binding of libeda_b200.so (C ABI: include/eda_b200.h).

The product path has NO fallback: if the library is missing or a call fails, a RuntimeError is
raised (the reference raises RuntimeError through TORCH_CHECK, pointnet2/_ext_src/include/utils.h:10-30,
and exit(-1)s on launch errors, cuda_utils.h:35-44 — here launch errors also become RuntimeError).
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
SO_PATH = os.path.join(_HERE, "lib", "libeda_b200.so")

_c_int, _c_float, _vp, _sz = ctypes.c_int, ctypes.c_float, ctypes.c_void_p, ctypes.c_size_t

# name -> (restype, argtypes); must list every prototype of include/eda_b200.h (tests check this)
PROTOTYPES = {
    "eda_version": (_c_int, []),
    "eda_error_string": (ctypes.c_char_p, [_c_int]),
    "eda_last_cuda_error": (ctypes.c_char_p, []),
    "eda_launch_count": (ctypes.c_ulonglong, []),
    "eda_fps_scratch_bytes": (_sz, [_c_int, _c_int, _c_int]),
    "eda_fps_identity_check": (_c_int, [_vp, _c_int, _c_int, _c_int, _vp, _vp, _vp]),
    "eda_furthest_point_sampling_ex": (_c_int, [_vp, _c_int, _c_int, _c_int, _vp, _vp, _vp, _c_int, _vp, _vp]),
    "eda_furthest_point_sampling": (_c_int, [_vp, _c_int, _c_int, _c_int, _vp, _vp, _vp]),
    "eda_furthest_point_sampling_progress": (_c_int, [_vp, _c_int, _c_int, _c_int, _vp, _vp, _vp, _c_int, _vp]),
    "eda_stream_wait_value32": (_c_int, [_vp, _vp, _c_int]),
    "eda_selftest_fps_exchange": (_c_int, [_c_int, _c_int, _c_int, _c_int, _vp, _vp]),
    "eda_fps_plan": (_c_int, [_c_int, _c_int, _c_int, _vp, _vp, _vp]),
    "eda_ball_query_range": (_c_int, [_vp, _vp, _c_int, _c_int, _c_int, _c_int, _c_int, _c_float, _c_int, _vp, _vp]),
    "eda_sa_mlp_forward_range": (_c_int, [_vp, _vp, _vp, _c_int, _vp, _vp, _vp, _vp, _vp, _c_int, _c_int, _c_int,
                                          _c_int, _c_int, _c_int, _c_int, _c_int, _c_int, _c_int, _c_float, _c_int,
                                          _vp, _vp]),
    "eda_ball_query": (_c_int, [_vp, _vp, _c_int, _c_int, _c_int, _c_float, _c_int, _vp, _vp]),
    "eda_group_points": (_c_int, [_vp, _vp, _c_int, _c_int, _c_int, _c_int, _c_int, _vp, _vp]),
    "eda_group_points_grad": (_c_int, [_vp, _vp, _c_int, _c_int, _c_int, _c_int, _c_int, _vp, _vp]),
    "eda_gather_points": (_c_int, [_vp, _vp, _c_int, _c_int, _c_int, _c_int, _vp, _vp]),
    "eda_gather_points_grad": (_c_int, [_vp, _vp, _c_int, _c_int, _c_int, _c_int, _vp, _vp]),
    "eda_three_nn": (_c_int, [_vp, _vp, _c_int, _c_int, _c_int, _vp, _vp, _vp]),
    "eda_three_interpolate": (_c_int, [_vp, _vp, _vp, _c_int, _c_int, _c_int, _c_int, _vp, _vp]),
    "eda_three_interpolate_grad": (_c_int, [_vp, _vp, _vp, _c_int, _c_int, _c_int, _c_int, _vp, _vp]),
    "eda_sa_mlp_packed_floats": (_sz, [_c_int, _c_int, _c_int, _c_int]),
    "eda_sa_mlp_pack": (_c_int, [_vp, _vp, _vp, _vp, _vp, _vp, _c_int, _c_int, _c_int, _c_int, _c_int, _vp, _vp]),
    "eda_sa_mlp_forward": (_c_int, [_vp, _vp, _vp, _c_int, _vp, _vp, _vp, _vp, _vp, _c_int, _c_int, _c_int, _c_int,
                                    _c_int, _c_int, _c_int, _c_int, _c_float, _c_int, _c_int, _vp, _vp, _vp]),
    "eda_bn_finalize": (_c_int, [_vp, ctypes.c_double, _vp, _vp, _c_float, _c_float, _vp, _vp, _c_int, _c_int, _vp,
                                 _vp, _vp, _vp, _vp]),
    "eda_transpose_last2": (_c_int, [_vp, _c_int, _c_int, _c_int, _vp, _vp]),
    "eda_transpose_strided": (_c_int, [_vp, _c_int, _c_int, _c_int, _c_int, _c_int, _vp, _vp]),
    "eda_linear_packed_floats": (_sz, [_c_int, _c_int]),
    "eda_linear_pack": (_c_int, [_vp, _vp, _c_int, _c_int, _vp, _vp]),
    "eda_linear_pack_batch": (_c_int, [_vp, _c_int, _c_int, _vp]),
    "eda_linear_forward": (_c_int, [_vp, _c_int, _c_int, _c_int, _c_int, _vp, _vp, _c_float, _c_int, _c_float,
                                    ctypes.c_uint, _vp, _vp]),
    "eda_dropout_mask": (_c_int, [ctypes.c_uint, _vp, _c_float, ctypes.c_longlong, _c_int, ctypes.c_uint, ctypes.c_uint,
                                  _vp, _vp]),
    "eda_dropout_apply": (_c_int, [_vp, ctypes.c_uint, _vp, _c_float, ctypes.c_longlong, _c_int, ctypes.c_uint, ctypes.c_uint,
                                   _vp, _vp]),
    "eda_debug_timestamps": (_c_int, [_vp, _c_int]),
    "eda_debug_timestamps_attn": (_c_int, [_vp, _c_int]),
    "eda_attention_forward": (_c_int, [_vp, _vp, _vp, _c_int, _vp, _c_int, _c_int, _c_int, _c_int, _c_int, _c_float,
                                       _c_float, ctypes.c_uint, _vp, _vp, _vp]),
    "eda_linear_pack_strided": (_c_int, [_vp, ctypes.c_longlong, ctypes.c_longlong, _c_int, _c_int, _vp, _vp]),
    "eda_attention_forward_lse": (_c_int, [_vp, _vp, _vp, _c_int, _vp, _c_int, _c_int, _c_int, _c_int, _c_int, _c_float,
                                           _c_float, ctypes.c_uint, _vp, _vp, _vp, _vp]),
    "eda_attention_backward": (_c_int, [_vp, _vp, _vp, ctypes.c_longlong, _vp, _c_int, _vp, _vp, _vp, _vp, _c_int, _c_int, _c_int,
                                        _c_int, _c_int, _c_float, _c_float, ctypes.c_uint, _vp, _vp, _vp, _vp, _vp, _vp]),
    "eda_attention_backward_tc": (_c_int, [_vp, _vp, _vp, ctypes.c_longlong, _vp, _c_int, _vp, _vp, _c_int, _vp, _vp, _vp,
                                           _vp, _c_int, _c_int, _c_int, _c_int, _c_int, _c_float, _c_float, ctypes.c_uint,
                                           _vp, _vp, _vp, _vp, _vp, _vp]),
    "eda_wgrad": (_c_int, [_vp, _c_int, _c_int, _c_int, _vp]),
    "eda_wgrad_small": (_c_int, [_vp, _c_int, _vp, _c_int, ctypes.c_longlong, _c_int, _c_int, _vp, _c_int, _vp, _vp]),
    "eda_fp_gather_rows": (_c_int, [_vp, _c_int, _vp, _c_int, _vp, _vp, _c_int, _c_int, _c_int, _c_int, _c_int, _vp, _vp,
                                    _vp]),
    "eda_fp_scatter_rows": (_c_int, [_vp, _c_int, _vp, _vp, _c_int, _c_int, _c_int, _c_int, _vp, _vp]),
    "eda_layernorm_backward": (_c_int, [_vp, _vp, _vp, _c_float, ctypes.c_longlong, _c_int, _vp, _vp, _vp, _vp, _c_float,
                                        ctypes.c_uint, _vp, _vp]),
    "eda_relu_backward": (_c_int, [_vp, _vp, _c_float, ctypes.c_longlong, _vp, _vp]),
    "eda_rows_gemm": (_c_int, [_vp, _c_int, _vp, _vp, _vp, ctypes.c_longlong, ctypes.c_longlong, ctypes.c_longlong, _c_int,
                               _c_int, _vp, _c_int, _vp]),
    "eda_rows_gemm_stats": (_c_int, [_vp, _c_int, _vp, _vp, _vp, ctypes.c_longlong, ctypes.c_longlong, ctypes.c_longlong,
                                     _c_int, _c_int, _vp, _c_int, _vp, _vp]),
    "eda_sa_gather_rows": (_c_int, [_vp, _vp, _vp, _c_int, _vp, _c_int, _c_int, _c_int, _c_int, _c_int, _c_int, _c_float,
                                    _c_int, _vp, _vp]),
    "eda_bn_relu_apply": (_c_int, [_vp, _vp, _vp, ctypes.c_longlong, _c_int, _vp, _vp]),
    "eda_sa_pool_backward": (_c_int, [_vp, _vp, _vp, _vp, _vp, _vp, ctypes.c_longlong, _c_int, _c_int, _vp, _vp, _vp]),
    "eda_sa_pool_backward_apply": (_c_int, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, ctypes.c_double, _c_int, ctypes.c_longlong,
                                            _c_int, _c_int, _vp]),
    "eda_bn_relu_backward_stats": (_c_int, [_vp, _vp, _vp, _vp, _vp, _vp, ctypes.c_longlong, _c_int, _vp, _vp]),
    "eda_bn_relu_backward_apply": (_c_int, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, ctypes.c_double, _c_int, ctypes.c_longlong,
                                            _c_int, _vp]),
    "eda_col_stats": (_c_int, [_vp, ctypes.c_longlong, _c_int, _vp, _vp]),
    "eda_sa_pool_forward": (_c_int, [_vp, _vp, _vp, ctypes.c_longlong, _c_int, _c_int, _vp, _vp, _vp]),
    "eda_sa_pool_backward_stats": (_c_int, [_vp, _vp, _vp, _vp, _vp, ctypes.c_longlong, _c_int, _c_int, _vp, _vp]),
    "eda_sa_scatter_rows": (_c_int, [_vp, _vp, _c_int, _c_int, _c_int, _c_int, _c_int, _c_int, _vp, _vp]),
    "eda_peer_buffer_bytes": (_sz, [_c_int, _c_int]),
    "eda_peer_allreduce": (_c_int, [_vp, _c_int, _c_int, _c_int, _vp, _c_int, _c_int, _vp]),
    "eda_bn_finalize_peer": (_c_int, [_vp, _c_int, _c_int, _c_int, _vp, ctypes.c_double, _vp, _vp, _c_float, _c_float,
                                      _vp, _vp, _c_int, _c_int, _vp, _vp, _vp, _vp, _vp]),
    "eda_selftest_umma": (_c_int, [_vp, _vp, _c_int, _c_int, _c_int, _vp, _vp]),
    "eda_selftest_umma_rate": (_c_int, [_c_int, _c_int, _c_int, _c_int, _c_int, _c_int, _c_int, _vp, _vp]),
    "eda_selftest_umma_probe": (_c_int, [_c_int, _c_int, _c_int, _c_int, _c_int, _c_int, _vp, _vp]),
}



class LinearProblem(ctypes.Structure):
    """struct eda_linear_problem (include/eda_b200.h)."""
    _fields_ = [("x", _vp), ("pos", _vp), ("w_packed", _vp), ("bias", _vp), ("residual", _vp), ("y", _vp),
                ("rows", _c_int), ("y_batch_rows", _c_int), ("y_ld", _c_int), ("round_tf32", _c_int), ("pre_ln", _vp),
                ("y_row_stride", _c_int), ("reserved", _c_int)]


class WgradProblem(ctypes.Structure):
    """struct eda_wgrad_problem (include/eda_b200.h)."""
    _fields_ = [("dy", _vp), ("x", _vp), ("dw", _vp), ("db", _vp), ("rows", ctypes.c_longlong),
                ("ldy", _c_int), ("ldx", _c_int), ("ldw", _c_int), ("x_scale", _vp), ("x_shift", _vp)]


_lib = None


def load():
    """Load the library once.  Raises RuntimeError (never falls back) when it is absent."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(SO_PATH):
        raise RuntimeError(
            f"eda_b200: {SO_PATH} not found — build it with `python -m eda_b200.build` "
            "(or __graft_entry__.build()); there is no CPU/PyTorch fallback for this path")
    lib = ctypes.CDLL(SO_PATH)
    for name, (res, args) in PROTOTYPES.items():
        fn = getattr(lib, name)  # AttributeError here = header/library mismatch: fail loudly
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc, what):
    if rc != 0:
        lib = load()
        msg = lib.eda_error_string(rc).decode()
        cuda = lib.eda_last_cuda_error().decode()
        raise RuntimeError(f"eda_b200.{what} failed: {msg}" + (f" [{cuda}]" if cuda else ""))
