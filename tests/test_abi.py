"""The C-ABI boundary: include/eda_b200.h <-> libeda_b200.so <-> eda_b200._lib (no GPU needed:
nothing here launches a kernel)."""
import ctypes
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "eda_b200.h")).read()
    return sorted(set(re.findall(r"EDA_API[^;(]*?\b(eda_\w+)\s*\(", text)))


@pytest.fixture(scope="module")
def built_so():
    from eda_b200 import build

    return build.build()


def test_header_declares_something():
    syms = declared_symbols()
    assert "eda_furthest_point_sampling" in syms and "eda_ball_query" in syms and len(syms) >= 13


def test_library_exports_every_declared_symbol(built_so):
    lib = ctypes.CDLL(built_so)
    for s in declared_symbols():
        assert hasattr(lib, s), f"{s} declared in include/eda_b200.h but not exported"


def test_binding_covers_header(built_so):
    from eda_b200 import _lib

    assert sorted(_lib.PROTOTYPES) == declared_symbols()
    lib = _lib.load()
    assert lib.eda_version() >= 100
    assert lib.eda_error_string(0) == b"ok"
    assert b"invalid" in lib.eda_error_string(-1)


def test_no_torch_types_in_abi():
    text = open(os.path.join(ROOT, "include", "eda_b200.h")).read()
    code = re.sub(r"/\*.*?\*/", "", text, flags=re.S)  # comments cite the reference's at::Tensor signatures
    assert "at::" not in code and "torch" not in code.lower() and "#include <cuda" not in code


def test_argument_validation_without_gpu(built_so):
    # invalid arguments are rejected before any CUDA call, so this runs on a CPU-only box
    from eda_b200 import _lib

    lib = _lib.load()
    assert lib.eda_furthest_point_sampling(None, 1, 16, 4, None, None, None) == -1
    assert lib.eda_furthest_point_sampling(None, -1, 16, 4, None, None, None) == -1
    assert lib.eda_furthest_point_sampling(None, 0, 16, 4, None, None, None) == 0  # empty batch: nothing to do
    assert lib.eda_ball_query(None, None, 1, 8, 8, 0.2, 4, None, None) == -1
    assert lib.eda_gather_points(None, None, 1, 3, 8, 4, None, None) == -1
    assert lib.eda_three_nn(None, None, 1, 4, 4, None, None, None) == -1
    with pytest.raises(RuntimeError, match="invalid argument"):
        _lib.check(-1, "x")


def test_ext_surface_matches_reference_bindings():
    # pointnet2/_ext_src/src/bindings.cpp:11-24
    from eda_b200.pointnet2 import _ext

    names = {"furthest_point_sampling", "gather_points", "gather_points_grad", "ball_query", "group_points",
             "group_points_grad", "three_nn", "three_interpolate", "three_interpolate_grad"}
    assert names == set(_ext.__all__)
    for n in names:
        assert callable(getattr(_ext, n))


def test_ext_rejects_cpu_and_bad_tensors_like_the_reference():
    from eda_b200.pointnet2 import _ext

    x = torch.rand(1, 8, 3)
    with pytest.raises(RuntimeError, match="CPU not supported"):
        _ext.furthest_point_sampling(x, 4)
    with pytest.raises(RuntimeError, match="CPU not supported"):
        _ext.ball_query(x, x, 0.2, 4)
    with pytest.raises(RuntimeError, match="must be a float tensor"):
        _ext.furthest_point_sampling(x.double(), 4)
    with pytest.raises(RuntimeError, match="must be a contiguous tensor"):
        _ext.furthest_point_sampling(torch.rand(1, 3, 8).transpose(1, 2), 4)
    with pytest.raises(RuntimeError, match="must be an int tensor"):
        _ext.gather_points(torch.rand(1, 3, 8), torch.zeros(1, 4, dtype=torch.int64))


def test_product_package_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "eda_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), f"{f} imports oracle/"
                assert "liboracle" not in src
