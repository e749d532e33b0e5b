"""Loader for the reference's own compiled `pointnet2._ext` (TEST INFRASTRUCTURE).

oracle/build_ref.py compiles the unmodified reference sources where they lie under
/root/reference into oracle/_ref/pointnet2/_ext.<abi>.so; that file travels to the GPU box with the
gpurun snapshot.  It is CUDA-only (every op raises "CPU not supported" on CPU tensors).
"""
import glob
import importlib.util
import os

_HERE = os.path.dirname(os.path.abspath(__file__))


def reference_ext_path():
    hits = sorted(glob.glob(os.path.join(_HERE, "_ref", "pointnet2", "_ext*.so")))
    return hits[0] if hits else None


def load_reference_ext():
    import torch  # noqa: F401  (libtorch must be loaded before the extension)

    path = reference_ext_path()
    if path is None:
        return None
    spec = importlib.util.spec_from_file_location("_ext", path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod
