import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box via gpurun)")


@pytest.fixture(scope="session")
def oracle():
    """The CPU oracle (test infrastructure).  Built on demand with gcc."""
    from oracle import pointnet2_oracle

    pointnet2_oracle.build()
    return pointnet2_oracle


@pytest.fixture(scope="session")
def ext():
    """The product op surface (CUDA).  Importing never falls back to anything."""
    import torch

    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from eda_b200.pointnet2 import _ext

    return _ext


@pytest.fixture(scope="session")
def ref_ext():
    """The reference's own compiled `_ext` (oracle/_ref, built by oracle/build_ref.py), if it travelled."""
    import torch

    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from oracle import ref_loader

    mod = ref_loader.load_reference_ext()
    if mod is None:
        pytest.skip("oracle/_ref/pointnet2/_ext*.so not present")
    return mod
