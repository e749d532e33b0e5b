"""eda_b200 — B200-native (sm_100a) hot path of yanmin-wu/EDA behind the reference's module API.

Layout (only what the path needs):
  csrc/            hand-written CUDA kernels + the C ABI (include/eda_b200.h)
  lib/             built libeda_b200.so (git-ignored, in-tree so it travels to the GPU box)
  _lib.py          ctypes binding of the C ABI (fails loudly if the library is missing)
  pointnet2/       host-side mirror of the reference's pointnet2/ package:
                   _ext (9 ops), pointnet2_utils, pytorch_utils, pointnet2_modules
  backbone_module.py, encoder_decoder_layers.py   mirrors of the reference's models/ files on the path

Nothing in this package imports oracle/ (test infrastructure).
"""
__version__ = "0.1.0"
