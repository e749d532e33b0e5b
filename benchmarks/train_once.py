"""Two eager training steps of the hot path at configs[3] shapes (for ncu: kernel launch list / per-kernel metrics of
the backward kernels; numbers printed under a profiler are never bench values)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from eda_b200 import ddp, hotpath  # noqa: E402

dev = torch.device("cuda", 0)
torch.cuda.set_device(dev)
torch.manual_seed(0)
model = hotpath.HotPath().to(dev).train()
fg = ddp.FlatGradients(model)
inputs = [t.to(dev) for t in hotpath.synthetic_inputs(8)]
for _ in range(int(sys.argv[1]) if len(sys.argv) > 1 else 2):
    fg.zero()
    hotpath.quadratic_loss(model(*inputs)).backward()
    fg.sync()
torch.cuda.synchronize()
print("ok")
