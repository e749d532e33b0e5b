"""cProfile of the HOST side of one eager training step (where the ~38 ms of Python / ctypes / autograd time go)."""
import cProfile
import os
import pstats
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


def step():
    fg.zero()
    hotpath.quadratic_loss(model(*inputs)).backward()
    fg.sync()


for _ in range(3):
    step()
torch.cuda.synchronize()
pr = cProfile.Profile()
pr.enable()
for _ in range(3):
    step()
pr.disable()
torch.cuda.synchronize()
st = pstats.Stats(pr)
st.sort_stats("tottime").print_stats(35)
