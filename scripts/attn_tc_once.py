import os, sys, torch
sys.path.insert(0, os.getcwd())
from eda_b200 import attn_ops as ops
dev = torch.device("cuda", 0)
B, E, H, Nq, Nk = 8, 288, 8, 1024, 1024
g = torch.Generator().manual_seed(0)
r = lambda *s: torch.randn(*s, generator=g).to(dev)
q, k = r(B * Nq, E), r(B * Nk, E)
vt = r(B, E, Nk); dctx = r(B * Nq, E)
lse = torch.empty(B, H, Nq, device=dev)
c = ops.attention_raw(q, k, vt, None, B, Nq, Nk, H, lse=lse)
for _ in range(2):
    ops.attention_backward_raw(q, k, vt, dctx, c, lse, None, B, Nq, Nk, H, impl="tc")
torch.cuda.synchronize(); print("ok")
