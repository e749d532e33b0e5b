import os, sys, torch
sys.path.insert(0, "/root/repo")
from eda_b200 import attn_ops as ops
torch.manual_seed(0)
for (R, N, K, off) in ((256, 128, 32, 0),):
    dy = torch.randn(R, N).cuda(); x = torch.randn(R, K).cuda()
    buf = torch.zeros(N * (K + 4) + 8, device="cuda")
    dw = buf[off:off + N * (K + 4)].view(N, K + 4)[:, :K] if off else torch.zeros(N, K, device="cuda")
    db = torch.zeros(N, device="cuda")
    ops.wgrad([dict(dy=dy, x=x, dw=dw, db=db)], N, K)
    torch.cuda.synchronize()
    ref = dy.double().t() @ x.double()
    print(R, N, K, off, "nonzero", int((dw != 0).sum()), "of", dw.numel(), "rel", float((dw.double() - ref).norm() / ref.norm()),
          "db rel", float((db.double() - dy.double().sum(0)).norm() / dy.double().sum(0).norm()))
    print(dw[:2, :5].cpu().numpy().round(2)); print(ref[:2, :5].cpu().numpy().round(2))
