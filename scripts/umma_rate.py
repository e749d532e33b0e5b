"""Tensor-pipe rate probe: cycles per kind::tf32 MMA (128 x N x 8) for operand sources / layouts / issue-loop styles /
number of other warps waiting on an mbarrier in the same CTA."""
import ctypes, json, os, sys
import torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from eda_b200 import _lib
lib = _lib.load()
out = {}
for iters in (4, 8, 2000):
    for ctas in (1, 148):
        cyc = torch.zeros(2 * ctas, dtype=torch.int64, device="cuda")
        for N in (64, 144, 256):
            for a_mode in (1, 2):
                for waiting in (0, 14, 28):
                    rc = lib.eda_selftest_umma_rate(N, a_mode, 0, 1, iters, ctas, waiting, ctypes.c_void_p(cyc.data_ptr()), None)
                    torch.cuda.synchronize()
                    if rc != 0:
                        continue
                    c = cyc.view(ctas, 2).double()
                    out[f"iters{iters}_ctas{ctas}_N{N}_a{a_mode}_wait{waiting}"] = [round(c[:, 0].mean().item() / iters, 1), round(c[:, 1].mean().item() / iters, 1)]
print("key: [issue cycles per MMA, completion cycles per MMA]; peak tf32 = 1934 MAC/clk/SM -> N=144: 76, N=256: 136 cycles")
for k, v in out.items():
    print(k, v)
