"""Tensor-pipe rate probe: cycles per kind::tf32 MMA (128 x N x 8) for operand sources / layouts; `fresh` = every MMA
reads new shared-memory addresses (a real K loop), `fixed` = the same descriptors every time, `cycled` = 4 K-steps of one
tile, `4 descriptors ahead` = fresh addresses with the descriptors of four MMAs in distinct registers before the four
issues (the form umma::mma4_tf32_ss_w gives the production kernels), `warp-converged` = whole warp, uniform datapath, one
MMA per statement.  Second table: pairs of MMAs sharing A into two accumulators (the linear kernel's N = 288 = 2 x 144 case)."""
import ctypes, os, sys
import torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from eda_b200 import _lib
lib = _lib.load()
ctas = 148
cyc = torch.zeros(2 * ctas, dtype=torch.int64, device="cuda")
def run(N, a_mode, b_swz, pre, iters):
    rc = lib.eda_selftest_umma_rate(N, a_mode, b_swz, pre, iters, ctas, 0, ctypes.c_void_p(cyc.data_ptr()), None)
    torch.cuda.synchronize()
    if rc != 0:
        return None
    c = cyc.view(ctas, 2).double()
    return c[:, 0].mean().item() / iters, c[:, 1].mean().item() / iters
iters = 2000
for N in (144, 256):
    for a_mode, an in ((1, "A tmem"), (2, "A smem swz128")):
        for b_swz in (0, 1):
            for pre, pn in ((1, "fixed"), (0, "cycled"), (2, "fresh"), (3, "fresh, 4 descriptors ahead"), (4, "fresh, warp-converged")):
                r = run(N, a_mode, b_swz, pre, iters)
                if r:
                    print(f"N {N:3d} {an:14s} B {'swz128' if b_swz else 'noswz '} {pn:6s}: issue {r[0]:7.1f}  complete {r[1]:7.1f} cycles/MMA")
for N in (128, 144):
    for acc_off in (0, 128, 144, 160, 256):
        for b_rows in (0, 2 * N):
            for it in (12, 2000):
                r = run(N, 2, 0, 2 | (acc_off << 8) | (b_rows << 20), it)
                if r:
                    print(f"pairs N {N:3d} acc_off {acc_off:3d} b_rows {b_rows or N:3d} iters {it:4d}: issue {r[0]:7.1f}  complete {r[1]:7.1f} cycles/MMA")
