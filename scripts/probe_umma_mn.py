"""Layout probe for MN-major B operands (kind::tf32): prints which shared-memory float the tensor core reads as B(n, k)
for the swizzle layout types (0 none, 1 = 128B with 32-byte atoms, 2 = 128B, 4 = 64B, 6 = 32B)."""
import ctypes, sys, os, torch
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
from eda_b200 import _lib
lib = _lib.load()
def probe(N, lbo, sbo, mn, layout, off):
    D = torch.full((8, N), -1.0, device="cuda")
    rc = lib.eda_selftest_umma_probe(N, lbo, sbo, mn, layout, off, ctypes.c_void_p(D.data_ptr()), None)
    _lib.check(rc, "probe"); torch.cuda.synchronize()
    return D.cpu().int()
for layout, lbo, sbo in ((1, 4096, 1024), (1, 1024, 4096), (0, 128, 1024), (2, 4096, 1024), (4, 4096, 1024), (6, 4096, 1024)):
    N = 64
    D = probe(N, lbo, sbo, 1, layout, 0)
    print(f"MN-major layout={layout} lbo={lbo} sbo={sbo}: rows k=0..7, cols n=0..{N-1}")
    for k in range(8):
        print("  k=%d:" % k, " ".join("%4d" % v for v in D[k].tolist()))
    if layout == 1:
        exp = [[(n // 32) * (lbo // 4) + (k // 4) * (sbo // 4) + (k % 4) * 32 + ((((n % 32) >> 3) ^ (k & 3)) * 8) + n % 8
                for n in range(N)] for k in range(8)]
        print("  matches: 32 MN elements per 128-byte row, 32-byte units XOR (k & 3), 4 k rows per group (SBO), next 32 MN at LBO:",
              D.tolist() == exp)
