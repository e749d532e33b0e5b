import ctypes, sys, os, torch
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
from eda_b200 import _lib
lib = _lib.load()
def probe(N, lbo, sbo, mn, layout, off):
    D = torch.full((8, N), -1.0, device="cuda")
    rc = lib.eda_selftest_umma_probe(N, lbo, sbo, mn, layout, off, ctypes.c_void_p(D.data_ptr()), None)
    _lib.check(rc, "probe"); torch.cuda.synchronize()
    return D.cpu().int()
for mn, lbo, sbo, layout, off in [(0, 16, 1024, 2, 0), (0, 16, 1024, 2, 32), (0, 16, 1024, 2, 96), (0, 1024, 1024, 2, 0)]:
    N = 16
    D = probe(N, lbo, sbo, mn, layout, off)
    print(f"mn={mn} lbo={lbo} sbo={sbo} layout={layout} off={off}: rows k=0..7, cols n=0..{N-1}")
    for k in range(8):
        print("  k=%d:" % k, " ".join("%4d" % v for v in D[k].tolist()))
    exp = [[n * 32 + ((((off // 16) + k // 4)) ^ (n & 7)) * 4 + k % 4 for n in range(N)] for k in range(8)]
    print("  matches n*32 + ((off/16 + k/4) ^ (n&7))*4 + k%4 :", D.tolist() == exp)
