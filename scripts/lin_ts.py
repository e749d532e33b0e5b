import ctypes, sys, os, torch
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
from eda_b200 import _lib, attn_ops as ops
lib = _lib.load()
for R in (640, 8192):
    x = torch.randn(R, 288, device="cuda"); W = torch.randn(288, 288, device="cuda") / 17; b = torch.randn(288, device="cuda")
    pw = ops.pack_weight(W)
    for ln in (False, True):
        for it in range(3):
            ops.linear_raw([dict(x=x, w_packed=pw, bias=b, residual=x if ln else None)], 288, 288, ln=(b, b, 1e-5) if ln else None)
        torch.cuda.synchronize()
        ts = (ctypes.c_longlong * 128)()
        lib.eda_debug_timestamps(ts, 128)
        t = list(ts)
        print(f"R={R} ln={ln}: setup {t[1]-t[0]} staging_done {t[2]-t[1]} mma_done {t[3]-t[1]} epilogue: tmem_pass {t[9]-t[8]} barrier {t[10]-t[9]} copy_out {t[16]-t[10]} (ln: first-batch loads issued {t[13]-t[10]}, passA {t[11]-t[10]}, sync {t[12]-t[11]}, passB {t[16]-t[12]}) | epilogue_end {t[16]-t[1]} total {t[17]-t[0]} cycles")
        print("   per K block (cycles after setup): A staged / W landed / MMAs issued:", " ".join(f"[{t[32+3*k]-t[1]} {t[33+3*k]-t[1]} {t[34+3*k]-t[1]}]" for k in range(9)))
        print("   staging warp 0 per K block (after setup): top / issued / landed / fixed / fenced / arrived:", " ".join("[" + " ".join(str(t[64+6*k+j]-t[1]) for j in range(6)) + "]" for k in range(9)))
        print("   MMA thread, K block 3: fence->step0 %d, step1 +%d, step2 +%d, step3 +%d, commit +%d" % (t[121]-t[120], t[122]-t[121], t[123]-t[122], t[124]-t[123], t[125]-t[124]))
