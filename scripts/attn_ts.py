import ctypes, sys, os, torch
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
from eda_b200 import _lib, attn_ops as ops
lib = _lib.load()
B, H, E = 8, 8, 288
for Nq, Nk in ((1024, 1024), (256, 1024), (128, 256)):
    q = torch.randn(B * Nq, E, device="cuda"); k = torch.randn(B * Nk, E, device="cuda"); vt = torch.randn(B, E, Nk, device="cuda")
    for it in range(3):
        ops.attention_raw(q, k, vt, None, B, Nq, Nk, H)
    torch.cuda.synchronize()
    ts = (ctypes.c_longlong * 32)()
    lib.eda_debug_timestamps_attn(ts, 32)
    t = list(ts)
    names = ["issue+waitgroup", "fixup", "fence+sync", "QK mma", "tmem ld", "max", "sync", "exp+st", "O rescale+stwait", "sync", "PV mma"]
    print(f"Nq={Nq} Nk={Nk}: block total {t[11]-t[0]} cycles: " + ", ".join(f"{n} {t[i+1]-t[i]}" for i, n in enumerate(names)))
