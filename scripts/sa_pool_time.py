"""SA max-pool passes at the SA1 / SA2 shapes of the step: device time against the compulsory bytes."""
import ctypes, os, sys, torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from eda_b200 import _lib
from benchmarks.kernels import time_ms
lib = _lib.load()
vp = lambda t: ctypes.c_void_p(t.data_ptr())
dev = torch.device("cuda", 0)
for centres, S, C in ((16384, 64, 128), (8192, 32, 256)):
    R = centres * S
    z = torch.randn(R, C, device=dev)
    scale, shift = torch.rand(C, device=dev) + 0.5, torch.randn(C, device=dev)
    mean, invstd = torch.randn(C, device=dev) * 0.1, torch.rand(C, device=dev) + 0.5
    out = torch.empty(centres, C, device=dev); amax = torch.empty(centres, C, dtype=torch.int32, device=dev)
    gout = torch.randn(centres, C, device=dev); stats = torch.zeros(2 * C, device=dev)
    f = time_ms(lambda: lib.eda_sa_pool_forward(vp(z), vp(scale), vp(shift), centres, S, C, vp(out), vp(amax), None), 2, 10)
    zw = z.clone()
    b = time_ms(lambda: lib.eda_sa_pool_backward_apply(vp(zw), vp(amax), vp(gout), vp(scale), vp(mean), vp(invstd), vp(stats),
                                                       ctypes.c_double(float(R)), 1, centres, S, C, None), 2, 10)
    mb = R * C * 4 / 1e6
    print(f"centres {centres} S {S} C {C}: pool forward {f * 1e3:6.1f} us ({mb / f / 1e3:5.2f} TB/s over {mb:.0f} MB)   "
          f"pool backward apply {b * 1e3:6.1f} us ({2 * mb / b / 1e3:5.2f} TB/s over {2 * mb:.0f} MB)")
