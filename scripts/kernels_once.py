"""Runs each roofline kernel of bench.py a few times (for ncu captures; numbers printed under a profiler are never
bench values):  python scripts/kernels_once.py [linear|wgrad|fps|ballq|attention|sa_mlp|all]"""
import os, sys
import torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from benchmarks import kernels as kn

which = sys.argv[1] if len(sys.argv) > 1 else "all"
dev = torch.device("cuda", 0)
_, _, _, sm_mhz, _ = kn.peaks()
if which in ("linear", "all"):
    print(kn.linear_point(8192, 288, 288, dev))
    print(kn.linear_point(8192, 288, 288, dev, ln=True))
if which in ("wgrad", "all"):
    print(kn.wgrad_point(8192, 288, 288, 3, dev))
if which in ("fps", "all"):
    print(kn.fps_point(8, 50000, 2048, dev, sm_mhz, with_floor=False))
if which in ("ballq", "all"):
    print(kn.ball_query_point(8, 50000, 2048, 0.2, 64, dev))
if which in ("attention", "all"):
    print(kn.attention_point(8, 1024, 1024, dev))
    print(kn.attention_point(8, 256, 1024, dev))
if which in ("sa_mlp", "all"):
    print(kn.sa_mlp_point(8, 50000, 2048, 64, 3, [64, 64, 128], 0.2, dev))
