"""Profiling driver (ncu): a few PT iterations of the 512^3 diffusion step. Usage: prof_diffusion.py [n] [iters] [variant]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import b200stencil  # noqa
from b200stencil import part1, capi
n = int(sys.argv[1]) if len(sys.argv) > 1 else 512
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 8
kv = {"auto": 0, "direct": 1, "tma": 2}[sys.argv[3] if len(sys.argv) > 3 else "auto"]
s = part1.Diffusion3D(n, n, n, kernel_variant=kv)
s.init_gaussian()
e = s.iterate(iters)
print("err", e[-1], "ms/iter", s.stats()[1] / iters)
s.close()
