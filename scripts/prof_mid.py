"""Phase stamps (clock64 of block 0's SM) of the cluster kernel / the collapsed coarse kernel:
B2S_MG_PROF=1 B2S_MG_CLUSTER=16 python scripts/prof_mid.py [n]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import b200stencil  # noqa
from b200stencil import part2
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1025
b = part2.to_device(np.random.default_rng(1).random((n, n)))
x = part2.zeros(n, n)
hd = part2.MGHandle(n, n, part2.MGOpt(use_graph=False))
for _ in range(5):
    hd.vcycle(x, b, 1.0 / (n - 1), 0.0, 1e-6, False)
print("coarse sweeps", hd.last_coarse_sweeps())
hd.close()
