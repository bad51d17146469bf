"""Profiling driver (ncu): a few V-cycles of the 2-D multigrid solve. Usage: prof_mg.py [n] [cycles] [graph] [a|b]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import b200stencil  # noqa
from b200stencil import part2
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1025
cycles = int(sys.argv[2]) if len(sys.argv) > 2 else 3
graph = (sys.argv[3] != "0") if len(sys.argv) > 3 else True
b = part2.to_device(np.random.default_rng(1).random((n, n)))
x = part2.zeros(n, n)
variant_b = len(sys.argv) > 4 and sys.argv[4] == "b"
hd = part2.MGHandle(n, n, part2.MGOpt(use_graph=graph, smoother=1 if variant_b else 0, restriction=1 if variant_b else 0))
r, ms = hd.cycles(x, b, 1.0 / (n - 1), 0.0, 1e-6, cycles)
print("r_rms", r, "ms/cycle", ms / max(1, cycles - (1 if graph else 0)))
hd.close()
