"""Per-kernel breakdown of one fused V-cycle (b2s_mg_profile_kernels): isolated average launch times vs the cycle time.
Usage: mg_kernel_breakdown.py [n ...]"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import b200stencil  # noqa
from b200stencil import part2
for n in [int(a) for a in sys.argv[1:]] or [1025]:
    b = part2.to_device(np.random.default_rng(1).random((n, n)))
    x = part2.zeros(n, n)
    hd = part2.MGHandle(n, n, part2.MGOpt())
    hd.cycles(x, b, 1.0 / (n - 1), 0.0, 1e-6, 200)
    _, ms = hd.cycles(x, b, 1.0 / (n - 1), 0.0, 1e-6, 100)
    ks = hd.profile_kernels(x, b, 1.0 / (n - 1), 0.0, reps=50)
    print(json.dumps({"n": n, "env": {k: v for k, v in os.environ.items() if k.startswith("B2S_MG")}, "ms_per_vcycle": ms / 100,
                      "sum_isolated_ms": sum(k["ms"] for k in ks),
                      "kernels": [[k["kernel"], k["level"], k["grid"], round(k["ms"] * 1e3, 2)] for k in ks]}))
    hd.close()
