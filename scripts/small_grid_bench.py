"""PT-iteration time of L2-resident grids (configs[0] 32^3 and the reference's 128^3 benchmark shape) with and without the
CUDA-graph batches (B2S_DIFF_GRAPH=0/1 in the environment). One JSON line."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import b200stencil  # noqa
from b200stencil import part1
out = {"graph": os.environ.get("B2S_DIFF_GRAPH", "auto")}
for n in (32, 64, 128, 192):
    g = part1.Diffusion3D(n, n, n)
    g.init_gaussian()
    g.iterate(1024, want_hist=False)
    ms = 0.0
    for _ in range(5):
        g.iterate(2048, want_hist=False)
        ms += g.stats()[1]
    out[str(n)] = {"us_per_iteration": ms / (5 * 2048) * 1e3, "T_eff_GBs": 24.0 * (n - 2.0) ** 3 / (ms / (5 * 2048) * 1e-3) / 1e9}
    g.close()
print(json.dumps(out))
