"""Variant-A (damped Jacobi + injection) V-cycle bench, one JSON line: ms per V-cycle at the given sizes under the
current environment switches (B2S_MG_CLUSTER, B2S_MG_TILE, ...). Usage: mgbench_a.py [sizes...]"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import b200stencil  # noqa
from b200stencil import part2
sizes = tuple(int(a) for a in sys.argv[1:]) or (1025, 2049, 4097)
out = part2.bench_vcycle(sizes=sizes, opt=part2.MGOpt(), e2e=False)
brief = {n: {"ms_per_vcycle": round(v["ms_per_vcycle"], 5), "gdof_s": round(v["dof_per_s"] / 1e9, 2), "cycles": v["vcycles_to_1e-6"],
             "solve_ms": round(v["solve_ms"], 4), "launches": v["kernel_launches_per_vcycle"]} for n, v in out["sizes"].items()}
print(json.dumps({"label": os.environ.get("B2S_LABEL", ""), "env": {k: v for k, v in os.environ.items() if k.startswith("B2S_MG")},
                  "sizes": brief}))
