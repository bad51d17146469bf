"""Variant-B (red-black Gauss-Seidel + full weighting) V-cycle bench: fused tile kernels (every tile shape) vs the unfused
path. One JSON line per configuration. Usage: mgbench_variants.py [sizes...]"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import b200stencil  # noqa
from b200stencil import part2
sizes = tuple(int(a) for a in sys.argv[1:]) or (1025, 2049, 4097)
label = os.environ.get("B2S_LABEL", "")
fuse = int(os.environ.get("B2S_FUSE", "1"))
out = part2.bench_vcycle(sizes=sizes, opt=part2.MGOpt(smoother=1, restriction=1, fuse_sweeps=fuse))
brief = {n: {"ms_per_vcycle": round(v["ms_per_vcycle"], 4), "gdof_s": round(v["dof_per_s"] / 1e9, 2), "cycles": v["vcycles_to_1e-6"],
             "launches": v["kernel_launches_per_vcycle"]} for n, v in out["sizes"].items()}
print(json.dumps({"label": label, "fuse_sweeps": fuse, "rb_tile": os.environ.get("B2S_MG_RB_TILE", "0"), "sizes": brief}))
