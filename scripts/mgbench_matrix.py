"""Config #2 matrix (multigrid_bench.jl sweeps coarse sizes and coarse solvers): per-V-cycle time and cycle count at n^2 for
coarse_solve_size in {5, 9}, coarse solver in {Jacobi, CG}, variant A (Jacobi + injection) and B (RB-GS + full weighting)."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import b200stencil  # noqa
from b200stencil import part2
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1025
for variant in ("A", "B"):
    for cs in (5, 9):
        for solver in (0, 1):
            opt = part2.MGOpt(coarse_solve_size=cs, coarse_solver=solver, smoother=1 if variant == "B" else 0,
                              restriction=1 if variant == "B" else 0)
            d = part2.bench_vcycle(sizes=(n,), opt=opt)["sizes"][str(n)]
            print(json.dumps({"n": n, "variant": variant, "coarse_solve_size": cs, "coarse_solver": ["jacobi", "cg"][solver],
                              "ms_per_vcycle": round(d["ms_per_vcycle"], 4), "vcycles": d["vcycles_to_1e-6"],
                              "solve_ms": round(d["solve_ms"], 3)}))
