"""One fully converged time step of the 512^3 problem (config #3): iteration count, wall time, T_eff. The reference
remarks that 2^9 is 'very hard to converge' (part1_scaling_experiments.jl:31); here it takes seconds."""
import sys, os, json, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import b200stencil  # noqa
from b200stencil import part1, capi
n = int(sys.argv[1]) if len(sys.argv) > 1 else 512
tol = float(sys.argv[2]) if len(sys.argv) > 2 else 1e-8
out = {}
for name, kv in (("tma", capi.KERNEL_TMA), ("direct", capi.KERNEL_DIRECT)):
    s = part1.Diffusion3D(n, n, n, kernel_variant=kv)
    s.init_gaussian()
    t0 = time.perf_counter()
    it, err = s.solve_timestep(tol)
    wall = time.perf_counter() - t0
    ms = s.stats()[1]
    out[name] = {"n": n, "tol": tol, "iterations": it, "err": err, "device_ms": ms, "wall_s": wall,
                 "T_eff_GBs": 24.0 * (n - 2) ** 3 * it / (ms * 1e-3) / 1e9}
    s.close()
print(json.dumps(out))
