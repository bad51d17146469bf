set -x
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q > gpurun_out/r1u_pytest_all.log 2>&1
echo "pytest exit $?" >> gpurun_out/r1u_pytest_all.log
timeout 300 python - > gpurun_out/r1u_solve_overhead.log 2>&1 <<'PY'
import time, numpy as np, torch
import b200stencil
from b200stencil import part2
for n in (1025, 2049):
    b = part2.to_device(np.random.default_rng(1).random((n, n)))
    hd = part2.MGHandle(n, n, part2.MGOpt())
    x = part2.zeros(n, n)
    hd.solve(x, b, 1.0/(n-1), 0.0, 1e-6, 100, False)
    for rep in range(3):
        x.zero_(); torch.cuda.synchronize()
        t0 = time.perf_counter(); r, nc = hd.solve(x, b, 1.0/(n-1), 0.0, 1e-6, 100, False); wall = time.perf_counter() - t0
        print(n, "cycles", nc, "wall_ms", wall*1e3, "device_ms", hd.stats()[1], "launches", hd.stats()[0])
    hd.close()
PY
true
