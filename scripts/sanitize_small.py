"""Small end-to-end exercise of every kernel for compute-sanitizer (memcheck / racecheck / synccheck): tiny grids only."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import b200stencil  # noqa: F401
from b200stencil import capi, part1, part2

# path 1: both kernel variants, multi-slab pushes, solve loop
for kv, shape, nsl in ((capi.KERNEL_TMA, (64, 32, 18), 1), (capi.KERNEL_DIRECT, (33, 17, 9), 1), (capi.KERNEL_TMA, (64, 16, 10), 2),
                       (capi.KERNEL_DIRECT, (20, 12, 8), 3)):
    for halo in (0, 1):
        g = part1.Diffusion3D(*shape, nslabs=nsl, devices=[0] * nsl, halo_mode=halo, kernel_variant=kv)
        g.init_gaussian()
        g.iterate(5)
        g.solve_timestep(1e-3, 40)
        g.advance_time()
        g.gather()
        g.close()
# path 2: V-cycles in every configuration, L0 calls, CG, Navier-Stokes steps
rng = np.random.default_rng(0)
for shape in ((129, 129), (257, 65)):
    b = part2.to_device(rng.random(shape))
    for cfg in (dict(), dict(fuse_sweeps=2), dict(fuse_sweeps=0), dict(smoother=1, restriction=1), dict(coarse_solver=1),
                dict(smem_levels=False, use_graph=False)):
        for bcs in (False, True):
            x = part2.zeros(*shape)
            hd = part2.MGHandle(shape[0], shape[1], part2.MGOpt(**cfg))
            hd.solve(x, b, 1.0 / (min(shape) - 1), 3.0, 1e-6, 3, bcs)
            hd.close()
n = 66
bb = np.zeros((n, n)); bb[1:-1, 1:-1] = 1.0
part2.cg(part2.zeros(n, n), part2.to_device(bb), 1 / 65, 1 / 65, 3.14, 1e-6, 30)
for beta in (0.0, 0.5):
    sim = part2.NavierStokes2D(part2.SimIn_t(nx=129, ny=33, beta=beta, Pr=0.1, tol=1e-6, niters=5))
    sim.init_cosine("T"); sim.set_field("W", rng.random((129, 33)))
    sim.step(); sim.step()
    sim.close()
print("sanitize_small done")
