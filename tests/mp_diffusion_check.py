"""One process per GPU (torchrun): z-slab diffusion with the fused NVLink halo push + peer-store norm exchange, checked
against the oracle's rank emulation. Every rank verifies its own slab bit-for-bit. Usage (2+ GPUs):
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
      tests/mp_diffusion_check.py [nx ny nz] [halo_mode] [auto|direct|tma] [dimx dimy dimz]
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist

import b200stencil  # noqa: F401
from b200stencil import capi, part1, dist as D
from oracle import oracle_lib as O  # checker only

rank, world, local = D.env_rank_world()
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
shape = tuple(int(a) for a in sys.argv[1:4]) if len(sys.argv) >= 4 else (64, 64, 34)
halo = int(sys.argv[4]) if len(sys.argv) >= 5 else 0
variant = {"auto": 0, "direct": 1, "tma": 2}[sys.argv[5]] if len(sys.argv) >= 6 else 0
dims = tuple(int(a) for a in sys.argv[6:9]) if len(sys.argv) >= 9 else (1, 1, world)  # general decomposition (optional)
assert dims[0] * dims[1] * dims[2] == world, (dims, world)
nsl, begin, count = D.slab_layout(rank, world)
g = part1.Diffusion3D(*shape, nslabs=nsl, devices=[local], slab_begin=begin, slab_count=count, halo_mode=halo,
                      scale_physical_size=True, kernel_variant=variant, dims=dims if dims[0] * dims[1] > 1 else None)
g.init_gaussian()
D.connect(g, dist)
o = O.Diffusion3D(*shape, dims=dims, halo_mode=halo, scale_physical_size=True)
assert np.array_equal(g.get("Ht"), o.get("Ht", rank))
done = 0
for chunk in (1, 1, 1, 2, 20):
    eo = o.iterate(chunk)
    eg = g.iterate(chunk)
    done += chunk
    assert np.allclose(eg, eo, rtol=1e-12, atol=0), (rank, done, eg, eo)
    assert np.array_equal(g.get("Htau"), o.get("Htau", rank)), (rank, done)
it_o, err_o = o.solve_timestep(1e-5)
it_g, err_g = g.solve_timestep(1e-5)
assert it_g == it_o, (rank, it_g, it_o)
o.advance_time(); g.advance_time()
assert np.array_equal(g.get("Ht"), o.get("Ht", rank))
H = D.gather_global(g, dist)
if rank == 0:
    assert np.array_equal(H, o.gather())
# all ranks took the same decision on the same iteration
its = [None] * world
dist.all_gather_object(its, (it_g, err_g))
assert all(v == its[0] for v in its), its
dist.barrier()
g.close()
dist.destroy_process_group()
os.write(1, ("rank%d ok iters=%d\n" % (rank, it_g)).encode())
