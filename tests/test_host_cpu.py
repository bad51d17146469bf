"""CPU-side checks of the product's host logic (no GPU, no compute calls): the C-ABI library loads and exports every
symbol include/b200stencil.h declares, fails loudly without a device, its host-side geometry matches the oracle, and
the one-process-per-GPU plumbing works across 2 gloo ranks."""
import ctypes as C
import os
import subprocess
import sys
import textwrap

import numpy as np
import pytest

from conftest import ROOT


def test_library_exports_every_declared_symbol(b2s):
    from b200stencil import capi
    declared = capi.header_symbols()
    assert len(declared) >= 50
    assert capi.missing_symbols() == []
    assert sorted(capi.SIGNATURES) == declared, "ctypes table and header disagree"
    assert capi.lib().b2s_version() == 100


def test_header_is_plain_c_and_matches_the_ctypes_mirror(b2s, tmp_path):
    """include/b200stencil.h must compile as C (what cgo / ccall / ctypes bind against) and its structs must have the
    layout the Python mirror assumes (sizes and the offsets of the fields appended last)."""
    import ctypes as C
    import subprocess
    from b200stencil import capi
    src = tmp_path / "abi.c"
    src.write_text(r'''
#include <stdio.h>
#include <stddef.h>
#include "b200stencil.h"
int main(void) {
  printf("%zu %zu %zu %zu %zu %zu %zu %zu\n", sizeof(b2s_diff3d_config), offsetof(b2s_diff3d_config, dimx),
         offsetof(b2s_diff3d_config, dimy), sizeof(b2s_diff3d_params), sizeof(b2s_mg_config), offsetof(b2s_mg_config, fuse_sweeps),
         sizeof(b2s_ns2d_params), sizeof(b2s_ns2d_stepinfo));
  return 0;
}
''')
    exe = tmp_path / "abi"
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), "-o", str(exe), str(src)], check=True)
    got = [int(v) for v in subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.split()]
    want = [C.sizeof(capi.Diff3DConfig), capi.Diff3DConfig.dimx.offset, capi.Diff3DConfig.dimy.offset, C.sizeof(capi.Diff3DParams),
            C.sizeof(capi.MGConfig), capi.MGConfig.fuse_sweeps.offset, C.sizeof(capi.NS2DParams), C.sizeof(capi.NS2DStepInfo)]
    assert got == want, (got, want)


def test_no_cpu_fallback(b2s):
    from b200stencil import capi, part1
    if capi.device_count() > 0:
        pytest.skip("a GPU is present")
    with pytest.raises(capi.B2SError) as e:
        part1.Diffusion3D(32, 32, 32)
    assert e.value.code == capi.ERR_NO_DEVICE
    cfg = capi.MGConfig(129, 129, 5, 0, 0, 0, 0, 1, 1, 1)
    h = C.c_void_p()
    assert capi.lib().b2s_mg_create(C.byref(h), C.byref(cfg)) == capi.ERR_NO_DEVICE


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "finalprojectrepo.jl_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".jl", ".h")):
                txt = open(os.path.join(dirpath, f), errors="replace").read()
                assert "oracle_lib" not in txt and "liboracle" not in txt and "import oracle" not in txt, f


@pytest.mark.parametrize("n,nslabs,scale", [((32, 32, 32), 1, False), ((128, 128, 64), 2, False), ((512, 512, 512), 8, True),
                                            ((64, 48, 34), 3, True)])
def test_host_geometry_matches_oracle(b2s, oracle, n, nslabs, scale):
    """dx, dy, dz, dtau, total_N of part1_kernel_programming.jl:104-131 -- bit-identical to the oracle's."""
    from b200stencil import capi
    cfg = capi.Diff3DConfig(n[0], n[1], n[2], nslabs, 0, nslabs, None, 0, 0, int(scale), 0, 0)
    p = capi.Diff3DParams()
    capi.check(capi.lib().b2s_diff3d_params_for(C.byref(cfg), C.byref(p)))
    small = tuple(min(v, 8) for v in n)  # geometry only depends on n through the formulas; keep the oracle tiny
    o = oracle.Diffusion3D(*n, dims=(1, 1, nslabs), scale_physical_size=scale) if np.prod(n) * nslabs < 3e6 else None
    if o is not None:
        assert (p.dx, p.dy, p.dz, p.dt, p.dtau, p.lx, p.ly, p.lz) == (o.dx, o.dy, o.dz, o.dt, o.dtau, o.lx, o.ly, o.lz)
    nzg = nslabs * (n[2] - 2) + 2
    lz = 10.0 * nslabs if scale else 10.0
    assert p.nz_g == nzg and p.dz == lz / nzg and p.total_N == float(nslabs) * n[0] * n[1] * n[2]
    assert p.dtau == min(p.dx, p.dy, p.dz) ** 2 / 1.0 / 8.1
    del small


@pytest.mark.parametrize("n,dims,scale", [((16, 12, 10), (2, 2, 1), False), ((16, 12, 10), (2, 2, 2), True),
                                          ((10, 10, 18), (3, 1, 2), True), ((64, 64, 128), (2, 2, 1), False)])
def test_host_geometry_general_decomposition(b2s, oracle, n, dims, scale):
    """The same geometry for ImplicitGlobalGrid's general rank grids dims = (dimx, dimy, dimz) (2x2x1, 2x2x2 of the
    published scaling runs, part1_scaling_experiments.jl): bit-identical to the oracle's rank emulation."""
    from b200stencil import capi
    nr = dims[0] * dims[1] * dims[2]
    cfg = capi.Diff3DConfig(n[0], n[1], n[2], nr, 0, nr, None, 0, 0, int(scale), 0, 0, dims[0], dims[1])
    p = capi.Diff3DParams()
    capi.check(capi.lib().b2s_diff3d_params_for(C.byref(cfg), C.byref(p)))
    if np.prod(n) * nr < 3e6:
        o = oracle.Diffusion3D(*n, dims=dims, scale_physical_size=scale)
        assert (p.dx, p.dy, p.dz, p.dt, p.dtau, p.lx, p.ly, p.lz) == (o.dx, o.dy, o.dz, o.dt, o.dtau, o.lx, o.ly, o.lz)
    assert (p.nx_g, p.ny_g, p.nz_g) == tuple(d * (m - 2) + 2 for d, m in zip(dims, n))
    assert p.total_N == float(nr) * n[0] * n[1] * n[2]
    bad = capi.Diff3DConfig(16, 16, 16, 6, 0, 6, None, 0, 0, 0, 0, 0, 4, 1)  # 4 does not divide 6 ranks
    h = C.c_void_p()
    assert capi.lib().b2s_diff3d_create(C.byref(h), C.byref(bad)) != 0


def test_mg_algorithmic_bytes(b2s):
    from b200stencil import part2
    # 132 B x 1,402,168 points = 185.1 MB per V-cycle at 1025^2
    assert part2.mg_algorithmic_bytes(1025, 1025) == 132.0 * 1402168
    assert part2.mg_algorithmic_bytes(5, 5) == 0.0


def test_slab_layout_helpers(b2s):
    from b200stencil import dist as D
    assert D.slab_layout(3, 8) == (8, 3, 1)
    assert D.global_nz(512, 8) == 8 * 510 + 2 and D.z_offset(2, 512) == 1020
    with pytest.raises(ValueError):
        D.slab_layout(2, 2)
    assert D.all_gather_blobs(b"x") == [b"x"] and D.max_over_ranks(3.5) == 3.5


WORKER = textwrap.dedent("""
    import os, sys
    sys.path.insert(0, {root!r})
    import torch.distributed as dist
    import b200stencil
    from b200stencil import dist as D, capi
    import ctypes as C
    dist.init_process_group("gloo")
    rank, world, _ = D.env_rank_world()
    assert (rank, world) == (dist.get_rank(), dist.get_world_size())
    nslabs, begin, count = D.slab_layout(rank, world)
    n = capi.lib().b2s_diff3d_ipc_blob_bytes()
    blob = bytes([rank]) * n
    blobs = D.all_gather_blobs(blob, dist)
    assert [b[0] for b in blobs] == list(range(world)) and all(len(b) == n for b in blobs)
    # every rank derives identical global numerics from the same config (lock-step exit decisions rely on it)
    cfg = capi.Diff3DConfig(64, 64, 34, nslabs, begin, count, None, 0, 0, 1, 0, 0)
    p = capi.Diff3DParams()
    capi.check(capi.lib().b2s_diff3d_params_for(C.byref(cfg), C.byref(p)))
    vals = [None] * world
    dist.all_gather_object(vals, (p.dz, p.dtau, p.total_N, p.nz_g))
    assert all(v == vals[0] for v in vals), vals
    assert D.max_over_ranks(10.0 + rank, dist) == 10.0 + world - 1
    # gather!-equivalent across processes: z-slabs are concatenated, a general decomposition (here dims = (2,1,1): the
    # ranks split x) is assembled from the blocks every rank placed into its own copy of the global array
    import numpy as np

    class Fake:
        def __init__(self, dims):
            self.dims = dims
        def gather(self):
            if self.dims[0] * self.dims[1] == 1:
                return np.full((3, 2, 4), float(rank + 1), order="F")
            a = np.zeros((3 * self.dims[0], 2, 4), order="F")
            a[3 * rank:3 * rank + 3] = rank + 1
            return a
    Hz = D.gather_global(Fake((1, 1, world)), dist)
    Hx = D.gather_global(Fake((world, 1, 1)), dist)
    if rank == 0:
        assert Hz.shape == (3, 2, 4 * world) and all((Hz[:, :, 4 * r:4 * r + 4] == r + 1).all() for r in range(world))
        assert Hx.shape == (3 * world, 2, 4) and all((Hx[3 * r:3 * r + 3] == r + 1).all() for r in range(world))
    else:
        assert Hz is None and Hx is None
    # the same geometry on every rank for a general decomposition, too
    cfg2 = capi.Diff3DConfig(20, 18, 16, world, rank, 1, None, 0, 0, 1, 0, 0, world, 1)
    capi.check(capi.lib().b2s_diff3d_params_for(C.byref(cfg2), C.byref(p)))
    vals = [None] * world
    dist.all_gather_object(vals, (p.dx, p.dtau, p.total_N, p.nx_g))
    assert all(v == vals[0] for v in vals) and vals[0][3] == world * 18 + 2, vals
    dist.barrier()
    dist.destroy_process_group()
    os.write(1, ("rank%dok" % rank).encode() + bytes([10]))
""")


def test_two_rank_plumbing_gloo(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER.format(root=ROOT))
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", OMP_NUM_THREADS="1")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                        "--master-addr", "127.0.0.1", "--master-port", "29531", str(script)],
                       capture_output=True, text=True, env=env, timeout=240)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "rank0ok" in r.stdout and "rank1ok" in r.stdout, r.stdout


def test_bench_reference_arm_contract(tmp_path):
    """bench.py --impl reference prints the contract's JSON line (tiny grid so it runs in seconds here)."""
    import json
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--n", "48", "--steps", "2",
                        "--warmup", "1", "--iters", "5"], capture_output=True, text=True, timeout=240)
    assert r.returncode == 0, r.stderr
    d = json.loads(r.stdout.strip().splitlines()[-1])
    assert d["impl"] == "reference" and d["unit"] == "GB/s" and d["value"] > 0
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1
