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


# ---- the Julia boundary, checked mechanically (Julia cannot run in this image) ----------------------------------------
_JL = os.path.join(ROOT, "finalprojectrepo.jl_b200", "julia", "B200Stencil.jl")
_HDR = os.path.join(ROOT, "include", "b200stencil.h")
_JL_STRUCT_OF = {"Diff3DConfig": "b2s_diff3d_config", "Diff3DParams": "b2s_diff3d_params", "MGConfig": "b2s_mg_config",
                 "NS2DParams": "b2s_ns2d_params", "NS2DStepInfo": "b2s_ns2d_stepinfo"}
_OPAQUE = ("b2s_diff3d", "b2s_mg", "b2s_ns2d")


def _c_type(decl):
    """'const double *Ht_dev' -> canonical C type without the parameter name: 'double*'."""
    import re
    d = re.sub(r"\bconst\b", " ", decl).strip()
    stars = d.count("*")
    d = d.replace("*", " ")
    words = d.split()
    base = words[:-1] if len(words) > 1 else words  # the last word is the parameter name (prototypes here always name them)
    if not base:
        base = words
    return " ".join(base) + "*" * stars


def _header_prototypes():
    import re
    txt = re.sub(r"/\*.*?\*/", " ", open(_HDR).read(), flags=re.S)
    protos = {}
    for m in re.finditer(r"([A-Za-z_][\w \*]*?)\b(b2s_[a-z0-9_]+)\s*\(([^;{]*?)\)\s*;", txt):
        ret, name, args = m.group(1).strip(), m.group(2), m.group(3).strip()
        ret = re.sub(r"\bconst\b", " ", ret).replace(" ", "")
        params = [] if args in ("void", "") else [_c_type(a) for a in args.split(",")]
        protos[name] = (ret, params)
    structs = {}
    for m in re.finditer(r"typedef\s+struct\s*\{(.*?)\}\s*(b2s_\w+)\s*;", txt, flags=re.S):
        fields = []
        for stmt in m.group(1).split(";"):
            stmt = stmt.strip()
            if not stmt:
                continue
            first, *rest = [p.strip() for p in stmt.split(",")]
            t = _c_type(first)
            fields.append(t)
            fields.extend([t] * len(rest))  # "int nx, ny, nz;"
        structs[m.group(2)] = fields
    return protos, structs


def _julia_matches_c(jl, c):
    """Is the Julia ccall argument type `jl` a correct binding of the C parameter type `c`?"""
    scalars = {"Cint": "int", "Cdouble": "double", "Csize_t": "size_t", "Clonglong": "long long", "Cstring": "char*"}
    if jl in scalars:
        return scalars[jl] == c
    if jl in ("CuPtr{Cdouble}", "Ptr{Cdouble}", "Ref{Cdouble}"):
        return c == "double*"
    if jl in ("Ptr{Cint}", "Ref{Cint}"):
        return c == "int*"
    if jl in ("Ptr{Clonglong}", "Ref{Clonglong}"):
        return c == "long long*"
    if jl == "Ptr{Cvoid}":  # opaque handle or void* (stream, blob)
        return c == "void*" or c in tuple(o + "*" for o in _OPAQUE)
    if jl == "Ptr{UInt8}":
        return c == "void*"
    if jl == "Ref{Ptr{Cvoid}}":
        return c in tuple(o + "**" for o in _OPAQUE) or c == "double**"
    if jl.startswith("Ref{") and jl[4:-1] in _JL_STRUCT_OF:
        return c == _JL_STRUCT_OF[jl[4:-1]] + "*"
    return False


def test_julia_ccall_signatures_match_header():
    """Every ccall in julia/B200Stencil.jl names a function the header declares, with the same arity, a matching return
    type and argument types that are correct bindings of the C parameter types; every mirrored struct has the header's
    field sequence. (The reference's host language cannot be executed here, so the boundary is checked mechanically.)"""
    import re
    src = open(_JL).read()
    protos, structs = _header_prototypes()
    assert len(protos) >= 50
    calls = re.findall(r"ccall\(\(:(b2s_\w+),\s*lib\),\s*(\w+),\s*\(([^()]*)\)", src, flags=re.S)
    assert len(calls) >= 30
    seen = set()
    for name, ret, argt in calls:
        assert name in protos, f"{name} is not declared in b200stencil.h"
        cret, cparams = protos[name]
        jl_args = [a.strip() for a in argt.replace("\n", " ").split(",") if a.strip()]
        assert len(jl_args) == len(cparams), (name, jl_args, cparams)
        assert _julia_matches_c(ret, cret if cret != "constchar*" else "char*"), (name, ret, cret)
        for i, (j, c) in enumerate(zip(jl_args, cparams)):
            assert _julia_matches_c(j, c), f"{name} argument {i + 1}: Julia {j} does not bind C `{c}`"
        seen.add(name)
    # the entry points of SURVEY 8(b) are all reachable from Julia
    for need in ("b2s_diff3d_create", "b2s_diff3d_solve_timestep", "b2s_diff3d_gather", "b2s_diff3d_ipc_connect",
                 "b2s_diffusion3d_step_tau", "b2s_mg_create", "b2s_mg_solve", "b2s_mg_vcycle", "b2s_mg_pcg_solve2", "b2s_cg_solve",
                 "b2s_iteration2d", "b2s_residual2d", "b2s_restrict_inject2d", "b2s_prolongate2d", "b2s_matvec2d",
                 "b2s_apply_bc2d", "b2s_ns2d_create", "b2s_ns2d_set_solver", "b2s_ns2d_step", "b2s_ns2d_get_field", "b2s_ns2d_set_field"):
        assert need in seen, need
    for fn in ("diffusion_3D_kernel_programming", "diffusion_3D_array_programming", "main", "MGsolve_2DPoisson!",
               "Vcycle_2DPoisson!", "iteration_2DPoisson!", "residual_2DPoisson_wrapper!", "restrict_wrapper!",
               "prolongate_wrapper!", "preallocate_buffers", "cg!", "matrix_free_matvec_prod_wrapper!", "navier_stokes_2D"):
        assert re.search(r"function\s+" + re.escape(fn) + r"\(", src), fn
    for st in ("SimIn_t", "SimOut_t", "MGOpt", "BenchResults"):
        assert re.search(r"struct\s+" + st + r"\b", src), st
    # struct layouts
    jl_ct = {"Cint": "int", "Cdouble": "double", "Ptr{Cint}": "int*"}
    for jname, cname in _JL_STRUCT_OF.items():
        m = re.search(r"^struct\s+" + jname + r"\b(.*?)^end", src, flags=re.S | re.M)
        assert m, jname
        fields = [jl_ct[t] for t in re.findall(r"^\s*\w+::([\w{}]+)", m.group(1), flags=re.M)]
        assert fields == structs[cname], (jname, fields, structs[cname])


def test_ctypes_signatures_match_header(b2s):
    """The same mechanical check for the Python mirror: arity and pointer/scalar kinds of every SIGNATURES entry."""
    from b200stencil import capi
    protos, structs = _header_prototypes()
    for name, (res, args) in capi.SIGNATURES.items():
        cret, cparams = protos[name]
        assert len(args) == len(cparams), (name, len(args), cparams)
        for a, c in zip(args, cparams):
            if a is C.c_int:
                assert c == "int", (name, c)
            elif a is C.c_double:
                assert c == "double", (name, c)
            elif a is C.c_size_t:
                assert c == "size_t", (name, c)
            else:
                assert c.endswith("*"), (name, a, c)
    for py, cname in ((capi.Diff3DConfig, "b2s_diff3d_config"), (capi.Diff3DParams, "b2s_diff3d_params"),
                      (capi.MGConfig, "b2s_mg_config"), (capi.NS2DParams, "b2s_ns2d_params"), (capi.NS2DStepInfo, "b2s_ns2d_stepinfo")):
        kinds = ["int" if t is C.c_int else "double" if t is C.c_double else "int*" for _, t in py._fields_]
        assert kinds == structs[cname], (cname, kinds, structs[cname])


def test_experiment_csv_schemas_are_the_references(b2s, tmp_path):
    """The header lines of benchmark-results/bench_diffusion_scaling_gpu.csv:1, bench_multigrid_gpu.csv:1 and
    part2_semi_implicit_vs_explicit_experiment_results.csv:1 of the reference, and Julia's spelling of Bool / Float64."""
    from b200stencil import experiments as E
    assert ",".join(E.SCALING_COLUMNS) == ("delta_t,Work,Performance,Memory,Intensity,Throughput,use_shared_memory,use_gpu,"
                                           "strong_scaling,n_threads,n_mpi_ranks")
    assert ",".join(E.MULTIGRID_COLUMNS) == "execution_policy,coarse_solver,k,l,median_time,mean_time,std_time,seed,use_gpu,nthreads"
    assert ",".join(E.SEMI_IMPLICIT_COLUMNS) == "nx,ny,Pr,beta,t_elapsed,timed_iters"
    assert E.DIMS_DICT == {1: (1, 1, 1), 2: (2, 1, 1), 4: (2, 2, 1), 8: (2, 2, 2)}  # part1_scaling_experiments.jl:35-40
    fn = str(tmp_path / "r" / "bench.csv")
    row = dict(delta_t=49.015100955963135, Work=6.9700101156e11, Performance=1.4220128041482763e10, Memory=1.44563172768e12,
               Intensity=0.48214285714285715, Throughput=2.949359890085314e10, use_shared_memory=True, use_gpu=True,
               strong_scaling=True, n_threads=1, n_mpi_ranks=1)
    E.append_row(fn, E.SCALING_COLUMNS, row)
    E.append_row(fn, E.SCALING_COLUMNS, dict(row, use_shared_memory=False))
    lines = open(fn).read().splitlines()
    assert lines[0] == ",".join(E.SCALING_COLUMNS) and len(lines) == 3
    # the reference's own first data row (bench_diffusion_scaling_gpu.csv:2), digit for digit up to Julia's e-notation
    assert lines[1].split(",")[0] == "49.015100955963135" and lines[1].endswith(",true,true,true,1,1")
    assert lines[2].endswith(",false,true,true,1,1")
    import pandas as pd
    df = pd.read_csv(fn)
    assert df.Work[0] == 6.9700101156e11 and list(df.columns) == E.SCALING_COLUMNS
    with pytest.raises(ValueError):
        E.append_row(fn, E.MULTIGRID_COLUMNS, {c: 0 for c in E.MULTIGRID_COLUMNS})
