"""Experiment glue of the reference, writing the reference's own CSV schemas so that its plotting scripts read B200
results unchanged.  Mirrors (reference file:line):

  part1_scaling_experiments      scripts-part1/part1_scaling_experiments.jl:26-75
       columns  delta_t,Work,Performance,Memory,Intensity,Throughput,use_shared_memory,use_gpu,strong_scaling,n_threads,n_mpi_ranks
  multigrid_bench                scripts-part2/multigrid_bench.jl:27-60
       columns  execution_policy,coarse_solver,k,l,median_time,mean_time,std_time,seed,use_gpu,nthreads
  semi_implicit_vs_explicit      scripts-part2/part2_semi_implicit_vs_explicit_experiments.jl
       columns  nx,ny,Pr,beta,t_elapsed,timed_iters

Rows are appended to an existing file like the reference does (read, push!, rewrite). The MPI ranks of the reference
are the in-process ranks of a handle here (`devices`: one CUDA ordinal per rank, may repeat).
"""
import csv
import os
import statistics
import time

import numpy as np

from . import _capi as capi

SCALING_COLUMNS = ["delta_t", "Work", "Performance", "Memory", "Intensity", "Throughput", "use_shared_memory", "use_gpu",
                   "strong_scaling", "n_threads", "n_mpi_ranks"]
MULTIGRID_COLUMNS = ["execution_policy", "coarse_solver", "k", "l", "median_time", "mean_time", "std_time", "seed", "use_gpu",
                     "nthreads"]
SEMI_IMPLICIT_COLUMNS = ["nx", "ny", "Pr", "beta", "t_elapsed", "timed_iters"]

# part1_scaling_experiments.jl:35-40 (what MPI.Dims_create gives the reference for 1, 2, 4, 8 ranks)
DIMS_DICT = {1: (1, 1, 1), 2: (2, 1, 1), 4: (2, 2, 1), 8: (2, 2, 2)}


def _jl(v):
    """Julia's CSV.write spelling of a value."""
    if isinstance(v, (bool, np.bool_)):
        return "true" if v else "false"
    if isinstance(v, float):
        return repr(v)
    return str(v)


def append_row(filename, columns, row):
    """`DataFrame(CSV.File(filename))`; `push!`; `CSV.write` -- part1_scaling_experiments.jl:62-73."""
    rows = []
    if os.path.isfile(filename):
        with open(filename, newline="") as f:
            r = list(csv.reader(f))
        if r:
            if r[0] != columns:
                raise ValueError(f"{filename} has columns {r[0]}, expected {columns}")
            rows = r[1:]
    rows.append([_jl(row[c]) for c in columns])
    os.makedirs(os.path.dirname(os.path.abspath(filename)), exist_ok=True)
    with open(filename, "w", newline="") as f:
        w = csv.writer(f, lineterminator="\n")
        w.writerow(columns)
        w.writerows(rows)


def part1_scaling_experiments(n_mpi_ranks=1, filename=None, devices=None, n_global=2 ** 7, ttot=2.0, tol=1e-6,
                              layouts=("reference",), verbose=False):
    """The reference's scaling benchmark for one rank count: strong (n_global^3 in total, divided over the rank grid of
    DIMS_DICT) and weak (n_global^3 per rank, scale_physical_size) scaling, both kernel variants; one CSV row each.
    `layouts`: "reference" = the 2x1x1 / 2x2x1 / 2x2x2 grids of the reference, "zslab" = dims (1, 1, N) with the fused
    NVLink halo push (its rows are written to <filename>.zslab.csv: same schema, different decomposition)."""
    from . import part1
    if n_mpi_ranks not in DIMS_DICT:
        raise ValueError("the reference benchmarks 1, 2, 4 or 8 ranks")
    filename = filename or os.path.join("benchmark-results", "bench_diffusion_scaling_gpu.csv")
    devices = list(devices) if devices is not None else [r % max(1, capi.device_count()) for r in range(n_mpi_ranks)]
    out = []
    for layout in layouts:
        dims = DIMS_DICT[n_mpi_ranks] if layout == "reference" else (1, 1, n_mpi_ranks)
        fn = filename if layout == "reference" else filename[:-4] + ".zslab.csv"
        for strong_scaling in (True, False):
            n = tuple(n_global // d for d in dims) if strong_scaling else (n_global,) * 3
            for use_shared_memory in (True, False):
                kw = dict(dims=dims) if dims[0] * dims[1] > 1 else dict(nslabs=n_mpi_ranks)
                _, _, b, iters = part1.diffusion_3D_kernel_programming(
                    nx=n[0], ny=n[1], nz=n[2], ttot=ttot, tol=tol, use_shared_memory=use_shared_memory, verbose=verbose,
                    scale_physical_size=not strong_scaling, devices=devices, return_iters=True, **kw)
                row = dict(delta_t=b.dt, Work=b.Work, Performance=b.Performance, Memory=b.Memory, Intensity=b.Intensity,
                           Throughput=b.Throughput, use_shared_memory=use_shared_memory, use_gpu=True,
                           strong_scaling=strong_scaling, n_threads=1, n_mpi_ranks=n_mpi_ranks)
                append_row(fn, SCALING_COLUMNS, row)
                out.append(dict(row, layout=layout, dims=list(dims), local_grid=list(n), timed_iters=sum(iters[3:]),
                                iters_per_step=iters, file=fn))
    return out


def multigrid_bench(ks=(7, 8, 9, 10), filename=None, samples=5, seed=1, device=0, ls=None, smoother=0, restriction=0):
    """multigrid_bench.jl:27-60: for k, l in 2:min(k-4, 8), solver in [jacobi, conjugate_gradient], execution policy in
    [parallel, parallel_shmem]: time MGsolve_2DPoisson!(x = 0, b ~ U[0,1), h = 1/(n-1), c = 0, tol 1e-6, Nmax = 100).
    Like the reference's @benchmark the handle (its prealloc_dict is `nothing`, multigrid.jl:51) is created inside the
    timed call; median / mean / std over `samples` evaluations."""
    import torch
    from . import part2
    filename = filename or os.path.join("benchmark-results", "bench_multigrid_gpu.csv")
    rows = []
    for k in ks:
        for l in (ls if ls is not None else range(2, min(k - 4, 8) + 1)):
            for solver in (part2.jacobi, part2.conjugate_gradient):
                for policy in (part2.parallel, part2.parallel_shmem):
                    n = 2 ** k + 1
                    opt = part2.MGOpt(coarse_solve_size=2 ** l + 1, coarse_solver=solver, execution_policy=policy,
                                      smoother=smoother, restriction=restriction)
                    b = part2.to_device(np.random.default_rng(seed).random((n, n)), device)
                    times = []
                    for s in range(samples + 1):
                        x = part2.zeros(n, n, device)
                        torch.cuda.synchronize()
                        t0 = time.perf_counter()
                        part2.MGsolve_2DPoisson(x, b, 1.0 / (n - 1), 0.0, 1e-6, 100, False, opt=opt)
                        torch.cuda.synchronize()
                        if s > 0:  # BenchmarkTools' warm-up evaluation
                            times.append(time.perf_counter() - t0)
                    row = dict(execution_policy="parallel" if policy == part2.parallel else "parallel_shmem",
                               coarse_solver="jacobi" if solver == part2.jacobi else "conjugate_gradient", k=k, l=l,
                               median_time=statistics.median(times), mean_time=statistics.fmean(times),
                               std_time=statistics.stdev(times) if len(times) > 1 else 0.0, seed=seed, use_gpu=True,
                               nthreads=1)
                    append_row(filename, MULTIGRID_COLUMNS, row)
                    rows.append(row)
    return rows


def semi_implicit_vs_explicit(nx=2049, ny=513, Pr=1.0e-3, betas=(0.5,), ttot=None, filename=None, device=0, tol=1.0e-7):
    """part2_semi_implicit_vs_explicit_experiments.jl: one row per beta (the explicit beta = 0 run needs ~9000 steps at
    the published shape and is left to the caller)."""
    from . import part2
    filename = filename or os.path.join("benchmark-results", "part2_semi_implicit_vs_explicit_experiment_results.csv")
    rows = []
    for beta in betas:
        opt = part2.SimIn_t(nx=nx, ny=ny, Pr=Pr, beta=beta, tol=tol)
        if ttot is not None:
            opt.ttot = ttot
        out = part2.navier_stokes_2D(opt=opt, verbose=False, device=device)
        row = dict(nx=nx, ny=ny, Pr=Pr, beta=beta, t_elapsed=out.t_elapsed, timed_iters=float(out.timed_iters))
        append_row(filename, SEMI_IMPLICIT_COLUMNS, row)
        rows.append(row)
    return rows
