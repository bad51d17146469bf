"""The reference's scaling benchmark on B200s (scripts-part1/part1_scaling_experiments.jl:26-75): 128^3 in total (strong) and
128^3 per rank (weak), ttot = 2, tol = 1e-6, over 1/2/4/8 ranks = GPUs of one box, in the reference's rank grids
(2x1x1, 2x2x1, 2x2x2) and as z-slabs (fused NVLink halo push). Writes the reference's CSV schema and prints one JSON line
per row with the timed iteration counts (known answers: 12,905 / 13,074 / 13,242 / 13,008 strong; 12,499 / 12,079 / 11,477 weak).
Usage: strong_scaling_table.py OUTDIR [ranks...]"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import b200stencil  # noqa
from b200stencil import capi, experiments as E
out = sys.argv[1] if len(sys.argv) > 1 else "gpurun_out"
ranks = [int(a) for a in sys.argv[2:]] or [1, 2, 4, 8]
ngpu = capi.device_count()
for n in ranks:
    if n > ngpu:
        print(json.dumps({"n_mpi_ranks": n, "skipped": f"only {ngpu} GPUs"}))
        continue
    rows = E.part1_scaling_experiments(n_mpi_ranks=n, filename=os.path.join(out, "bench_diffusion_scaling_b200.csv"),
                                       devices=list(range(n)), layouts=("reference", "zslab") if n > 1 else ("reference",))
    for r in rows:
        print(json.dumps({k: r[k] for k in ("n_mpi_ranks", "layout", "dims", "local_grid", "strong_scaling", "use_shared_memory",
                                            "timed_iters", "delta_t", "Performance", "Throughput")}), flush=True)
