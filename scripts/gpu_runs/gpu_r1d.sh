set -x
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q > gpurun_out/r1d_pytest_all.log 2>&1
echo "pytest exit $?" >> gpurun_out/r1d_pytest_all.log
# launch lists (cold-cache, serialised: shares only)
python bench.py --steps 2 --warmup 3 --iters 20 --no-cpu-baseline --no-e2e --no-mg > gpurun_out/r1d_bench_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r1d_launches_bench.csv \
    python bench.py --steps 2 --warmup 3 --iters 20 --no-cpu-baseline --no-e2e --no-mg > gpurun_out/r1d_ncu_bench.log 2>&1
python scripts/prof_mg.py 1025 4 0 > gpurun_out/r1d_mg1025_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r1d_launches_mg1025.csv \
    python scripts/prof_mg.py 1025 4 0 > gpurun_out/r1d_ncu_mg1025.log 2>&1
python scripts/prof_mg.py 4097 3 0 > gpurun_out/r1d_mg4097_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r1d_launches_mg4097.csv \
    python scripts/prof_mg.py 4097 3 0 > gpurun_out/r1d_ncu_mg4097.log 2>&1
# full capture of the dominant kernel of each path
python scripts/prof_diffusion.py 512 8 tma > gpurun_out/r1d_diff_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:step_tma -s 3 -c 2 -o gpurun_out/r1d_prof_diffusion \
    python scripts/prof_diffusion.py 512 8 tma > gpurun_out/r1d_ncu_diff.log 2>&1
python scripts/prof_mg.py 4097 2 0 > gpurun_out/r1d_mg4097b_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:mg_ -s 40 -c 34 -o gpurun_out/r1d_prof_mg4097 \
    python scripts/prof_mg.py 4097 2 0 > gpurun_out/r1d_ncu_mg4097b.log 2>&1
true
