set -x
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q > gpurun_out/r2f_pytest_all.log 2>&1
echo "pytest exit $?" >> gpurun_out/r2f_pytest_all.log
timeout 120 python scripts/full_timestep_512.py 128 1e-6 > gpurun_out/r2f_ts128.json 2>&1
timeout 120 python scripts/full_timestep_512.py 256 1e-5 > gpurun_out/r2f_ts256.json 2>&1
timeout 120 python scripts/full_timestep_512.py 32 1e-8 > gpurun_out/r2f_ts32.json 2>&1
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/r2f_bench.json 2> gpurun_out/r2f_bench.err
python scripts/prof_mg.py 1025 4 0 > gpurun_out/r2f_mg1025_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r2f_launches_mg1025.csv \
    python scripts/prof_mg.py 1025 4 0 > gpurun_out/r2f_ncu_mg1025.log 2>&1
python scripts/prof_mg.py 4097 3 0 > gpurun_out/r2f_mg4097_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r2f_launches_mg4097.csv \
    python scripts/prof_mg.py 4097 3 0 > gpurun_out/r2f_ncu_mg4097.log 2>&1
true
