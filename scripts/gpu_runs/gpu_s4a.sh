# round 2: full suite + bench N=1 (final single-GPU kernel instantiation) + ncu launch list and --set full of the step kernel
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/s4a_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/s4a_pytest.log
tail -4 gpurun_out/s4a_pytest.log
B2S_MG_CLUSTER=16 timeout 600 python -m pytest tests/test_gpu_multigrid.py -x -q > gpurun_out/s4a_pytest_mg_cluster16.log 2>&1; echo "pytest exit $?" >> gpurun_out/s4a_pytest_mg_cluster16.log
tail -2 gpurun_out/s4a_pytest_mg_cluster16.log
timeout 900 python bench.py > gpurun_out/s4a_bench.json 2> gpurun_out/s4a_bench.err; echo "bench exit $?"
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/s4a_smoke.log 2>&1; echo "smoke exit $?"
python bench.py --steps 2 --warmup 1 --iters 20 --no-mg --no-cpu-baseline --no-e2e > gpurun_out/s4a_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/s4a_launches_bench.csv \
    python bench.py --steps 2 --warmup 1 --iters 20 --no-mg --no-cpu-baseline --no-e2e > gpurun_out/s4a_ncu_launches.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:step_tma -s 40 -c 3 -o gpurun_out/s4a_ncu_full_step \
    python bench.py --steps 2 --warmup 1 --iters 20 --no-mg --no-cpu-baseline --no-e2e > gpurun_out/s4a_ncu_full.log 2>&1
ls -la gpurun_out/s4a_*
true
