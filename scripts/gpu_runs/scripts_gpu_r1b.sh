set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_multigrid.py -m gpu -q -x > gpurun_out/r1b_pytest_mg.log 2>&1
echo "pytest exit $?" >> gpurun_out/r1b_pytest_mg.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r1b_smoke.log 2>&1
echo "smoke exit $?" >> gpurun_out/r1b_smoke.log
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/r1b_bench.json 2> gpurun_out/r1b_bench.err
echo "bench exit $?" >> gpurun_out/r1b_bench.err
