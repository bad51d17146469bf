set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_multigrid.py -m gpu -q -x > gpurun_out/r4g_pytest_mg.log 2>&1
echo "pytest exit $?" >> gpurun_out/r4g_pytest_mg.log
python scripts/mgbench_matrix.py 1025 > gpurun_out/r4g_matrix.jsonl 2> gpurun_out/r4g_matrix.err
true
