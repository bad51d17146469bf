set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_diffusion.py -m gpu -q -x > gpurun_out/r5g_pytest_diff.log 2>&1
echo "pytest exit $?" >> gpurun_out/r5g_pytest_diff.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/r5g_smoke.log 2>&1
timeout 300 python bench.py --no-mg --no-cpu-baseline > gpurun_out/r5g_bench.json 2> gpurun_out/r5g_bench.err
true
