set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_multi.py -m gpu -q -x > gpurun_out/r4i_pytest_multi_4gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/r4i_pytest_multi_4gpu.log
true
