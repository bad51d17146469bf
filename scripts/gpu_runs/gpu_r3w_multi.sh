set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -q -x > gpurun_out/r3w_pytest_multi_2gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/r3w_pytest_multi_2gpu.log
true
