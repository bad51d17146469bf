set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_diffusion.py -m gpu -q -x > gpurun_out/r3v_pytest_diff.log 2>&1
echo "pytest exit $?" >> gpurun_out/r3v_pytest_diff.log
true
