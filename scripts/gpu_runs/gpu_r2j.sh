set -x
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q > gpurun_out/r2j_pytest_all.log 2>&1
echo "pytest exit $?" >> gpurun_out/r2j_pytest_all.log
