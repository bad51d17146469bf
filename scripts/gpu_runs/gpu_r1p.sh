set -x
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q > gpurun_out/r1p_pytest_all.log 2>&1
echo "pytest exit $?" >> gpurun_out/r1p_pytest_all.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r1p_smoke.log 2>&1
echo "smoke exit $?" >> gpurun_out/r1p_smoke.log
