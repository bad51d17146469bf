set -x
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q > gpurun_out/r5e_pytest_all.log 2>&1
echo "pytest exit $?" >> gpurun_out/r5e_pytest_all.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/r5e_smoke.log 2>&1
( time timeout 900 python bench.py ) > gpurun_out/r5e_bench.json 2> gpurun_out/r5e_bench.err
true
