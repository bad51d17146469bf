# round 2, run 1: full GPU suite on the new z-slab protocol + bench N=1
set -x
mkdir -p gpurun_out
nvidia-smi -L
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/s1a_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/s1a_pytest.log
tail -5 gpurun_out/s1a_pytest.log
timeout 600 python bench.py > gpurun_out/s1a_bench.json 2> gpurun_out/s1a_bench.err; echo "bench exit $?"
tail -c 1500 gpurun_out/s1a_bench.json
true
