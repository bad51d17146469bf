set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_multigrid.py -m gpu -q > gpurun_out/r1n_pytest_mg.log 2>&1
echo "pytest exit $?" >> gpurun_out/r1n_pytest_mg.log
B2S_MG_PROF=1 timeout 300 python scripts/prof_coarse.py 1025 > gpurun_out/r1n_coarse_prof.log 2>&1
timeout 600 python -c "
import json, b200stencil
from b200stencil import part2
print(json.dumps(part2.bench_vcycle(sizes=(1025,2049,4097,8193))))
print(json.dumps(part2.bench_navier_stokes()))" > gpurun_out/r1n_mgbench.json 2> gpurun_out/r1n_mgbench.err
true
