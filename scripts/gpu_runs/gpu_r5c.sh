set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_multigrid.py -m gpu -q -x > gpurun_out/r5c_pytest_mg.log 2>&1
echo "pytest exit $?" >> gpurun_out/r5c_pytest_mg.log
B2S_MG_PROF=1 timeout 300 python scripts/prof_coarse.py > gpurun_out/r5c_coarse_phases.log 2>&1
timeout 600 python -c "
import json, b200stencil
from b200stencil import part2
d=part2.bench_vcycle(sizes=(1025,2049,4097), e2e=False)
print(json.dumps({'ms': {k: round(v['ms_per_vcycle'],4) for k,v in d['sizes'].items()}, 'solve_ms': {k: round(v['solve_ms'],3) for k,v in d['sizes'].items()}}))
d=part2.bench_vcycle(sizes=(1025,2049,4097), opt=part2.MGOpt(smoother=1, restriction=1), e2e=False)
print(json.dumps({'variant': 'B', 'ms': {k: round(v['ms_per_vcycle'],4) for k,v in d['sizes'].items()}}))
" >> gpurun_out/r5c_bench.jsonl 2>> gpurun_out/r5c_bench.err
true
