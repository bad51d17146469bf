set -x
mkdir -p gpurun_out
cp finalprojectrepo.jl_b200/libb200stencil.so /tmp/lib_keep.so
for rep in 1 2; do
for v in old new; do
cp scripts/ab/lib_$v.so finalprojectrepo.jl_b200/libb200stencil.so
timeout 300 python -c "
import json, b200stencil
from b200stencil import part2
d=part2.bench_vcycle(sizes=(1025,2049), e2e=False)
print(json.dumps({'lib': '$v', 'A': {k: round(v['ms_per_vcycle'],4) for k,v in d['sizes'].items()}}))
" >> gpurun_out/r5a_ab.jsonl 2>> gpurun_out/r5a_ab.err
echo "{\"lib\": \"$v\"}" >> gpurun_out/r5a_ab_diff.jsonl
timeout 300 python bench.py --steps 5 --warmup 3 --no-mg --no-cpu-baseline >> gpurun_out/r5a_ab_diff.jsonl 2>> gpurun_out/r5a_ab.err
done
done
cp /tmp/lib_keep.so finalprojectrepo.jl_b200/libb200stencil.so
timeout 1200 python -m pytest tests/test_gpu_diffusion.py tests/test_gpu_multigrid.py -m gpu -q -x > gpurun_out/r5a_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/r5a_pytest.log
true
