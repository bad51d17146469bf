set -x
mkdir -p gpurun_out
cp finalprojectrepo.jl_b200/libb200stencil.so /tmp/lib_keep.so
for rep in 1 2; do
for v in t256 t384 t512; do
cp scripts/ab/lib_$v.so finalprojectrepo.jl_b200/libb200stencil.so
timeout 300 python -c "
import json, b200stencil
from b200stencil import part2
d=part2.bench_vcycle(sizes=(1025,2049), e2e=False)
e=part2.bench_vcycle(sizes=(1025,2049), opt=part2.MGOpt(smoother=1, restriction=1), e2e=False)
print(json.dumps({'lib': '$v', 'A': {k: round(v['ms_per_vcycle'],4) for k,v in d['sizes'].items()}, 'B': {k: round(v['ms_per_vcycle'],4) for k,v in e['sizes'].items()}}))
" >> gpurun_out/r4o_ab.jsonl 2>> gpurun_out/r4o_ab.err
done
done
cp /tmp/lib_keep.so finalprojectrepo.jl_b200/libb200stencil.so
true
