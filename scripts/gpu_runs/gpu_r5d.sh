set -x
mkdir -p gpurun_out
for mp in 2000 600 200; do
B2S_MG_SMEM_MAXPTS=$mp timeout 600 python -c "
import json, b200stencil
from b200stencil import part2
d=part2.bench_vcycle(sizes=(1025,2049), e2e=False)
print(json.dumps({'maxpts': $mp, 'ms': {k: round(v['ms_per_vcycle'],4) for k,v in d['sizes'].items()}}))" >> gpurun_out/r5d_smemlevels.jsonl 2>> gpurun_out/r5d.err
done
true
