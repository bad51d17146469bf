set -x
mkdir -p gpurun_out
for sm in 1500000 1000000 250000; do
B2S_MG_STREAM_MIN=$sm timeout 600 python -c "
import json, b200stencil
from b200stencil import part2
d=part2.bench_vcycle(sizes=(1025,2049,4097), e2e=False)
print(json.dumps({'stream_min': $sm, 'ms': {k: round(v['ms_per_vcycle'],4) for k,v in d['sizes'].items()}}))" >> gpurun_out/r4h_streammin.jsonl 2>> gpurun_out/r4h_streammin.err
done
true
