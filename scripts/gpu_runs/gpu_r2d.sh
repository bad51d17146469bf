set -x
mkdir -p gpurun_out
for m in 1500000 500000 200000 50000; do
B2S_MG_STREAM_MIN=$m timeout 600 python -c "
import json, b200stencil
from b200stencil import part2
print(json.dumps(part2.bench_vcycle(sizes=(1025,2049,4097))))" >> gpurun_out/r2d_mgbench.jsonl 2>> gpurun_out/r2d_mgbench.err
done
true
