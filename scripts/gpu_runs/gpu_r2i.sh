set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_multigrid.py -m gpu -q -x > gpurun_out/r2i_pytest_mg.log 2>&1
echo "pytest exit $?" >> gpurun_out/r2i_pytest_mg.log
for kind in warp 2col; do
B2S_MG_STREAM_KIND=$kind timeout 600 python -c "
import json, b200stencil
from b200stencil import part2
print(json.dumps(part2.bench_vcycle(sizes=(1025,2049,4097,8193))))" >> gpurun_out/r2i_mgbench.jsonl 2>> gpurun_out/r2i_mgbench.err
done
true
