set -x
mkdir -p gpurun_out
for t in 0 1 2 3; do
B2S_MG_TILE=$t timeout 600 python -c "
import json, b200stencil
from b200stencil import part2
print(json.dumps(part2.bench_vcycle(sizes=(1025,2049,4097,8193))))" >> gpurun_out/r1h_mgbench_tiles.jsonl 2>> gpurun_out/r1h_mgbench.err
done
B2S_MG_TILE=1 timeout 900 python -m pytest tests/test_gpu_multigrid.py -m gpu -q -k "vcycle or bench_shape or full_size" > gpurun_out/r1h_pytest_tile1.log 2>&1
B2S_MG_TILE=3 timeout 900 python -m pytest tests/test_gpu_multigrid.py -m gpu -q -k "vcycle or bench_shape or full_size" > gpurun_out/r1h_pytest_tile3.log 2>&1
timeout 300 python bench.py --steps 3 --iters 200 --no-mg --no-e2e --no-cpu-baseline >> gpurun_out/r1h_diff.jsonl 2>> gpurun_out/r1h_diff.err
true
