# round 2: tile kernels with the rhs read from global memory (2 shared arrays instead of 3 -> more resident blocks, fewer waves):
# parity + A/B against a build that stages it
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_multigrid.py -x -q > gpurun_out/s11_pytest_mg.log 2>&1; echo "pytest exit $?" >> gpurun_out/s11_pytest_mg.log
tail -3 gpurun_out/s11_pytest_mg.log
B2S_LABEL=fg1 timeout 300 python scripts/mgbench_a.py 1025 2049 4097 >> gpurun_out/s11_ab.jsonl 2>> gpurun_out/s11.err
timeout 120 python scripts/mg_kernel_breakdown.py 1025 2>>gpurun_out/s11.err | cut -c1-420 >> gpurun_out/s11_breakdown.txt
cd finalprojectrepo.jl_b200/csrc && touch multigrid2d.cu && make EXTRA="-DB2S_TILE_FG=0" > /dev/null 2>&1; cd ../..
B2S_LABEL=fg0 timeout 300 python scripts/mgbench_a.py 1025 2049 4097 >> gpurun_out/s11_ab.jsonl 2>> gpurun_out/s11.err
timeout 120 python scripts/mg_kernel_breakdown.py 1025 2>>gpurun_out/s11.err | cut -c1-420 >> gpurun_out/s11_breakdown.txt
cat gpurun_out/s11_ab.jsonl gpurun_out/s11_breakdown.txt
true
