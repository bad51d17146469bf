# round 2: tile sweeps with two vertically adjacent points per thread: parity + A/B against a build with one point per thread
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_multigrid.py -x -q > gpurun_out/s10_pytest_mg.log 2>&1; echo "pytest exit $?" >> gpurun_out/s10_pytest_mg.log
tail -3 gpurun_out/s10_pytest_mg.log
B2S_LABEL=yb2 timeout 300 python scripts/mgbench_a.py 1025 2049 4097 >> gpurun_out/s10_ab.jsonl 2>> gpurun_out/s10.err
timeout 120 python scripts/mg_kernel_breakdown.py 1025 2>>gpurun_out/s10.err | cut -c1-420 >> gpurun_out/s10_breakdown.txt
cd finalprojectrepo.jl_b200/csrc && touch multigrid2d.cu && make EXTRA="-DB2S_TILE_YB=1" > /dev/null 2>&1; cd ../..
B2S_LABEL=yb1 timeout 300 python scripts/mgbench_a.py 1025 2049 4097 >> gpurun_out/s10_ab.jsonl 2>> gpurun_out/s10.err
timeout 120 python scripts/mg_kernel_breakdown.py 1025 2>>gpurun_out/s10.err | cut -c1-420 >> gpurun_out/s10_breakdown.txt
cat gpurun_out/s10_ab.jsonl gpurun_out/s10_breakdown.txt
true
