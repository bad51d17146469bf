# round 2: shuffle-based x neighbours in the two-column streaming multigrid kernels: parity, then A/B against a build without them
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_multigrid.py -x -q > gpurun_out/s6b_pytest_mg.log 2>&1; echo "pytest exit $?" >> gpurun_out/s6b_pytest_mg.log
tail -3 gpurun_out/s6b_pytest_mg.log
B2S_LABEL=shfl1 timeout 300 python scripts/mgbench_a.py 2049 4097 8193 >> gpurun_out/s6b_ab.jsonl 2>> gpurun_out/s6b.err
timeout 120 python scripts/mg_kernel_breakdown.py 4097 2>>gpurun_out/s6b.err | cut -c1-330 >> gpurun_out/s6b_breakdown.txt
cd finalprojectrepo.jl_b200/csrc && touch multigrid2d.cu && make EXTRA="-DB2S_S2_SHFL=0" > /dev/null 2>&1; cd ../..
B2S_LABEL=shfl0 timeout 300 python scripts/mgbench_a.py 2049 4097 8193 >> gpurun_out/s6b_ab.jsonl 2>> gpurun_out/s6b.err
timeout 120 python scripts/mg_kernel_breakdown.py 4097 2>>gpurun_out/s6b.err | cut -c1-330 >> gpurun_out/s6b_breakdown.txt
cat gpurun_out/s6b_ab.jsonl gpurun_out/s6b_breakdown.txt
true
