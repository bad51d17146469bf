set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_multigrid.py -m gpu -q -x > gpurun_out/r3x_pytest_mg.log 2>&1
echo "pytest exit $?" >> gpurun_out/r3x_pytest_mg.log
for t in "" 8 9 0 2; do
B2S_MG_RB_TILE=$t B2S_LABEL=fused timeout 600 python scripts/mgbench_variants.py 1025 2049 4097 8193 >> gpurun_out/r3x_mgbench_rb.jsonl 2>> gpurun_out/r3x_mgbench_rb.err
done
true
