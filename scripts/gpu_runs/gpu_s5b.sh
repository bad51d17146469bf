# round 2: CUDA-graph batches of PT iterations on small grids: parity suite, then A/B
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_diffusion.py tests/test_gpu_lifecycle.py -x -q > gpurun_out/s5b_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/s5b_pytest.log
tail -3 gpurun_out/s5b_pytest.log
for g in 0 1 0 1; do B2S_DIFF_GRAPH=$g python scripts/small_grid_bench.py >> gpurun_out/s5b_small.jsonl 2>>gpurun_out/s5b.err; done
cat gpurun_out/s5b_small.jsonl
true
