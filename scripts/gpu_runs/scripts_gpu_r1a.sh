set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/r1a_smi.txt 2>&1
nproc > gpurun_out/r1a_nproc.txt; lscpu | head -20 >> gpurun_out/r1a_nproc.txt; free -g >> gpurun_out/r1a_nproc.txt
timeout 900 python -m pytest tests/test_gpu_diffusion.py -m gpu -x -q > gpurun_out/r1a_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/r1a_pytest.log
timeout 600 python bench.py --steps 3 --warmup 3 --no-mg > gpurun_out/r1a_bench.json 2> gpurun_out/r1a_bench.err
for cfg in 0 1 2 3 4 5 6; do
  B2S_TMA_CFG=$cfg timeout 300 python bench.py --steps 2 --iters 100 --no-mg --no-e2e --no-cpu-baseline --variant tma >> gpurun_out/r1a_sweep.jsonl 2>> gpurun_out/r1a_sweep.err
done
timeout 300 python bench.py --steps 2 --iters 100 --no-mg --no-e2e --no-cpu-baseline --variant direct >> gpurun_out/r1a_sweep.jsonl 2>> gpurun_out/r1a_sweep.err
for zc in 32 64 128 255; do
  B2S_ZCHUNK=$zc timeout 300 python bench.py --steps 2 --iters 100 --no-mg --no-e2e --no-cpu-baseline --variant tma >> gpurun_out/r1a_sweep_zc.jsonl 2>> gpurun_out/r1a_sweep.err
done
true
