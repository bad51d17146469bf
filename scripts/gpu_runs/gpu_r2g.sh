set -x
mkdir -p gpurun_out
timeout 900 python bench.py --steps 5 --warmup 3 --no-mg --no-cpu-baseline > gpurun_out/r2g_bench.json 2> gpurun_out/r2g_bench.err
nvidia-smi --query-gpu=timestamp,clocks.sm --format=csv,noheader,nounits -i 0 > gpurun_out/r2g_smi_format.txt 2>&1
true
