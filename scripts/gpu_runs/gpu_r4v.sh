set -x
mkdir -p gpurun_out
timeout 600 python bench.py --no-mg --no-cpu-baseline --steps 10 > gpurun_out/r4v_bench.json 2> gpurun_out/r4v_bench.err
true
