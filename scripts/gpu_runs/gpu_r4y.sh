set -x
mkdir -p gpurun_out
( time timeout 900 python bench.py ) > gpurun_out/r4y_bench.json 2> gpurun_out/r4y_bench.err
true
