set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_multigrid.py tests/test_gpu_diffusion.py -m gpu -q > gpurun_out/r1g_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/r1g_pytest.log
timeout 600 python -c "
import json, b200stencil
from b200stencil import part2
print(json.dumps(part2.bench_vcycle(sizes=(1025,2049,4097,8193))))" > gpurun_out/r1g_mgbench.json 2> gpurun_out/r1g_mgbench.err
for zc in 0 64; do
B2S_ZCHUNK=$zc timeout 300 python bench.py --steps 3 --iters 200 --no-mg --no-e2e --no-cpu-baseline --variant tma >> gpurun_out/r1g_diff.jsonl 2>> gpurun_out/r1g_diff.err
done
python scripts/prof_mg.py 4097 3 0 > gpurun_out/r1g_mg4097_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r1g_launches_mg4097.csv \
    python scripts/prof_mg.py 4097 3 0 > gpurun_out/r1g_ncu_mg4097.log 2>&1
python scripts/prof_mg.py 4097 2 0 > gpurun_out/r1g_mg4097b_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:mg_ -s 16 -c 4 -o gpurun_out/r1g_prof_mg4097 \
    python scripts/prof_mg.py 4097 2 0 > gpurun_out/r1g_ncu_mg4097b.log 2>&1
true
