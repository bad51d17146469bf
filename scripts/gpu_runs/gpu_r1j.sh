set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_multigrid.py -m gpu -q -x > gpurun_out/r1j_pytest_mg.log 2>&1
echo "pytest exit $?" >> gpurun_out/r1j_pytest_mg.log
for chv in 0 16 32 64 128; do
B2S_MG_CH=$chv timeout 600 python -c "
import json, b200stencil
from b200stencil import part2
print(json.dumps(part2.bench_vcycle(sizes=(1025,2049,4097,8193))))" >> gpurun_out/r1j_mgbench_ch.jsonl 2>> gpurun_out/r1j_mgbench.err
done
python scripts/prof_mg.py 4097 3 0 > gpurun_out/r1j_mg4097_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r1j_launches_mg4097.csv \
    python scripts/prof_mg.py 4097 3 0 > gpurun_out/r1j_ncu_mg4097.log 2>&1
python scripts/prof_mg.py 1025 4 0 > gpurun_out/r1j_mg1025_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r1j_launches_mg1025.csv \
    python scripts/prof_mg.py 1025 4 0 > gpurun_out/r1j_ncu_mg1025.log 2>&1
python scripts/prof_mg.py 4097 2 0 > gpurun_out/r1j_mg4097b_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:mg_ -s 13 -c 3 -o gpurun_out/r1j_prof_mg4097 \
    python scripts/prof_mg.py 4097 2 0 > gpurun_out/r1j_ncu_mg4097b.log 2>&1
true
