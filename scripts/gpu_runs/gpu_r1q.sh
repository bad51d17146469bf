set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_multigrid.py -m gpu -q > gpurun_out/r1q_pytest_mg.log 2>&1
echo "pytest exit $?" >> gpurun_out/r1q_pytest_mg.log
timeout 600 python -c "
import json, b200stencil
from b200stencil import part2
print(json.dumps(part2.bench_vcycle(sizes=(1025,2049,4097,8193))))
print(json.dumps(part2.bench_navier_stokes()))" > gpurun_out/r1q_mgbench.json 2> gpurun_out/r1q_mgbench.err
python scripts/prof_mg.py 1025 4 0 > gpurun_out/r1q_mg1025_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r1q_launches_mg1025.csv \
    python scripts/prof_mg.py 1025 4 0 > gpurun_out/r1q_ncu_mg1025.log 2>&1
true
