set -x
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q > gpurun_out/r3r_pytest_all.log 2>&1
echo "pytest exit $?" >> gpurun_out/r3r_pytest_all.log
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/r3r_bench.json 2> gpurun_out/r3r_bench.err
for v in a b; do
python scripts/prof_mg.py 1025 4 0 $v > gpurun_out/r3r_mg1025_${v}_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/r3r_launches_mg1025_$v.csv \
    python scripts/prof_mg.py 1025 4 0 $v > gpurun_out/r3r_ncu_mg1025_$v.log 2>&1
done
python scripts/prof_mg.py 4097 3 0 b > gpurun_out/r3r_mg4097_b_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/r3r_launches_mg4097_b.csv \
    python scripts/prof_mg.py 4097 3 0 b > gpurun_out/r3r_ncu_mg4097_b.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:mg_.*_rb_kernel -c 4 -o gpurun_out/r3r_ncu_full_rb4097 \
    python scripts/prof_mg.py 4097 1 0 b > gpurun_out/r3r_ncu_full_rb.log 2>&1
true
