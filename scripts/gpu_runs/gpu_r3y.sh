set -x
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q > gpurun_out/r3y_pytest_all.log 2>&1
echo "pytest exit $?" >> gpurun_out/r3y_pytest_all.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/r3y_smoke.log 2>&1
( time timeout 900 python bench.py ) > gpurun_out/r3y_bench.json 2> gpurun_out/r3y_bench.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r3y_bench_reference.json 2> gpurun_out/r3y_bench_reference.err
for v in a b; do
for n in 1025 4097; do
python scripts/prof_mg.py $n 4 0 $v > gpurun_out/r3y_mg${n}_${v}_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/r3y_launches_mg${n}_$v.csv \
    python scripts/prof_mg.py $n 4 0 $v > gpurun_out/r3y_ncu_mg${n}_$v.log 2>&1
done
done
ncu --set full --clock-control none --import-source on -k regex:"mg_.*_rb_kernel" -c 4 -o gpurun_out/r3y_ncu_full_rb4097 \
    python scripts/prof_mg.py 4097 1 0 b > gpurun_out/r3y_ncu_full_rb.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"mg_(up|down)_kernel" -c 12 -o gpurun_out/r3y_ncu_full_tiles1025 \
    python scripts/prof_mg.py 1025 1 0 a > gpurun_out/r3y_ncu_full_tiles.log 2>&1
ncu --set full --clock-control none -k regex:"mg_coarse_kernel" -c 1 -o gpurun_out/r3y_ncu_full_coarse1025 \
    python scripts/prof_mg.py 1025 1 0 a > gpurun_out/r3y_ncu_full_coarse.log 2>&1
python bench.py --steps 2 --warmup 1 > gpurun_out/r3y_bench_short.json 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r3y_launches_bench.csv \
    python bench.py --steps 2 --warmup 1 > gpurun_out/r3y_ncu_bench.log 2>&1
true
