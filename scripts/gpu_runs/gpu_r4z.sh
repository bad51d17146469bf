set -x
mkdir -p gpurun_out
for v in a b; do
python scripts/prof_mg.py 1025 4 0 $v > gpurun_out/r4z_mg1025_${v}_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/r4z_launches_mg1025_$v.csv \
    python scripts/prof_mg.py 1025 4 0 $v > gpurun_out/r4z_ncu_mg1025_$v.log 2>&1
done
ncu --set full --clock-control none --import-source on -k regex:"mg_(up|down)_kernel" -c 12 -o gpurun_out/r4z_ncu_full_tiles1025 \
    python scripts/prof_mg.py 1025 1 0 a > gpurun_out/r4z_ncu_full_tiles.log 2>&1
ncu --set full --clock-control none -k regex:"mg_coarse_kernel" -c 1 -o gpurun_out/r4z_ncu_full_coarse1025 \
    python scripts/prof_mg.py 1025 1 0 a > gpurun_out/r4z_ncu_full_coarse.log 2>&1
true
