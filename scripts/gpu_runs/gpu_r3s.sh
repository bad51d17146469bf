set -x
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:"mg_(up|down)_kernel" -c 12 -o gpurun_out/r3s_ncu_full_tiles1025 \
    python scripts/prof_mg.py 1025 1 0 a > gpurun_out/r3s_ncu_full_tiles.log 2>&1
true
