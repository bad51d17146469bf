set -x
mkdir -p gpurun_out
python scripts/prof_mg.py 4097 2 0 > gpurun_out/r2a_mg4097_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:stream2 -s 1 -c 2 -o gpurun_out/r2a_prof_mg4097 \
    python scripts/prof_mg.py 4097 2 0 > gpurun_out/r2a_ncu_mg4097.log 2>&1
true
