# round 2: ncu launch lists of the final multigrid V-cycle at 1025^2 and 4097^2 (shares per kernel)
set -x
mkdir -p gpurun_out
for n in 1025 4097; do
python scripts/prof_mg.py $n 4 0 a > gpurun_out/s13_plain_$n.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/s13_launches_mg$n.csv python scripts/prof_mg.py $n 4 0 a > gpurun_out/s13_ncu_$n.log 2>&1
done
ls -la gpurun_out/s13_*
true
