set -x
mkdir -p gpurun_out
for c in 0 8 16; do
  B2S_MG_PROF=1 B2S_MG_CLUSTER=$c python scripts/prof_mid.py 1025 > gpurun_out/s2c_midprof_$c.log 2>&1
  tail -3 gpurun_out/s2c_midprof_$c.log
done
true
