# round 2: cluster kernel v2 (st.async + mbarrier dataflow): parity, A/B timing, phase stamps
set -x
mkdir -p gpurun_out
timeout 120 python scripts/prof_mid.py 257 > gpurun_out/s3a_first.log 2>&1; echo "first exit $?"
tail -2 gpurun_out/s3a_first.log | cut -c1-300
timeout 600 python -m pytest tests/test_gpu_multigrid.py -x -q > gpurun_out/s3a_pytest_mg.log 2>&1; echo "pytest exit $?" >> gpurun_out/s3a_pytest_mg.log
tail -5 gpurun_out/s3a_pytest_mg.log
for c in 0 16 8; do
  B2S_MG_CLUSTER=$c B2S_LABEL=cluster$c timeout 300 python scripts/mgbench_a.py 1025 2049 >> gpurun_out/s3a_mgbench.jsonl 2>> gpurun_out/s3a_mgbench.err
done
B2S_MG_CLUSTER=16 B2S_MG_CLUSTER_MAXPTS=70000 B2S_LABEL=cluster16_257 timeout 300 python scripts/mgbench_a.py 1025 >> gpurun_out/s3a_mgbench.jsonl 2>> gpurun_out/s3a_mgbench.err
cat gpurun_out/s3a_mgbench.jsonl
for c in 16 8; do
  B2S_MG_PROF=1 B2S_MG_CLUSTER=$c timeout 120 python scripts/prof_mid.py 1025 > gpurun_out/s3a_midprof_$c.log 2>&1
  tail -2 gpurun_out/s3a_midprof_$c.log | cut -c1-1200
done
true
