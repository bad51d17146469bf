set -x
mkdir -p gpurun_out
B2S_MG_PROF=1 timeout 300 python scripts/prof_coarse.py 1025 > gpurun_out/r1m_coarse_prof.log 2>&1
timeout 900 python scripts/sanitize_small.py > gpurun_out/r1m_sanitize_plain.log 2>&1 && \
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 7 python scripts/sanitize_small.py > gpurun_out/r1m_memcheck.log 2>&1
echo "memcheck exit $?" >> gpurun_out/r1m_memcheck.log
true
