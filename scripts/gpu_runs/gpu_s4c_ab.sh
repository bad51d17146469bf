set -x
mkdir -p gpurun_out
for f in 0 1 0 1; do B2S_FORCE_MULTI_KERNEL=$f python scripts/ab_multi_kernel.py >> gpurun_out/s4c_ab.jsonl 2>> gpurun_out/s4c_ab.err; done
cat gpurun_out/s4c_ab.jsonl
true
