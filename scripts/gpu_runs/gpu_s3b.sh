set -x
mkdir -p gpurun_out
for c in 0 16; do B2S_MG_CLUSTER=$c timeout 200 python scripts/mg_kernel_breakdown.py 1025 2049 4097 >> gpurun_out/s3b_breakdown.jsonl 2>>gpurun_out/s3b.err; done
cat gpurun_out/s3b_breakdown.jsonl
true
