# round 2: prefetch depth of the streaming multigrid kernels (ring 8/D 6/DC 4 = default vs ring 16 with deeper look-ahead);
# each variant is built on the box, checked bit for bit (full-size 5-implementation agreement test) and timed
set -x
mkdir -p gpurun_out
cd finalprojectrepo.jl_b200/csrc
for v in "16 14 12" "16 12 8" "12 10 8" "8 6 4"; do
  set -- $v
  touch multigrid2d.cu
  make EXTRA="-DB2S_S2_RING=$1 -DB2S_S2_D=$2 -DB2S_S2_DC=$3" > /dev/null 2>&1 || { echo "build failed $v"; continue; }
  grep -A2 "stream2_kernel" build/multigrid2d.ptxas.log | grep -E "Used" | head -2
  ( cd ../.. && timeout 300 python -m pytest tests/test_gpu_multigrid.py -x -q -k "full_size_properties and not variant_b" 2>&1 | tail -1
    B2S_LABEL="ring$1_d$2_dc$3" timeout 300 python scripts/mgbench_a.py 2049 4097 8193 >> gpurun_out/s5a_ring.jsonl 2>> gpurun_out/s5a_ring.err
    timeout 120 python scripts/mg_kernel_breakdown.py 4097 2>>gpurun_out/s5a_ring.err | cut -c1-420 >> gpurun_out/s5a_breakdown.txt )
done
cd ../..
cat gpurun_out/s5a_ring.jsonl gpurun_out/s5a_breakdown.txt
true
