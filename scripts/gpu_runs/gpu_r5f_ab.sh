set -x
mkdir -p gpurun_out
cp finalprojectrepo.jl_b200/libb200stencil.so /tmp/lib_keep.so
for rep in 1 2; do
for v in old new; do
cp scripts/ab/lib_$v.so finalprojectrepo.jl_b200/libb200stencil.so
echo "{\"lib\": \"$v\"}" >> gpurun_out/r5f_ab.jsonl
timeout 300 python bench.py --steps 5 --warmup 3 --no-mg --no-cpu-baseline >> gpurun_out/r5f_ab.jsonl 2>> gpurun_out/r5f_ab.err
done
done
cp /tmp/lib_keep.so finalprojectrepo.jl_b200/libb200stencil.so
true
