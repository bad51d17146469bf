set -x
mkdir -p gpurun_out
cp finalprojectrepo.jl_b200/libb200stencil.so /tmp/lib_keep.so
for rep in 1 2 3; do
for v in old new; do
cp scripts/ab/lib_$v.so finalprojectrepo.jl_b200/libb200stencil.so
echo "{\"lib\": \"$v\"}" >> gpurun_out/r4l_ab.jsonl
timeout 300 python bench.py --steps 5 --warmup 3 --no-mg --no-cpu-baseline >> gpurun_out/r4l_ab.jsonl 2>> gpurun_out/r4l_ab.err
done
done
cp /tmp/lib_keep.so finalprojectrepo.jl_b200/libb200stencil.so
timeout 900 python -m pytest tests/test_gpu_diffusion.py -m gpu -q -x > gpurun_out/r4l_pytest_diff.log 2>&1
echo "pytest exit $?" >> gpurun_out/r4l_pytest_diff.log
true
