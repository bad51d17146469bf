set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_diffusion.py -m gpu -q -x > gpurun_out/r4u_pytest_diff.log 2>&1
echo "pytest exit $?" >> gpurun_out/r4u_pytest_diff.log
timeout 600 python bench.py --no-mg --no-cpu-baseline > gpurun_out/r4u_bench.json 2> gpurun_out/r4u_bench.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/r4u_bench_n2.json 2> gpurun_out/r4u_bench_n2.err
true
