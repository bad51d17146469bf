set -x
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/r1e_topo.txt 2>&1
timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -q -x > gpurun_out/r1e_pytest_multi.log 2>&1
echo "pytest exit $?" >> gpurun_out/r1e_pytest_multi.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29551 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/r1e_bench_n2.json 2> gpurun_out/r1e_bench_n2.err
echo "bench2 exit $?" >> gpurun_out/r1e_bench_n2.err
timeout 600 python bench.py --gpus 1 --steps 3 --warmup 3 --no-mg --no-cpu-baseline > gpurun_out/r1e_bench_n1.json 2> gpurun_out/r1e_bench_n1.err
true
