# round 2, run 2 (2 GPUs): multi-GPU parity tests of the neighbour-flag protocol + bench N=1 / N=2 on the same box
set -x
mkdir -p gpurun_out
nvidia-smi -L
timeout 900 python -m pytest tests/test_gpu_multi.py -x -q > gpurun_out/s1b_pytest_multi.log 2>&1; echo "pytest exit $?" >> gpurun_out/s1b_pytest_multi.log
tail -5 gpurun_out/s1b_pytest_multi.log
timeout 300 python bench.py --no-mg --no-cpu-baseline > gpurun_out/s1b_bench_n1.json 2> gpurun_out/s1b_bench_n1.err; echo "bench1 exit $?"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 > gpurun_out/s1b_bench_n2.json 2> gpurun_out/s1b_bench_n2.err; echo "bench2 exit $?"
tail -c 600 gpurun_out/s1b_bench_n2.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --impl reference --steps 2 --warmup 1 > gpurun_out/s1b_bench_n2_ref.json 2> gpurun_out/s1b_bench_n2_ref.err; echo "ref exit $?"
python - <<'PY'
import json
for f in ("s1b_bench_n1","s1b_bench_n2","s1b_bench_n2_ref"):
    try:
        d=json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
        print(f, d.get("value"), d.get("ms_per_step"), (d.get("e2e") or {}).get("value"), d.get("parity_check"), (d.get("cpu_baseline") or {}).get("cores"))
    except Exception as e:
        print(f, "ERR", e)
PY
true
