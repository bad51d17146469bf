# round 2, final validation (2 GPUs): full GPU suite, smoke, default bench N = 1, bench N = 2 and its reference arm
set -x
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -x -q > gpurun_out/s7b_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/s7b_pytest.log
tail -4 gpurun_out/s7b_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/s7b_smoke.log 2>&1; echo "smoke exit $?"
timeout 900 python bench.py > gpurun_out/s7b_bench_n1.json 2> gpurun_out/s7b_bench_n1.err; echo "bench exit $?"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 > gpurun_out/s7b_bench_n2.json 2> gpurun_out/s7b_bench_n2.err; echo "bench2 exit $?"
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --impl reference --steps 2 --warmup 1 > gpurun_out/s7b_bench_n2_ref.json 2> gpurun_out/s7b_bench_n2_ref.err; echo "ref exit $?"
python - <<'PY'
import json
for f in ("s7b_bench_n1","s7b_bench_n2","s7b_bench_n2_ref"):
    try:
        d=json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
        print(f, d.get("value"), d.get("ms_per_step"), (d.get("e2e") or {}).get("value"), d.get("parity_check"), (d.get("cpu_baseline") or {}).get("cores"), (d.get("roofline") or {}).get("frac"))
    except Exception as e:
        print(f, "ERR", e)
PY
true
