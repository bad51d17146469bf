# round 2: uploads and downloads of the pipelined e2e path on separate copy streams (full duplex): state-I/O tests, then
# bench with e2e at N = 8 and N = 1 on the same box
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_diffusion.py tests/test_gpu_lifecycle.py -x -q > gpurun_out/s8a_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/s8a_pytest.log
tail -2 gpurun_out/s8a_pytest.log
B="--no-mg --no-cpu-baseline --steps 6"
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus 8 $B > gpurun_out/s8a_n8.json 2>gpurun_out/s8a_n8.err
python bench.py $B > gpurun_out/s8a_n1.json 2>/dev/null
python - <<'PY'
import json
for f in ["s8a_n1","s8a_n8"]:
    try:
        d=json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
        e=d.get("e2e") or {}
        print(f, round(d["value"],1), round(d["ms_per_step"],3), d.get("parity_check"), "e2e", e.get("value"), e.get("ms_per_step"), e.get("iterations_device_ms_per_step"), e.get("passes_ms_per_step"))
    except Exception as ex:
        print(f, "ERR", ex)
PY
true
