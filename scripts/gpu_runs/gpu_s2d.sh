# round 2 (2 GPUs): full GPU suite on the current tree (merged halo flags, reordered boundary chunks, NS MG-PCG, CSV glue) + N=1/N=2
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/s2d_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/s2d_pytest.log
tail -4 gpurun_out/s2d_pytest.log
B="--no-mg --no-cpu-baseline --no-e2e --steps 8"
CUDA_VISIBLE_DEVICES=0 python bench.py $B > gpurun_out/s2d_g0.json 2>/dev/null
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 $B > gpurun_out/s2d_n2.json 2>gpurun_out/s2d_n2.err
CUDA_VISIBLE_DEVICES=1 python bench.py $B > gpurun_out/s2d_g1.json 2>/dev/null
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 $B > gpurun_out/s2d_n2b.json 2>gpurun_out/s2d_n2b.err
python - <<'PY'
import json
for f in ("s2d_g0","s2d_n2","s2d_g1","s2d_n2b"):
    try:
        d=json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
        print(f, round(d["value"],1), round(d["ms_per_step"],3), d["roofline"].get("per_rank_ms_per_step"), d["clocks"]["sm_mhz"], d.get("parity_check"))
    except Exception as e:
        print(f, "ERR", e)
PY
true
