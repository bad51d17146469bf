# where do the 3.6 % at N=2 go? GPU0 alone, GPU1 alone, both as independent replicas at the same time, then coupled
set -x
mkdir -p gpurun_out
B="--no-mg --no-cpu-baseline --no-e2e --steps 8"
CUDA_VISIBLE_DEVICES=0 python bench.py $B > gpurun_out/s1c_g0.json 2>/dev/null
CUDA_VISIBLE_DEVICES=1 python bench.py $B > gpurun_out/s1c_g1.json 2>/dev/null
CUDA_VISIBLE_DEVICES=0 python bench.py $B > gpurun_out/s1c_both_g0.json 2>/dev/null &
CUDA_VISIBLE_DEVICES=1 python bench.py $B > gpurun_out/s1c_both_g1.json 2>/dev/null
wait
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 $B > gpurun_out/s1c_n2.json 2>/dev/null
CUDA_VISIBLE_DEVICES=0 python bench.py $B > gpurun_out/s1c_g0_again.json 2>/dev/null
python - <<'PY'
import json
for f in ("s1c_g0","s1c_g1","s1c_both_g0","s1c_both_g1","s1c_n2","s1c_g0_again"):
    try:
        d=json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
        print(f, round(d["value"],1), round(d["ms_per_step"],3), d["roofline"].get("per_rank_ms_per_step"), d["clocks"]["sm_mhz"], d["clocks"]["reasons"])
    except Exception as e:
        print(f, "ERR", e)
PY
true
