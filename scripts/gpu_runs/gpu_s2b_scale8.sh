# round 2: 1/2/4/8-GPU weak scaling with the neighbour-flag protocol + fence-free mailbox (device-resident value; e2e at N=8 and N=1)
set -x
mkdir -p gpurun_out
B="--no-mg --no-cpu-baseline --steps 6"
for n in 8 4 2; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n bench.py --gpus $n $B > gpurun_out/s2b_n$n.json 2>gpurun_out/s2b_n$n.err
done
python bench.py $B > gpurun_out/s2b_n1.json 2>/dev/null
for g in 3 7; do CUDA_VISIBLE_DEVICES=$g python bench.py $B --no-e2e > gpurun_out/s2b_g$g.json 2>/dev/null; done
python - <<'PY'
import json
for f in ("s2b_n1","s2b_n2","s2b_n4","s2b_n8","s2b_g3","s2b_g7"):
    try:
        d=json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
        e=d.get("e2e") or {}
        print(f, round(d["value"],1), round(d["ms_per_step"],3), [round(x,2) for x in d["roofline"].get("per_rank_ms_per_step")], d["clocks"]["sm_mhz"], d.get("parity_check"), "e2e", e.get("value"), e.get("iterations_device_ms_per_step"), e.get("passes_ms_per_step"), e.get("numa_binding"))
    except Exception as ex:
        print(f, "ERR", ex)
PY
true
