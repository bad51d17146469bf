# round 2, final 8-GPU run: weak scaling 8/4/2/1 (parity check inside bench), 8 independent replicas at the same time (the
# floor a coupled run can reach on this box), the reference's strong/weak scaling table (128^3) in its CSV schema
set -x
mkdir -p gpurun_out
B="--no-mg --no-cpu-baseline --steps 6"
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus 8 $B > gpurun_out/s4b_n8.json 2>gpurun_out/s4b_n8.err
for n in 4 2; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n bench.py --gpus $n $B --no-e2e > gpurun_out/s4b_n$n.json 2>gpurun_out/s4b_n$n.err
done
python bench.py $B > gpurun_out/s4b_n1.json 2>/dev/null
for g in 0 1 2 3 4 5 6 7; do CUDA_VISIBLE_DEVICES=$g python bench.py $B --no-e2e > gpurun_out/s4b_replica_g$g.json 2>/dev/null & done
wait
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus 8 $B --no-e2e > gpurun_out/s4b_n8b.json 2>gpurun_out/s4b_n8b.err
timeout 600 python scripts/strong_scaling_table.py gpurun_out 1 2 4 8 > gpurun_out/s4b_strong_scaling.jsonl 2> gpurun_out/s4b_strong_scaling.err
tail -3 gpurun_out/s4b_strong_scaling.err
python - <<'PY'
import json
for f in ["s4b_n1","s4b_n2","s4b_n4","s4b_n8","s4b_n8b"]+[f"s4b_replica_g{g}" for g in range(8)]:
    try:
        d=json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
        e=d.get("e2e") or {}
        print(f, round(d["value"],1), round(d["ms_per_step"],3), [round(x,2) for x in d["roofline"].get("per_rank_ms_per_step")], d["clocks"]["sm_mhz"], d.get("parity_check"), "e2e", e.get("value"))
    except Exception as ex:
        print(f, "ERR", ex)
PY
cat gpurun_out/s4b_strong_scaling.jsonl | cut -c1-400
true
