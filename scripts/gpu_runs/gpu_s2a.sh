# round 2, run 4 (2 GPUs): MG cluster kernel parity + A/B timing; diffusion N=2 with the fence-free mailbox
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_multigrid.py -x -q > gpurun_out/s2a_pytest_mg.log 2>&1; echo "pytest exit $?" >> gpurun_out/s2a_pytest_mg.log
tail -5 gpurun_out/s2a_pytest_mg.log
for c in 0 8 16; do
  B2S_MG_CLUSTER=$c B2S_LABEL=cluster$c timeout 300 python scripts/mgbench_a.py 1025 2049 >> gpurun_out/s2a_mgbench.jsonl 2>> gpurun_out/s2a_mgbench.err
done
cat gpurun_out/s2a_mgbench.jsonl
timeout 600 python -m pytest tests/test_gpu_multi.py tests/test_gpu_diffusion.py -x -q > gpurun_out/s2a_pytest_diff.log 2>&1; echo "pytest exit $?" >> gpurun_out/s2a_pytest_diff.log
tail -3 gpurun_out/s2a_pytest_diff.log
B="--no-mg --no-cpu-baseline --no-e2e --steps 8"
CUDA_VISIBLE_DEVICES=0 python bench.py $B > gpurun_out/s2a_g0.json 2>/dev/null
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 $B > gpurun_out/s2a_n2.json 2>gpurun_out/s2a_n2.err
CUDA_VISIBLE_DEVICES=1 python bench.py $B > gpurun_out/s2a_g1.json 2>/dev/null
python - <<'PY'
import json
for f in ("s2a_g0","s2a_n2","s2a_g1"):
    try:
        d=json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
        print(f, round(d["value"],1), round(d["ms_per_step"],3), d["roofline"].get("per_rank_ms_per_step"), d["clocks"]["sm_mhz"], d.get("parity_check"))
    except Exception as e:
        print(f, "ERR", e)
PY
true
