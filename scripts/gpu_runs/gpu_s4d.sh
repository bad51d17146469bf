# round 2 (2 GPUs): halo push hoisted out of the plane loop -- parity (diffusion + multi-GPU suites), A/B lean vs exchange-capable
# instantiation on one slab, N=1 / N=2 weak scaling
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_multi.py tests/test_gpu_diffusion.py -x -q > gpurun_out/s4d_pytest_diff.log 2>&1; echo "pytest exit $?" >> gpurun_out/s4d_pytest_diff.log
tail -3 gpurun_out/s4d_pytest_diff.log
for f in 0 1 0 1; do B2S_FORCE_MULTI_KERNEL=$f python scripts/ab_multi_kernel.py >> gpurun_out/s4d_ab.jsonl 2>> gpurun_out/s4d_ab.err; done
cat gpurun_out/s4d_ab.jsonl
B="--no-mg --no-cpu-baseline --no-e2e --steps 8"
CUDA_VISIBLE_DEVICES=0 python bench.py $B > gpurun_out/s4d_g0.json 2>/dev/null
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 $B > gpurun_out/s4d_n2.json 2>gpurun_out/s4d_n2.err
CUDA_VISIBLE_DEVICES=1 python bench.py $B > gpurun_out/s4d_g1.json 2>/dev/null
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 $B > gpurun_out/s4d_n2b.json 2>gpurun_out/s4d_n2b.err
python - <<'PY'
import json
for f in ("s4d_g0","s4d_n2","s4d_g1","s4d_n2b"):
    try:
        d=json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
        print(f, round(d["value"],1), round(d["ms_per_step"],3), d["roofline"].get("per_rank_ms_per_step"), d["clocks"]["sm_mhz"], d.get("parity_check"))
    except Exception as e:
        print(f, "ERR", e)
PY
true
