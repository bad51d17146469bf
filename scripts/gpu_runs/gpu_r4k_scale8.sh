set -x
mkdir -p gpurun_out
NG=$(nvidia-smi -L | wc -l)
echo "gpus $NG" > gpurun_out/r4k_info.txt
for N in 8 4 2; do
  if [ "$N" -gt "$NG" ]; then continue; fi
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2957$N bench.py --gpus $N --steps 3 --warmup 3 >> gpurun_out/r4k_scale.jsonl 2>> gpurun_out/r4k_scale.err
  echo "N=$N exit $?" >> gpurun_out/r4k_info.txt
done
timeout 300 python bench.py --gpus 1 --steps 3 --warmup 3 --no-mg --no-cpu-baseline >> gpurun_out/r4k_scale.jsonl 2>> gpurun_out/r4k_scale.err
echo "N=1 exit $?" >> gpurun_out/r4k_info.txt
# general decomposition 2x2x2, one process per GPU, against the oracle's rank emulation
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29590 tests/mp_diffusion_check.py 64 32 18 0 tma 2 2 2 > gpurun_out/r4k_mpcheck_2x2x2.log 2>&1
echo "mpcheck 2x2x2 exit $?" >> gpurun_out/r4k_info.txt
true
