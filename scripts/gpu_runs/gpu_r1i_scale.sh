set -x
mkdir -p gpurun_out
NG=$(nvidia-smi -L | wc -l)
echo "gpus $NG" > gpurun_out/r1i_info.txt
for N in $NG 2 1; do
  if [ "$N" -gt "$NG" ]; then continue; fi
  if [ "$N" -eq 1 ]; then
    timeout 600 python bench.py --gpus 1 --steps 3 --warmup 3 --no-mg --no-cpu-baseline >> gpurun_out/r1i_scale.jsonl 2>> gpurun_out/r1i_scale.err
  else
    timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2956$N bench.py --gpus $N --steps 3 --warmup 3 >> gpurun_out/r1i_scale.jsonl 2>> gpurun_out/r1i_scale.err
  fi
  echo "N=$N exit $?" >> gpurun_out/r1i_info.txt
done
timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -q > gpurun_out/r1i_pytest_multi.log 2>&1
true
