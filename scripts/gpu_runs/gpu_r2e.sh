set -x
mkdir -p gpurun_out
for cfg in 0 1 3; do for zc in 0 8 16 32; do
echo "cfg $cfg zc $zc" >> gpurun_out/r2e_small.log
B2S_TMA_CFG=$cfg B2S_ZCHUNK=$zc timeout 120 python scripts/full_timestep_512.py 128 1e-6 >> gpurun_out/r2e_small.log 2>&1
done; done
for cfg in 0 1 3; do for zc in 0 16 32; do
echo "256: cfg $cfg zc $zc" >> gpurun_out/r2e_small.log
B2S_TMA_CFG=$cfg B2S_ZCHUNK=$zc timeout 120 python scripts/full_timestep_512.py 256 1e-5 >> gpurun_out/r2e_small.log 2>&1
done; done
true
