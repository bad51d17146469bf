set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_diffusion.py -m gpu -q > gpurun_out/r1r_pytest_diff.log 2>&1
echo "pytest exit $?" >> gpurun_out/r1r_pytest_diff.log
timeout 600 python scripts/full_timestep_512.py 512 1e-8 > gpurun_out/r1r_full_timestep_512.json 2> gpurun_out/r1r_full_timestep_512.err
timeout 300 python scripts/full_timestep_512.py 128 1e-6 > gpurun_out/r1r_full_timestep_128.json 2>> gpurun_out/r1r_full_timestep_512.err
true
