# round 2: full GPU suite with the small-grid changes (graph batches, short z chunks); small-grid bench; ncu --set full of the
# multigrid level kernels at 4097^2 and 1025^2 (DRAM traffic per launch for bench.py's mg_roofline)
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/s6a_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/s6a_pytest.log
tail -4 gpurun_out/s6a_pytest.log
python scripts/small_grid_bench.py > gpurun_out/s6a_small.json 2>gpurun_out/s6a_small.err; cat gpurun_out/s6a_small.json
python scripts/prof_mg.py 4097 3 0 a > gpurun_out/s6a_plain4097.log 2>&1 && \
ncu --set full --clock-control none -k regex:"mg_(up|down)_stream2" -c 6 -o gpurun_out/s6a_ncu_full_mg4097 python scripts/prof_mg.py 4097 3 0 a > gpurun_out/s6a_ncu4097.log 2>&1
python scripts/prof_mg.py 1025 3 0 a > gpurun_out/s6a_plain1025.log 2>&1 && \
ncu --set full --clock-control none -k regex:"mg_(up|down)_kernel" -c 12 -o gpurun_out/s6a_ncu_full_mg1025 python scripts/prof_mg.py 1025 3 0 a > gpurun_out/s6a_ncu1025.log 2>&1
ls -la gpurun_out/s6a_*
true
