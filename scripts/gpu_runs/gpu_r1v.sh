set -x
mkdir -p gpurun_out
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/r1v_bench.json 2> gpurun_out/r1v_bench.err
echo "bench exit $?" >> gpurun_out/r1v_bench.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r1v_bench_reference.json 2> gpurun_out/r1v_bench_reference.err
python bench.py --steps 2 --warmup 3 --iters 20 --no-cpu-baseline --no-e2e --no-mg > gpurun_out/r1v_bench_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r1v_launches_bench.csv \
    python bench.py --steps 2 --warmup 3 --iters 20 --no-cpu-baseline --no-e2e --no-mg > gpurun_out/r1v_ncu_bench.log 2>&1
python scripts/prof_diffusion.py 512 8 tma > gpurun_out/r1v_diff_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:step_tma -s 3 -c 2 -o gpurun_out/r1v_prof_diffusion \
    python scripts/prof_diffusion.py 512 8 tma > gpurun_out/r1v_ncu_diff.log 2>&1
true
