# round 2, last call: the final tree on one GPU -- full GPU suite, smoke, default bench (what the driver runs)
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/s9_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/s9_pytest.log
tail -3 gpurun_out/s9_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/s9_smoke.log 2>&1; echo "smoke exit $?"
timeout 900 python bench.py > gpurun_out/s9_bench.json 2> gpurun_out/s9_bench.err; echo "bench exit $?"
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/s9_bench_ref.json 2>/dev/null; echo "ref exit $?"
tail -c 400 gpurun_out/s9_bench_ref.json
true
