# z-chunk sweep on L2-resident grids (graph batches on)
set -x
mkdir -p gpurun_out
for zc in 0 2 3 4 6; do B2S_ZCHUNK=$zc B2S_LABEL=zc$zc python scripts/small_grid_bench.py | sed "s/^{/{\"zchunk\": $zc, /" >> gpurun_out/s5c_zchunk.jsonl 2>>gpurun_out/s5c.err; done
cat gpurun_out/s5c_zchunk.jsonl
true
