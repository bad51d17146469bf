set -x
mkdir -p gpurun_out
for mb in 1 100 200 296 400 600 1200; do
B2S_MG_TILE_MINBLOCKS=$mb timeout 600 python -c "
import json, b200stencil
from b200stencil import part2
d=part2.bench_vcycle(sizes=(1025,2049,4097))
print(json.dumps({'minblocks': $mb, 'ms': {k: round(v['ms_per_vcycle'],4) for k,v in d['sizes'].items()}}))" >> gpurun_out/r3n_tiles.jsonl 2>> gpurun_out/r3n_tiles.err
done
true
