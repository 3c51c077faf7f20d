#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -x -q -m gpu > gpurun_out/r02_pytest3.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r02_pytest3.log
for cfg in 4 5 2; do
 for v in tiles rows; do
  SEQJ_KNN=$v timeout 400 python bench.py --config $cfg --steps 3 --warmup 3 > gpurun_out/r02_bench_cfg${cfg}_$v.json 2> gpurun_out/r02_bench_cfg${cfg}_$v.err; echo "bench cfg $cfg $v rc=$?"
  python - <<PY
import json
d=json.loads(open('gpurun_out/r02_bench_cfg${cfg}_$v.json').read())
print('cfg$cfg $v', round(d['ms_per_step'],1), d['wall_ms_each_step'], {k: round(x,1) for k,x in (d.get('phases_ms_per_step') or {}).items() if k in ('prim','total')})
PY
 done
done
timeout 600 python tools/mst_structured.py "lists=rows" "lists=tiles" > gpurun_out/r02_mst_structured_tiles.log 2>&1; cat gpurun_out/r02_mst_structured_tiles.log
