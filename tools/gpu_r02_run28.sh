#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -x -q -m gpu > gpurun_out/r02_pytest3.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r02_pytest3.log
timeout 400 python bench.py --steps 8 --warmup 3 > gpurun_out/r02_bench_n1.json 2> gpurun_out/r02_bench_n1.err; echo "bench rc=$?"
python - <<PY
import json
d=json.loads(open('gpurun_out/r02_bench_n1.json').read())
print(round(d['ms_per_step'],1), d['wall_ms_each_step'], {k: round(x,2) for k,x in (d.get('phases_ms_per_step') or {}).items()})
print('e2e', d['e2e']['sequence_wall_s'], 'gemm frac', d.get('roofline_gemm',{}).get('frac'))
PY
timeout 600 python tools/mst_structured.py "lists=rows" "lists=tiles" > gpurun_out/r02_mst_structured_tiles.log 2>&1; cat gpurun_out/r02_mst_structured_tiles.log
