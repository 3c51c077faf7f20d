#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_fullsize.py tests/test_gpu_units.py -x -q -m gpu -k "tile_list or measure_dm or config4 or config5" > gpurun_out/r02_pytest_tiles.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r02_pytest_tiles.log
timeout 400 python bench.py --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/r02_bench_n1b.json 2> gpurun_out/r02_bench_n1b.err; echo "bench rc=$?"
python - <<PY
import json
d=json.loads(open('gpurun_out/r02_bench_n1b.json').read())
print(round(d['ms_per_step'],1), d['wall_ms_each_step'], {k: round(x,2) for k,x in (d.get('phases_ms_per_step') or {}).items() if k in ('prim','total')})
PY
timeout 400 python bench.py --config 4 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r02_bench_cfg4_n1.json 2> gpurun_out/r02_bench_cfg4_n1.err; echo "bench cfg4 rc=$?"
python - <<PY
import json
d=json.loads(open('gpurun_out/r02_bench_cfg4_n1.json').read())
print(round(d['ms_per_step'],1), d['wall_ms_each_step'], 'e2e', d['e2e']['wall_ms_each_step'], {k: round(x,1) for k,x in (d.get('phases_ms_per_step') or {}).items()})
PY
