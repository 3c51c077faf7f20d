#!/bin/bash
set -u
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r02_pytest_final.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/r02_pytest_final.log
tail -3 gpurun_out/r02_pytest_final.log
python bench.py > gpurun_out/r02_bench_final.json 2> gpurun_out/r02_bench_final.err; echo "bench rc=$?"
python - <<PY
import json
d=json.loads(open('gpurun_out/r02_bench_final.json').read())
print(round(d['ms_per_step'],2), d['wall_ms_each_step'], 'e2e', d['e2e']['wall_ms_each_step'], {k: round(v,2) for k,v in d['phases_ms_per_step'].items() if k in ('prim','total')})
PY
