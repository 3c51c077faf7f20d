#!/bin/bash
set -u
mkdir -p gpurun_out
for v in tiles rows; do
BENCH_STEP_PHASES=1 SEQJ_KNN=$v timeout 400 python bench.py --steps 14 --warmup 3 --no-cpu-baseline > gpurun_out/r02_bench_steps_$v.json 2> gpurun_out/r02_bench_steps_$v.err; echo "bench $v rc=$?"
grep "step phases" gpurun_out/r02_bench_steps_$v.err | cut -c1-330
python - <<PY
import json
d=json.loads(open('gpurun_out/r02_bench_steps_$v.json').read())
print(round(d['ms_per_step'],1), d['wall_ms_each_step'], 'e2e', d['e2e'].get('wall_ms_each_step'))
PY
done
