#!/bin/bash
set -u
mkdir -p gpurun_out
BENCH_STEP_PHASES=1 timeout 400 python bench.py --steps 14 --warmup 3 --no-cpu-baseline > gpurun_out/r02_bench_steps_nvml.json 2> gpurun_out/r02_bench_steps_nvml.err; echo "bench rc=$?"
grep "step phases" gpurun_out/r02_bench_steps_nvml.err | cut -c1-40 | tr '\n' ' '; echo
python - <<PY
import json
d=json.loads(open('gpurun_out/r02_bench_steps_nvml.json').read())
print(round(d['ms_per_step'],1), d['wall_ms_each_step'], 'e2e', d['e2e'].get('wall_ms_each_step'), d['clocks'])
PY
timeout 400 python bench.py > gpurun_out/r02_bench_default.json 2> gpurun_out/r02_bench_default.err; echo "default bench rc=$?"
python - <<PY
import json
d=json.loads(open('gpurun_out/r02_bench_default.json').read())
print(round(d['ms_per_step'],1), d['wall_ms_each_step'], 'e2e', d['e2e'].get('wall_ms_each_step'), d['clocks'], d['cpu_baseline'])
PY
