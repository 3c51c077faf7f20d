#!/bin/bash
# Is the nvidia-smi sampler behind the sporadic +25..50 ms steps?  Same bench, three sampler settings.
set -u
mkdir -p gpurun_out
for mode in "BENCH_SMI_MS=250" "BENCH_SMI_MS=1000" "BENCH_NO_SAMPLER=1" "BENCH_SMI_MS=250"; do
  env $mode timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r02_bench_noise.json 2>/dev/null
  python -c "
import json,sys
d=json.load(open('gpurun_out/r02_bench_noise.json'))
w=d['device_ms_each_step']; print('$mode', round(d['ms_per_step'],1), 'median', sorted(w)[len(w)//2], 'max', max(w), 'outliers>235:', sum(1 for x in w if x>235), w)
"
done
