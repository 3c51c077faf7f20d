#!/bin/bash
# HEAD of round 2: first call of a fresh process against later calls (10.8 M sampled matrix entries, all chunk / group trees)
set -u
mkdir -p gpurun_out
for rep in 1 2; do
  echo "== fresh process $rep"; timeout 200 python tools/determinism_stress.py 4 2>&1 | tail -6
done > gpurun_out/r02_stress_head.log 2>&1
cat gpurun_out/r02_stress_head.log
