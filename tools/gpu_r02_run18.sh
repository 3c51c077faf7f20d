#!/bin/bash
set -u
mkdir -p gpurun_out
for rep in 1 2 3; do
  echo "== fresh process $rep"; timeout 300 python tools/determinism_stress.py 4 2>&1 | tail -25
done > gpurun_out/r02_stress2.log 2>&1
grep -c "identical$" gpurun_out/r02_stress2.log; grep "flaky" gpurun_out/r02_stress2.log
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/r02_pytest2.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/r02_pytest2.log
timeout 600 python bench.py > gpurun_out/r02_bench_fix.json 2> gpurun_out/r02_bench_fix.err; echo "bench rc=$?"; cut -c1-600 gpurun_out/r02_bench_fix.json
