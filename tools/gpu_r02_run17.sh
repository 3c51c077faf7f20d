#!/bin/bash
set -u
mkdir -p gpurun_out
for v in "X=1" "SEQJ_MST=prim"; do
  echo "== $v"; env $v timeout 400 python tools/determinism_stress.py 30 2>&1 | grep -v "identical$" | tail -60
done > gpurun_out/r02_stress.log 2>&1
tail -80 gpurun_out/r02_stress.log
