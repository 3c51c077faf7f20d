#!/bin/bash
set -u
mkdir -p gpurun_out
for v in "X=1" "SEQJ_MST=prim" "SEQJ_HIER=0" "SEQJ_FUSE_MST=single" "SEQJ_DENSE_TILES=0"; do
  echo "== $v"; env $v timeout 300 python tools/determinism_check2.py 2>&1 | tail -12
done > gpurun_out/r02_determinism2.log 2>&1
cat gpurun_out/r02_determinism2.log
