#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29577 tools/mgpu_determinism.py > gpurun_out/r02_mgpu_determinism.log 2>&1; echo "rc=$?"; grep -v "OMP_NUM\|^\*\*\*\*\|^$" gpurun_out/r02_mgpu_determinism.log | cut -c1-700
