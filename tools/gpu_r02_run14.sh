#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 600 python tools/determinism_check.py > gpurun_out/r02_determinism.log 2>&1; echo "rc=$?"; cat gpurun_out/r02_determinism.log
