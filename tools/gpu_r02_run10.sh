#!/bin/bash
# round 2, GPU call 10 (1 GPU): rescan prefilter + element-wise metrics: MST back-end check, unit tests, structured timing
set -u
mkdir -p gpurun_out
timeout 300 python tools/mst_check.py > gpurun_out/r02_mst_check.log 2>&1; echo "mst_check rc=$?"; tail -2 gpurun_out/r02_mst_check.log
timeout 600 python -m pytest tests/test_gpu_units.py tests/test_gpu_sequence.py -m gpu -q > gpurun_out/r02_pytest_units.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/r02_pytest_units.log
timeout 300 python tools/mst_structured.py prim "" "budget=1000000000" > gpurun_out/r02_mst_structured.log 2>&1; echo "mst_structured rc=$?"; cat gpurun_out/r02_mst_structured.log
for kind in clusters; do
  python tools/mst_profile_case.py $kind > gpurun_out/r02_mstprof_$kind.log 2>&1 &&
  timeout 500 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"knn_rows|boruvka_knn|prim_kernel|prim_cluster|root_tree|tree_kernel" -c 400 --csv --log-file gpurun_out/r02_mstprof_$kind.csv python tools/mst_profile_case.py $kind > gpurun_out/r02_mstprof_ncu_$kind.log 2>&1
  echo "$kind rc=$?"; cat gpurun_out/r02_mstprof_$kind.log
done
