#!/bin/bash
# round 2, GPU call 9 (1 GPU): where does the MST stage spend its time on structured data? (ncu launch list)
set -u
mkdir -p gpurun_out
for kind in clusters smooth; do
  python tools/mst_profile_case.py $kind > gpurun_out/r02_mstprof_$kind.log 2>&1 &&
  timeout 500 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"knn_rows|boruvka_knn|prim_kernel|prim_cluster|root_tree|tree_kernel" -c 400 --csv --log-file gpurun_out/r02_mstprof_$kind.csv python tools/mst_profile_case.py $kind > gpurun_out/r02_mstprof_ncu_$kind.log 2>&1
  echo "$kind rc=$?"; cat gpurun_out/r02_mstprof_$kind.log
done
du -sh gpurun_out
