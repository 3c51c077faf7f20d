#!/bin/bash
# MST kernels launch list with the tile list pass (random / smooth inputs)
set -u
mkdir -p gpurun_out
for kind in random smooth; do
  timeout 500 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"knn_|boruvka_knn|prim_kernel|prim_cluster|root_tree|tree_kernel" -c 400 --csv --log-file gpurun_out/r02_mstprof_tiles_$kind.csv python tools/mst_profile_case.py $kind > gpurun_out/r02_mstprof_tiles_ncu_$kind.log 2>&1
  echo "$kind rc=$?"
  python - <<PY
import csv,collections
rows=[r for r in csv.reader(open('gpurun_out/r02_mstprof_tiles_$kind.csv')) if len(r)>10 and r[0].isdigit()]
agg=collections.OrderedDict()
for r in rows:
    name=r[4].split('(')[0][:40]; grid=r[7] if len(r)>7 else ''
    v=float(r[-1].replace(',',''))
    print(name, r[7], r[8], r[-2], v)
PY
done 2>&1 | head -150
