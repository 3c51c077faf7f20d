#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 300 python tools/mst_check.py > gpurun_out/r02_mst_check.log 2>&1; echo "mst_check rc=$?"; tail -2 gpurun_out/r02_mst_check.log
timeout 300 python tools/mst_structured.py prim "" > gpurun_out/r02_mst_structured.log 2>&1; echo "mst_structured rc=$?"; cat gpurun_out/r02_mst_structured.log
python tools/mst_profile_case.py clusters > gpurun_out/r02_mstprof_clusters.log 2>&1 &&
timeout 500 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"knn_rows|boruvka_knn" -c 60 --csv --log-file gpurun_out/r02_mstprof_clusters.csv python tools/mst_profile_case.py clusters > gpurun_out/r02_mstprof_ncu_clusters.log 2>&1
python - <<'PY'
import csv
rows=list(csv.reader(open('gpurun_out/r02_mstprof_clusters.csv')))
hi=[i for i,r in enumerate(rows) if r and r[0]=='ID'][0]
col={h:i for i,h in enumerate(rows[hi])}
for r in rows[hi+1:hi+10]:
    print(r[col['Kernel Name']].split('(')[0][:30], r[col['Grid Size']], float(r[col['Metric Value']].replace(',',''))/1e6)
PY
timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r02_bench_quick.json 2>/dev/null; python -c "import json;d=json.load(open('gpurun_out/r02_bench_quick.json'));print(d['ms_per_step'],d['phases_ms_per_step']['prim'],d['wall_ms_each_step'],d['device_ms_each_step'])"
