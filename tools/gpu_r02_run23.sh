#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_units.py tests/test_gpu_sequence.py -x -q -m gpu -k "measure_dm or mst_back_ends" > gpurun_out/r02_pytest_tiles.log 2>&1; echo "pytest units rc=$?"; tail -3 gpurun_out/r02_pytest_tiles.log
timeout 600 python -m pytest tests/test_gpu_fullsize.py -x -q -m gpu -k "tile_list or planted" >> gpurun_out/r02_pytest_tiles.log 2>&1; echo "pytest fullsize rc=$?"; tail -3 gpurun_out/r02_pytest_tiles.log
for kt in 0 1; do
SEQJ_KT=$kt SEQJ_KNN=tiles timeout 300 python bench.py --steps 5 --warmup 3 > gpurun_out/r02_bench_knn_tiles$kt.json 2> gpurun_out/r02_bench_knn_tiles$kt.err; echo "bench kt=$kt rc=$?"
python - <<PY
import json
d=json.loads(open('gpurun_out/r02_bench_knn_tiles$kt.json').read())
print(d['ms_per_step'], d['wall_ms_each_step'])
print({k: round(x,2) for k,x in (d.get('phases_ms_per_step') or {}).items()})
PY
done
timeout 500 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"knn_|boruvka_knn" -c 24 --csv --log-file gpurun_out/r02_mstprof_tiles_random.csv python tools/mst_profile_case.py random > gpurun_out/r02_mstprof_tiles_ncu_random.log 2>&1
python - <<PY
import csv
rows=[r for r in csv.reader(open('gpurun_out/r02_mstprof_tiles_random.csv')) if len(r)>10 and r[0].isdigit()]
for r in rows[:24]:
    print(r[4].split('(')[0][:40], r[7], r[8], r[-1])
PY
timeout 600 python tools/mst_structured.py "lists=rows" "lists=tiles" > gpurun_out/r02_mst_structured_tiles.log 2>&1; cat gpurun_out/r02_mst_structured_tiles.log
