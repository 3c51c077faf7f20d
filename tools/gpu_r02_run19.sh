#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_units.py tests/test_gpu_sequence.py -x -q -m gpu -k "measure_dm or mst_back_ends" > gpurun_out/r02_pytest_tiles.log 2>&1; echo "pytest units rc=$?"; tail -5 gpurun_out/r02_pytest_tiles.log
timeout 600 python -m pytest tests/test_gpu_fullsize.py -x -q -m gpu -k "tile_list or planted or cold_and_warm_identical" >> gpurun_out/r02_pytest_tiles.log 2>&1; echo "pytest fullsize rc=$?"; tail -5 gpurun_out/r02_pytest_tiles.log
for v in tiles rows; do
  SEQJ_KNN=$v timeout 300 python bench.py --steps 5 --warmup 3 > gpurun_out/r02_bench_knn_$v.json 2> gpurun_out/r02_bench_knn_$v.err; echo "bench $v rc=$?"
  python - <<PY
import json
d=json.loads(open('gpurun_out/r02_bench_knn_$v.json').read())
print('$v', d['ms_per_step'], d['wall_ms_each_step'], {k: round(x,2) for k,x in d['e2e']['device_ms_per_step'].items()} )
print({k: round(x,2) for k,x in (d.get('phases_ms_per_step') or {}).items()})
PY
done
timeout 600 python tools/mst_structured.py prim "lists=rows" "lists=tiles" > gpurun_out/r02_mst_structured_tiles.log 2>&1; cat gpurun_out/r02_mst_structured_tiles.log
