#!/bin/bash
# round 2, GPU call 7 (1 GPU): the cluster Boruvka of the fuse and the shared-memory rescan, each step under its own
# timeout (a barrier bug must not hang the box), then bench + structured-data MST timing
set -u
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_fullsize.py -m gpu -q -k "fuse_cluster" > gpurun_out/r02_pytest_fuse.log 2>&1; echo "fuse test rc=$?"; tail -5 gpurun_out/r02_pytest_fuse.log
timeout 600 python -m pytest tests -m gpu -q > gpurun_out/r02_pytest.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/r02_pytest.log
timeout 300 python tools/mst_check.py > gpurun_out/r02_mst_check.log 2>&1; echo "mst_check rc=$?"; tail -2 gpurun_out/r02_mst_check.log
timeout 300 python tools/mst_structured.py prim "" "budget=0" "budget=1000000000" > gpurun_out/r02_mst_structured.log 2>&1; echo "mst_structured rc=$?"; cat gpurun_out/r02_mst_structured.log
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02_bench_n1.json 2> gpurun_out/r02_bench_n1.err; echo "bench rc=$?"
python -c "import json;d=json.load(open('gpurun_out/r02_bench_n1.json'));print(d['ms_per_step'],d['e2e']['sequence_wall_s'],d['phases_ms_per_step'])"
du -sh gpurun_out
