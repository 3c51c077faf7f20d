#!/bin/bash
# round 2, GPU call 8 (1 GPU): full parity suite (incl. config 4 at full size), bench with CPU baselines, launch list
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r02_pytest.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/r02_pytest.log
timeout 300 python tools/mst_structured.py prim "" > gpurun_out/r02_mst_structured.log 2>&1; echo "mst_structured rc=$?"; cat gpurun_out/r02_mst_structured.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/r02_bench_n1.json 2> gpurun_out/r02_bench_n1.err; echo "bench rc=$?"
python -c "import json;d=json.load(open('gpurun_out/r02_bench_n1.json'));print(d['ms_per_step'],d['e2e']['sequence_wall_s'],d['e2e']['pageable']['sequence_wall_s'],d['phases_ms_per_step'],d['roofline']['traffic'])"
CMD="python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e"
BENCH_NO_SAMPLER=1 $CMD > gpurun_out/r02_plain.log 2>&1 &&
BENCH_NO_SAMPLER=1 timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/r02_launches_config3.csv $CMD > gpurun_out/r02_ncu_list.log 2>&1
echo "ncu list rc=$?"
du -sh gpurun_out
