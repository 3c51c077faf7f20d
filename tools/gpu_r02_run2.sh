#!/bin/bash
# round 2, GPU call 2 (1 GPU): parity tests, MST back-end check, bench, structured-data MST timing, ncu launch list +
# a SHORT full capture (12 launches; the 75-launch capture of call 1 ran into its 20-minute limit)
set -u
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r02_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r02_pytest.log
python tools/mst_check.py > gpurun_out/r02_mst_check.log 2>&1; echo "mst_check rc=$?"; tail -3 gpurun_out/r02_mst_check.log
python bench.py > gpurun_out/r02_bench_n1.json 2> gpurun_out/r02_bench_n1.err; echo "bench rc=$?"; tail -c 600 gpurun_out/r02_bench_n1.err
python tools/mst_structured.py prim "" > gpurun_out/r02_mst_structured.log 2>&1; echo "mst_structured rc=$?"; cat gpurun_out/r02_mst_structured.log
CMD="python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e"
BENCH_NO_SAMPLER=1 $CMD > gpurun_out/r02_plain.log 2>&1 &&
BENCH_NO_SAMPLER=1 timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/r02_launches_config3.csv $CMD > gpurun_out/r02_ncu_list.log 2>&1
echo "ncu list rc=$?"
PCMD="python tools/profile_case.py --N 10000 --L 4096 --scales 8 16"
$PCMD > gpurun_out/r02_plain2.log 2>&1 &&
timeout 600 ncu --set full --clock-control none \
  -k regex:"direct_dist_kernel|gemm_dist_kernel|derive_kernel|combine_sym_kernel|knn_rows_kernelILb0" \
  -c 12 -o gpurun_out/r02_full $PCMD > gpurun_out/r02_ncu_full.log 2>&1
echo "ncu full rc=$?"; tail -3 gpurun_out/r02_ncu_full.log
# keep gpurun_out under the 64 MiB copy-back limit
for f in gpurun_out/*.ncu-rep; do [ -f "$f" ] && [ $(stat -c %s "$f") -gt 45000000 ] && { echo "dropping $f ($(stat -c %s "$f") bytes)"; rm -f "$f"; }; done
du -sh gpurun_out
