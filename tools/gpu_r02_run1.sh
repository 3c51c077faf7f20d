#!/bin/bash
# round 2, GPU call 1: parity tests, bench, structured-data MST check, ncu launch list + full capture (1 GPU)
set -u
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r02_pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/r02_pytest.log
tail -5 gpurun_out/r02_pytest.log
python bench.py > gpurun_out/r02_bench_n1.json 2> gpurun_out/r02_bench_n1.err; echo "bench rc=$?"
tail -c 1500 gpurun_out/r02_bench_n1.err
python tools/mst_structured.py prim "" > gpurun_out/r02_mst_structured.log 2>&1; echo "mst_structured rc=$?"; cat gpurun_out/r02_mst_structured.log
CMD="python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e"
BENCH_NO_SAMPLER=1 $CMD > gpurun_out/r02_plain.log 2>&1 &&
BENCH_NO_SAMPLER=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/r02_launches_config3.csv $CMD > gpurun_out/r02_ncu_list.log 2>&1
echo "ncu list rc=$?"
PCMD="python tools/profile_case.py --N 10000 --L 4096 --scales 1 2 4 8 16"
$PCMD > gpurun_out/r02_plain2.log 2>&1 &&
timeout 1200 ncu --set full --clock-control none --import-source on \
  -k regex:"direct_dist_kernel|gemm_dist_kernel|derive_kernel|combine_sym_kernel|knn_rows_kernel|boruvka_knn_kernel|prep_kernel|tree_kernel" \
  -c 75 -o gpurun_out/r02_full $PCMD > gpurun_out/r02_ncu_full.log 2>&1
echo "ncu full rc=$?"; tail -3 gpurun_out/r02_ncu_full.log
ls -la gpurun_out | tail -15
