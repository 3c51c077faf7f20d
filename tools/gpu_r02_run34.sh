#!/bin/bash
# final 1-GPU pass of round 2: parity tests, smoke, bench, ncu launch list of the bench command, ncu --set full of the kernels of HEAD
set -u
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r02_pytest_final.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/r02_pytest_final.log
tail -5 gpurun_out/r02_pytest_final.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/r02_smoke.log
python bench.py > gpurun_out/r02_bench_final.json 2> gpurun_out/r02_bench_final.err; echo "bench rc=$?"
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r02_bench_reference.json 2> gpurun_out/r02_bench_reference.err; echo "reference arm rc=$?"; cut -c1-300 gpurun_out/r02_bench_reference.json
CMD="python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e"
BENCH_NO_SAMPLER=1 $CMD > gpurun_out/r02_plain.log 2>&1 &&
BENCH_NO_SAMPLER=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/r02_launches_config3.csv $CMD > gpurun_out/r02_ncu_list.log 2>&1
echo "ncu list rc=$?"
PCMD="python tools/profile_case.py --N 10000 --L 4096 --scales 1 2 4 8 16"
$PCMD > gpurun_out/r02_plain2.log 2>&1 &&
timeout 1500 ncu --set full --clock-control none --import-source on \
  -k regex:"direct_dist_kernel|gemm_dist_kernel|derive_kernel|combine_sym_kernel|knn_sample_kernel|knn_tiles_kernel|knn_merge_kernel|boruvka_knn_kernel|prep_kernel|tree_kernel|root_tree_kernel" \
  -c 60 -o /tmp/r02_full_final $PCMD > gpurun_out/r02_ncu_full_final.log 2>&1
echo "ncu full rc=$?"; tail -2 gpurun_out/r02_ncu_full_final.log
if [ -f /tmp/r02_full_final.ncu-rep ]; then
  ncu -i /tmp/r02_full_final.ncu-rep --page raw --csv > gpurun_out/r02_full_final_raw.csv 2>/dev/null
  ls -la /tmp/r02_full_final.ncu-rep gpurun_out/r02_full_final_raw.csv
fi
