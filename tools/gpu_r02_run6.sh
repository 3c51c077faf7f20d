#!/bin/bash
# round 2, GPU call 6 (1 GPU): parity suite (Float32 path, prep flags), bench, rescan-budget sweep on structured data,
# second ncu --set full capture (MST / combine / prep / tree kernels), exported to CSV on the box
set -u
mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/r02_pytest.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/r02_pytest.log
python tools/mst_structured.py prim "" "budget=0" "budget=5000" "budget=10000" "budget=1000000000" > gpurun_out/r02_mst_structured.log 2>&1; echo "mst_structured rc=$?"; cat gpurun_out/r02_mst_structured.log
python bench.py --steps 10 --warmup 3 > gpurun_out/r02_bench_n1.json 2> gpurun_out/r02_bench_n1.err; echo "bench rc=$?"
python -c "import json;d=json.load(open('gpurun_out/r02_bench_n1.json'));print(d['ms_per_step'],d['e2e']['sequence_wall_s'],d['e2e']['pageable'],d['phases_ms_per_step'])"
PCMD="python tools/profile_case.py --N 10000 --L 4096 --scales 8 16"
$PCMD > gpurun_out/r02_plain2.log 2>&1 &&
timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base demangled \
  -k regex:"combine_sym_kernel|knn_rows_kernel<false>|knn_rows_kernel<\(bool\)0>|boruvka_knn_kernel|tree_kernel|prep_kernel|boruvka_sparse|root_tree" \
  -c 14 -o /tmp/r02_full2 $PCMD > gpurun_out/r02_ncu_full2.log 2>&1
echo "ncu full rc=$?"; tail -2 gpurun_out/r02_ncu_full2.log
if [ -f /tmp/r02_full2.ncu-rep ]; then
  ncu -i /tmp/r02_full2.ncu-rep --page raw --csv > gpurun_out/r02_full2_raw.csv 2>/dev/null
  ncu -i /tmp/r02_full2.ncu-rep --page source --csv --kernel-name-base demangled -k regex:knn_rows_kernel -c 1 > gpurun_out/r02_src_knn_rows.csv 2>/dev/null
  ls -la /tmp/r02_full2.ncu-rep
fi
du -sh gpurun_out
