#!/bin/bash
# round 2, GPU call 3 (1 GPU): full parity suite on the planner build, the other BASELINE configs, structured-data MST
# timing, and a 12-launch ncu --set full capture exported to CSV on the box (the report itself is too large to bring back)
set -u
mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/r02_pytest.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/r02_pytest.log
python tools/mst_structured.py "" > gpurun_out/r02_mst_structured.log 2>&1; echo "mst_structured rc=$?"; cat gpurun_out/r02_mst_structured.log
for cfg in 2 4 5; do
  python bench.py --config $cfg --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r02_bench_cfg$cfg.json 2> gpurun_out/r02_bench_cfg$cfg.err; echo "bench cfg $cfg rc=$?"
  python -c "import json;d=json.load(open('gpurun_out/r02_bench_cfg$cfg.json'));print(d['ms_per_step'],d['phases_ms_per_step'])"
done
PCMD="python tools/profile_case.py --N 10000 --L 4096 --scales 8 16"
$PCMD > gpurun_out/r02_plain2.log 2>&1 &&
timeout 600 ncu --set full --clock-control none --import-source on \
  -k regex:"direct_dist_kernel|gemm_dist_kernel|derive_kernel|combine_sym_kernel|knn_rows_kernelILb0" \
  -c 12 -o /tmp/r02_full $PCMD > gpurun_out/r02_ncu_full.log 2>&1
echo "ncu full rc=$?"; tail -2 gpurun_out/r02_ncu_full.log
if [ -f /tmp/r02_full.ncu-rep ]; then
  ncu -i /tmp/r02_full.ncu-rep --page raw --csv > gpurun_out/r02_full_raw.csv 2>/dev/null
  ncu -i /tmp/r02_full.ncu-rep --page details --csv > gpurun_out/r02_full_details.csv 2>/dev/null
  ncu -i /tmp/r02_full.ncu-rep --page source --csv -k regex:direct_dist_kernel -c 1 > gpurun_out/r02_src_emd.csv 2>/dev/null
  ncu -i /tmp/r02_full.ncu-rep --page source --csv -k regex:knn_rows_kernel -c 1 > gpurun_out/r02_src_knn_rows.csv 2>/dev/null
  ls -la /tmp/r02_full.ncu-rep
fi
du -sh gpurun_out; ls -la gpurun_out | tail -12
