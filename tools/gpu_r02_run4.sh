#!/bin/bash
# round 2, GPU call 4 (2 GPUs): full parity suite incl. the multi-GPU tests, structured-data MST timing, bench at 1 and 2
# GPUs (torchrun), single-process 2-device handle
set -u
mkdir -p gpurun_out
nvidia-smi -L
python -m pytest tests -m gpu -q > gpurun_out/r02_pytest_2gpu.log 2>&1; echo "pytest rc=$?"; tail -8 gpurun_out/r02_pytest_2gpu.log
python tools/mst_structured.py prim "" > gpurun_out/r02_mst_structured.log 2>&1; echo "mst_structured rc=$?"; cat gpurun_out/r02_mst_structured.log
python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r02_bench_n1b.json 2> gpurun_out/r02_bench_n1b.err; echo "bench n1 rc=$?"
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r02_bench_n2.json 2> gpurun_out/r02_bench_n2.err; echo "bench n2 rc=$?"; tail -c 800 gpurun_out/r02_bench_n2.err
python -c "
import json
for f in ('r02_bench_n1b','r02_bench_n2'):
    d=json.load(open('gpurun_out/'+f+'.json')); print(f, d['ms_per_step'], d['e2e']['sequence_wall_s'] if d.get('e2e') else None, d.get('parity_vs_single_gpu'), d.get('phases_ms_per_step_rank_max'))
"
python tools/multidev_check.py 2 10000 4096 > gpurun_out/r02_multidev2.log 2>&1; echo "multidev rc=$?"; cat gpurun_out/r02_multidev2.log
du -sh gpurun_out
