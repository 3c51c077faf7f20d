#!/bin/bash
# 2 GPUs: multi-GPU pytest, sharded determinism, bench N=2
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_multi.py -x -q -m gpu > gpurun_out/r02_pytest_2gpu.log 2>&1; echo "pytest multi rc=$?"; tail -4 gpurun_out/r02_pytest_2gpu.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29577 tools/mgpu_determinism.py > gpurun_out/r02_mgpu_determinism.log 2>&1; echo "determinism rc=$?"; grep -v "OMP_NUM\|^\*\*\*\*\|^$" gpurun_out/r02_mgpu_determinism.log | cut -c1-400
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29545 bench.py --gpus 2 --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/r02_bench_cfg3_n2.json 2> gpurun_out/r02_bench_cfg3_n2.err; echo "bench n2 rc=$?"
python - <<PY
import json
ls=[l for l in open('gpurun_out/r02_bench_cfg3_n2.json').read().splitlines() if l.startswith('{')]
d=json.loads(ls[-1]); print(round(d['ms_per_step'],2), 'e2e', round(d['e2e']['sequence_wall_s']*1e3,2), d.get('parity_vs_single_gpu'), d.get('wall_ms_each_step'), d.get('phases_ms_per_step_rank_max'), d['clocks'])
PY
