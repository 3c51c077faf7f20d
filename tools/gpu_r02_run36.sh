#!/bin/bash
# 2 GPUs: config 4 (streamed + sharded) with the sharded upload in the timed region (the 8-GPU run of the day ran out of memory there)
set -u
mkdir -p gpurun_out
SEQJ_TRACE_SHARD=1 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29549 bench.py --gpus 2 --config 4 --steps 3 --warmup 2 --no-cpu-baseline > gpurun_out/r02_bench_cfg4_n2.json 2> gpurun_out/r02_bench_cfg4_n2.err; echo "bench cfg4 n2 rc=$?"
python - <<PY
import json
ls=[l for l in open('gpurun_out/r02_bench_cfg4_n2.json').read().splitlines() if l.startswith('{')]
if ls:
    d=json.loads(ls[-1]); print(round(d['ms_per_step'],1), d['wall_ms_each_step'], 'e2e', d['e2e']['wall_ms_each_step'], d.get('parity_vs_single_gpu'), d.get('phases_ms_per_step_rank_max'))
else:
    print(open('gpurun_out/r02_bench_cfg4_n2.err').read()[-2000:])
PY
grep "shard 0\|shard 1" gpurun_out/r02_bench_cfg4_n2.json gpurun_out/r02_bench_cfg4_n2.err | cut -c1-200 | head -40
