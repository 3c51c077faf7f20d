#!/bin/bash
# 8 GPUs: final scaling numbers of config 3 (4 / 8 GPUs, torchrun), config 4 on 8 GPUs, single-process 8-device handle
set -u
mkdir -p gpurun_out
run() { # ngpu config steps tag
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $1 --master-addr 127.0.0.1 --master-port $((29540 + $1 + $2)) bench.py --gpus $1 --config $2 --steps $3 --warmup 3 --no-cpu-baseline > gpurun_out/r02_bench_$4.json 2> gpurun_out/r02_bench_$4.err
  echo "bench $4 rc=$?"
  python - <<PY
import json
txt=open('gpurun_out/r02_bench_$4.json').read()
ls=[l for l in txt.splitlines() if l.startswith('{')]
if ls:
    d=json.loads(ls[-1]); print('$4', round(d['ms_per_step'],2), 'e2e', round(d['e2e']['sequence_wall_s']*1e3,2), d['e2e'].get('wall_ms_each_step'), d.get('parity_vs_single_gpu',{}).get('ok'), d.get('wall_ms_each_step'), d.get('phases_ms_per_step_rank_max'))
else:
    print(open('gpurun_out/r02_bench_$4.err').read()[-1500:])
PY
}
run 8 3 20 cfg3_n8
run 4 3 12 cfg3_n4
run 8 4 3 cfg4_n8
python tools/multidev_check.py 8 10000 4096 > gpurun_out/r02_multidev8.log 2>&1; echo "multidev8 rc=$?"; cat gpurun_out/r02_multidev8.log
python tools/multidev_check.py 4 10000 4096 > gpurun_out/r02_multidev4.log 2>&1; echo "multidev4 rc=$?"; cat gpurun_out/r02_multidev4.log
