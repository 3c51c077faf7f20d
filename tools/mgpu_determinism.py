"""torchrun --nproc-per-node 2 tools/mgpu_determinism.py : config 3 on random data.  (1) Do GPU 0 and GPU 1 give the same
unsharded result?  (2) Does the sharded run reproduce it, step after step?  Prints the groups that differ."""
import os
import sys

import hashlib

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import sequencerj_b200 as s  # noqa: E402
from sequencerj_b200 import api, sharding  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
device = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=device)
N, L = 10000, 4096
scales, metrics = (1, 2, 4, 8, 16), s.ALL_METRICS
host = torch.empty((N, L), dtype=torch.float64, pin_memory=True)
host.numpy()[:] = np.maximum(np.random.default_rng(2732).random((N, L)), 2.220446049250313e-16)
dev = host.cuda()
At = host.numpy()
h = s.Handle(device=local)


def summ(r):
    return {(k[0].id, k[1]): (v[0], tuple(r.EOSeg[k][0]), hashlib.sha1(r.EOAlgScale[k][1].parents.tobytes()).hexdigest()) for k, v in r.EOAlgScale.items()}


def gather(obj):
    out = [None] * world
    dist.all_gather_object(out, obj)
    return out


ref = api._sequence(At, scales, metrics, h, device_ptr=dev.data_ptr())
refs = gather((summ(ref), float(ref.eta), hashlib.sha1(ref.order.tobytes()).hexdigest()))
if rank == 0:
    a, b = refs[0], refs[1]
    bad = [k for k in a[0] if a[0][k] != b[0][k]]
    print("unsharded GPU0 vs GPU1: groups that differ:", bad, "eta equal:", a[1] == b[1], "order equal:", a[2] == b[2], flush=True)
sref = refs[0][0]
for step in range(6):
    r = sharding.sequence_sharded(At, scales, metrics, rank, world, h, device_ptr=dev.data_ptr(), dist=dist, device=device)
    mine = summ(r)
    bad = [(k, mine[k][0] - sref[k][0], sum(1 for x, y in zip(mine[k][1], sref[k][1]) if x != y), mine[k][2] != sref[k][2])
           for k in mine if mine[k] != sref[k]]
    allbad = gather((rank, bad, float(r.eta) == refs[0][1], hashlib.sha1(r.order.tobytes()).hexdigest() == refs[0][2]))
    if rank == 0:
        print(f"sharded step {step}:", allbad, flush=True)
ref2 = api._sequence(At, scales, metrics, h, device_ptr=dev.data_ptr())
again = gather((rank, [k for k, v in summ(ref2).items() if v != sref[k]], float(ref2.eta) == refs[0][1]))
if rank == 0:
    print("unsharded again (after the sharded runs, workspace re-planned):", again, flush=True)
dist.destroy_process_group()
