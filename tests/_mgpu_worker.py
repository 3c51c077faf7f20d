"""Worker of tests/test_gpu_multi.py (torchrun --nproc-per-node W tests/_mgpu_worker.py N L): the group-major sharded
sequence() -- device-side exchange (seqj_sequence_shard + all_gather_into_tensor + seqj_fuse_dev), with a resident
input and with the sharded upload -- and the chunk-major one (NCCL send/recv reduce of the N x N partials) must agree
with the single-GPU call on every rank: same order, eta, group etas / trees, chunk etas."""
import os
import sys

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
N, L = int(sys.argv[1]), int(sys.argv[2])
scales = tuple(int(x) for x in sys.argv[3].split(",")) if len(sys.argv) > 3 else (1, 2, 4, 8)
metrics = s.ALL_METRICS
A = np.maximum(np.random.default_rng(5).random((N, L)), 2.3e-16)
dev = torch.from_numpy(A).to(device)
h = s.Handle(device=local)
ref = api._sequence(A, scales, metrics, h)                      # every rank computes the single-GPU answer

runs = {
    "group/resident": lambda: sharding.sequence_sharded(A, scales, metrics, rank, world, h, device_ptr=dev.data_ptr(), dist=dist, device=device),
    "group/sharded-upload": lambda: sharding.sequence_sharded(A, scales, metrics, rank, world, h, dist=dist, device=device),
    "chunk": lambda: sharding.sequence_sharded_chunks((N, L), scales, metrics, rank, world, h, dist, device, dev.data_ptr()),
}
ok = True
for name, fn in runs.items():
    r = fn()
    same = np.array_equal(r.order, ref.order) and abs(r.eta - ref.eta) <= 1e-12 * abs(ref.eta)
    worst = 0.0
    for key, (eta, tree) in r.EOAlgScale.items():
        e0, t0 = ref.EOAlgScale[key]
        worst = max(worst, abs(eta - e0) / abs(e0))
        same = same and tree.root == t0.root and np.array_equal(tree.parents, t0.parents)
    for key, (etas, trees) in r.EOSeg.items():
        a, b = r.chunk_ranges[key] if name == "chunk" else (0, len(etas))
        same = same and np.allclose(etas, ref.EOSeg[key][0][a:b], rtol=1e-12)
    if name != "chunk":                                          # group-major runs whole groups: the same trees, so the
        same = same and worst == 0.0 and r.eta == ref.eta        # (integer-derived) etas are bit-identical
    print(f"rank {rank} {name}: match={same} worst group-eta rel diff={worst:.2e} groups here={len(r.EOAlgScale)} "
          f"timings={ {k: round(v, 1) for k, v in r.timings.items() if k.startswith('host')} }", flush=True)
    ok = ok and same
flag = torch.tensor([1 if ok else 0], device=device)
dist.all_reduce(flag, op=dist.ReduceOp.MIN)
dist.destroy_process_group()
sys.exit(0 if int(flag.item()) == 1 else 1)
