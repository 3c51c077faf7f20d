"""torchrun --nproc-per-node W tools/mgpu_check.py N L : group-major and chunk-major sharded sequence() must agree
with the single-GPU call (same order, eta; group etas within rounding)."""
import os, sys
import numpy as np
import torch
import torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import sequencerj_b200 as s
from sequencerj_b200 import api, sharding

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
N, L = int(sys.argv[1]), int(sys.argv[2])
scales, metrics = (1, 2, 4, 8), s.ALL_METRICS
A = np.maximum(np.random.default_rng(5).random((N, L)), 2.3e-16)
dev = torch.from_numpy(A).cuda()
h = s.Handle(device=local)
device = torch.device("cuda", local)
ref = api._sequence(A, scales, metrics, h)                      # every rank computes the single-GPU answer


def gather_groups(mine, owner):
    return sharding.torch_allgather_groups(mine, owner, metrics, scales, N, rank, world, dist, device)


rg = sharding.sequence_sharded(A, scales, metrics, rank, world, h, None, device_ptr=dev.data_ptr(), gather_groups=gather_groups)
rc = sharding.sequence_sharded_chunks((N, L), scales, metrics, rank, world, h, dist, device, dev.data_ptr())
ok = True
for name, r in (("group", rg), ("chunk", rc)):
    same = np.array_equal(r.order, ref.order) and abs(r.eta - ref.eta) <= 1e-12 * abs(ref.eta)
    worst = 0.0
    for key, (eta, tree) in r.EOAlgScale.items():
        e0, t0 = ref.EOAlgScale[key]
        worst = max(worst, abs(eta - e0) / abs(e0))
        same = same and tree.root == t0.root and np.array_equal(tree.parents, t0.parents)
    for key, (etas, trees) in r.EOSeg.items():
        a, b = r.chunk_ranges[key] if name == "chunk" else (0, len(etas))
        e0 = ref.EOSeg[key][0][a:b]
        same = same and np.allclose(etas, e0, rtol=1e-12)
    print(f"rank {rank} {name}-major: match={same} worst group-eta rel diff={worst:.2e} "
          f"timings={ {k: round(v, 1) for k, v in r.timings.items() if k.startswith('host')} }", flush=True)
    ok = ok and same
flag = torch.tensor([1 if ok else 0], device=device)
dist.all_reduce(flag, op=dist.ReduceOp.MIN)
dist.destroy_process_group()
sys.exit(0 if int(flag.item()) == 1 else 1)
