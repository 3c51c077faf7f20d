"""Worker of tests/test_sharding_gloo.py: one rank of a world_size-2 gloo job.  The GPU calls of the product
(`api._sequence`, `api.fuse`) are replaced by oracle-backed fakes so that the HOST logic of the multi-GPU
path (group assignment, masks, all-gather, metric-major merge, fuse inputs) runs on CPU."""
import os
import sys
import types

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch.distributed as dist
    import sequencerj_b200 as s
    from sequencerj_b200 import api, sharding
    from oracle import seq_oracle as O

    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    dist.init_process_group("gloo", rank=rank, world_size=world)
    A = np.random.default_rng(77).random((40, 60))               # M x N
    At = np.ascontiguousarray(A.T)
    metrics, scales = s.ALL_METRICS, (1, 2, 4)
    Ac = O.clamp_input(A)
    calls = {"seq": 0, "fuse": 0, "mask": None}

    def fake_sequence(At_, scales_, metrics_, handle=None, group_mask=None, device_ptr=None, probes=None, pack=None):
        calls["seq"] += 1
        calls["mask"] = None if group_mask is None else np.array(group_mask).reshape(len(metrics_), len(scales_))
        r = types.SimpleNamespace(EOSeg={}, EOAlgScale={}, group_msts={}, timings={}, launches=1)
        for k, m in enumerate(metrics_):
            for l, sc in enumerate(scales_):
                if group_mask is not None and not calls["mask"][k, l]:
                    continue
                g, _ = O.compute_group(Ac, m.id, sc)
                r.EOAlgScale[(m, sc)] = (g.eta, api.BFSTree(g.root, g.parents))
                r.EOSeg[(m, sc)] = (g.eta_chunk, g.parents_chunk)
                r.group_msts[(m, sc)] = api.WeightedTree(60, g.mst[0].astype(np.int32), g.mst[1].astype(np.int32), g.mst[2])
                if pack is not None:              # the device exchange format of seqj_sequence_shard, on a CPU tensor
                    Nn = 60
                    slot = pack[len(r.group_msts) - 1].numpy()
                    slot[0] = g.eta
                    slot[1:Nn] = g.mst[2]
                    ij = slot[Nn:2 * Nn - 1].view(np.int32).reshape(Nn - 1, 2)
                    ij[:, 0], ij[:, 1] = g.mst[0], g.mst[1]
        return r

    def fake_fuse_dev(N, G, allp, slots, eps_floor=1e-6, handle=None):
        calls["fuse"] += 1
        groups = []
        for g_ in range(G):
            slot = allp[int(slots[g_])].numpy()
            ij = slot[N:2 * N - 1].view(np.int32).reshape(N - 1, 2)
            groups.append(types.SimpleNamespace(eta=float(slot[0]), mst=(ij[:, 0].astype(np.int64), ij[:, 1].astype(np.int64),
                                                                        slot[1:N].copy())))
        f = O.fuse_groups(N, groups)
        return types.SimpleNamespace(eta=f.eta, order=f.order, root=f.root, timings={}, launches=1, EOSeg={},
                                     EOAlgScale={}, group_msts={})

    def fake_fuse(N, etas, msts, eps_floor=1e-6, handle=None):
        calls["fuse"] += 1
        groups = [types.SimpleNamespace(eta=e, mst=(t.src.astype(np.int64), t.dst.astype(np.int64), t.weight))
                  for e, t in zip(etas, msts)]
        f = O.fuse_groups(N, groups)
        return types.SimpleNamespace(eta=f.eta, order=f.order, root=f.root, timings={}, launches=1, EOSeg={},
                                     EOAlgScale={}, group_msts={})

    api._sequence, api.fuse, api.fuse_dev = fake_sequence, fake_fuse, fake_fuse_dev

    def allgather(obj):
        out = [None] * world
        dist.all_gather_object(out, obj)
        return out

    r = sharding.sequence_sharded(At, scales, metrics, rank, world, None, allgather)
    ref = O.sequence_oracle(A, scales)
    owner = sharding.plan_groups(40, metrics, scales, world)
    ok = (np.array_equal(r.order, ref.order) and r.eta == ref.eta and calls["seq"] == 1 and calls["fuse"] == 1
          and np.array_equal(calls["mask"], (owner == rank).astype(np.uint8))
          and len(r.EOAlgScale) == int((owner == rank).sum()))
    # device path of the same driver (packed slots, ONE all_gather_into_tensor, fuse from the gathered buffer) on CPU
    # tensors over gloo; the input is "uploaded" 1/world per rank and all-gathered
    import torch
    r2 = sharding.sequence_sharded(At, scales, metrics, rank, world, None, dist=dist, device=torch.device("cpu"))
    ok = ok and np.array_equal(r2.order, ref.order) and r2.eta == ref.eta and calls["seq"] == 2 and calls["fuse"] == 2
    up = sharding.upload_sharded(At, rank, world, dist, torch.device("cpu"))
    ok = ok and np.array_equal(up.numpy()[:At.shape[0]], At)
    slots = sharding.slots_of_groups(owner, int(max((owner == q).sum() for q in range(world))))
    ok = ok and len(set(slots.tolist())) == owner.size
    flags = [None] * world
    dist.all_gather_object(flags, bool(ok))
    dist.destroy_process_group()
    if rank == 0:
        with open(os.environ["SEQJ_TEST_OUT"], "w") as f:
            f.write("OK" if all(flags) else f"FAIL {flags}")
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
