"""Multi-GPU sharding of the (metric, scale) groups, one process per GPU (group-major).

Every group is an independent unit: its chunk matrices, MSTs, eta values and the eta-weighted
combination stay on the GPU that owns it.  The only exchange is the per-group result
(eta_kl and the N-1 MST edges, ~160 KB at N = 10k) gathered to every rank for the final fuse
(`_weighted_p_matrix`, reference src/sequencer.jl:266-273).  No N x N matrix crosses NVLink.
"""
from __future__ import annotations

import os
import time

import numpy as np

from . import api

# relative FP64 issue cost per pair-element (EMD = 2 DADD, contractions = 1 DMMA lane-FMA)
METRIC_COST = {0: 1.0, 1: 2.0, 2: 1.0, 3: 1.0, 4: 2.0, 5: 1.0, 6: 2.0}     # direct (FP64 ALU) tiles cost twice the DMMA ones
PRIM_COST_ROWS = 600.0     # one Prim row pass costs about as much as this many k-elements of distance work


def nchunks(M: int, scale: int) -> int:
    cl = -(-M // scale)
    return -(-M // cl)


def group_costs(M: int, metrics, scales) -> np.ndarray:
    """Estimated cost of each (metric, scale) group, metric-major, in units of N^2."""
    c = np.zeros((len(metrics), len(scales)))
    for k, m in enumerate(metrics):
        for l, s in enumerate(scales):
            c[k, l] = METRIC_COST[m.id] * M + PRIM_COST_ROWS * (nchunks(M, s) + 1)
    return c


def scales_chain_nested(M: int, scales) -> bool:
    """True when every scale's chunk length is a multiple of the next finer one's, i.e. the library can derive
    the contraction metrics of each scale from the previous one (csrc/kernels_derive.cuh)."""
    cls = sorted({-(-M // int(s)) for s in scales})
    return len(cls) > 1 and all(b % a == 0 for a, b in zip(cls, cls[1:]))


def assign_groups(costs: np.ndarray, world: int, bundles=None) -> np.ndarray:
    """Longest-processing-time greedy assignment of groups to ranks. Returns owner[k, l].
    `bundles`: optional list of (cost, [(k, l), ...]) units that must stay on one rank."""
    owner = np.zeros(costs.shape, dtype=np.int64)
    load = np.zeros(world)
    if bundles is None:
        bundles = [(costs[k, l], [(k, l)]) for k in range(costs.shape[0]) for l in range(costs.shape[1])]
    for cst, members in sorted(bundles, key=lambda t: (-t[0], t[1][0])):
        r = int(np.argmin(load))
        for k, l in members:
            owner[k, l] = r
        load[r] += cst
    return owner


def plan_groups(M: int, metrics, scales, world: int, with_load: bool = False):
    """Group-major plan, owner[k, l]: the library's cost-model LPT (seqj_plan_shards, csrc/host_shard.cuh -- pure
    host code, also used by the single-process multi-device handle).  EMD groups are individual units; a contraction
    metric over nested scales stays on one rank (only its finest scale runs DMMA tiles, the others are derived
    there) unless cutting the chain balances the ranks better."""
    from . import _lib
    lib = _lib.load()
    mids = np.array([m.id for m in metrics], dtype=np.int32)
    scs = np.array([int(s) for s in scales], dtype=np.int32)
    owner = np.zeros(len(mids) * len(scs), dtype=np.int32)
    load = np.zeros(world)
    st = lib.seqj_plan_shards(int(M), _lib.iptr(mids), len(mids), _lib.iptr(scs), len(scs), int(world), _lib.iptr(owner),
                              _lib.dptr(load))
    if st != 0:
        raise ValueError("seqj_plan_shards: bad arguments")
    owner = owner.reshape(len(mids), len(scs)).astype(np.int64)
    return (owner, load) if with_load else owner


def merge_group_results(parts, metrics, scales):
    """parts: list (one per rank) of {(metric_id, scale): (eta, src, dst, w)}. Returns (etas, msts) in the
    reference's group order (metric-major), which fixes the summation order of the fuse."""
    allg = {}
    for p in parts:
        allg.update(p)
    etas, msts = [], []
    for m in metrics:
        for s in scales:
            eta, src, dst, w = allg[(m.id, int(s))]
            etas.append(eta)
            msts.append(api.WeightedTree(len(src) + 1, np.asarray(src, dtype=np.int32), np.asarray(dst, dtype=np.int32),
                                         np.asarray(w, dtype=np.float64)))
    return etas, msts


def torch_allgather_groups(mine, owner, metrics, scales, N, rank, world, dist, device):
    """All-gather of the per-group (eta, MST edges) as fixed-shape tensors over torch.distributed
    (NCCL on GPUs): every rank contributes `cap` group slots (cap = max groups per rank), so no pickling
    and no size negotiation.  Returns the merged dict {(metric_id, scale): (eta, src, dst, w)}."""
    import torch
    cap = int(max(int((owner == r).sum()) for r in range(world)))
    ne = N - 1
    ints = torch.zeros((cap, 2 + 2 * ne), dtype=torch.int32)          # metric id, scale, src[ne], dst[ne]
    flts = torch.zeros((cap, 1 + ne), dtype=torch.float64)            # eta, w[ne]
    ints[:, 0] = -1
    for slot, ((mid, sc), (eta, src, dst, w)) in enumerate(sorted(mine.items())):
        ints[slot, 0], ints[slot, 1] = mid, sc
        ints[slot, 2:2 + ne] = torch.from_numpy(np.ascontiguousarray(src, dtype=np.int32))
        ints[slot, 2 + ne:] = torch.from_numpy(np.ascontiguousarray(dst, dtype=np.int32))
        flts[slot, 0] = eta
        flts[slot, 1:] = torch.from_numpy(np.ascontiguousarray(w, dtype=np.float64))
    ints_d, flts_d = ints.to(device), flts.to(device)
    all_i = torch.empty((world,) + tuple(ints.shape), dtype=torch.int32, device=device)
    all_f = torch.empty((world,) + tuple(flts.shape), dtype=torch.float64, device=device)
    dist.all_gather_into_tensor(all_i, ints_d)
    dist.all_gather_into_tensor(all_f, flts_d)
    all_i, all_f = all_i.cpu().numpy(), all_f.cpu().numpy()
    merged = {}
    for r in range(world):
        for slot in range(cap):
            mid = int(all_i[r, slot, 0])
            if mid < 0:
                continue
            merged[(mid, int(all_i[r, slot, 1]))] = (float(all_f[r, slot, 0]), all_i[r, slot, 2:2 + ne],
                                                     all_i[r, slot, 2 + ne:], all_f[r, slot, 1:])
    return merged


def slots_of_groups(owner: np.ndarray, cap: int) -> np.ndarray:
    """Slot of every group (metric-major) in the all-gathered exchange buffer [world][cap]: rank r's q-th group
    (metric-major among its own) sits in slot r * cap + q (the order seqj_sequence_shard packs them in)."""
    flat = owner.reshape(-1)
    seen = {}
    slots = np.zeros(flat.size, dtype=np.int32)
    for g, r in enumerate(flat.tolist()):
        slots[g] = r * cap + seen.get(r, 0)
        seen[r] = seen.get(r, 0) + 1
    return slots


def upload_sharded(At, rank: int, world: int, dist, device):
    """Every rank uploads 1/world of the objects over its own PCIe link and the ranks all-gather the slabs over
    NVLink, instead of `world` full copies of A crossing PCIe at once.  Returns the [>= N, M] device tensor."""
    import torch
    N, M = At.shape
    rows = -(-N // world)
    lo, hi = min(N, rank * rows), min(N, (rank + 1) * rows)
    shard = torch.zeros((rows, M), dtype=torch.float64, device=device)
    if hi > lo:
        shard[:hi - lo].copy_(torch.from_numpy(At[lo:hi]), non_blocking=True)
    full = torch.empty((world * rows, M), dtype=torch.float64, device=device)
    dist.all_gather_into_tensor(full, shard)
    return full


def sequence_sharded(At, scales, metrics, rank: int, world: int, handle=None, allgather=None, device_ptr=None,
                     gather_groups=None, dist=None, device=None):
    """`_sequence` over `world` ranks, group-major.  The only communication is the gather of per-group results.

    Device path (`dist` + `device` given; NCCL on GPUs, gloo in the CPU test): every rank packs its groups'
    (eta, MST edges) on the device (seqj_sequence_shard), ONE all_gather_into_tensor moves the slots over NVLink and
    every rank fuses straight from the gathered device buffer (seqj_fuse_dev): no host bounce between the last group
    kernel and the fuse.  Without `device_ptr` the input is uploaded 1/world per rank and all-gathered.

    Generic path (`gather_groups(mine, owner)` or `allgather(obj)`): the same exchange through host objects.
    Every rank returns the same final SequencerResult; per-group details are kept for the groups it owns.
    `timings` gains host-side wall-clock entries (host_local_ms, host_gather_ms, host_fuse_ms)."""
    N, M = At.shape
    scales = tuple(int(s) for s in scales)
    if world == 1:
        return api._sequence(At, scales, metrics, handle, device_ptr=device_ptr)
    owner = plan_groups(M, metrics, scales, world)
    mask = (owner == rank).astype(np.uint8)
    if dist is not None and device is not None:
        return _sequence_sharded_device(At, scales, metrics, rank, world, handle, device_ptr, dist, device, owner, mask)
    mine = {}
    local = None
    t0 = time.perf_counter()
    if mask.any():
        local = api._sequence(At, scales, metrics, handle, group_mask=mask, device_ptr=device_ptr)
        for (m, s), (eta, _) in local.EOAlgScale.items():
            t = local.group_msts[(m, s)]
            mine[(m.id, s)] = (eta, t.src, t.dst, t.weight)
    t1 = time.perf_counter()
    if gather_groups is not None:
        merged = gather_groups(mine, owner)
        etas, msts = merge_group_results([merged], metrics, scales)
    else:
        etas, msts = merge_group_results(allgather(mine), metrics, scales)
    t2 = time.perf_counter()
    final = api.fuse(N, etas, msts, handle=handle)
    t3 = time.perf_counter()
    return _finish_sharded(final, local, t0, t1, t2, t3)


def _finish_sharded(final, local, t0, t1, t2, t3, t_up=0.0):
    if local is not None:
        final.EOSeg, final.EOAlgScale, final.group_msts = local.EOSeg, local.EOAlgScale, local.group_msts
        for k, v in local.timings.items():
            final.timings[k] = final.timings.get(k, 0.0) + v
        final.launches += local.launches
    final.timings["host_upload_ms"] = t_up * 1e3
    final.timings["host_local_ms"] = (t1 - t0) * 1e3
    final.timings["host_gather_ms"] = (t2 - t1) * 1e3
    final.timings["host_fuse_ms"] = (t3 - t2) * 1e3
    return final


def _torch_cached_free(device) -> int:
    """Bytes torch's caching allocator holds without using them (it serves new tensors from there first)."""
    import torch
    return int(torch.cuda.memory_reserved(device) - torch.cuda.memory_allocated(device))


def _sequence_sharded_device(At, scales, metrics, rank, world, handle, device_ptr, dist, device, owner, mask):
    import torch
    N, M = At.shape
    G = owner.size
    cuda = getattr(device, "type", str(device)) == "cuda"

    def sync():
        if cuda:
            torch.cuda.current_stream(device).synchronize()      # the library runs on its own streams

    tu = time.perf_counter()
    keep = None
    if device_ptr is None:
        # The library keeps its workspace cached between calls and sizes it from the memory that was free when it was
        # planned; if the upload buffers do not fit next to it (config 4: 0.74 GB per rank), every rank gives the cache
        # back and the next call plans around them.  The decision is collective: the all-gather below needs every rank.
        if cuda:
            h = handle or api.default_handle()
            rows = -(-N // world)
            need = (rows + world * rows) * M * 8 + (16 << 20)         # shard + gathered copy (+ allocator slack)
            tight = torch.zeros(1, dtype=torch.int32, device=device)
            if torch.cuda.mem_get_info(device)[0] + _torch_cached_free(device) < need:
                tight += 1
            dist.all_reduce(tight, op=dist.ReduceOp.MAX)
            if int(tight.item()):
                h.check(h.lib.seqj_release_workspace(h.h))
            if os.environ.get("SEQJ_TRACE_SHARD"):
                print(f"[shard {rank}] free {torch.cuda.mem_get_info(device)[0] >> 20} MiB, torch cache {_torch_cached_free(device) >> 20} MiB, "
                      f"need {need >> 20} MiB, released {int(tight.item())}, {1e3 * (time.perf_counter() - tu):.1f} ms", flush=True)
        keep = upload_sharded(At, rank, world, dist, device)
        device_ptr = keep.data_ptr()
        if cuda and os.environ.get("SEQJ_TRACE_SHARD"):
            torch.cuda.synchronize(device)
            print(f"[shard {rank}] upload done at {1e3 * (time.perf_counter() - tu):.1f} ms", flush=True)
    cap = int(max(int((owner == r).sum()) for r in range(world)))
    stride = 2 * N                                                  # seqj_pack_stride(N)
    pack = torch.zeros((cap, stride), dtype=torch.float64, device=device)
    sync()
    t0 = time.perf_counter()
    local = None
    if mask.any():
        local = api._sequence(At, scales, metrics, handle, group_mask=mask, device_ptr=device_ptr, pack=pack)
    t1 = time.perf_counter()
    allp = torch.empty((world * cap, stride), dtype=torch.float64, device=device)
    dist.all_gather_into_tensor(allp, pack)
    sync()
    t2 = time.perf_counter()
    final = api.fuse_dev(N, G, allp, slots_of_groups(owner, cap), handle=handle)
    t3 = time.perf_counter()
    del keep
    return _finish_sharded(final, local, t0, t1, t2, t3, t_up=t0 - tu)


# ---- chunk-major sharding: all-gather of chunk etas + reduce of N x N partials ----
def chunk_lengths(M: int, scale: int):
    cl = -(-M // scale)
    return [min(cl, M - r0) for r0 in range(0, M, cl)]


def chunk_partition(M: int, metrics, scales, world: int):
    """Contiguous equal-cost split of the metric-major list of (metric, scale, chunk) tasks over `world` ranks.
    Returns (lo, hi): int arrays [world, G], rank r computes chunks [lo[r,g], hi[r,g]) of group g."""
    G = len(metrics) * len(scales)
    tasks = []
    for k, m in enumerate(metrics):
        for l, s in enumerate(scales):
            for c, mc in enumerate(chunk_lengths(M, s)):
                # Prim is one latency-bound launch per rank whatever the number of graphs, so a chunk's cost is its
                # distance work plus a small per-graph term
                tasks.append((k * len(scales) + l, c, METRIC_COST[m.id] * mc + 32.0))
    total = sum(t[2] for t in tasks)
    lo = np.zeros((world, G), dtype=np.int32)
    hi = np.zeros((world, G), dtype=np.int32)
    seen = np.zeros((world, G), dtype=bool)
    cum = 0.0
    for g, c, cost in tasks:
        r = min(world - 1, int((cum + 0.5 * cost) * world / total))
        if not seen[r, g]:
            lo[r, g] = c
            seen[r, g] = True
        hi[r, g] = c + 1
        cum += cost
    return lo, hi


def sequence_sharded_chunks(shape, scales, metrics, rank: int, world: int, handle, dist, device, device_ptr):
    """`_sequence` with the chunk tasks split evenly over the ranks.  Communication (torch.distributed, NCCL over
    NVLink): (1) all-gather of the chunk etas (a [G, max_chunks] table, all-reduce of disjoint entries), (2) for
    every group computed by more than one rank, a reduce (sum) of the eta-weighted N x N partials to the group's
    owner, in rank = chunk order, (3) the all-gather of per-group (eta, MST) before the fuse.  The summation
    order of step (2) differs from the reference's sequential loop by rounding only."""
    import torch
    N, M = shape
    scales = tuple(int(s) for s in scales)
    G = len(metrics) * len(scales)
    lo, hi = chunk_partition(M, metrics, scales, world)
    nch = [nchunks(M, s) for _ in metrics for s in scales]
    active = [g for g in range(G) if lo[rank, g] < hi[rank, g]]
    slot_of = {g: i for i, g in enumerate(active)}
    ldD = -(-N // 16) * 16
    stride = N * ldD
    t0 = time.perf_counter()
    S = torch.empty((max(1, len(active)), stride), dtype=torch.float64, device=device)
    local = api._sequence_partial((N, M), scales, metrics, lo[rank], hi[rank], S.data_ptr(), stride, device_ptr, handle)
    t1 = time.perf_counter()
    # (1) chunk etas
    eta = torch.zeros((G, max(nch)), dtype=torch.float64, device=device)
    for (m, s), (etas, _) in local.EOSeg.items():
        g = metrics.index(m) * len(scales) + scales.index(s)
        a, b = local.chunk_ranges[(m, s)]
        eta[g, a:b] = torch.tensor(etas, dtype=torch.float64)
    dist.all_reduce(eta)
    eta_h = eta.cpu().numpy()
    # (2) reduce the partial sums to the owners
    owned = []
    tmp = None
    for g in range(G):
        parts = [r for r in range(world) if lo[r, g] < hi[r, g]]
        owner = parts[0]
        if rank == owner:
            for r in parts[1:]:
                if tmp is None:
                    tmp = torch.empty(stride, dtype=torch.float64, device=device)
                dist.recv(tmp, src=r)
                S[slot_of[g]].add_(tmp)
            owned.append(g)
        elif rank in parts:
            dist.send(S[slot_of[g]], dst=owner)
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    # owners: Dkl = sum / sum(etas); group-level _measure_dm
    mine = {}
    if owned:
        sums = []
        for g in owned:
            tot = 0.0
            for c in range(nch[g]):
                tot += float(eta_h[g, c])
            sums.append(tot)
        res = api.groups_finish(S.data_ptr(), stride, [slot_of[g] for g in owned], N, sums, handle=handle)
        for g, (e, tree, mst) in zip(owned, res):
            m, s = metrics[g // len(scales)], scales[g % len(scales)]
            mine[(m.id, s)] = (e, mst.src, mst.dst, mst.weight)
            local.EOAlgScale[(m, s)] = (e, tree)
            local.group_msts[(m, s)] = mst
    t3 = time.perf_counter()
    owner_tbl = np.zeros((len(metrics), len(scales)), dtype=np.int64)
    for g in range(G):
        owner_tbl[g // len(scales), g % len(scales)] = [r for r in range(world) if lo[r, g] < hi[r, g]][0]
    merged = torch_allgather_groups(mine, owner_tbl, metrics, scales, N, rank, world, dist, device)
    etas, msts = merge_group_results([merged], metrics, scales)
    t4 = time.perf_counter()
    final = api.fuse(N, etas, msts, handle=handle)
    t5 = time.perf_counter()
    final.EOSeg, final.EOAlgScale, final.group_msts = local.EOSeg, local.EOAlgScale, local.group_msts
    final.chunk_ranges = local.chunk_ranges
    for k, v in local.timings.items():
        final.timings[k] = final.timings.get(k, 0.0) + v
    final.launches += local.launches
    final.timings.update(host_local_ms=(t1 - t0) * 1e3, host_reduce_ms=(t2 - t1) * 1e3, host_finish_ms=(t3 - t2) * 1e3,
                         host_gather_ms=(t4 - t3) * 1e3, host_fuse_ms=(t5 - t4) * 1e3)
    return final
