"""Multi-GPU sharding of the (metric, scale) groups, one process per GPU (group-major).

Every group is an independent unit: its chunk matrices, MSTs, eta values and the eta-weighted
combination stay on the GPU that owns it.  The only exchange is the per-group result
(eta_kl and the N-1 MST edges, ~160 KB at N = 10k) gathered to every rank for the final fuse
(`_weighted_p_matrix`, reference src/sequencer.jl:266-273).  No N x N matrix crosses NVLink.
"""
from __future__ import annotations

import numpy as np

from . import api

# relative FP64 issue cost per pair-element (EMD = 2 DADD, contractions = 1 DMMA lane-FMA)
METRIC_COST = {0: 1.0, 1: 2.0, 2: 1.0, 3: 1.0}
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


def assign_groups(costs: np.ndarray, world: int) -> np.ndarray:
    """Longest-processing-time greedy assignment of groups to ranks. Returns owner[k, l]."""
    owner = np.zeros(costs.shape, dtype=np.int64)
    load = np.zeros(world)
    flat = sorted(((costs[k, l], k, l) for k in range(costs.shape[0]) for l in range(costs.shape[1])),
                  key=lambda t: (-t[0], t[1], t[2]))
    for cst, k, l in flat:
        r = int(np.argmin(load))
        owner[k, l] = r
        load[r] += cst
    return owner


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


def sequence_sharded(At, scales, metrics, rank: int, world: int, handle=None, allgather=None, device_ptr=None):
    """`_sequence` over `world` ranks. `allgather(obj) -> list of obj` is the only communication
    (torch.distributed.all_gather_object over NCCL/gloo in practice).  Every rank returns the same
    final SequencerResult; per-group details are kept for the groups it owns."""
    N, M = At.shape
    scales = tuple(int(s) for s in scales)
    if world == 1:
        return api._sequence(At, scales, metrics, handle, device_ptr=device_ptr)
    owner = assign_groups(group_costs(M, metrics, scales), world)
    mask = (owner == rank).astype(np.uint8)
    mine = {}
    local = None
    if mask.any():
        local = api._sequence(At, scales, metrics, handle, group_mask=mask, device_ptr=device_ptr)
        for (m, s), (eta, _) in local.EOAlgScale.items():
            t = local.group_msts[(m, s)]
            mine[(m.id, s)] = (eta, t.src, t.dst, t.weight)
    parts = allgather(mine) if world > 1 else [mine]
    etas, msts = merge_group_results(parts, metrics, scales)
    final = api.fuse(N, etas, msts, handle=handle)
    if local is not None:
        final.EOSeg, final.EOAlgScale, final.group_msts = local.EOSeg, local.EOAlgScale, local.group_msts
        for k, v in local.timings.items():
            final.timings[k] = final.timings.get(k, 0.0) + v
        final.launches += local.launches
    return final
