"""Config 3 on random data: (1) is the unsharded call bit-reproducible run to run, (2) do the two masked calls of the
2-rank plan (run one after the other on ONE GPU) reproduce its per-group results exactly?  Prints which groups differ."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import sequencerj_b200 as s  # noqa: E402
from sequencerj_b200 import api, sharding  # noqa: E402

N, L = 10000, 4096
scales, metrics = (1, 2, 4, 8, 16), s.ALL_METRICS
At = np.maximum(np.random.default_rng(2732).random((N, L)), 2.220446049250313e-16)
h = s.default_handle()


def summary(r):
    return {k: (v[0], tuple(np.asarray(r.EOSeg[k][0]).tolist()), r.EOAlgScale[k][1].parents.tobytes()) for k, v in r.EOAlgScale.items()}


def diff(a, b, tag):
    bad = []
    for k in a:
        if k not in b:
            continue
        if a[k][0] != b[k][0] or a[k][1] != b[k][1] or a[k][2] != b[k][2]:
            ce = sum(1 for x, y in zip(a[k][1], b[k][1]) if x != y)
            bad.append((str(k[0]), k[1], a[k][0] - b[k][0], ce, a[k][2] != b[k][2]))
    print(tag, "groups that differ:", bad if bad else "none", flush=True)
    return bad


ref = api._sequence(At, scales, metrics, h)
sref = summary(ref)
for rep in range(3):
    r = api._sequence(At, scales, metrics, h)
    diff(sref, summary(r), f"unsharded run {rep + 1} vs run 0:")
    print("   order identical:", np.array_equal(r.order, ref.order), "eta identical:", r.eta == ref.eta)
for world in (2, 4):
    owner = sharding.plan_groups(L, metrics, scales, world)
    for rank in range(world):
        mask = (owner == rank).astype(np.uint8)
        for rep in range(2):
            p = api._sequence(At, scales, metrics, h, group_mask=mask)
            diff(summary(p), sref, f"world {world} rank {rank} rep {rep} vs unsharded:")
