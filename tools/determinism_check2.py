"""Fresh process: unsharded (A) -> two masked calls of the 2-rank plan (different workspace sizes) -> unsharded (B).
A and B must be identical; prints the groups that differ and whether sampled entries of their group matrices differ."""
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
rng = np.random.default_rng(3)
G = 20
pg = np.repeat(np.arange(G), 400).astype(np.int32)
pc = np.full(len(pg), -1, dtype=np.int32)
a, b = rng.integers(0, N, len(pg)), rng.integers(0, N, len(pg))
pi, pj = np.minimum(a, b).astype(np.int32), np.maximum(a, b).astype(np.int32)
probes = (pg, pc, pi, pj)


def summary(r):
    return {k: (v[0], tuple(np.asarray(r.EOSeg[k][0]).tolist()), r.EOAlgScale[k][1].parents.tobytes(),
                r.group_msts[k].weight.tobytes()) for k, v in r.EOAlgScale.items()}


A = api._sequence(At, scales, metrics, h, probes=probes)
owner = sharding.plan_groups(L, metrics, scales, 2)
for rank in range(2):
    api._sequence(At, scales, metrics, h, group_mask=(owner == rank).astype(np.uint8))
B = api._sequence(At, scales, metrics, h, probes=probes)
sa, sb = summary(A), summary(B)
for k in sa:
    g = metrics.index(k[0]) * len(scales) + scales.index(k[1])
    sel = pg == g
    nd = int(np.sum(A.probes[sel] != B.probes[sel]))
    flags = (sa[k][0] != sb[k][0], sa[k][1] != sb[k][1], sa[k][2] != sb[k][2], sa[k][3] != sb[k][3])
    if any(flags) or nd:
        print("DIFF", str(k[0]), k[1], "eta/chunk-eta/parents/mst-weights differ:", flags, "probe entries that differ:", nd,
              "max |d|:", float(np.max(np.abs(A.probes[sel] - B.probes[sel]))), flush=True)
print("env", {k: v for k, v in os.environ.items() if k.startswith("SEQJ_")}, "order identical:", np.array_equal(A.order, B.order),
      "eta identical:", A.eta == B.eta, flush=True)
