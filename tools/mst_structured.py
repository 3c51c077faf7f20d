"""Prim phase of sequence() on STRUCTURED data at config-3 size, per MST mode (SEQJ_MST = prim | default | knn):
a smooth ordered family (the Sequencer's use case) and tightly clustered objects (the adversarial case for the
kNN lists: every list fills up with cluster mates, so whole clusters have to be rescanned).  GPU box only."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import sequencerj_b200 as s  # noqa: E402

N, L = 10000, 4096
MODES = sys.argv[1:] or ["prim", "", "knn"]          # "" = default policy; "knn,rounds=24" = knn everywhere, 24 launches
rng = np.random.default_rng(7)


def smooth():
    x = np.arange(L)[:, None]
    mu = np.linspace(0.15 * L, 0.85 * L, N)[None, :]
    A = np.exp(-0.5 * ((x - mu) / (0.05 * L)) ** 2) + 1e-3
    return np.asfortranarray(A[:, rng.permutation(N)])


def clustered(k=50):
    centers = rng.random((L, k))
    A = np.repeat(centers, N // k, axis=1) * (1.0 + 1e-3 * rng.standard_normal((L, N // k * k)))
    return np.asfortranarray(np.abs(A[:, rng.permutation(A.shape[1])]))


for name, A in (("smooth family", smooth()), ("50 tight clusters", clustered())):
    ref = None
    for mode in MODES:
        os.environ.pop("SEQJ_MST", None)
        os.environ.pop("SEQJ_KNN_ROUNDS", None)
        os.environ.pop("SEQJ_KNN_BUDGET", None)
        os.environ.pop("SEQJ_KNN", None)
        for part in filter(None, mode.split(",")):
            if part.startswith("rounds="):
                os.environ["SEQJ_KNN_ROUNDS"] = part[7:]
            elif part.startswith("lists="):           # first list pass: rows (full matrices) | tiles (upper triangle)
                os.environ["SEQJ_KNN"] = part[6:]
            elif part.startswith("budget="):          # rows a graph of a large batch may rescan before Prim takes it
                os.environ["SEQJ_KNN_BUDGET"] = part[7:]
            else:
                os.environ["SEQJ_MST"] = part
        for _ in range(2):
            r = s.sequence(A, scales=(1, 2, 4, 8, 16), metrics=s.ALL_METRICS)
        if ref is None:
            ref = r
        same = np.array_equal(r.order, ref.order) and r.eta == ref.eta and all(
            r.EOAlgScale[k][0] == ref.EOAlgScale[k][0] for k in ref.EOAlgScale)
        print(f"{name:18s} SEQJ_MST={mode or 'default':16s} total {r.timings['total']:.1f} ms prim {r.timings['prim']:.1f} ms "
              f"identical_to_prim={same}", flush=True)
