"""One sequence() call of a given shape, for ncu captures (see profiles/)."""
import argparse
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import sequencerj_b200 as s  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--N", type=int, default=4096)
ap.add_argument("--L", type=int, default=2048)
ap.add_argument("--scales", type=int, nargs="+", default=[1, 8])
ap.add_argument("--metrics", type=int, nargs="+", default=[0, 1, 2, 3])
ap.add_argument("--reps", type=int, default=1)
a = ap.parse_args()
A = np.random.default_rng(2732).random((a.L, a.N))
for _ in range(a.reps):
    r = s.sequence(A, scales=tuple(a.scales), metrics=tuple(s.ALL_METRICS[i] for i in a.metrics))
print("eta", r.eta, {k: round(v, 3) for k, v in r.timings.items()})
