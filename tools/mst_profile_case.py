"""One sequence() on the clustered / smooth structured inputs of tools/mst_structured.py, for an ncu launch list of
the MST kernels (which of list pass, rescans, Boruvka rounds, Prim takes the time).  python tools/mst_profile_case.py clusters|smooth"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import sequencerj_b200 as s  # noqa: E402

N, L = 10000, 4096
rng = np.random.default_rng(7)
kind = sys.argv[1] if len(sys.argv) > 1 else "clusters"
if kind == "random":
    A = np.asfortranarray(np.maximum(rng.random((L, N)), 2.220446049250313e-16))
elif kind == "smooth":
    x = np.arange(L)[:, None]
    mu = np.linspace(0.15 * L, 0.85 * L, N)[None, :]
    A = np.exp(-0.5 * ((x - mu) / (0.05 * L)) ** 2) + 1e-3
    A = np.asfortranarray(A[:, rng.permutation(N)])
else:
    k = 50
    centers = rng.random((L, k))
    A = np.repeat(centers, N // k, axis=1) * (1.0 + 1e-3 * rng.standard_normal((L, N // k * k)))
    A = np.asfortranarray(np.abs(A[:, rng.permutation(A.shape[1])]))
r = s.sequence(A, scales=(1, 2, 4, 8, 16), metrics=s.ALL_METRICS)
print(kind, "total", round(r.timings["total"], 1), "prim", round(r.timings["prim"], 1))
print({k: round(v, 1) for k, v in r.timings.items()})
