"""kNN-Boruvka (SEQJ_MST=knn) against the Prim kernels (SEQJ_MST=prim) through seqj_measure_dm: the two must return
the same tree bit for bit (edge set, weights, root, eta, BFS parents, order) on random, tied, clustered and smooth
inputs.  Run on the GPU box: python tools/mst_check.py"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import sequencerj_b200 as s  # noqa: E402

rng = np.random.default_rng(2732)


def sym(W):
    W = np.triu(W, 1)
    return W + W.T


def cases():
    for N in (2, 3, 5, 17, 64, 100, 257, 1000, 1031, 2500, 4099, 10000):
        yield f"random N={N}", sym(rng.random((N, N)) + 0.01)
    for N, lv in ((7, 2), (300, 4), (1031, 3), (2500, 5), (4099, 2)):
        yield f"integer ties N={N} levels={lv}", sym(rng.integers(1, lv + 1, (N, N)).astype(np.float64))
    yield "all equal N=500", sym(np.full((500, 500), 1e-6))
    for N, k in ((1500, 5), (3000, 40)):
        pts = np.concatenate([rng.normal(3.0 * c, 0.01, (6, N // k)) for c in range(k)], axis=1)
        G = pts.T @ pts
        sq = np.diag(G)
        yield f"{k} tight clusters N={N}", sym(np.sqrt(np.maximum(sq[:, None] + sq[None, :] - 2 * G, 0.0)) + 1e-6)
    N = 3000
    x = np.arange(128)[:, None]
    mu = np.linspace(15, 113, N)[None, :]
    F = np.exp(-0.5 * ((x - mu) / 8.0) ** 2) + 1e-3
    F = (F / F.sum(0))[:, rng.permutation(N)]
    U = np.cumsum(F, axis=0)
    blocks = [np.abs(U[:, i:i + 250, None] - U[:, None, :]).sum(0) for i in range(0, N, 250)]
    yield f"smooth family EMD N={N}", sym(np.concatenate(blocks, axis=0) + 1e-6)
    base = rng.random((900, 900)) + 0.01
    yield "duplicated objects N=2700", sym(np.kron(np.ones((3, 3)), sym(base)) + 1e-6)


def run(mode, D):
    os.environ["SEQJ_MST"] = mode
    t0 = time.perf_counter()
    m = s.measure_dm(D, want_order=True)
    return m, time.perf_counter() - t0


bad = 0
for name, D in cases():
    a, ta = run("prim", D)
    b, tb = run("knn", D)
    ea = sorted(zip(a.mst.src.tolist(), a.mst.dst.tolist(), a.mst.weight.view(np.uint64).tolist()))
    eb = sorted(zip(b.mst.src.tolist(), b.mst.dst.tolist(), b.mst.weight.view(np.uint64).tolist()))
    same = (ea == eb and a.root == b.root and a.eta == b.eta and np.array_equal(a.bfs.parents, b.bfs.parents)
            and np.array_equal(a.order, b.order))
    print(f"{'ok ' if same else 'BAD'} {name}: eta={a.eta:.6g}/{b.eta:.6g} root={a.root}/{b.root} "
          f"edges_equal={ea == eb} wall prim {ta * 1e3:.1f} ms knn {tb * 1e3:.1f} ms", flush=True)
    bad += not same
os.environ.pop("SEQJ_MST", None)
print("FAILED" if bad else "ALL OK", bad)
sys.exit(1 if bad else 0)
