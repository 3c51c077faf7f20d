"""CPU restatement (ORACLE) of SequencerJ's hot path -- TEST INFRASTRUCTURE ONLY.

This module is the checker, never the product: only ``tests/``, ``__graft_entry__.smoke()``
and ``bench.py``'s CPU-baseline / ``--impl reference`` legs may import it.  The product path
(``sequencerj.jl_b200``) never imports anything from ``oracle/``.

It restates, in NumPy/SciPy, the algorithm of ``/root/reference`` (pure Julia; Julia is not
installed in this image, so the reference itself cannot be executed here):

* ``src/sequencer.jl:184-276``  (_sequence: metric x scale x chunk loop, eta-weighted combine, fuse)
* ``src/sequencer.jl:284-290,331-374`` (_measure_dm, _leastcentralpt, elongation, _mst)
* ``src/sequencer.jl:383-407``  (_splitnorm), ``:310-328`` (_weighted_p_matrix)
* ``src/distancemetrics.jl:46-50,82-93,143-162,191-223`` (EMD, Energy, _cdf_distance)
* ``src/utils.jl:15-24`` (unroll), ``src/bfs.jl:6-9`` (bfs tree)

The arithmetic of several steps lives in un-vendored Julia packages (no Manifest.toml in the
reference; compat ranges only, ``Project.toml:20-28``).  Their published algorithms are restated:

* Distances.jl 0.8-0.10: ``pairwise`` (SemiMetric: i>j + mirror, diag 0; PreMetric: full),
  ``Euclidean(thresh)`` (Gram trick + direct recompute when v < thresh*(|a|^2+|b|^2)),
  ``KLDivergence`` (sum a*log(a/b) over a>0).
* StatsBase.jl 0.33: ``ecdf(x; weights)`` (right-continuous weighted ECDF, sequential running
  sum divided by the cached weight sum; values >= last grid point evaluate to exactly 1).
* LightGraphs.jl 1.3: ``kruskal_mst`` (stable sort of the upper-triangle edges enumerated in
  column-major order, union-find), ``closeness_centrality`` (weighted Dijkstra from each vertex),
  ``dijkstra_shortest_paths(g, s, DefaultDistance())`` (hop counts), ``bfs_tree``.

Parity pins (see tests/test_oracle_golden.py): every known answer in the reference's
``test/algotests.jl`` (Euclidean sqrt(14), KL 0.5108256237659907, EMD 4.0781331438047861, Energy
1.0 / 2.3674181638546394 / 0.8800334097615822) and the reference's only exact end-to-end test
(``test/runtests.jl:167-175``: shuffled Hummingbird columns are recovered exactly).  MST edges,
root index, eta values and order vectors are pinned by NO reference test: for those this oracle
is "parity unpinned" beyond independent SciPy/NetworkX cross-checks.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field

import numpy as np

EPS_FLOOR = 1e-6        # distancemetrics.jl:3   (const ϵ)
L2THR = 1e-7            # distancemetrics.jl:6
JULIA_EPS = 2.220446049250313e-16   # eps(Float64), sequencer.jl:130
FLOATMAX32 = 3.4028234663852886e38  # floatmax(Float32), sequencer.jl:351-352

# metric ids shared with the C ABI (include/seqj.h)
L2, EMD, KLD, ENERGY = 0, 1, 2, 3
EMD_GRID, ENERGY_GRID = 4, 5                   # the TYPES EMD / Energy + explicit grid (sequencer.jl:217-219)
PERIODIC = 6                                   # Distances.PeriodicEuclidean(periods) (test/runtests.jl:189-236)
# further element-wise Distances.jl metrics a caller may pass (`pairwise` takes any PreMetric, sequencer.jl:226)
CITYBLOCK, TOTALVARIATION, CHEBYSHEV, SQEUCLIDEAN = 7, 8, 9, 10
ALL_METRICS = (L2, EMD, KLD, ENERGY)           # distancemetrics.jl:257 (order matters)
METRIC_NAMES = {L2: "Euclidean", EMD: "EMD", KLD: "KLDivergence", ENERGY: "Energy"}
FIB = [1, 2, 3, 5, 8, 13, 21, 34, 55, 89, 144, 233, 377, 610, 987, 1597, 2584, 5168]  # sequencer.jl:148


# --------------------------------------------------------------------------------------
# pair-level metrics (faithful; used for the algotests known answers and small-N checks)
# --------------------------------------------------------------------------------------
class ECDF:
    """StatsBase.ecdf(x; weights) restated (sorted values + weights with cached sum)."""

    def __init__(self, x, weights):
        x = np.asarray(x, dtype=np.float64)
        w = np.asarray(weights, dtype=np.float64)
        ordr = np.argsort(x, kind="stable")
        self.sorted_values = x[ordr]
        self.weights = w[ordr]
        self.wsum = float(np.sum(self.weights))      # Weights(...).sum (cached)

    def __call__(self, v):
        v = np.asarray(v, dtype=np.float64)
        ordr = np.argsort(v, kind="stable")
        m = len(v)
        r = np.empty(m, dtype=np.float64)
        r0 = 0.0
        i = 0
        done = False
        for j, x in enumerate(self.sorted_values):
            while i < m and x > v[ordr[i]]:
                r[ordr[i]] = r0
                i += 1
            r0 += self.weights[j]
            if i >= m:
                done = True
                break
        while i < m:
            r[ordr[i]] = self.wsum
            i += 1
        return r / self.wsum


def cdf_distance(u: ECDF, v: ECDF, p: int = 1) -> float:
    """distancemetrics.jl:191-223 (_cdf_distance)."""
    uv, vv = u.sorted_values, v.sorted_values
    uplusv = np.sort(np.concatenate(([0.0], uv, vv)))          # :199-200
    dx = np.diff(uplusv)                                       # :202
    if len(uv) == len(vv) and np.array_equal(uv, vv):          # :206 same-grid shortcut
        dx = dx[dx != 0.0]
        UA, VA = u(uv), v(vv)
    else:                                                      # :210-214
        A = uplusv[:-1]
        UA, VA = u(A), v(A)
    if p == 1:
        return float(np.sum(np.abs(UA - VA) * dx))             # :217
    if p == 2:
        return float(np.sqrt(np.sum((UA - VA) ** 2 * dx)))     # :219
    return float(np.sum(np.abs(UA - VA) ** p * dx) ** 1 / p)   # :221 (precedence quirk kept)


def emd(u, v, uw=None, vw=None) -> float:
    """EMD(u,v) distancemetrics.jl:46-50 and EMD(u,v,uw,vw) :82-87."""
    if uw is None:
        uw, vw = u, v
        u = np.arange(1, len(uw) + 1, dtype=np.float64)
        v = np.arange(1, len(vw) + 1, dtype=np.float64)
    if len(u) != len(uw) or len(v) != len(vw):
        raise ValueError("u and u weights must have same length")
    return cdf_distance(ECDF(u, uw), ECDF(v, vw), 1)


def energy(u, v, uw=None, vw=None) -> float:
    """Energy(u,v) distancemetrics.jl:158-162 and Energy(u,v,uw,vw) :143-148."""
    if uw is None:
        uw, vw = u, v
        u = np.arange(1, len(uw) + 1, dtype=np.float64)
        v = np.arange(1, len(vw) + 1, dtype=np.float64)
    if len(u) != len(uw) or len(v) != len(vw):
        raise ValueError("u and u weights must have same length")
    return math.sqrt(2.0) * cdf_distance(ECDF(u, uw), ECDF(v, vw), 2)


def euclidean(a, b) -> float:
    a = np.asarray(a, dtype=np.float64); b = np.asarray(b, dtype=np.float64)
    return float(np.sqrt(np.sum((a - b) ** 2)))


def kl_divergence(a, b) -> float:
    """Distances.KLDivergence: sum_k a_k>0 ? a_k*log(a_k/b_k) : 0."""
    a = np.asarray(a, dtype=np.float64); b = np.asarray(b, dtype=np.float64)
    s = 0.0
    for x, y in zip(a, b):
        if x > 0:
            s += x * math.log(x / y)
    return s


# --------------------------------------------------------------------------------------
# _splitnorm and CDFs
# --------------------------------------------------------------------------------------
def clamp_input(A: np.ndarray) -> np.ndarray:
    """sequence(): assert + clamp!(A, eps(), typemax) (sequencer.jl:111,130). Returns a copy."""
    A = np.array(A, dtype=np.float64)
    if A.ndim != 2:
        raise ValueError("A must be M x N")
    if not (np.all(np.isfinite(A)) and np.all(A >= 0)):
        raise AssertionError("Input data cannot contain negative, NaN or infinite values.")
    return np.maximum(A, JULIA_EPS)


def clamp_input_f32(A: np.ndarray) -> np.ndarray:
    """Float32 input: `clamp!(A, eps(), typemax(Float32))` stays in Float32 (sequencer.jl:130)."""
    A = np.array(A, dtype=np.float32)
    if A.ndim != 2:
        raise ValueError("A must be M x N")
    if not (np.all(np.isfinite(A)) and np.all(A >= 0)):
        raise AssertionError("Input data cannot contain negative, NaN or infinite values.")
    return np.maximum(A, np.float32(JULIA_EPS))


def splitnorm_f32(A32: np.ndarray, scale: int):
    """_splitnorm on a Float32 matrix: `S ./ sum(S, dims=1)` evaluated in Float32 (sequencer.jl:402-403)."""
    out = []
    for r0, r1 in chunk_bounds(A32.shape[0], scale):
        S = A32[r0:r1, :]
        out.append((S / np.sum(S, axis=0, keepdims=True, dtype=np.float32)).astype(np.float32))
    return out


def pairwise_f32(metric: int, P: np.ndarray, l2_thresh: float = L2THR, block: int = 64) -> np.ndarray:
    """The reference's arithmetic on Float32 chunks (quirk Q2): Distances evaluates Euclidean (Gram + threshold
    recompute) and KLDivergence in Float32; EMD / Energy build ECDFs on Float32 weights (cumulative sums and the
    division in Float32) and only the multiplication by the Float64 grid steps promotes (distancemetrics.jl:199-219).
    Restated from the published algorithms; no Julia run pins the summation orders (parity unpinned)."""
    P = np.asarray(P, dtype=np.float32)
    m, N = P.shape
    if metric == L2:
        G = P.T @ P
        sa2 = np.sum(P * P, axis=0, dtype=np.float32)
        self_ = sa2[:, None] + sa2[None, :]
        v = self_ - np.float32(2) * G
        bad = v < np.float32(l2_thresh) * self_
        np.fill_diagonal(bad, False)
        bi, bj = np.nonzero(bad)
        for i0 in range(0, len(bi), 4096):
            ii, jj = bi[i0:i0 + 4096], bj[i0:i0 + 4096]
            v[ii, jj] = np.sum((P[:, ii] - P[:, jj]) ** 2, axis=0, dtype=np.float32)
        r = np.sqrt(np.maximum(v, np.float32(0)))
        r = np.tril(r, -1)
        return (r + r.T).astype(np.float64)
    if metric == KLD:
        r = np.zeros((N, N), dtype=np.float32)
        for i in range(N):
            col = P[:, i:i + 1]
            r[i, :] = np.sum(col * np.log(col / P), axis=0, dtype=np.float32)
        np.fill_diagonal(r, 0.0)
        return r.astype(np.float64)
    U = (np.cumsum(P, axis=0, dtype=np.float32) / np.sum(P, axis=0, keepdims=True, dtype=np.float32)).astype(np.float32)
    U[-1, :] = 1.0
    UT = np.ascontiguousarray(U.T)
    r = np.zeros((N, N))
    for i0 in range(0, N, block):
        d = UT[i0:i0 + block, None, :] - UT[None, :, :]              # Float32 differences
        if metric == EMD:
            r[i0:i0 + block, :] = np.sum(np.abs(d).astype(np.float64), axis=2)
        else:
            r[i0:i0 + block, :] = math.sqrt(2.0) * np.sqrt(np.sum((d * d).astype(np.float64), axis=2))
    np.fill_diagonal(r, 0.0)
    return r


def chunk_bounds(M: int, scale: int):
    """_splitnorm chunking (sequencer.jl:389-397): chunklen = cld(M,scale); starts 1:chunklen:M."""
    if scale > M:
        raise AssertionError(f"Scale ({scale}) cannot be larger than the number of data elements ({M}).")
    if scale < 1:
        raise AssertionError("scale must be >= 1")
    cl = -(-M // scale)
    return [(r0, min(r0 + cl, M)) for r0 in range(0, M, cl)]


def splitnorm(A: np.ndarray, scale: int):
    """sequencer.jl:383-407: list of chunk matrices (m_c x N), each column normalised to sum 1."""
    out = []
    for r0, r1 in chunk_bounds(A.shape[0], scale):
        S = A[r0:r1, :]
        out.append(S / np.sum(S, axis=0, keepdims=True))
    return out


def chunk_cdf(P: np.ndarray) -> np.ndarray:
    """ECDF of every column on the shared unit grid 1..m, evaluated on that grid:
    sequential running sum / cached sum, last entry exactly 1 (StatsBase ecdf semantics)."""
    U = np.cumsum(P, axis=0) / np.sum(P, axis=0, keepdims=True)
    U[-1, :] = 1.0
    return U


# --------------------------------------------------------------------------------------
# pairwise distance matrices:  abs.(pairwise(alg, m; dims=2)) .+ ϵ   (sequencer.jl:226)
# --------------------------------------------------------------------------------------
def pairwise_faithful(metric: int, P: np.ndarray, l2_thresh: float = L2THR) -> np.ndarray:
    """Per-pair evaluation exactly as Distances.pairwise drives the reference functors. O(N^2) Python."""
    m, N = P.shape
    r = np.zeros((N, N))
    if metric == KLD:                                # PreMetric: full matrix
        for i in range(N):
            for j in range(N):
                r[i, j] = kl_divergence(P[:, i], P[:, j])
        return r
    if metric == L2:                                 # Gram trick + threshold recompute
        G = P.T @ P
        sa2 = np.sum(P * P, axis=0)
        for j in range(N):
            for i in range(j + 1, N):
                selfterms = sa2[i] + sa2[j]
                v = selfterms - 2 * G[i, j]
                if v < l2_thresh * selfterms:
                    v = float(np.sum((P[:, i] - P[:, j]) ** 2))
                r[i, j] = r[j, i] = math.sqrt(max(v, 0.0)) if v > 0 else 0.0
        return r
    fn = emd if metric == EMD else energy
    for j in range(N):
        for i in range(j + 1, N):
            r[i, j] = r[j, i] = fn(P[:, i], P[:, j])
    return r


def pairwise_vec(metric: int, P: np.ndarray, l2_thresh: float = L2THR, block: int = 64) -> np.ndarray:
    """Vectorised restatement of the same matrices (direct differences, no Gram cancellation
    for EMD/Energy/KL; Gram + threshold for Euclidean as Distances does)."""
    m, N = P.shape
    r = np.zeros((N, N))
    if metric == L2:
        G = P.T @ P
        sa2 = np.sum(P * P, axis=0)
        self_ = sa2[:, None] + sa2[None, :]
        v = self_ - 2 * G
        bad = v < l2_thresh * self_
        np.fill_diagonal(bad, False)
        bi, bj = np.nonzero(bad)
        for i0 in range(0, len(bi), 4096):
            ii, jj = bi[i0:i0 + 4096], bj[i0:i0 + 4096]
            v[ii, jj] = np.sum((P[:, ii] - P[:, jj]) ** 2, axis=0)
        r = np.sqrt(np.maximum(v, 0.0))
        r = np.tril(r, -1); r = r + r.T                       # i>j evaluated, mirrored
        return r
    if metric == KLD:
        logP = np.log(P)
        for i0 in range(0, N, block):
            a = P[:, i0:i0 + block]                              # m x b
            # sum_k a_k * log(a_k / b_k), evaluated with the division as the reference does
            for ii in range(a.shape[1]):
                col = a[:, ii:ii + 1]
                r[i0 + ii, :] = np.sum(col * np.log(col / P), axis=0)
        np.fill_diagonal(r, 0.0)
        return r
    U = chunk_cdf(P)
    UT = np.ascontiguousarray(U.T)                               # N x m
    for i0 in range(0, N, block):
        d = UT[i0:i0 + block, None, :] - UT[None, :, :]          # b x N x m
        if metric == EMD:
            r[i0:i0 + block, :] = np.sum(np.abs(d), axis=2)
        else:
            r[i0:i0 + block, :] = math.sqrt(2.0) * np.sqrt(np.sum(d * d, axis=2))
    np.fill_diagonal(r, 0.0)
    return r


def pairwise_grid(metric: int, P: np.ndarray, lgrid: np.ndarray, faithful: bool = False, block: int = 64) -> np.ndarray:
    """EMD((lgrid, lgrid)) / Energy((lgrid, lgrid)) as the reference builds them from the chunk-local grid slice
    (sequencer.jl:217-219): same-grid branch of _cdf_distance with deltas g_k - g_{k-1}, g_0 = 0
    (distancemetrics.jl:199-209,217-219)."""
    m, N = P.shape
    r = np.zeros((N, N))
    if faithful:
        fn = emd if metric == EMD_GRID else energy
        for j in range(N):
            for i in range(j + 1, N):
                r[i, j] = r[j, i] = fn(lgrid, lgrid, P[:, i], P[:, j])
        return r
    dx = np.diff(np.concatenate(([0.0], lgrid)))
    UT = np.ascontiguousarray(chunk_cdf(P).T)
    for i0 in range(0, N, block):
        d = UT[i0:i0 + block, None, :] - UT[None, :, :]
        if metric == EMD_GRID:
            r[i0:i0 + block, :] = np.sum(np.abs(d) * dx, axis=2)
        else:
            r[i0:i0 + block, :] = math.sqrt(2.0) * np.sqrt(np.sum(d * d * dx, axis=2))
    np.fill_diagonal(r, 0.0)
    return r


def periodic_euclidean(a, b, periods) -> float:
    """Distances.PeriodicEuclidean(periods)(a, b) (Distances.jl 0.10 metrics.jl, `eval_op`/`eval_end`; un-vendored
    dependency, Project.toml compat "0.10"):  s1 = |a - b|; s2 = mod(s1, p); s3 = min(s2, p - s2); sqrt(sum s3^2)."""
    a = np.asarray(a, dtype=np.float64); b = np.asarray(b, dtype=np.float64)
    p = np.asarray(periods, dtype=np.float64)
    if not (len(a) == len(b) == len(p)):
        raise ValueError("DimensionMismatch: PeriodicEuclidean needs one period per element")
    s = np.mod(np.abs(a - b), p)
    s = np.minimum(s, p - s)
    return float(np.sqrt(np.sum(s * s)))


def pairwise_periodic(P: np.ndarray, periods, faithful: bool = False) -> np.ndarray:
    m, N = P.shape
    p = np.asarray(periods, dtype=np.float64)
    if len(p) != m:
        raise ValueError("DimensionMismatch: PeriodicEuclidean needs one period per row of the chunk")
    r = np.zeros((N, N))
    if faithful:
        for j in range(N):
            for i in range(j + 1, N):
                r[i, j] = r[j, i] = periodic_euclidean(P[:, i], P[:, j], p)
        return r
    for i in range(N):
        s = np.mod(np.abs(P[:, i:i + 1] - P), p[:, None])
        s = np.minimum(s, p[:, None] - s)
        r[i, :] = np.sqrt(np.sum(s * s, axis=0))
    np.fill_diagonal(r, 0.0)
    return r


def pairwise_elementwise(metric: int, P: np.ndarray, block: int = 64) -> np.ndarray:
    """Distances.jl Cityblock / TotalVariation / Chebyshev / SqEuclidean (published definitions: sum |a-b|,
    sum |a-b| / 2, max |a-b|, sum (a-b)^2), evaluated pair by pair over the columns as `pairwise` does."""
    m, N = P.shape
    PT = np.ascontiguousarray(P.T)
    r = np.zeros((N, N))
    for i0 in range(0, N, block):
        d = PT[i0:i0 + block, None, :] - PT[None, :, :]
        if metric == CITYBLOCK:
            r[i0:i0 + block] = np.abs(d).sum(axis=2)
        elif metric == TOTALVARIATION:
            r[i0:i0 + block] = np.abs(d).sum(axis=2) / 2
        elif metric == CHEBYSHEV:
            r[i0:i0 + block] = np.abs(d).max(axis=2)
        else:
            r[i0:i0 + block] = (d * d).sum(axis=2)
    np.fill_diagonal(r, 0.0)
    return r


def chunk_distance(metric: int, P: np.ndarray, faithful: bool = False,
                   eps_floor: float = EPS_FLOOR, l2_thresh: float = L2THR, lgrid=None, periods=None) -> np.ndarray:
    if metric in (CITYBLOCK, TOTALVARIATION, CHEBYSHEV, SQEUCLIDEAN):
        return np.abs(pairwise_elementwise(metric, P)) + eps_floor
    if metric == PERIODIC:
        return np.abs(pairwise_periodic(P, periods, faithful)) + eps_floor
    if metric in (EMD_GRID, ENERGY_GRID):
        return np.abs(pairwise_grid(metric, P, np.asarray(lgrid, dtype=np.float64), faithful)) + eps_floor
    f = pairwise_faithful if faithful else pairwise_vec
    return np.abs(f(metric, P, l2_thresh)) + eps_floor


# --------------------------------------------------------------------------------------
# _measure_dm: MST (Kruskal), least central point, elongation, BFS tree
# --------------------------------------------------------------------------------------
@dataclass
class Measure:
    mst_i: np.ndarray          # N-1 edge endpoints (0-based, i<j), in Kruskal acceptance order
    mst_j: np.ndarray
    mst_w: np.ndarray
    root: int                  # least central vertex (0-based)
    eta: float
    parents: np.ndarray        # BFS-tree parent of each vertex (root -> itself)
    depth: np.ndarray          # hop depth from root
    root_score_gap: float = 0.0    # relative gap between best and second-best distance sum
    min_cut_gap: float = float("nan")


def kruskal_mst(D: np.ndarray):
    """_mst (sequencer.jl:358-374): clamp to [ϵ, Inf], Symmetric(dm) = upper triangle,
    LightGraphs.kruskal_mst = stable sort of edges in column-major upper-triangle order + union-find."""
    N = D.shape[0]
    W = np.maximum(D, EPS_FLOOR)
    jj, ii = np.tril_indices(N)              # (j outer, i<=j inner): column-major upper triangle
    w = W[ii, jj]
    order = np.argsort(w, kind="stable")
    parent = list(range(N))

    def find(x):
        root = x
        while parent[root] != root:
            root = parent[root]
        while parent[x] != root:
            parent[x], x = root, parent[x]
        return root

    ei, ej, ew = [], [], []
    ii_s, jj_s, w_s = ii[order], jj[order], w[order]
    for a, b, ww in zip(ii_s.tolist(), jj_s.tolist(), w_s.tolist()):
        ra, rb = find(a), find(b)
        if ra != rb:
            parent[ra] = rb
            ei.append(a); ej.append(b); ew.append(ww)
            if len(ei) >= N - 1:
                break
    return np.array(ei, dtype=np.int64), np.array(ej, dtype=np.int64), np.array(ew, dtype=np.float64)


def _tree_adjacency(N, ei, ej, ew):
    adj = [[] for _ in range(N)]
    for a, b, w in zip(ei.tolist(), ej.tolist(), ew.tolist()):
        adj[a].append((b, w)); adj[b].append((a, w))
    return adj


def tree_distance_sums(N, ei, ej, ew, faithful=True):
    """sum_v d_w(u, v) for every u on the tree. faithful: N single-source traversals with path sums
    accumulated from the source outward (what Dijkstra does on a tree); else O(N) rerooting."""
    if faithful:
        from scipy.sparse import coo_matrix
        from scipy.sparse.csgraph import dijkstra
        g = coo_matrix((np.concatenate([ew, ew]), (np.concatenate([ei, ej]), np.concatenate([ej, ei]))),
                       shape=(N, N)).tocsr()
        sums = np.empty(N)
        for u0 in range(0, N, 512):
            idx = np.arange(u0, min(u0 + 512, N))
            d = dijkstra(g, directed=False, indices=idx)
            sums[idx] = d.sum(axis=1)
        return sums
    adj = _tree_adjacency(N, ei, ej, ew)
    order, par, parw = [0], [-1] * N, [0.0] * N
    seen = [False] * N; seen[0] = True
    for u in order:
        for v, w in adj[u]:
            if not seen[v]:
                seen[v] = True; par[v] = u; parw[v] = w; order.append(v)
    size = [1] * N
    wd = [0.0] * N
    for u in order[1:]:
        wd[u] = wd[par[u]] + parw[u]
    for u in reversed(order[1:]):
        size[par[u]] += size[u]
    S = [0.0] * N
    S[0] = float(sum(wd))
    for u in order[1:]:
        S[u] = S[par[u]] + parw[u] * (N - 2 * size[u])
    return np.array(S)


def bfs_from(N, ei, ej, ew, root):
    adj = _tree_adjacency(N, ei, ej, ew)
    par = np.full(N, -1, dtype=np.int64); depth = np.zeros(N, dtype=np.int64)
    par[root] = root
    q = [root]
    for u in q:
        for v, _ in adj[u]:
            if par[v] < 0:
                par[v] = u; depth[v] = depth[u] + 1; q.append(v)
    return par, depth


def elongation_from_depths(depth: np.ndarray) -> float:
    """sequencer.jl:334-354: hlen = mean(h); hwid = mean(count per distinct level)/2; eta = hlen/hwid + 1."""
    N = len(depth)
    hlen = float(np.sum(depth)) / N
    nlev = len(np.unique(depth))
    hwid = (N / nlev) / 2
    hlen = min(max(hlen, JULIA_EPS), FLOATMAX32)
    hwid = min(max(hwid, JULIA_EPS), FLOATMAX32)
    return hlen / hwid + 1


def measure_dm(D: np.ndarray, faithful_root: bool = True) -> Measure:
    """_measure_dm (sequencer.jl:284-290)."""
    N = D.shape[0]
    ei, ej, ew = kruskal_mst(D)
    if N == 1:
        return Measure(ei, ej, ew, 0, elongation_from_depths(np.zeros(1, dtype=np.int64)),
                       np.zeros(1, dtype=np.int64), np.zeros(1, dtype=np.int64))
    sums = tree_distance_sums(N, ei, ej, ew, faithful=faithful_root)
    closeness = ((N - 1) / sums) * ((N - 1) / (N - 1))           # LightGraphs closeness_centrality, normalised
    root = int(np.argmin(closeness))                              # last(findmin(.)): first minimum
    srt = np.sort(sums)
    gap = float((srt[-1] - srt[-2]) / srt[-1]) if N > 1 else 0.0
    par, depth = bfs_from(N, ei, ej, ew, root)
    return Measure(ei, ej, ew, root, elongation_from_depths(depth), par, depth, root_score_gap=gap)


def unroll(parents: np.ndarray, root: int) -> np.ndarray:
    """utils.jl:15-24: reverse(DFS preorder over the directed BFS tree, children ascending)."""
    N = len(parents)
    children = [[] for _ in range(N)]
    for v in range(N):
        if v != root:
            children[int(parents[v])].append(v)       # ascending because v ascends
    out, stack = [], [root]
    while stack:
        u = stack.pop()
        out.append(u)
        stack.extend(reversed(children[u]))
    return np.array(out[::-1], dtype=np.int64)


# --------------------------------------------------------------------------------------
# _sequence
# --------------------------------------------------------------------------------------
@dataclass
class GroupResult:
    metric: int
    scale: int
    eta_chunk: list
    parents_chunk: list
    root_chunk: list
    eta: float
    parents: np.ndarray
    root: int
    mst: tuple
    D: np.ndarray = None       # kept only when keep_matrices
    D_chunks: list = None


@dataclass
class OracleResult:
    groups: list
    eta: float
    root: int
    order: np.ndarray          # 0-based
    parents: np.ndarray
    mst: tuple
    D_edges: dict              # {(i,j): value} for i<j on the edge union
    min_root_gap: float = 0.0


def weighted_p_edges(N, msts, etas):
    """_weighted_p_matrix (sequencer.jl:310-328) + D = inv.(Pw) (:269), on the edge union only."""
    P = {}
    for (ei, ej, ew), eta in zip(msts, etas):
        for a, b, w in zip(ei.tolist(), ej.tolist(), ew.tolist()):
            key = (a, b) if a < b else (b, a)
            P[key] = P.get(key, 0.0) + eta * (1.0 / w)
    tot = 0.0
    for e in etas:
        tot += e
    return {k: 1.0 / (v / tot) for k, v in P.items()}


def dense_from_edges(N, edges):
    D = np.full((N, N), np.inf)
    np.fill_diagonal(D, 0.0)
    for (a, b), v in edges.items():
        D[a, b] = D[b, a] = v
    return D


def compute_group(A, k, l, faithful=False, faithful_root=True, keep_matrices=False,
                  eps_floor=EPS_FLOOR, l2_thresh=L2THR, grid=None, periods=None, f32=False):
    """One (metric k, scale l) iteration of the double loop in _sequence (sequencer.jl:196-258).
    A is already clamped. Returns (GroupResult, min root-score gap)."""
    M, N = A.shape
    min_gap = float("inf")
    S = np.zeros((N, N))
    eta_c, par_c, root_c, Dcs = [], [], [], []
    if grid is None:
        grid = np.arange(1, M + 1, dtype=np.float64)          # ensuregrid! default (sequencer.jl:298)
    for P, (r0, r1) in zip(splitnorm_f32(A, l) if f32 else splitnorm(A, l), chunk_bounds(M, l)):
        if f32:
            Dc = np.abs(pairwise_f32(k, P, l2_thresh)) + eps_floor
        else:
            Dc = chunk_distance(k, P, faithful, eps_floor, l2_thresh, lgrid=grid[r0:r1], periods=periods)
        if k == KLD:                     # Symmetric(dm): only the upper triangle is ever read
            Dsym = np.triu(Dc) + np.triu(Dc, 1).T
        else:
            Dsym = Dc
        ms = measure_dm(Dsym, faithful_root)
        min_gap = min(min_gap, ms.root_score_gap)
        S += np.maximum(ms.eta * Dc, JULIA_EPS)          # sequencer.jl:231-235
        eta_c.append(ms.eta); par_c.append(ms.parents); root_c.append(ms.root)
        if keep_matrices:
            Dcs.append(Dc)
    tot = 0.0
    for e in eta_c:
        tot += e
    Dkl = S / tot                                        # :247
    Dsym = np.triu(Dkl) + np.triu(Dkl, 1).T if k == KLD else Dkl
    ms = measure_dm(Dsym, faithful_root)
    min_gap = min(min_gap, ms.root_score_gap)
    mst = (ms.mst_i, ms.mst_j, ms.mst_w)
    return GroupResult(k, l, eta_c, par_c, root_c, ms.eta, ms.parents, ms.root, mst,
                       Dkl if keep_matrices else None, Dcs if keep_matrices else None), min_gap


def fuse_groups(N, groups, faithful_root=True, min_gap=float("inf")) -> OracleResult:
    """sequencer.jl:266-275: proximity fuse, final _measure_dm, unroll."""
    msts = [g.mst for g in groups]
    etas = [g.eta for g in groups]
    edges = weighted_p_edges(N, msts, etas)
    D = dense_from_edges(N, edges)
    ms = measure_dm(D, faithful_root)
    min_gap = min(min_gap, ms.root_score_gap)
    order = unroll(ms.parents, ms.root)
    return OracleResult(groups, ms.eta, ms.root, order, ms.parents, (ms.mst_i, ms.mst_j, ms.mst_w),
                        edges, min_gap)


def _norm_args(scales, metrics):
    if isinstance(scales, (int, np.integer)):
        scales = (int(scales),)
    if isinstance(metrics, (int, np.integer)):
        metrics = (int(metrics),)
    return tuple(scales), tuple(metrics)


def sequence_oracle(A, scales, metrics=ALL_METRICS, faithful=False, faithful_root=True,
                    keep_matrices=False, eps_floor=EPS_FLOOR, l2_thresh=L2THR, grid=None, periods=None,
                    f32=False) -> OracleResult:
    """_sequence (sequencer.jl:184-276) on an M x N matrix (columns = objects).  f32: the input is a Float32 matrix
    and the chunk distances follow the reference's Float32 arithmetic (quirk Q2, pairwise_f32)."""
    A = clamp_input_f32(A) if f32 else clamp_input(A)
    M, N = A.shape
    scales, metrics = _norm_args(scales, metrics)
    groups = []
    min_gap = float("inf")
    for k in metrics:
        for l in scales:
            g, gap = compute_group(A, k, l, faithful, faithful_root, keep_matrices, eps_floor, l2_thresh,
                                   None if grid is None else np.asarray(grid, dtype=np.float64), periods, f32)
            groups.append(g)
            min_gap = min(min_gap, gap)
    return fuse_groups(N, groups, faithful_root, min_gap)


_PAR_A = None


def _par_group(args):
    k, l, faithful_root, faithful = args
    return compute_group(_PAR_A, k, l, faithful, faithful_root)[0]


def sequence_oracle_parallel(A, scales, metrics=ALL_METRICS, procs=None, faithful_root=True, faithful=False) -> OracleResult:
    """Same result as sequence_oracle, with the independent (metric, scale) groups spread over host
    processes (fork). Used only as the CPU baseline of bench.py: the reference itself is single-threaded."""
    import multiprocessing as mp
    import os
    global _PAR_A
    A = clamp_input(A)
    M, N = A.shape
    scales, metrics = _norm_args(scales, metrics)
    tasks = [(k, l, faithful_root, faithful) for k in metrics for l in scales]
    procs = procs or os.cpu_count() or 1
    _PAR_A = A
    try:
        if procs <= 1 or len(tasks) == 1:
            groups = [_par_group(t) for t in tasks]
        else:
            with mp.get_context("fork").Pool(min(procs, len(tasks))) as pool:
                groups = pool.map(_par_group, tasks, chunksize=1)
    finally:
        _PAR_A = None
    return fuse_groups(N, groups, faithful_root)
