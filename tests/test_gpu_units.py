"""GPU parity tests of each layer through the C ABI (libseqj_b200.so) against the CPU oracle.

Tolerances (north_star): distance matrices and eta within 1e-9 relative; MST / root / order
identical (the seeded inputs below have no ties).  CDF-based metrics add an absolute floor of
64*m*2^-53 for the inherent prefix-sum rounding noise (the reference's own sequential cumsum has
the same order of noise)."""
import numpy as np
import pytest

from conftest import rel_close

pytestmark = pytest.mark.gpu


def _rand(M, N, seed):
    return np.random.default_rng(seed).random((M, N))


@pytest.mark.parametrize("M,N,scale", [(50, 100, 1), (50, 100, 2), (50, 100, 4), (10, 7, 6), (137, 33, 5),
                                       (1000, 64, 8), (4096, 16, 16)])
def test_splitnorm_matches_oracle(seqj, oracle, handle, M, N, scale):
    A = oracle.clamp_input(_rand(M, N, 1))
    P, U, nch = seqj.splitnorm(A, scale)
    chunks = oracle.splitnorm(A, scale)
    assert nch == len(chunks) == len(oracle.chunk_bounds(M, scale))
    Pref = np.vstack(chunks)
    Uref = np.vstack([oracle.chunk_cdf(c) for c in chunks])
    rel_close(P, Pref, 1e-13)
    rel_close(U, Uref, 1e-13, atol=1e-15)
    for (r0, r1) in oracle.chunk_bounds(M, scale):
        assert np.all(U[r1 - 1] == 1.0)          # last CDF entry is exactly 1 (StatsBase ecdf)


def test_splitnorm_scale_too_large(seqj, handle):
    with pytest.raises(AssertionError):
        seqj.splitnorm(np.ones((5, 4)), 6)


@pytest.mark.parametrize("metric", [0, 1, 2, 3])
@pytest.mark.parametrize("m,N", [(50, 100), (13, 100), (11, 37), (137, 300), (256, 1000), (1000, 513)])
def test_pairwise_matches_oracle(seqj, oracle, handle, metric, m, N):
    A = oracle.clamp_input(_rand(m, N, 7 + metric))
    P = oracle.splitnorm(A, 1)[0]
    ref = oracle.chunk_distance(metric, P)
    got = seqj.pairwise(metric, A, normalize=True, kl_full=True)
    rel_close(got, ref, 1e-9, atol=64 * m * 2.0 ** -53)
    if metric != oracle.KLD:
        assert np.array_equal(got, got.T)
    assert np.all(np.diag(got) == 1e-6)


@pytest.mark.parametrize("metric", [0, 1, 2, 3])
def test_pairwise_faithful_small(seqj, oracle, handle, metric):
    """Per-pair evaluation exactly as the reference's functors (ECDF objects, sorts) at small N."""
    A = oracle.clamp_input(_rand(25, 40, 3))
    P = oracle.splitnorm(A, 1)[0]
    ref = oracle.chunk_distance(metric, P, faithful=True)
    got = seqj.pairwise(metric, A, normalize=True, kl_full=True)
    rel_close(got, ref, 1e-9, atol=64 * 25 * 2.0 ** -53)


def test_kl_mirrored_upper(seqj, oracle, handle):
    A = oracle.clamp_input(_rand(64, 150, 5))
    full = seqj.pairwise(2, A, kl_full=True)
    sym = seqj.pairwise(2, A, kl_full=False)
    iu = np.triu_indices(150, 1)
    rel_close(sym[iu], full[iu], 1e-12)
    assert np.array_equal(sym, sym.T)        # Symmetric(dm): upper triangle wins (sequencer.jl:365)
    assert not np.allclose(full, full.T)


@pytest.mark.parametrize("metric", [0, 2, 3])
def test_near_duplicate_columns_take_fixup_path(seqj, oracle, handle, metric):
    """Gram-form cancellation regime: columns that differ by ~1e-7 relative must still match the
    direct formulas of the reference (recomputed by the fix-up kernel)."""
    rng = np.random.default_rng(11)
    base = rng.random((200, 1)) + 0.1
    A = base * (1.0 + 1e-7 * rng.standard_normal((200, 90)))
    A[:, 10] = A[:, 3]                        # exact duplicate -> distance exactly 0 (+eps)
    A = oracle.clamp_input(A)
    P = oracle.splitnorm(A, 1)[0]
    if metric == 0:                           # direct formula (what Distances recomputes below thresh)
        ref = np.sqrt(((P[:, :, None] - P[:, None, :]) ** 2).sum(axis=0)) + 1e-6
    else:
        ref = oracle.chunk_distance(metric, P)
    got = seqj.pairwise(metric, A, kl_full=True)
    # Absolute floor 4*m*2^-53 for Energy and KL: when two objects differ by ~1e-7 the result is dominated by
    # rounding that BOTH implementations carry in their inputs -- a few ulp(1) per CDF entry from the prefix sum
    # (sequential cumsum vs warp scan), and the column normalisation sum(p) = 1 +- few ulp, to which KL is not
    # invariant (KL(s*p || t*q) ~ KL + log(s/t)).  The reference's own values carry the same noise.
    rel_close(got, ref, 1e-9, atol=(4 * 200 * 2.0 ** -53) if metric in (2, 3) else 1e-16)
    assert got[10, 3] == 1e-6 and got[3, 10] == 1e-6


@pytest.mark.parametrize("metric", [0, 1, 2, 3])
def test_pairwise_smooth_ordered_family(seqj, oracle, handle, metric):
    """Smooth 1-parameter family (a bump moving across the grid, the Sequencer paper's toy data): neighbouring
    objects differ by ~1 %, i.e. the regime between the Gram form and the direct fix-up.  Checked against the
    oracle's direct formulas."""
    N, m = 400, 96
    x = np.arange(m)[:, None]
    mu = np.linspace(0.2 * m, 0.8 * m, N)[None, :]
    A = np.exp(-0.5 * ((x - mu) / (0.08 * m)) ** 2) + 1e-3
    A = A[:, np.random.default_rng(3).permutation(N)]
    P = oracle.splitnorm(oracle.clamp_input(A), 1)[0]
    if metric == 0:
        ref = np.sqrt(((P[:, :, None] - P[:, None, :]) ** 2).sum(axis=0)) + 1e-6       # direct, no Gram form
    else:
        ref = oracle.chunk_distance(metric, P)
    got = seqj.pairwise(metric, A, kl_full=True)
    rel_close(got, ref, 1e-9, atol=4 * m * 2.0 ** -53)


@pytest.mark.parametrize("metric", [0, 2, 3])
def test_dense_near_duplicate_tiles(seqj, oracle, handle, metric):
    """Whole tiles of near-duplicate objects (500 perturbed copies of one profile + 200 unrelated objects): the
    DMMA epilogue flags more than 1/32 of such tiles, which are then recomputed by the direct tile kernels
    (sum (a-b)^2, sum p (log p - log q)) instead of entry by entry.  Same 1e-9 bar against the direct formulas."""
    rng = np.random.default_rng(77)
    m, n_dup, n_rand = 64, 500, 200
    base = rng.random((m, 1)) + 0.2
    A = np.concatenate([base * (1.0 + 1e-6 * rng.standard_normal((m, n_dup))), rng.random((m, n_rand)) + 0.05], axis=1)
    A = oracle.clamp_input(A[:, rng.permutation(n_dup + n_rand)])
    P = oracle.splitnorm(A, 1)[0]
    if metric == 0:
        ref = np.sqrt(((P[:, :, None] - P[:, None, :]) ** 2).sum(axis=0)) + 1e-6
    else:
        ref = oracle.chunk_distance(metric, P)
    got = seqj.pairwise(metric, A, kl_full=False)
    if metric == 2:
        ref = np.triu(ref) + np.triu(ref, 1).T           # mirrored upper triangle (what the pipeline consumes)
    rel_close(got, ref, 1e-9, atol=4 * m * 2.0 ** -53)


def test_pairwise_algotests_known_answers(seqj, handle):
    """reference test/algotests.jl:6-25 (Euclidean sqrt(14); KL 0.5108256237659907)."""
    A = np.array([[1.0, 2.0], [1.0, 3.0], [1.0, 4.0]])
    Dm = seqj.pairwise(0, A, normalize=False, eps_floor=0.0)
    assert abs(Dm[0, 1] - np.sqrt(14.0)) < 1e-14 and Dm[0, 0] == 0.0
    K = seqj.pairwise(2, np.array([[0.5, 0.9], [0.5, 0.1]]), normalize=False, eps_floor=0.0, kl_full=True)
    assert abs(K[0, 1] - 0.5108256237659907) < 1e-15


def test_cdf_distance_algotests_known_answers(seqj, handle):
    """Reference test/algotests.jl:36-107: explicit, DIFFERENT grids (the merge branch of _cdf_distance), values
    checked there against scipy.stats.wasserstein_distance / energy_distance."""
    g1, g2, u, v = [3.4, 3.9, 7.5, 7.8], [4.5, 1.4], [1.4, 0.9, 3.1, 7.2], [3.2, 3.5]
    rel_close([seqj.emd(g1, g2, u, v)], [4.0781331438047861], 1e-12)                  # :36-41
    rel_close([seqj.EMD((g1, g2))(u, v)], [4.0781331438047861], 1e-12)                # :50-57
    rel_close([seqj.Energy()(g1, g2, u, v)], [2.3674181638546394], 1e-12)             # :80-85
    rel_close([seqj.Energy()([0, 8], [0, 8], [3, 1], [2, 2])], [1.0], 1e-12)          # :74-78
    h1, h2, a, b = [0.7, 7.4, 2.4, 6.8], [1.4, 8.0], [2.1, 4.2, 7.4, 8.0], [7.6, 8.8]
    rel_close([seqj.energy(h1, h2, a, b)], [0.8800334097615822], 1e-12)               # :88-92
    rel_close([seqj.Energy((h1, h2))(a, b)], [0.8800334097615822], 1e-12)             # :100-107
    g = [0.5, 1.0, 1.5, 2.0]
    assert np.isfinite(seqj.EMD((g, g))([3.4, 3.9, 7.5, 7.8], [1.4, 0.9, 3.1, 7.2]))  # :43-48
    assert seqj.EMD()([1, 1, 1], [2, 3, 4]) is not None                               # :30-33
    assert seqj.Energy()([1, 1, 1], [2, 3, 4]) is not None                            # :62-65


@pytest.mark.parametrize("nu,nv", [(1, 1), (2, 5), (17, 17), (100, 37), (1000, 1500), (9000, 9000), (20000, 3)])
@pytest.mark.parametrize("kind", ["different", "same", "overlap"])
def test_cdf_distance_matches_oracle(seqj, oracle, handle, nu, nv, kind):
    rng = np.random.default_rng(nu * 31 + nv)
    if kind == "same":
        nv = nu
        gu = rng.permutation(np.cumsum(rng.random(nu) + 0.01))        # distinct, positive, unsorted
        gv = gu.copy()
    elif kind == "overlap":                                            # grids share values and repeat some
        pool = np.round(rng.random(max(nu, nv)) * 20, 1) + 0.1
        gu, gv = rng.choice(pool, nu), rng.choice(pool, nv)
        if nu == nv and np.array_equal(np.sort(gu), np.sort(gv)):
            gv = gv + 0.05
    else:
        gu, gv = rng.random(nu) * 10, rng.random(nv) * 12
    wu, wv = rng.random(nu) + 1e-3, rng.random(nv) + 1e-3
    rel_close([seqj.emd(gu, gv, wu, wv)], [oracle.emd(gu, gv, wu, wv)], 1e-11)
    rel_close([seqj.energy(gu, gv, wu, wv)], [oracle.energy(gu, gv, wu, wv)], 1e-11)


def test_cdf_distance_default_grids_of_different_length(seqj, oracle, handle):
    """`emd(u, v)` with length(u) != length(v): grids axes(u,1), axes(v,1) differ -> general branch."""
    rng = np.random.default_rng(4)
    u, v = rng.random(40), rng.random(25)
    rel_close([seqj.emd(u, v)], [oracle.emd(u, v)], 1e-12)
    rel_close([seqj.energy(u, v)], [oracle.energy(u, v)], 1e-12)
    u2 = rng.random(40)
    rel_close([seqj.emd(u, u2)], [oracle.emd(u, u2)], 1e-12)           # equal lengths: the tile path
    rel_close([seqj.emd(u, u2)], [seqj.emd(np.arange(1, 41.0), np.arange(1, 41.0), u, u2)], 1e-12)


def test_cdf_distance_errors(seqj, handle):
    with pytest.raises(ValueError):
        seqj.emd([1.0, 2.0], [1.0], [0.5], [1.0])                      # u and uw lengths differ
    with pytest.raises(ValueError):
        seqj.emd(np.ones((2, 2)), [1.0], np.ones((2, 2)), [1.0])
    with pytest.raises(AssertionError):                                # identical grids with a repeated value:
        seqj.emd([1.0, 1.0, 2.0], [1.0, 1.0, 2.0], [1, 1, 1], [1, 2, 3])   # the reference throws DimensionMismatch
    with pytest.raises(AssertionError):
        seqj.emd([1.0, np.nan], [1.0, 2.0], [1, 1], [1, 1])


def test_emd_energy_pair_functions(seqj, oracle, handle):
    """EMD()(u,v) / Energy()(u,v) on the default unit grid (distancemetrics.jl:46-50,158-162)
    against the faithful ECDF restatement and SciPy."""
    from scipy.stats import energy_distance, wasserstein_distance
    rng = np.random.default_rng(5)
    u, v = rng.random(37), rng.random(37)
    g = np.arange(1, 38, dtype=float)
    assert abs(seqj.emd(u, v) - oracle.emd(u, v)) < 1e-12
    assert abs(seqj.energy(u, v) - oracle.energy(u, v)) < 1e-12
    assert abs(seqj.emd(u, v) - wasserstein_distance(g, g, u, v)) < 1e-12
    assert abs(seqj.energy(u, v) - energy_distance(g, g, u, v)) < 1e-12
    assert seqj.EMD()(u, v) == seqj.emd(u, v) and seqj.Energy()(u, v) == seqj.energy(u, v)


def _random_dist(N, seed):
    rng = np.random.default_rng(seed)
    X = rng.random((N, 3))
    Dm = np.sqrt(((X[:, None, :] - X[None, :, :]) ** 2).sum(-1)) + 1e-6
    return Dm


def _edge_set(t):
    return set(zip(np.minimum(t.src, t.dst).tolist(), np.maximum(t.src, t.dst).tolist()))


# auto = kNN-filtered Boruvka with the row list kernel at these sizes (kernels_mst.cuh); tiles = the same with the
# upper-triangle list pass (knn_tiles_kernel + knn_merge_kernel, the default from N = 6144) forced; prim = the Prim kernels
MST_MODES = ("auto", "tiles", "prim")


def _measure_modes(seqj, monkeypatch, Dm, **kw):
    """`_measure_dm` through both MST back ends of the library (SEQJ_MST is read at every call)."""
    out = []
    for mode in MST_MODES:
        monkeypatch.delenv("SEQJ_MST", raising=False)
        monkeypatch.setenv("SEQJ_KNN", "tiles" if mode == "tiles" else "rows")
        if mode == "prim":
            monkeypatch.setenv("SEQJ_MST", mode)
        out.append(seqj.measure_dm(Dm, **kw))
    monkeypatch.delenv("SEQJ_MST", raising=False)
    monkeypatch.delenv("SEQJ_KNN", raising=False)
    return out


@pytest.mark.parametrize("N", [2, 3, 17, 100, 1000, 1024, 1031, 2500, 4099])
def test_measure_dm_matches_oracle(seqj, oracle, handle, monkeypatch, N):
    Dm = _random_dist(N, N)
    ref = oracle.measure_dm(Dm)
    for got in _measure_modes(seqj, monkeypatch, Dm, want_order=True):
        assert _edge_set(got.mst) == set(zip(ref.mst_i.tolist(), ref.mst_j.tolist()))
        rel_close(np.sort(got.mst.weight), np.sort(ref.mst_w), 1e-15)
        assert got.root == ref.root
        assert got.eta == ref.eta                      # pure function of integers -> bit-identical
        assert np.array_equal(got.bfs.parents, ref.parents)
        assert np.array_equal(got.order, oracle.unroll(ref.parents, ref.root))


@pytest.mark.parametrize("kind", ["clusters", "duplicates", "smooth"])
def test_measure_dm_structured_inputs_take_the_rescan_path(seqj, oracle, handle, monkeypatch, kind):
    """Inputs whose neighbour lists fill up with vertices of the same component (tight clusters larger than the
    list, exact duplicates, a smooth one-parameter family): the kNN-filtered Boruvka has to rescan rows, and both
    back ends must still return the reference's Kruskal tree."""
    rng = np.random.default_rng(5)
    if kind == "clusters":
        pts = np.concatenate([rng.normal(3.0 * c, 0.01, (6, 150)) for c in range(5)], axis=1)
        Dm = np.sqrt(((pts[:, :, None] - pts[:, None, :]) ** 2).sum(0)) + 1e-6
    elif kind == "duplicates":
        base = rng.random((250, 250)) + 0.01
        base = np.triu(base, 1); base = base + base.T
        Dm = np.kron(np.ones((3, 3)), base) + 1e-6
    else:
        x = np.arange(96)[:, None]
        mu = np.linspace(10, 86, 500)[None, :]
        F = np.exp(-0.5 * ((x - mu) / 6.0) ** 2) + 1e-3
        U = np.cumsum(F / F.sum(0), axis=0)[:, rng.permutation(500)]
        Dm = np.abs(U[:, :, None] - U[:, None, :]).sum(0) + 1e-6
    np.fill_diagonal(Dm, 1e-6)
    ref = oracle.measure_dm(Dm, faithful_root=False)
    for got in _measure_modes(seqj, monkeypatch, Dm, want_order=True):
        assert _edge_set(got.mst) == set(zip(ref.mst_i.tolist(), ref.mst_j.tolist()))
        assert got.eta == ref.eta
    a, *rest = _measure_modes(seqj, monkeypatch, Dm, want_order=True)
    for b in rest:
        assert a.root == b.root and np.array_equal(a.bfs.parents, b.bfs.parents) and np.array_equal(a.order, b.order)


@pytest.mark.parametrize("N,levels", [(7, 2), (64, 3), (300, 4), (1031, 3), (2500, 5), (4099, 2)])
def test_measure_dm_tied_weights_follow_kruskal_order(seqj, oracle, handle, monkeypatch, N, levels):
    """Heavily tied weights (a handful of distinct integer values): the reference's Kruskal visits equal-weight
    edges in column-major order of the upper triangle; the device MST uses the same strict total order, so
    the tree -- not only its weight -- is the reference's."""
    rng = np.random.default_rng(N + levels)
    Dm = rng.integers(1, levels + 1, size=(N, N)).astype(np.float64)
    Dm = np.triu(Dm, 1); Dm = Dm + Dm.T
    ref = oracle.measure_dm(Dm)
    for got in _measure_modes(seqj, monkeypatch, Dm, want_order=True):
        assert _edge_set(got.mst) == set(zip(ref.mst_i.tolist(), ref.mst_j.tolist()))
        assert got.root == ref.root
        assert got.eta == ref.eta
        assert np.array_equal(got.bfs.parents, ref.parents)
        assert np.array_equal(got.order, oracle.unroll(ref.parents, ref.root))


def test_measure_dm_all_equal_weights(seqj, oracle, handle, monkeypatch):
    """Every edge ties (duplicate objects: all distances collapse to the 1e-6 offset)."""
    N = 200
    Dm = np.full((N, N), 1e-6); np.fill_diagonal(Dm, 0.0)
    ref = oracle.measure_dm(Dm)
    for got in _measure_modes(seqj, monkeypatch, Dm):
        assert _edge_set(got.mst) == set(zip(ref.mst_i.tolist(), ref.mst_j.tolist()))
        assert got.root == ref.root and got.eta == ref.eta
        assert np.array_equal(got.bfs.parents, ref.parents)


def test_measure_dm_upper_triangle_wins_and_inf_nonedges(seqj, oracle, handle, monkeypatch):
    rng = np.random.default_rng(3)
    N = 60
    Dm = rng.random((N, N)) + 0.01                 # asymmetric: Symmetric(dm) keeps the upper triangle
    sym = np.triu(Dm) + np.triu(Dm, 1).T
    ref = oracle.measure_dm(sym)
    for got in _measure_modes(seqj, monkeypatch, Dm):
        assert _edge_set(got.mst) == set(zip(ref.mst_i.tolist(), ref.mst_j.tolist()))
        assert got.root == ref.root and got.eta == ref.eta
    # a path graph with Inf non-edges: the MST must be the path itself, eta = N
    P = np.full((N, N), np.inf); np.fill_diagonal(P, 0.0)
    for i in range(N - 1):
        P[i, i + 1] = P[i + 1, i] = 1.0 + 0.001 * i
    for got in _measure_modes(seqj, monkeypatch, P, want_order=True):
        assert _edge_set(got.mst) == {(i, i + 1) for i in range(N - 1)}
        assert got.eta == float(N)
        assert got.order.tolist() in (list(range(N)), list(range(N))[::-1])


def test_accumulate_matches_reference_formula(seqj, handle):
    rng = np.random.default_rng(9)
    Ds = rng.random((5, 40, 40)) + 1e-6
    etas = rng.random(5) * 5 + 1
    S = np.zeros((40, 40))
    for c in range(5):
        S += np.maximum(etas[c] * Ds[c], 2.220446049250313e-16)
    tot = 0.0
    for e in etas:
        tot += e
    assert np.array_equal(seqj.accumulate(Ds, etas), S / tot)     # same order of operations: bit-exact


@pytest.mark.parametrize("name", ["CITYBLOCK", "TOTALVARIATION", "CHEBYSHEV", "SQEUCLIDEAN"])
def test_elementwise_distances_metrics(seqj, oracle, handle, name):
    """the reference hands any Distances.PreMetric to `pairwise` (src/sequencer.jl:226).
    Cityblock, TotalVariation, Chebyshev and SqEuclidean run on the direct tile kernel; chunk matrices against the
    published definitions, and a whole `sequence` against the oracle (trees, eta, order)."""
    met = getattr(seqj, name)
    mid = getattr(oracle, name)
    rng = np.random.default_rng(17)
    for (m, N) in ((37, 150), (256, 333)):
        A = oracle.clamp_input(rng.random((m, N)))
        Dm = seqj.pairwise(met, A, normalize=True)
        P = A / A.sum(axis=0, keepdims=True)
        ref = oracle.chunk_distance(mid, P)
        assert np.array_equal(Dm, Dm.T) and np.all(np.diag(Dm) == 1e-6)
        rel_close(Dm, ref, 1e-9, atol=64 * m * 2.0 ** -53)
    A = rng.random((48, 90))
    r = seqj.sequence(A.copy(), scales=(1, 2, 3), metrics=(met, seqj.WASS1D), handle=handle)
    ref = oracle.sequence_oracle(A, (1, 2, 3), (mid, 1))
    assert np.array_equal(r.order, ref.order) and abs(r.eta - ref.eta) <= 1e-9 * ref.eta
    for g in ref.groups:
        key = (met if g.metric == mid else seqj.WASS1D, g.scale)
        rel_close(r.EOSeg[key][0], g.eta_chunk, 1e-9)
        assert np.array_equal(r.EOAlgScale[key][1].parents, g.parents)


def test_measure_dm_default_list_pass_at_a_ragged_size(seqj, handle, monkeypatch):
    """N = 6150 (>= 6144: the upper-triangle list pass is the default; 6150 is neither a multiple of 4 nor of the
    128-entry block edge, so the last block row / column and the row samples of the last rows are ragged): the tree, its
    root, eta and the order equal the Prim kernels' on a random matrix and on one with a block of duplicated objects."""
    rng = np.random.default_rng(17)
    N = 6150
    X = rng.random((N, 4))
    X[5000:5100] = X[100:200]                                  # 100 duplicated objects: exact ties at the eps floor
    G = X @ X.T
    sq = np.diag(G).copy()
    Dm = np.sqrt(np.maximum(sq[:, None] + sq[None, :] - 2.0 * G, 0.0)) + 1e-6
    np.fill_diagonal(Dm, 1e-6)
    Dm = np.triu(Dm, 1); Dm = Dm + Dm.T
    monkeypatch.delenv("SEQJ_KNN", raising=False)
    monkeypatch.delenv("SEQJ_MST", raising=False)
    a = seqj.measure_dm(Dm, want_order=True)
    monkeypatch.setenv("SEQJ_MST", "prim")
    b = seqj.measure_dm(Dm, want_order=True)
    monkeypatch.delenv("SEQJ_MST", raising=False)
    assert _edge_set(a.mst) == _edge_set(b.mst)
    assert a.root == b.root and a.eta == b.eta
    assert np.array_equal(a.bfs.parents, b.bfs.parents) and np.array_equal(a.order, b.order)
