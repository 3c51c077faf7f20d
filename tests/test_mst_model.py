"""CPU: the NumPy model of the kNN-filtered Boruvka (tools/knn_boruvka_proto.py, the algorithm of
sequencerj.jl_b200/csrc/kernels_mst.cuh) against the oracle's Kruskal (reference `_mst`, src/sequencer.jl:358-374):
every edge it adds belongs to the reference's MST, it always finishes once exhausted rows are rescanned, and the tree
is the reference's also when weights tie."""
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools"))
import knn_boruvka_proto as model  # noqa: E402


def _sym(W):
    W = np.triu(W, 1)
    return W + W.T


def _cases():
    rng = np.random.default_rng(5)
    yield "random", _sym(rng.random((150, 150)) + 0.01)
    yield "integer ties", _sym(rng.integers(1, 4, (120, 120)).astype(float))
    yield "all equal", _sym(np.full((60, 60), 1e-6))
    pts = np.concatenate([rng.normal(3.0 * c, 0.01, (4, 40)) for c in range(4)], axis=1)
    yield "clusters larger than the list", np.sqrt(((pts[:, :, None] - pts[:, None, :]) ** 2).sum(0)) + 1e-6
    base = _sym(rng.random((50, 50)) + 0.01)
    yield "triplicated objects", np.kron(np.ones((3, 3)), base) + 1e-6
    for n in (2, 3, 5, 33):
        yield f"tiny {n}", _sym(rng.random((n, n)) + 0.01)


@pytest.mark.parametrize("name,D", list(_cases()), ids=[c[0] for c in _cases()])
def test_model_returns_the_reference_tree(oracle, name, D):
    D = D.copy()
    np.fill_diagonal(D, 1e-6)
    lists, bound = model.knn_lists(D)
    for v, lv in enumerate(lists):                       # lists are sorted by (weight, j) and exclude the diagonal
        assert lv == sorted(lv) and all(j != v for _, j in lv) and len(lv) <= model.K
    edges, ncomp, _ = model.boruvka(D, lists, bound, rescan=True)
    ei, ej, ew = oracle.kruskal_mst(D)
    assert ncomp == 1
    assert {(a, b) for a, b, _ in edges} == set(zip(ei.tolist(), ej.tolist()))
    wref = dict(zip(zip(ei.tolist(), ej.tolist()), model.bits(ew).tolist()))
    assert all(wref[(a, b)] == w for a, b, w in edges)


@pytest.mark.parametrize("name,D", list(_cases()), ids=[c[0] for c in _cases()])
def test_round2_rescan_rules_return_the_reference_tree(oracle, name, D):
    """The round-2 kernels rescan a row only when its graph is completely blocked and keep ONE list entry per
    component in a rescanned list (kernels_mst.cuh): still the reference's tree, ties included, with fewer rescans."""
    D = D.copy()
    np.fill_diagonal(D, 1e-6)
    ei, ej, ew = oracle.kruskal_mst(D)
    ref = set(zip(ei.tolist(), ej.tolist()))
    counts = {}
    for dedupe, lazy in ((False, False), (True, False), (False, True), (True, True)):
        lists, bound = model.knn_lists(D)
        edges, ncomp, _ = model.boruvka(D, lists, bound, rescan=True, dedupe=dedupe, lazy=lazy)
        assert ncomp == 1 and {(a, b) for a, b, _ in edges} == ref, (dedupe, lazy)
        counts[(dedupe, lazy)] = model.boruvka.last_rescans
    assert counts[(True, True)] <= counts[(False, False)]


def test_dedupe_cuts_rescans_of_tight_clusters(oracle):
    """20 tight clusters of 30 near-equidistant objects (more than the 16 list entries): with one entry per component
    a rescanned row sees the lightest edge into each of the other clusters at once instead of running dry again after
    a merge -- one pass over the rows instead of ~1.5."""
    rng = np.random.default_rng(11)
    centers = rng.random((64, 20))
    A = np.abs(np.repeat(centers, 30, axis=1) * (1 + 0.01 * rng.standard_normal((64, 600))))
    D = np.sqrt(((A[:, :, None] - A[:, None, :]) ** 2).sum(0)) + 1e-6
    np.fill_diagonal(D, 1e-6)
    ei, ej, _ = oracle.kruskal_mst(D)
    res = {}
    for key, (dedupe, lazy) in {"r01": (False, False), "r02": (True, True)}.items():
        lists, bound = model.knn_lists(D)
        edges, ncomp, _ = model.boruvka(D, lists, bound, rescan=True, dedupe=dedupe, lazy=lazy)
        assert ncomp == 1 and {(a, b) for a, b, _ in edges} == set(zip(ei.tolist(), ej.tolist()))
        res[key] = model.boruvka.last_rescans
    assert res["r02"] < res["r01"], res
    assert res["r02"] <= D.shape[0], res


def test_model_without_rescans_only_adds_mst_edges(oracle):
    """Stuck components simply stop hooking: what has been added so far is still part of the reference's tree (this
    is what makes the Prim fallback of unfinished graphs safe)."""
    rng = np.random.default_rng(7)
    pts = np.concatenate([rng.normal(3.0 * c, 0.01, (4, 60)) for c in range(3)], axis=1)
    D = np.sqrt(((pts[:, :, None] - pts[:, None, :]) ** 2).sum(0)) + 1e-6
    np.fill_diagonal(D, 1e-6)
    lists, bound = model.knn_lists(D)
    edges, ncomp, _ = model.boruvka(D, lists, bound, rescan=False)
    ei, ej, _ = oracle.kruskal_mst(D)
    assert ncomp > 1                                       # the clusters exhaust their 16-entry lists
    assert {(a, b) for a, b, _ in edges} <= set(zip(ei.tolist(), ej.tolist()))


def test_streaming_threshold_variant_keeps_the_list_contract(oracle):
    """Round-2 candidate for the list kernel (a warp-wide threshold, tools/knn_boruvka_proto.py:row_list_streaming):
    its lists must satisfy the same contract -- sorted by (weight, j), every unlisted edge weighs >= the bound and
    comes after every listed one in (weight, j) order -- and drive Boruvka to the reference's tree, ties included."""
    rng = np.random.default_rng(3)
    for D in (_sym(rng.random((200, 200)) + 0.01), _sym(rng.integers(1, 4, (150, 150)).astype(float))):
        np.fill_diagonal(D, 1e-6)
        N = D.shape[0]
        W = model.bits(np.maximum(D, model.EPS))
        lists, bound = [], np.empty(N, dtype=np.uint64)
        for v in range(N):
            ex = np.zeros(N, dtype=bool); ex[v] = True
            lv, B, _ = model.row_list_streaming(W, v, ex)
            allv = sorted((int(W[v, j]), j) for j in range(N) if j != v)
            assert lv == allv[:len(lv)]                          # exactly the smallest entries, in order
            assert all(w >= B for w, _ in allv[len(lv):]) or B == int(model.NONE)
            lists.append(lv); bound[v] = B
        edges, ncomp, _ = model.boruvka(D, lists, bound, rescan=True)
        ei, ej, _ = oracle.kruskal_mst(D)
        assert ncomp == 1 and {(a, b) for a, b, _ in edges} == set(zip(ei.tolist(), ej.tolist()))


def test_threshold_lists_keep_the_list_contract(oracle):
    """Round 2, the upper-triangle list pass (kernels_mst.cuh: knn_sample / knn_tiles / knn_merge kernels, modelled by
    tools/knn_boruvka_proto.py:threshold_lists): every list holds exactly the smallest entries of its row in
    (weight, j) order, every unlisted edge weighs >= the bound, rows that overflow the candidate buffer come back empty
    with bound 0 (= rescan me), and Boruvka on these lists returns the reference's tree -- random weights, heavy ties
    (every row overflows) and duplicated objects."""
    rng = np.random.default_rng(11)
    base = _sym(rng.random((60, 60)) + 0.01)
    cases = (_sym(rng.random((300, 300)) + 0.01), _sym(rng.integers(1, 4, (130, 130)).astype(float)),
             np.kron(np.ones((3, 3)), base))
    for D in cases:
        np.fill_diagonal(D, 1e-6)
        N = D.shape[0]
        W = model.bits(np.maximum(D, model.EPS))
        for cap, div in ((96, 8), (8, 8), (96, 2)):
            lists, bound = model.threshold_lists(D, sample_div=div, cap=cap)
            for v in range(N):
                allv = sorted((int(W[v, j]), j) for j in range(N) if j != v)
                assert lists[v] == allv[:len(lists[v])]
                assert all(w >= int(bound[v]) for w, _ in allv[len(lists[v]):])
            edges, ncomp, _ = model.boruvka(D, lists, bound, rescan=True)
            ei, ej, _ = oracle.kruskal_mst(D)
            assert ncomp == 1 and {(a, b) for a, b, _ in edges} == set(zip(ei.tolist(), ej.tolist()))
