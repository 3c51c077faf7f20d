"""Parity at BASELINE.json's full sizes through size-independent properties (the oracle cannot finish these sizes):
sampled entries of full-size distance matrices against direct NumPy formulas, planted-order recovery, determinism."""
import numpy as np
import pytest

from conftest import rel_close

pytestmark = pytest.mark.gpu


def _planted(N, L, seed):
    """Smooth one-parameter family (a bump moving across the grid, as in the Sequencer paper), columns shuffled."""
    x = np.arange(L)[:, None]
    mu = np.linspace(0.15 * L, 0.85 * L, N)[None, :]
    A = np.exp(-0.5 * ((x - mu) / (0.05 * L)) ** 2) + 1e-3
    perm = np.random.default_rng(seed).permutation(N)
    return np.asfortranarray(A[:, perm]), perm


@pytest.mark.parametrize("metric", [0, 1, 2, 3])
def test_fullsize_chunk_matrix_sampled_entries(seqj, oracle, handle, metric):
    """One chunk matrix of config 3 (N = 10,000, chunk length 256 = scale 16): 4,000 random entries against the
    reference formulas evaluated directly in NumPy; symmetry and the diagonal over the whole matrix."""
    N, m = 10000, 256
    A = oracle.clamp_input(np.random.default_rng(2732).random((m, N)))
    Dm = seqj.pairwise(metric, A, normalize=True, kl_full=False)
    assert Dm.shape == (N, N) and np.all(np.diag(Dm) == 1e-6)
    assert np.array_equal(Dm, Dm.T)
    P = A / A.sum(axis=0, keepdims=True)
    rng = np.random.default_rng(1)
    ii, jj = rng.integers(0, N, 4000), rng.integers(0, N, 4000)
    keep = ii != jj
    ii, jj = ii[keep], jj[keep]
    lo, hi = np.minimum(ii, jj), np.maximum(ii, jj)
    if metric == 0:
        ref = np.sqrt(((P[:, lo] - P[:, hi]) ** 2).sum(axis=0))
    elif metric == 2:
        ref = (P[:, lo] * np.log(P[:, lo] / P[:, hi])).sum(axis=0)         # upper triangle: KL(p_lo || p_hi)
    else:
        U = oracle.chunk_cdf(P)
        d = U[:, lo] - U[:, hi]
        ref = np.abs(d).sum(axis=0) if metric == 1 else np.sqrt(2.0) * np.sqrt((d * d).sum(axis=0))
    rel_close(Dm[lo, hi], np.abs(ref) + 1e-6, 1e-9, atol=64 * m * 2.0 ** -53)


def test_config3_full_size_planted_order(seqj, handle):
    """BASELINE config 3 at full size (N = 10,000 x L = 4,096, ALL_METRICS, scales (1,2,4,8,16)) on an ordered
    family: the planted order must come back exactly (up to reversal), eta = N, and a second run is identical."""
    N, L = 10000, 4096
    A, perm = _planted(N, L, 7)
    r = seqj.sequence(A, scales=(1, 2, 4, 8, 16), metrics=seqj.ALL_METRICS)
    rec = perm[r.order]
    assert np.array_equal(rec, np.arange(N)) or np.array_equal(rec, np.arange(N)[::-1])
    assert r.eta == float(N)
    assert len(r.EOAlgScale) == 20 and sum(len(v[0]) for v in r.EOSeg.values()) == 124
    r2 = seqj.sequence(A, scales=(1, 2, 4, 8, 16), metrics=seqj.ALL_METRICS)
    assert np.array_equal(r.order, r2.order) and r.eta == r2.eta
    for key in r.EOAlgScale:
        assert r.EOAlgScale[key][0] == r2.EOAlgScale[key][0]            # bit-reproducible run to run


def test_config2_full_size_properties(seqj, handle):
    """BASELINE config 2 (N = 2,000 x L = 1,000, EMD + Energy, scales (1,2,4,8)) at full size: permutation
    equivariance (shuffling the input columns permutes the order consistently) and planted-order recovery."""
    N, L = 2000, 1000
    A, perm = _planted(N, L, 3)
    mets, scales = (seqj.WASS1D, seqj.ENERGY), (1, 2, 4, 8)
    r = seqj.sequence(A, scales=scales, metrics=mets)
    rec = perm[r.order]
    assert np.array_equal(rec, np.arange(N)) or np.array_equal(rec, np.arange(N)[::-1])
    q = np.random.default_rng(9).permutation(N)
    rq = seqj.sequence(np.asfortranarray(A[:, q]), scales=scales, metrics=mets)
    recq = perm[q[rq.order]]
    assert np.array_equal(recq, np.arange(N)) or np.array_equal(recq, np.arange(N)[::-1])
    assert rq.eta == r.eta
