"""Small invocations of every kernel family through the C ABI, sized for compute-sanitizer
(`compute-sanitizer --tool memcheck|synccheck|racecheck python tools/sanitize_case.py`); results are also
checked against the oracle so that a run that "passes" the tool did compute something.  See profiles/."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import sequencerj_b200 as s  # noqa: E402
from oracle import seq_oracle as O  # noqa: E402  (tools/ is test infrastructure)

which = sys.argv[1] if len(sys.argv) > 1 else "all"
rng = np.random.default_rng(2732)


def check_seq(A, scales, metrics, ids):
    r = s.sequence(A.copy(), scales=scales, metrics=metrics)
    ref = O.sequence_oracle(A, scales, metrics=ids)
    assert np.array_equal(r.order, ref.order)
    assert abs(r.eta - ref.eta) <= 1e-9 * abs(ref.eta)
    return r


if which in ("all", "seq"):
    # config 1: tiles + derived scales (1,2,4 of M=50 do not nest: tiles per scale), batched Prim, tree, combine, fuse
    check_seq(rng.random((50, 100)), (1, 2, 4), s.ALL_METRICS, (0, 1, 2, 3))
    # nested scales (derive kernels), N not a multiple of the tile edge, cluster Prim (few graphs)
    check_seq(rng.random((64, 203)), (1, 2, 4), s.ALL_METRICS, (0, 1, 2, 3))
    # smooth ordered family: fix-up lists and dense-tile direct kernels
    x = np.arange(96)[:, None]
    mu = np.linspace(10, 86, 150)[None, :]
    A = np.exp(-0.5 * ((x - mu) / 6.0) ** 2) + 1e-3
    check_seq(np.asfortranarray(A[:, rng.permutation(150)]), (1, 2), s.ALL_METRICS, (0, 1, 2, 3))
    # single-CTA Prim tier: many graphs (scales 1..8 of 4 metrics = 60 chunk graphs)
    check_seq(rng.random((64, 70)), (1, 2, 4, 8), s.ALL_METRICS, (0, 1, 2, 3))
    print("seq ok")

if which in ("all", "units"):
    A = O.clamp_input(rng.random((37, 131)))
    for metric in (0, 1, 2, 3):
        Dm = s.pairwise(metric, A, normalize=True)
        assert Dm.shape == (131, 131)
    chunks = s.splitnorm(A, 3)
    assert len(chunks) == 3
    Dm = s.pairwise(1, A)
    m = s.measure_dm(Dm, want_order=True)
    assert len(m.order) == 131
    S = s.accumulate([Dm, Dm * 1.5], [2.0, 3.0])
    assert S.shape == Dm.shape
    v = s.emd([3.4, 3.9, 7.5, 7.8], [4.5, 1.4], [1.4, 0.9, 3.1, 7.2], [3.2, 3.5])
    assert abs(v - 4.0781331438047861) < 1e-12
    v = s.energy([0.7, 7.4, 2.4, 6.8], [1.4, 8.0], [2.1, 4.2, 7.4, 8.0], [7.6, 8.8])
    assert abs(v - 0.8800334097615822) < 1e-12
    # explicit grid types and PeriodicEuclidean (direct tile kernel)
    r = s.sequence(rng.random((40, 90)), scales=(1, 2), metrics=(s.EMD, s.Energy), grid=np.cumsum(rng.random(40) + 0.5))
    assert len(r.order) == 90
    r = s.sequence(rng.random((24, 77)), scales=(1,), metrics=(s.PeriodicEuclidean(np.full(24, 0.75)),))
    assert len(r.order) == 77
    print("units ok")
