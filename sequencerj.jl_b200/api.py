"""Host-side mirror of the SequencerJ API above the C ABI (include/seqj.h).

Mirrors, name for name, what the reference exports (src/SequencerJ.jl:24-31) for the hot path:
``sequence(A; scales, metrics, grid)`` (src/sequencer.jl:105-145), ``SequencerResult`` and its
accessors ``D, mst, elong, order`` (:13-52), the metric constants ``L2, WASS1D, KLD, ENERGY,
ALL_METRICS`` (src/distancemetrics.jl:234-257), ``EMD`` / ``Energy`` / ``emd`` / ``energy`` on the
unit grid, ``elongation``, ``ensuregrid``, ``autoscale`` and ``prettyp``.  Julia is not installed
in this image, so this Python module plays the role of the Julia wrapper in
``sequencerj.jl_b200/julia/SequencerJB200.jl``; both are thin shells over the same C entry points.

Differences from Julia, by design: indices are 0-based (the Julia wrapper adds 1); the input is
converted to Float64 up front.  All compute happens in libseqj_b200.so on a
B200; nothing here falls back to the CPU.
"""
from __future__ import annotations

import ctypes as C
import logging
from dataclasses import dataclass, field

import numpy as np

from . import _lib

logger = logging.getLogger("sequencerj_b200")

EPS_FLOOR = 1e-6           # const ϵ            (src/distancemetrics.jl:3)
L2THR = 1e-7               # const L2THR        (src/distancemetrics.jl:6)
JULIA_EPS = 2.220446049250313e-16
FIB = [1, 2, 3, 5, 8, 13, 21, 34, 55, 89, 144, 233, 377, 610, 987, 1597, 2584, 5168]  # sequencer.jl:148


@dataclass(frozen=True)
class Metric:
    """A metric instance as the reference passes to `sequence` (Distances.PreMetric objects)."""
    id: int
    name: str

    def __repr__(self):
        return self.name


L2 = Metric(0, "Euclidean(1.0e-7)")        # Distances.Euclidean(L2THR)   distancemetrics.jl:234
WASS1D = Metric(1, "EMD(nothing)")         # EMD()                        distancemetrics.jl:247
KLD = Metric(2, "KLDivergence()")          # Distances.KLDivergence()     distancemetrics.jl:239
ENERGY = Metric(3, "Energy(nothing)")      # Energy()                     distancemetrics.jl:254
ALL_METRICS = (L2, WASS1D, KLD, ENERGY)    # distancemetrics.jl:257
# what the reference builds when the TYPES EMD / Energy are passed as metrics: EMD((lgrid, lgrid)) with the
# chunk-local slice of `grid` (sequencer.jl:217-219) -- delta-weighted CDF distances on the explicit grid
EMD_GRID = Metric(4, "EMD(grid)")
ENERGY_GRID = Metric(5, "Energy(grid)")
# further element-wise Distances.jl metrics (the reference hands any PreMetric to `pairwise`, sequencer.jl:226):
# evaluated on the chunk PDFs by the direct tile kernel
CITYBLOCK = Metric(7, "Cityblock()")
TOTALVARIATION = Metric(8, "TotalVariation()")
CHEBYSHEV = Metric(9, "Chebyshev()")
SQEUCLIDEAN = Metric(10, "SqEuclidean()")
_BY_ID = {m.id: m for m in ALL_METRICS + (EMD_GRID, ENERGY_GRID, CITYBLOCK, TOTALVARIATION, CHEBYSHEV, SQEUCLIDEAN)}


class PeriodicEuclidean(Metric):
    """`Distances.PeriodicEuclidean(periods)` as a `sequence` metric (reference test/runtests.jl:189-236):
    d(a, b) = sqrt(sum_k min(s_k, p_k - s_k)^2), s_k = |a_k - b_k| mod p_k.  One period per row; like Distances.jl
    it only evaluates vectors of exactly that length, i.e. scales whose single chunk is the whole column."""

    def __init__(self, periods):
        per = np.ascontiguousarray(periods, dtype=np.float64)
        if per.ndim != 1 or len(per) < 1:
            raise ValueError("periods must be a non-empty vector")
        object.__setattr__(self, "periods", per)
        # the reference's `show`: PeriodicEuclidean(n) [min,max]  (src/distancemetrics.jl:260)
        object.__setattr__(self, "id", 6)
        object.__setattr__(self, "name", f"PeriodicEuclidean({len(self.periods)}) [{self.periods.min()},{self.periods.max()}]")

    def __eq__(self, other):
        return isinstance(other, PeriodicEuclidean) and np.array_equal(self.periods, other.periods)

    def __hash__(self):
        return hash((6, self.periods.tobytes()))

_default_handle = None


def default_handle() -> _lib.Handle:
    global _default_handle
    if _default_handle is None:
        _default_handle = _lib.Handle()
    return _default_handle


# ------------------------------------------------------------------------------------------
@dataclass
class WeightedTree:
    """Minimum spanning tree as an edge list (stands in for the SimpleWeightedGraph of the reference)."""
    nv: int
    src: np.ndarray
    dst: np.ndarray
    weight: np.ndarray

    def edges(self):
        return list(zip(self.src.tolist(), self.dst.tolist(), self.weight.tolist()))


@dataclass
class BFSTree:
    """Directed BFS tree (parent -> child) rooted at the least central vertex; `parents[root] == root`."""
    root: int
    parents: np.ndarray

    def outneighbors(self, v):
        return [int(c) for c in np.nonzero(self.parents == v)[0] if c != self.root]

    def edges(self):
        return [(int(self.parents[c]), int(c)) for c in range(len(self.parents)) if c != self.root]


class SequencerResult:
    """Mirror of the reference's SequencerResult (src/sequencer.jl:13-20).

    EOSeg[(metric, scale)]      = (etas, [BFSTree per chunk])
    EOAlgScale[(metric, scale)] = (eta, BFSTree)
    D, mst, eta (η), order      = final distance matrix, MST, elongation, column order (0-based)
    """

    def __init__(self, N, EOSeg, EOAlgScale, D_triplets, mst, eta, order, root, parents, timings, launches,
                 group_msts):
        self.N = N
        self.EOSeg = EOSeg
        self.EOAlgScale = EOAlgScale
        self._D_triplets = D_triplets
        self.mst = mst
        self.eta = eta
        self.η = eta
        self.order = order
        self.root = root
        self.parents = parents
        self.timings = timings
        self.launches = launches
        self.group_msts = group_msts
        self._D = None
        self._lazy = None          # (handle, C result pointer) while the D triplets have not been fetched yet

    def _fetch_triplets(self):
        if self._lazy is not None:
            handle, res = self._lazy
            self._lazy = None
            try:
                lib = handle.lib
                nnz = C.c_int64()
                handle.check(lib.seqj_result_D_nnz(res, C.byref(nnz)))
                ti = np.empty(nnz.value, dtype=np.int32); tj = np.empty(nnz.value, dtype=np.int32); tv = np.empty(nnz.value)
                handle.check(lib.seqj_result_D_triplets(res, _lib.iptr(ti), _lib.iptr(tj), _lib.dptr(tv)))
                self._D_triplets = (ti, tj, tv)
            finally:
                handle.lib.seqj_result_free(res)

    def __del__(self):
        if getattr(self, "_lazy", None) is not None:
            handle, res = self._lazy
            self._lazy = None
            try:
                handle.lib.seqj_result_free(res)
            except Exception:
                pass

    @property
    def D(self):
        """Dense final distance matrix: +Inf off the edge union, 1e-6 on the diagonal (sequencer.jl:269,360)."""
        if self._D is None:
            self._fetch_triplets()
            i, j, v = self._D_triplets
            D = np.full((self.N, self.N), np.inf)
            np.fill_diagonal(D, EPS_FLOOR)
            D[i, j] = v
            D[j, i] = v
            self._D = D
        return self._D

    @property
    def D_triplets(self):
        """(i, j, value) of the final distance matrix on the edge union, i < j (fetched from the library and
        deduplicated on first access)."""
        self._fetch_triplets()
        return self._D_triplets

    def __repr__(self):
        return f"Sequencer Result: η = {self.eta:.4g}, order = {prettyp(self.order, 5)}"


def D(r: SequencerResult):
    return r.D


def mst(r: SequencerResult):
    return r.mst


def elong(r: SequencerResult):
    return r.eta


def order(r: SequencerResult):
    return r.order


def prettyp(v, length: int = 3) -> str:
    """src/utils.jl:48-53."""
    v = list(v)
    if len(v) < 7:
        return ",".join(str(x) for x in v)
    head = ",".join(str(x) for x in v[:length])
    tail = ",".join(str(x) for x in v[len(v) - length:])
    return head + "..." + tail


def ensuregrid(A, grid=None):
    """src/sequencer.jl:296-304. With the ALL_METRICS instances the grid is never used by the
    distance functors; only its length is checked."""
    M = A.shape[0]
    if grid is None:
        grid = np.arange(1, M + 1, dtype=np.float64)
    grid = np.asarray(grid, dtype=np.float64)
    assert len(grid) == M, f"Grid length must match dims 1 (row count) of A. Got grid:{len(grid)} and A:{M}"
    return grid


def _normalize_metrics(metrics):
    """Instances (L2, WASS1D, KLD, ENERGY) stay; the TYPES EMD / Energy select the explicit-grid variants, as the
    reference's `k in (EMD, Energy)` test does (sequencer.jl:217)."""
    if metrics is None:
        return ALL_METRICS
    if isinstance(metrics, (Metric, type)):
        metrics = (metrics,)
    out = []
    for m in metrics:
        if isinstance(m, Metric):
            out.append(m)
        elif m is EMD:
            out.append(EMD_GRID)
        elif m is Energy:
            out.append(ENERGY_GRID)
        elif isinstance(m, (EMD, Energy)) and m.grids is None:      # EMD() / Energy() instances == WASS1D / ENERGY
            out.append(WASS1D if isinstance(m, EMD) else ENERGY)
        else:
            raise TypeError(f"unsupported metric {m!r}: the B200 path implements L2, WASS1D, KLD, ENERGY, the EMD / Energy "
                            "types with an explicit grid, PeriodicEuclidean, Cityblock, TotalVariation, Chebyshev and "
                            "SqEuclidean (no CPU fallback)")
    return tuple(out)


def _tuplify(x):
    if isinstance(x, (Metric, int, np.integer)):
        return (x,)
    return tuple(x)


def _prepare_input(A):
    """Validation + clamp of `sequence` (src/sequencer.jl:111-130). Returns an (N, M) C-contiguous
    float64 array = Julia column-major M x N. Clamps the caller's array in place when it is a
    writable float64 array (the reference's clamp! is caller-visible)."""
    if isinstance(A, (list, tuple)):
        A = np.stack([np.asarray(c, dtype=np.float64) for c in A], axis=1)     # hcat(A...)
    A = np.asarray(A)
    if A.ndim == 1:
        A = A.reshape(1, -1)
    if A.ndim != 2:
        raise ValueError("A must be a vector or an M x N matrix")
    mn, mx = (A.min(), A.max()) if A.size else (0.0, 0.0)      # two streaming passes; NaN propagates into both
    if not (mn >= 0 and np.isfinite(mx)):
        raise AssertionError("Input data cannot contain negative, NaN or infinite values.")
    if A.dtype == np.float64 and A.flags.writeable:
        if mn < JULIA_EPS:
            np.maximum(A, JULIA_EPS, out=A)                                     # clamp!(A, eps(), typemax)
        Af = A
    else:
        Af = np.maximum(A.astype(np.float64), JULIA_EPS)
    return np.ascontiguousarray(Af.T)                                           # object-major (N, M)


def _layout_input(A):
    """Layout-only conditioning for the hot call: no pass over the data on the host.  The value checks of
    `sequence` (src/sequencer.jl:111) and the clamp (:130) run on the device copy inside seqj_sequence;
    `orig` is the caller's array when the reference's in-place `clamp!` can be made visible on it."""
    if isinstance(A, (list, tuple)):
        A = np.stack([np.asarray(c, dtype=np.float64) for c in A], axis=1)     # hcat(A...)
    A = np.asarray(A)
    if A.ndim == 1:
        A = A.reshape(1, -1)
    if A.ndim != 2:
        raise ValueError("A must be a vector or an M x N matrix")
    orig = A if (A.dtype in (np.float64, np.float32) and A.flags.writeable) else None
    # Float32 stays Float32 across the ABI (seqj_sequence_f32, quirk Q2); every other eltype is promoted to Float64
    dt = np.float32 if A.dtype == np.float32 else np.float64
    return np.ascontiguousarray(A.T, dtype=dt), orig              # no copy for a Fortran-ordered float64 / float32 A


def _collect(handle, res, N, metrics, scales, want_final=True, partial=False):
    lib = handle.lib
    chunk_ranges = {}
    ngroups = C.c_int()
    n64 = C.c_int64()
    handle.check(lib.seqj_result_dims(res, C.byref(n64), C.byref(ngroups)))
    EOSeg, EOAlgScale, group_msts, group_ms = {}, {}, {}, {}
    by_id = dict(_BY_ID)
    by_id.update({m.id: m for m in (metrics or ()) if isinstance(m, PeriodicEuclidean)})
    for g in range(ngroups.value):
        met, sc, nch, comp = C.c_int(), C.c_int(), C.c_int(), C.c_int()
        handle.check(lib.seqj_result_group_info(res, g, C.byref(met), C.byref(sc), C.byref(nch), C.byref(comp)))
        if not comp.value:
            continue
        key = (by_id[met.value], sc.value)
        lo, hi = C.c_int(), C.c_int()
        handle.check(lib.seqj_result_group_range(res, g, C.byref(lo), C.byref(hi)))
        nloc = hi.value - lo.value
        etas = np.empty(nloc)
        roots = np.empty(nloc, dtype=np.int32)
        pars = np.empty((nloc, N), dtype=np.int32)
        handle.check(lib.seqj_result_group_chunks(res, g, _lib.dptr(etas), _lib.iptr(roots), _lib.iptr(pars)))
        EOSeg[key] = (etas.tolist(), [BFSTree(int(roots[c]), pars[c]) for c in range(nloc)])
        chunk_ranges[key] = (lo.value, hi.value)
        gms = C.c_double()
        handle.check(lib.seqj_result_group_elapsed(res, g, C.byref(gms)))
        group_ms[key] = gms.value
        if partial:
            continue
        eta = C.c_double(); root = C.c_int32()
        par = np.empty(N, dtype=np.int32)
        mi = np.empty(N - 1, dtype=np.int32); mj = np.empty(N - 1, dtype=np.int32); mw = np.empty(N - 1)
        handle.check(lib.seqj_result_group(res, g, C.byref(eta), C.byref(root), _lib.iptr(par), _lib.iptr(mi),
                                           _lib.iptr(mj), _lib.dptr(mw)))
        EOAlgScale[key] = (eta.value, BFSTree(root.value, par))
        group_msts[key] = WeightedTree(N, mi, mj, mw)
    tm = np.zeros(_lib.SEQJ_NTIMERS)
    handle.check(lib.seqj_result_timings(res, _lib.dptr(tm)))
    launches = C.c_int64()
    handle.check(lib.seqj_result_launches(res, C.byref(launches)))
    timings = dict(zip(_lib.TIMER_NAMES, tm.tolist()))
    lo, hi = C.c_double(), C.c_double()
    handle.check(lib.seqj_result_input_range(res, C.byref(lo), C.byref(hi)))
    timings["input_min"] = lo.value
    nder = C.c_int()
    handle.check(lib.seqj_result_derived_groups(res, C.byref(nder)))
    timings["derived_groups"] = nder.value
    nw = C.c_int()
    handle.check(lib.seqj_result_waves(res, C.byref(nw)))
    timings["waves"] = nw.value
    probe_vals = None
    nprobes = getattr(handle, "_nprobes", 0)
    if nprobes:
        probe_vals = np.empty(nprobes)
        handle.check(lib.seqj_result_probes(res, _lib.dptr(probe_vals), nprobes))
    if not want_final:
        r = SequencerResult(N, EOSeg, EOAlgScale, None, None, None, None, None, None, timings, launches.value,
                            group_msts)
        r.chunk_ranges = chunk_ranges
        r.probes = probe_vals
        r.group_ms = group_ms
        return r
    eta = C.c_double(); root = C.c_int32()
    order_ = np.empty(N, dtype=np.int32); par = np.empty(N, dtype=np.int32)
    mi = np.empty(N - 1, dtype=np.int32); mj = np.empty(N - 1, dtype=np.int32); mw = np.empty(N - 1)
    handle.check(lib.seqj_result_final(res, C.byref(eta), C.byref(root), _lib.iptr(order_), _lib.iptr(par),
                                       _lib.iptr(mi), _lib.iptr(mj), _lib.dptr(mw)))
    r = SequencerResult(N, EOSeg, EOAlgScale, None, WeightedTree(N, mi, mj, mw), eta.value, order_,
                        root.value, par, timings, launches.value, group_msts)
    r._lazy = (handle, res)        # the result object now owns the C result until the triplets are fetched
    r.probes = probe_vals
    r.group_ms = group_ms
    return r


def _collect_and_release(handle, res, N, metrics, scales, want_final=True, partial=False):
    """_collect, then free the C result unless the returned object took ownership of it (lazy D triplets)."""
    r = None
    try:
        r = _collect(handle, res, N, metrics, scales, want_final=want_final, partial=partial)
        return r
    finally:
        if r is None or r._lazy is None:
            handle.lib.seqj_result_free(res)


def _sequence(At, scales, metrics, handle=None, group_mask=None, device_ptr=None, probes=None, pack=None):
    """`_sequence` (src/sequencer.jl:184-276) through the C ABI. At: (N, M) float64 C-contiguous host
    array, or (with device_ptr) the shape only.  `probes` = (group, chunk, i, j) int arrays (seqj_set_probes):
    the sampled matrix entries come back as `result.probes`.  `pack` = device tensor (anything with .data_ptr())
    of at least [computed groups, seqj_pack_stride(N)] doubles: with a group mask and a device input the computed
    groups' (eta, MST) are also left there for the device-side exchange (seqj_sequence_shard)."""
    handle = handle or default_handle()
    if probes is not None:
        pg, pc, pi, pj = (np.ascontiguousarray(x, dtype=np.int32) for x in probes)
        assert pg.shape == pc.shape == pi.shape == pj.shape and pg.ndim == 1
        handle.check(handle.lib.seqj_set_probes(handle.h, _lib.iptr(pg), _lib.iptr(pc), _lib.iptr(pi), _lib.iptr(pj), len(pg)))
        handle._nprobes = len(pg)
        try:
            return _sequence(At, scales, metrics, handle, group_mask, device_ptr, pack=pack)
        finally:
            handle._nprobes = 0
            handle.lib.seqj_set_probes(handle.h, None, None, None, None, 0)
    N, M = At.shape
    mids = np.array([m.id for m in metrics], dtype=np.int32)
    scs = np.array(scales, dtype=np.int32)
    res = C.c_void_p()
    mask = None
    if group_mask is not None:
        mask = np.ascontiguousarray(group_mask, dtype=np.uint8).reshape(-1)
        assert mask.size == len(mids) * len(scs)
    maskp = mask.ctypes.data_as(_lib.c_u8p) if mask is not None else None
    if pack is not None:
        assert device_ptr is not None and mask is not None, "the device-side exchange needs a device input and a group mask"
        st = handle.lib.seqj_sequence_shard(handle.h, C.c_void_p(device_ptr), M, N, _lib.iptr(mids), len(mids),
                                            _lib.iptr(scs), len(scs), EPS_FLOOR, L2THR, maskp,
                                            C.c_void_p(pack.data_ptr()), C.byref(res))
    elif device_ptr is not None:
        st = handle.lib.seqj_sequence_dev(handle.h, C.c_void_p(device_ptr), M, N, _lib.iptr(mids), len(mids),
                                          _lib.iptr(scs), len(scs), EPS_FLOOR, L2THR, maskp, C.byref(res))
    elif At.dtype == np.float32:
        st = handle.lib.seqj_sequence_f32(handle.h, At.ctypes.data_as(C.c_void_p), M, N, _lib.iptr(mids), len(mids),
                                          _lib.iptr(scs), len(scs), EPS_FLOOR, L2THR, maskp, C.byref(res))
    else:
        st = handle.lib.seqj_sequence(handle.h, At.ctypes.data_as(C.c_void_p), M, N, _lib.iptr(mids), len(mids),
                                      _lib.iptr(scs), len(scs), EPS_FLOOR, L2THR, maskp, C.byref(res))
    handle.check(st)
    return _collect_and_release(handle, res, N, metrics, scales, want_final=group_mask is None)


def _sequence_partial(shape, scales, metrics, chunk_lo, chunk_hi, S_out_ptr, S_stride, device_ptr, handle=None):
    """Chunk-major shard of `_sequence` (seqj_sequence_partial): chunks [lo, hi) of every group, eta-weighted
    partial sums left in the caller's device buffer.  Returns a result holding the chunk-level outputs."""
    handle = handle or default_handle()
    N, M = shape
    mids = np.array([m.id for m in metrics], dtype=np.int32)
    scs = np.array(scales, dtype=np.int32)
    lo = np.ascontiguousarray(chunk_lo, dtype=np.int32).reshape(-1)
    hi = np.ascontiguousarray(chunk_hi, dtype=np.int32).reshape(-1)
    assert lo.size == hi.size == len(mids) * len(scs)
    res = C.c_void_p()
    handle.check(handle.lib.seqj_sequence_partial(handle.h, C.c_void_p(device_ptr), M, N, _lib.iptr(mids), len(mids),
                                                  _lib.iptr(scs), len(scs), EPS_FLOOR, L2THR, _lib.iptr(lo),
                                                  _lib.iptr(hi), C.c_void_p(S_out_ptr), int(S_stride), C.byref(res)))
    return _collect_and_release(handle, res, N, metrics, scales, want_final=False, partial=True)


def groups_finish(S_ptr, S_stride, slots, N, eta_sums, eps_floor=EPS_FLOOR, handle=None):
    """Owner side of the chunk-major path (seqj_groups_finish): divide the reduced accumulators by sum(etas)
    and measure them.  Returns [(eta, BFSTree, WeightedTree)] per slot."""
    handle = handle or default_handle()
    ng = len(slots)
    slots_a = np.ascontiguousarray(slots, dtype=np.int32)
    sums = np.ascontiguousarray(eta_sums, dtype=np.float64)
    eta = np.empty(ng); root = np.empty(ng, dtype=np.int32); par = np.empty((ng, N), dtype=np.int32)
    mi = np.empty((ng, N - 1), dtype=np.int32); mj = np.empty((ng, N - 1), dtype=np.int32); mw = np.empty((ng, N - 1))
    handle.check(handle.lib.seqj_groups_finish(handle.h, C.c_void_p(S_ptr), int(S_stride), ng, _lib.iptr(slots_a), N,
                                               _lib.dptr(sums), eps_floor, _lib.dptr(eta), _lib.iptr(root),
                                               _lib.iptr(par), _lib.iptr(mi), _lib.iptr(mj), _lib.dptr(mw)))
    return [(float(eta[g]), BFSTree(int(root[g]), par[g]), WeightedTree(N, mi[g], mj[g], mw[g])) for g in range(ng)]


def _set_periods(metrics, handle):
    per = [m for m in metrics if isinstance(m, PeriodicEuclidean)]
    if len({m for m in per}) > 1:
        raise ValueError("one PeriodicEuclidean metric per call")
    if per:
        h_ = handle or default_handle()
        h_.check(h_.lib.seqj_set_periods(h_.h, _lib.dptr(per[0].periods), len(per[0].periods)))


def sequence(A, scales=None, metrics=ALL_METRICS, grid=None, handle=None, rng=None) -> SequencerResult:
    """Drop-in for `sequence(A; scales, metrics, grid)` (src/sequencer.jl:105-145).

    A: M x N (rows = samples, columns = objects), non-negative and finite.
    """
    At, orig = _layout_input(A)
    N, M = At.shape
    logger.info("Sequencing data with\n    shape: (%d, %d)\n    metric(s): %s\n    scale(s): %s", M, N, metrics, scales)
    ensuregrid(At.T, grid)
    metrics = _normalize_metrics(metrics)
    if any(m.id in (EMD_GRID.id, ENERGY_GRID.id) for m in metrics):
        h_ = handle or default_handle()
        g = np.ascontiguousarray(ensuregrid(At.T, grid), dtype=np.float64)
        h_.check(h_.lib.seqj_set_grid(h_.h, _lib.dptr(g), len(g)))
    _set_periods(metrics, handle)
    if scales is None:
        scales = autoscale(At.T, metrics=metrics, grid=grid, handle=handle, rng=rng, _At=At)
        logger.info("After autoscaling, best scale = %s", scales)
    scales = tuple(int(s) for s in _tuplify(scales))
    for s in scales:
        assert 1 <= s <= M, f"Scale ({s}) cannot be larger than the number of data elements ({M})."
    r = _sequence(At, scales, metrics, handle)
    if orig is not None and r.timings.get("input_min", 1.0) < JULIA_EPS:
        np.maximum(orig, orig.dtype.type(JULIA_EPS), out=orig)     # clamp!(A, eps(), typemax): caller-visible (sequencer.jl:130)
    for (m, l), (eta, _) in r.EOAlgScale.items():       # "$(k) at scale $(l): η = ... (...s)"  (src/sequencer.jl:260)
        logger.info("%s at scale %d: η = %.4g (%.2gs)", m, l, eta, r.group_ms.get((m, l), 0.0) * 1e-3)
    logger.info("Final: η = %.4g sequence = %s", r.eta, prettyp(r.order, 5))
    return r


def autoscale(A, scales=FIB, metrics=ALL_METRICS, grid=None, p=0.1, handle=None, rng=None, _At=None, trace=None):
    """src/sequencer.jl:158-179: run the Sequencer on a random p-sample of the columns for each
    Fibonacci scale < M ÷ 10 and keep the scale with the largest elongation.  The reference samples
    with an unseeded StatsBase.sample; pass `rng` (numpy Generator or seed) for reproducibility."""
    At = _At if _At is not None else _prepare_input(A)
    N, M = At.shape
    nsamp = int(np.floor(N * p))
    assert nsamp > 1, f"Matrix A contains too few columns ({N}) to sample with p = {p}. Use a larger p."
    maxscale = M // 10
    assert maxscale > 0, (f"Matrix A contains too few rows ({M}) for autoscale. The minimum is 10 rows. "
                          "It is recommended to set scale=(1,).")
    gen = rng if isinstance(rng, np.random.Generator) else np.random.default_rng(rng)
    # StatsBase.sample(a, dims; ordered=true) draws WITH replacement by default, so the reference's sample can hold
    # a column more than once (duplicates sit at distance 0 + eps of each other)
    idx = np.sort(gen.choice(N, size=nsamp, replace=True))
    Asamp = np.ascontiguousarray(At[idx])
    metrics = _normalize_metrics(metrics)
    best, max_eta = 1, 0.0
    for s in [f for f in FIB if f < maxscale]:
        r = _sequence(Asamp, (s,), metrics, handle)
        if trace is not None:
            trace.append((s, r.eta))          # (scale, elongation) of every scale tried, for the parity test
        if r.eta > max_eta:
            max_eta, best = r.eta, s
    return best


# ---- unit-level wrappers (each layer testable in isolation) -----------------------------------
def splitnorm(A, scale, handle=None):
    """`_splitnorm` (src/sequencer.jl:383-407) + per-chunk ECDF. Returns (P, U, nchunks), both M x N."""
    handle = handle or default_handle()
    A = np.asarray(A, dtype=np.float64)
    M, N = A.shape
    if scale > M:
        raise AssertionError(f"Scale ({scale}) cannot be larger than the number of data elements ({M}).")
    At = np.ascontiguousarray(A.T)
    P = np.empty((N, M)); U = np.empty((N, M))
    nch = C.c_int32()
    handle.check(handle.lib.seqj_splitnorm(handle.h, _lib.dptr(At), M, N, int(scale), _lib.dptr(P), _lib.dptr(U),
                                           C.byref(nch)))
    return P.T.copy(), U.T.copy(), nch.value


def pairwise(metric, chunk, normalize=True, kl_full=True, eps_floor=EPS_FLOOR, l2_thresh=L2THR, handle=None):
    """`abs.(pairwise(alg, m; dims=2)) .+ ϵ` (src/sequencer.jl:226) for one m x N chunk."""
    handle = handle or default_handle()
    mid = metric.id if isinstance(metric, Metric) else int(metric)
    if isinstance(metric, PeriodicEuclidean):
        _set_periods((metric,), handle)
    chunk = np.asarray(chunk, dtype=np.float64)
    m, N = chunk.shape
    Ct = np.ascontiguousarray(chunk.T)
    Dm = np.empty((N, N))
    handle.check(handle.lib.seqj_pairwise(handle.h, mid, _lib.dptr(Ct), m, N, int(bool(normalize)), int(bool(kl_full)),
                                          eps_floor, l2_thresh, _lib.dptr(Dm)))
    return Dm.T.copy()      # column-major N x N -> numpy [i, j]


@dataclass
class MeasureResult:
    mst: WeightedTree
    root: int
    eta: float
    bfs: BFSTree
    order: np.ndarray = None


def measure_dm(Dm, want_order=False, eps_floor=EPS_FLOOR, handle=None) -> MeasureResult:
    """`_measure_dm(D)` (src/sequencer.jl:284-290) [+ `unroll`, src/utils.jl:15-24]."""
    handle = handle or default_handle()
    Dm = np.asarray(Dm, dtype=np.float64)
    N = Dm.shape[0]
    Dcm = np.ascontiguousarray(Dm.T)         # column-major storage of D
    eta = C.c_double(); root = C.c_int32()
    par = np.empty(N, dtype=np.int32)
    mi = np.empty(N - 1, dtype=np.int32); mj = np.empty(N - 1, dtype=np.int32); mw = np.empty(N - 1)
    od = np.empty(N, dtype=np.int32) if want_order else None
    handle.check(handle.lib.seqj_measure_dm(handle.h, _lib.dptr(Dcm), N, eps_floor, C.byref(eta), C.byref(root),
                                            _lib.iptr(par), _lib.iptr(mi), _lib.iptr(mj), _lib.dptr(mw),
                                            _lib.iptr(od)))
    return MeasureResult(WeightedTree(N, mi, mj, mw), root.value, eta.value, BFSTree(root.value, par), od)


def elongation(Dm, handle=None) -> float:
    """Elongation of the MST of a distance matrix (reference `elongation(g, startidx)`, sequencer.jl:348-354,
    composed with `_mst` and `_leastcentralpt` as `_measure_dm` does)."""
    return measure_dm(Dm, handle=handle).eta


def accumulate(Dchunks, etas, handle=None):
    """eta-weighted combination of chunk matrices (src/sequencer.jl:231-235,247)."""
    handle = handle or default_handle()
    Dchunks = np.asarray(Dchunks, dtype=np.float64)
    n, N, _ = Dchunks.shape
    Dt = np.ascontiguousarray(np.transpose(Dchunks, (0, 2, 1)))
    etas = np.ascontiguousarray(etas, dtype=np.float64)
    out = np.empty((N, N))
    handle.check(handle.lib.seqj_accumulate(handle.h, _lib.dptr(Dt), _lib.dptr(etas), n, N, _lib.dptr(out)))
    return out.T.copy()


def fuse(N, etas, msts, eps_floor=EPS_FLOOR, handle=None) -> SequencerResult:
    """Final fuse (src/sequencer.jl:266-273) from per-group (eta, MST) lists."""
    handle = handle or default_handle()
    G = len(etas)
    etas = np.ascontiguousarray(etas, dtype=np.float64)
    mi = np.ascontiguousarray(np.concatenate([t.src for t in msts]), dtype=np.int32)
    mj = np.ascontiguousarray(np.concatenate([t.dst for t in msts]), dtype=np.int32)
    mw = np.ascontiguousarray(np.concatenate([t.weight for t in msts]), dtype=np.float64)
    res = C.c_void_p()
    handle.check(handle.lib.seqj_fuse(handle.h, N, G, _lib.dptr(etas), _lib.iptr(mi), _lib.iptr(mj), _lib.dptr(mw),
                                      eps_floor, C.byref(res)))
    return _collect_and_release(handle, res, N, None, None, want_final=True)


def fuse_dev(N, ngroups, pack, slot_of_group, eps_floor=EPS_FLOOR, handle=None) -> SequencerResult:
    """Final fuse (src/sequencer.jl:266-273) from group results in the device exchange format (seqj_fuse_dev):
    `pack` = device tensor of slots, group g (metric-major) in slot slot_of_group[g]."""
    handle = handle or default_handle()
    slots = np.ascontiguousarray(slot_of_group, dtype=np.int32)
    assert slots.size == ngroups
    res = C.c_void_p()
    handle.check(handle.lib.seqj_fuse_dev(handle.h, N, ngroups, C.c_void_p(pack.data_ptr()), _lib.iptr(slots), eps_floor,
                                          C.byref(res)))
    return _collect_and_release(handle, res, N, None, None, want_final=True)


def cdf_distance(metric, u, v, uw, vw, handle=None) -> float:
    """`EMD(u, v, uw, vw)` / `Energy(u, v, uw, vw)` (src/distancemetrics.jl:82-87, 143-148): distance between two
    weighted ECDFs on their own grids `u`, `v` (weights `uw`, `vw`), through `seqj_cdf_distance`."""
    handle = handle or default_handle()
    u, v, uw, vw = (np.ascontiguousarray(x, dtype=np.float64) for x in (u, v, uw, vw))
    if u.ndim != 1 or v.ndim != 1:
        raise ValueError("Only 1-d data vectors are supported.")
    if len(u) != len(uw):
        raise ValueError(f"u and u weights must have same length. Got u {len(u)} and uw {len(uw)}")
    if len(v) != len(vw):
        raise ValueError(f"v and v weights must have same length. Got v {len(v)} and vw {len(vw)}")
    mid = {WASS1D.id: 1, ENERGY.id: 3}[metric.id]
    out = C.c_double()
    handle.check(handle.lib.seqj_cdf_distance(handle.h, mid, _lib.dptr(u), _lib.dptr(uw), len(u), _lib.dptr(v),
                                              _lib.dptr(vw), len(v), C.byref(out)))
    return float(out.value)


class EMD:
    """`EMD()` / `EMD((gridu, gridv))` functor (src/distancemetrics.jl:16-21, 90-93): `EMD(grids)(u, v)` is
    `EMD(gridu, gridv, u, v)` with u, v the weights on those grids; without grids, the shared unit grid."""

    def __init__(self, grids=None):
        self.grids = None if grids is None else tuple(np.asarray(g, dtype=np.float64) for g in grids)

    def __call__(self, u, v=None, uw=None, vw=None):
        if uw is not None:
            return emd(u, v, uw, vw)
        v = u if v is None else v
        return emd(u, v) if self.grids is None else emd(self.grids[0], self.grids[1], u, v)


class Energy:
    """`Energy()` / `Energy((gridu, gridv))` functor (src/distancemetrics.jl:121-132)."""

    def __init__(self, grids=None):
        self.grids = None if grids is None else tuple(np.asarray(g, dtype=np.float64) for g in grids)

    def __call__(self, u, v=None, uw=None, vw=None):
        if uw is not None:
            return energy(u, v, uw, vw)
        v = u if v is None else v
        return energy(u, v) if self.grids is None else energy(self.grids[0], self.grids[1], u, v)


def _pair(metric, u, v, handle=None):
    u = np.asarray(u, dtype=np.float64)
    v = u if v is None else np.asarray(v, dtype=np.float64)
    if u.ndim != 1 or v.ndim != 1:
        raise ValueError("Only 1-d data vectors are supported.")
    if len(u) != len(v):          # default grids axes(u,1) and axes(v,1) differ: the general branch (:210-214)
        return cdf_distance(metric, np.arange(1, len(u) + 1.0), np.arange(1, len(v) + 1.0), u, v, handle)
    Dm = pairwise(metric, np.stack([u, v], axis=1), normalize=False, eps_floor=0.0, handle=handle)
    return float(Dm[0, 1])


def emd(u, v=None, uw=None, vw=None, handle=None):
    """`emd(u, v)` (src/distancemetrics.jl:101): EMD of the weights u, v on the default unit grids;
    `emd(u, v, uw, vw)` (:108): explicit grids u, v with weights uw, vw."""
    if uw is not None or vw is not None:
        return cdf_distance(WASS1D, u, v, uw, vw, handle)
    return _pair(WASS1D, u, v, handle)


def energy(u, v=None, uw=None, vw=None, handle=None):
    """`energy(u, v)` (src/distancemetrics.jl:172) and `energy(u, v, uw, vw)` (:167)."""
    if uw is not None or vw is not None:
        return cdf_distance(ENERGY, u, v, uw, vw, handle)
    return _pair(ENERGY, u, v, handle)
