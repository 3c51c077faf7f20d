"""sequencerj.jl_b200 -- B200-native (sm_100a) implementation of SequencerJ's data-parallel core.

The package directory name contains a dot, so it is imported through the top-level shim
``sequencerj_b200`` (repo root).  The compute lives in ``libseqj_b200.so`` (csrc/, C ABI in
include/seqj.h); this package is the host-side mirror of the reference's Julia API.
"""
from .api import (ALL_METRICS, CHEBYSHEV, CITYBLOCK, SQEUCLIDEAN, TOTALVARIATION, EMD, EMD_GRID, ENERGY, ENERGY_GRID, FIB, KLD, L2, WASS1D, BFSTree, Energy, MeasureResult, Metric, PeriodicEuclidean,
                  SequencerResult, WeightedTree, D, accumulate, autoscale, cdf_distance, default_handle, elong, elongation, emd,
                  energy, ensuregrid, fuse, measure_dm, mst, order, pairwise, prettyp, sequence, splitnorm)
from ._lib import Handle, SeqjError, load
from . import build as _build
from . import sharding

__all__ = ["ALL_METRICS", "CHEBYSHEV", "CITYBLOCK", "SQEUCLIDEAN", "TOTALVARIATION", "EMD", "EMD_GRID", "ENERGY", "ENERGY_GRID", "FIB", "KLD", "L2", "WASS1D", "BFSTree", "Energy", "MeasureResult", "PeriodicEuclidean",
           "Metric", "SequencerResult", "WeightedTree", "D", "accumulate", "autoscale", "cdf_distance", "default_handle", "elong",
           "elongation", "emd", "energy", "ensuregrid", "fuse", "measure_dm", "mst", "order", "pairwise", "prettyp",
           "sequence", "splitnorm", "Handle", "SeqjError", "load"]
