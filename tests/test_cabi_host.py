"""CPU-side checks: the C-ABI library loads and exports every symbol include/seqj.h declares, fails loudly
without a GPU (no CPU fallback), and the host-side mirror of the Julia API validates input as the reference."""
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    txt = open(os.path.join(ROOT, "include", "seqj.h")).read()
    return sorted(set(re.findall(r"\b(seqj_[a-zA-Z0-9_]+)\s*\(", txt)))


def test_library_exports_every_declared_symbol(seqj):
    lib = seqj.load()
    names = _declared_symbols()
    assert len(names) >= 20
    from sequencerj_b200 import _lib
    assert sorted(_lib.SIGNATURES) == names          # the ctypes binding covers the whole header
    for n in names:
        assert getattr(lib, n) is not None
    assert lib.seqj_version() == 200


def test_no_cpu_fallback(seqj):
    """Without a usable sm_100 device, creating a handle raises: the product never computes on the CPU."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(seqj.SeqjError):
        seqj.Handle()
    with pytest.raises(seqj.SeqjError):
        seqj.sequence(np.random.default_rng(0).random((20, 10)), scales=(1,))


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "sequencerj.jl_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".jl")):
                src = open(os.path.join(dirpath, f)).read()
                assert "seq_oracle" not in src and "import oracle" not in src and "from oracle" not in src, f


def test_prepare_input_matches_reference_semantics(seqj):
    from sequencerj_b200 import api
    A = np.array([[0.0, 1.0, 2.0], [3.0, 0.0, 5.0]])
    At = api._prepare_input(A)
    assert At.shape == (3, 2) and At.flags.c_contiguous
    assert A[0, 0] == 2.220446049250313e-16           # clamp! mutates the caller's array (sequencer.jl:130)
    assert At[0, 0] == 2.220446049250313e-16 and At[2, 1] == 5.0
    At32 = api._prepare_input(np.ones((4, 3), dtype=np.float32))
    assert At32.dtype == np.float64
    cols = api._prepare_input([[1.0, 2.0], [3.0, 4.0], [5.0, 6.0]])      # vector of vectors -> hcat
    assert cols.shape == (3, 2) and cols[1].tolist() == [3.0, 4.0]
    for bad in (-np.ones((2, 2)), np.array([[np.nan, 1.0]]), np.array([[np.inf, 1.0]])):
        with pytest.raises(AssertionError):
            api._prepare_input(bad)


def test_prettyp_and_show(seqj):
    assert seqj.prettyp(list(range(1, 11))) == "1,2,3...8,9,10"                 # utils.jl docstring
    assert seqj.prettyp(list(range(1, 11)), 5) == "1,2,3,4,5...6,7,8,9,10"
    assert seqj.prettyp([1, 2, 3]) == "1,2,3"


def test_constants_and_grid(seqj):
    assert [m.id for m in seqj.ALL_METRICS] == [0, 1, 2, 3]
    assert seqj.ALL_METRICS == (seqj.L2, seqj.WASS1D, seqj.KLD, seqj.ENERGY)
    assert repr(seqj.ALL_METRICS) == "(Euclidean(1.0e-7), EMD(nothing), KLDivergence(), Energy(nothing))"
    g = seqj.ensuregrid(np.ones((5, 3)))
    assert g.tolist() == [1.0, 2.0, 3.0, 4.0, 5.0]
    with pytest.raises(AssertionError):
        seqj.ensuregrid(np.ones((5, 3)), [1.0, 2.0])
    assert seqj.FIB[:6] == [1, 2, 3, 5, 8, 13] and seqj.FIB[-1] == 5168       # sequencer.jl:148 (sic)
