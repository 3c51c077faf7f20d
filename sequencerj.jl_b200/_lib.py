"""ctypes binding of include/seqj.h (the same symbols a Julia `ccall` binds)."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libseqj_b200.so")

SEQJ_NTIMERS = 13
TIMER_NAMES = ["total", "h2d", "prep", "dist_l2", "dist_emd", "dist_kld", "dist_energy", "prim", "tree",
               "combine", "fuse", "d2h", "derive"]

c_dp = C.POINTER(C.c_double)
c_ip = C.POINTER(C.c_int32)
c_u8p = C.POINTER(C.c_uint8)
vp = C.c_void_p

# name -> (restype, argtypes); every symbol declared in include/seqj.h
SIGNATURES = {
    "seqj_version": (C.c_int, []),
    "seqj_create": (C.c_int, [C.POINTER(vp), c_ip, C.c_int]),
    "seqj_destroy": (None, [vp]),
    "seqj_last_error": (C.c_char_p, [vp]),
    "seqj_plan_describe": (C.c_int, [C.c_int64, c_ip, C.c_int, c_ip, C.c_int, c_u8p, C.c_int64, C.c_int, C.c_char_p, C.c_int64]),
    "seqj_plan_shards": (C.c_int, [C.c_int64, c_ip, C.c_int, c_ip, C.c_int, C.c_int, c_ip, c_dp]),
    "seqj_set_workspace_limit": (C.c_int, [vp, C.c_uint64]),
    "seqj_release_workspace": (C.c_int, [vp]),
    "seqj_set_grid": (C.c_int, [vp, c_dp, C.c_int64]),
    "seqj_set_periods": (C.c_int, [vp, c_dp, C.c_int64]),
    "seqj_set_probes": (C.c_int, [vp, c_ip, c_ip, c_ip, c_ip, C.c_int64]),
    "seqj_result_probes": (C.c_int, [vp, c_dp, C.c_int64]),
    "seqj_sequence": (C.c_int, [vp, vp, C.c_int64, C.c_int64, c_ip, C.c_int, c_ip, C.c_int, C.c_double,
                                C.c_double, c_u8p, C.POINTER(vp)]),
    "seqj_sequence_f32": (C.c_int, [vp, vp, C.c_int64, C.c_int64, c_ip, C.c_int, c_ip, C.c_int, C.c_double,
                                    C.c_double, c_u8p, C.POINTER(vp)]),
    "seqj_sequence_dev": (C.c_int, [vp, vp, C.c_int64, C.c_int64, c_ip, C.c_int, c_ip, C.c_int, C.c_double,
                                    C.c_double, c_u8p, C.POINTER(vp)]),
    "seqj_sequence_partial": (C.c_int, [vp, vp, C.c_int64, C.c_int64, c_ip, C.c_int, c_ip, C.c_int, C.c_double,
                                        C.c_double, c_ip, c_ip, vp, C.c_int64, C.POINTER(vp)]),
    "seqj_groups_finish": (C.c_int, [vp, vp, C.c_int64, C.c_int, c_ip, C.c_int64, c_dp, C.c_double, c_dp, c_ip, c_ip,
                                     c_ip, c_ip, c_dp]),
    "seqj_pack_stride": (C.c_int64, [C.c_int64]),
    "seqj_sequence_shard": (C.c_int, [vp, vp, C.c_int64, C.c_int64, c_ip, C.c_int, c_ip, C.c_int, C.c_double,
                                      C.c_double, c_u8p, vp, C.POINTER(vp)]),
    "seqj_fuse_dev": (C.c_int, [vp, C.c_int64, C.c_int, vp, c_ip, C.c_double, C.POINTER(vp)]),
    "seqj_fuse": (C.c_int, [vp, C.c_int64, C.c_int, c_dp, c_ip, c_ip, c_dp, C.c_double, C.POINTER(vp)]),
    "seqj_result_dims": (C.c_int, [vp, C.POINTER(C.c_int64), C.POINTER(C.c_int)]),
    "seqj_result_group_info": (C.c_int, [vp, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int),
                                         C.POINTER(C.c_int)]),
    "seqj_result_group_elapsed": (C.c_int, [vp, C.c_int, c_dp]),
    "seqj_result_group_range": (C.c_int, [vp, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "seqj_result_group_chunks": (C.c_int, [vp, C.c_int, c_dp, c_ip, c_ip]),
    "seqj_result_group": (C.c_int, [vp, C.c_int, c_dp, c_ip, c_ip, c_ip, c_ip, c_dp]),
    "seqj_result_final": (C.c_int, [vp, c_dp, c_ip, c_ip, c_ip, c_ip, c_ip, c_dp]),
    "seqj_result_D_nnz": (C.c_int, [vp, C.POINTER(C.c_int64)]),
    "seqj_result_D_triplets": (C.c_int, [vp, c_ip, c_ip, c_dp]),
    "seqj_result_timings": (C.c_int, [vp, c_dp]),
    "seqj_result_derived_groups": (C.c_int, [vp, C.POINTER(C.c_int)]),
    "seqj_result_waves": (C.c_int, [vp, C.POINTER(C.c_int)]),
    "seqj_result_launches": (C.c_int, [vp, C.POINTER(C.c_int64)]),
    "seqj_result_input_range": (C.c_int, [vp, c_dp, c_dp]),
    "seqj_result_free": (None, [vp]),
    "seqj_splitnorm": (C.c_int, [vp, c_dp, C.c_int64, C.c_int64, C.c_int32, c_dp, c_dp, c_ip]),
    "seqj_pairwise": (C.c_int, [vp, C.c_int32, c_dp, C.c_int64, C.c_int64, C.c_int, C.c_int, C.c_double,
                                C.c_double, c_dp]),
    "seqj_measure_dm": (C.c_int, [vp, c_dp, C.c_int64, C.c_double, c_dp, c_ip, c_ip, c_ip, c_ip, c_dp, c_ip]),
    "seqj_accumulate": (C.c_int, [vp, c_dp, c_dp, C.c_int, C.c_int64, c_dp]),
    "seqj_cdf_distance": (C.c_int, [vp, C.c_int32, c_dp, c_dp, C.c_int64, c_dp, c_dp, C.c_int64, c_dp]),
}

_lib = None


class SeqjError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"seqj error {code}: {msg}")
        self.code = code


def load():
    """Load libseqj_b200.so. Fails loudly when it has not been built: there is no CPU fallback."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                           "(nvcc, sm_100a). The product path has no CPU fallback.")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def dptr(a):
    return a.ctypes.data_as(c_dp) if a is not None else None


def iptr(a):
    return a.ctypes.data_as(c_ip) if a is not None else None


class Handle:
    """Owns a seqj_handle. Creation fails (SeqjError) when no sm_100 device is usable."""

    def __init__(self, device=None):
        """device: None (current device), an int, or a list of ints (one process drives several GPUs)."""
        self.lib = load()
        self.h = vp()
        if device is None:
            st = self.lib.seqj_create(C.byref(self.h), None, 0)
        else:
            devs = [device] if isinstance(device, int) else list(device)
            dev = (C.c_int32 * len(devs))(*devs)
            st = self.lib.seqj_create(C.byref(self.h), dev, len(devs))
        if st != 0:
            msg = self.lib.seqj_last_error(self.h).decode() if self.h else "seqj_create failed"
            self.close()
            raise SeqjError(st, msg)

    def check(self, st):
        if st == -6:      # SEQJ_E_DOMAIN: the reference raises AssertionError (src/sequencer.jl:111)
            raise AssertionError(self.lib.seqj_last_error(self.h).decode())
        if st != 0:
            raise SeqjError(st, self.lib.seqj_last_error(self.h).decode())

    def close(self):
        if getattr(self, "h", None):
            self.lib.seqj_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def as_f64_contig(a):
    return np.ascontiguousarray(a, dtype=np.float64)
