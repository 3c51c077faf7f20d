"""The static schedule of a sequence() call (csrc/host_plan.cuh) checked on the CPU through seqj_plan_describe: every
chunk of the reference's loop (src/sequencer.jl:194-258) is produced exactly once, the pool never holds two live
matrices in one slot, every derived chunk finds its fine chunks resident, and the memory bound is respected."""
import ctypes as C

import numpy as np
import pytest


def _plan(M, metric_ids, scales, mats, hier=1, mask=None):
    from sequencerj_b200 import _lib
    lib = _lib.load()
    mids = np.array(metric_ids, dtype=np.int32)
    scs = np.array(scales, dtype=np.int32)
    buf = C.create_string_buffer(1 << 22)
    m = None if mask is None else np.ascontiguousarray(mask, dtype=np.uint8)
    st = lib.seqj_plan_describe(M, _lib.iptr(mids), len(mids), _lib.iptr(scs), len(scs),
                                None if m is None else m.ctypes.data_as(_lib.c_u8p), mats, hier, buf, len(buf))
    assert st == 0
    head, waves = None, []
    for ln in buf.value.decode().splitlines():
        f = ln.split()
        if f[0] == "error":
            return {"error": ln}, []
        if f[0] == "wave":                 # wave <i> tasks <n> out0 <o> done <a> <b>
            waves.append(({"index": int(f[1]), "tasks": int(f[3]), "out0": int(f[5]), "done": (int(f[7]), int(f[8]))}, []))
            continue
        d = {f[i]: int(f[i + 1]) for i in range(1, len(f), 2)}
        if f[0] == "plan":
            head = d
        else:
            waves[-1][1].append(d)
    return head, waves


def _nchunks(M, sc):
    cl = -(-M // sc)
    return -(-M // cl), cl


def _check(M, metric_ids, scales, mats, hier=1, mask=None):
    head, waves = _plan(M, metric_ids, scales, mats, hier, mask)
    assert "error" not in head, head
    ns = len(scales)
    assert head["pool_slots"] <= mats
    produced = {}                       # (g, c) -> count
    content = {}                        # slot -> (g, c) of the live matrix
    acc_owner = {}                      # accumulator slot -> group
    finished = set()
    derived_groups = set()
    for wd, segs in waves:
        live_this_wave = set()
        assert wd["tasks"] == sum(s["c1"] - s["c0"] for s in segs)
        for s in segs:
            g, l = s["g"], s["l"]
            assert g % ns == l
            nch, cl = _nchunks(M, scales[l])
            assert 0 <= s["c0"] < s["c1"] <= nch
            if s["src_slot0"] >= 0:                              # derived: the fine chunks must be resident
                derived_groups.add(g)
                assert metric_ids[g // ns] in (0, 2, 3)
                nf, clf = _nchunks(M, scales[s["src_l"]])
                assert cl % clf == 0 and nf > nch
                ratio = cl // clf
                gsrc = (g // ns) * ns + s["src_l"]
                for c in range(s["c0"], s["c1"]):
                    for f in range(c * ratio, min((c + 1) * ratio, nf)):
                        assert content.get(s["src_slot0"] + f - s["fbase"]) == (gsrc, f), (s, c, f)
            for c in range(s["c0"], s["c1"]):
                slot = s["slot0"] + c - s["c0"]
                assert 0 <= slot < head["pool_slots"]
                assert slot not in live_this_wave, "two matrices of one wave share a slot"
                assert slot not in acc_owner, "a chunk matrix overwrites a live accumulator"
                live_this_wave.add(slot)
                content[slot] = (g, c)
                produced[(g, c)] = produced.get((g, c), 0) + 1
            if s["first"]:
                assert s["sslot"] not in acc_owner and s["sslot"] not in live_this_wave
                assert 0 <= s["sslot"] < head["pool_slots"]
                acc_owner[s["sslot"]] = g
            assert acc_owner.get(s["sslot"]) == g
        # derivation sources of this wave may not be clobbered by accumulators claimed in it
        for s in segs:
            if s["last"]:
                finished.add(s["g"])
        for slot in [a for a, g in acc_owner.items() if g in finished]:
            del acc_owner[slot]
        assert wd["done"][1] - wd["done"][0] == len({s["g"] for s in segs if s["last"]})
    want = {}
    for k, mid in enumerate(metric_ids):
        for l, sc in enumerate(scales):
            if mask is not None and not np.asarray(mask).reshape(-1)[k * ns + l]:
                continue
            for c in range(_nchunks(M, sc)[0]):
                want[(k * ns + l, c)] = 1
    assert produced == want
    assert head["total_chunks"] == len(want) and head["derived_groups"] == len(derived_groups)
    return head, waves, derived_groups


def test_config3_single_wave_derives_12_groups():
    head, waves, der = _check(4096, (0, 1, 2, 3), (1, 2, 4, 8, 16), 212)
    assert head["waves"] == 1 and len(der) == 12
    head, waves, der = _check(4096, (0, 1, 2, 3), (1, 2, 4, 8, 16), 212, hier=0)
    assert head["waves"] == 1 and len(der) == 0


def test_config4_streams_with_every_coarse_energy_scale_derived():
    """N = 30,000 on one 180 GB GPU: 24 matrices of 7.2 GB fit.  Energy: blocks at scale 8, carry, top wave; all five
    coarse Energy scales derived although only 7 + 8 matrices of the metric are ever live."""
    head, waves, der = _check(2048, (1, 3), (1, 2, 4, 8, 16, 32), 24)
    assert len(der) == 5 and head["waves"] >= 9
    assert all(g >= 6 for g in der)                      # the Energy groups (metric index 1); EMD always runs tiles
    # the 8-GPU owner of the Energy bundle (group mask) gets the same chain schedule
    mask = np.zeros(12, dtype=np.uint8); mask[6:] = 1
    head, waves, der = _check(2048, (1, 3), (1, 2, 4, 8, 16, 32), 24, mask=mask)
    assert len(der) == 5 and head["waves"] == 9


@pytest.mark.parametrize("mats", [8, 12, 16, 24, 40, 64, 150])
def test_streaming_plans_are_consistent_for_any_memory(mats):
    _check(2048, (1, 3), (1, 2, 4, 8, 16, 32), mats)
    _check(4096, (0, 1, 2, 3), (1, 2, 4, 8, 16), mats)
    _check(512, (0, 2), (1, 2, 4), mats)


def test_ragged_and_non_nested_scales():
    _check(53, (0, 1, 2, 3), (3, 6, 7), 200)             # nothing nests: tiles everywhere
    head, waves, der = _check(10, (0, 3), (1, 2, 6), 100)   # chunk lengths 10 / 5 / 2: 5 does not nest in 2... 10 nests in 5
    _check(10, (0, 3), (1, 2, 6), 9)
    _check(50, (0, 1, 2, 3), (1, 2, 4), 7)               # chunk lengths 50 / 25 / 13: only 50 <- 25 nests
    _check(1000, (1, 3), (1, 2, 4, 8), 10)
    _check(4096, (0, 1, 2, 3), (1, 2, 4, 8, 16), 7)      # too tight for chains: everything flat, many waves


def test_too_little_memory_is_an_error():
    head, waves = _plan(4096, (0, 1), (1, 2), 1)
    assert "error" in head
