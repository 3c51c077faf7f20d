"""Run-to-run determinism of config 3 on random data, with enough probes to localise a flake.

K unsharded calls in ONE process.  Every call is compared with the first one: chunk eta / root / parents of all 124
chunk graphs, group eta / parents, and sampled entries of the chunk matrices (150k per finest-scale chunk matrix, so a
single glitched 64 x 128 tile is hit ~25 times; 20k per coarser chunk matrix).  Prints what differs and where."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import sequencerj_b200 as s  # noqa: E402
from sequencerj_b200 import api  # noqa: E402

K = int(sys.argv[1]) if len(sys.argv) > 1 else 30
N, L = 10000, 4096
scales, metrics = (1, 2, 4, 8, 16), s.ALL_METRICS
At = np.maximum(np.random.default_rng(2732).random((N, L)), 2.220446049250313e-16)
h = s.default_handle()
rng = np.random.default_rng(5)
pg, pc, pi, pj = [], [], [], []
for k in range(len(metrics)):
    for l, sc in enumerate(scales):
        g = k * len(scales) + l
        per = 150000 if sc == 16 else 20000
        for c in range(sc):
            pg.append(np.full(per, g, np.int32)); pc.append(np.full(per, c, np.int32))
            pi.append(rng.integers(0, N, per).astype(np.int32)); pj.append(rng.integers(0, N, per).astype(np.int32))
pg, pc, pi, pj = (np.concatenate(x) for x in (pg, pc, pi, pj))
print("probes:", len(pg), flush=True)


def snap(r):
    out = {}
    for key, (etas, trees) in r.EOSeg.items():
        out[key] = (np.array(etas), np.array([t.root for t in trees]), np.stack([t.parents for t in trees]),
                    r.EOAlgScale[key][0], r.EOAlgScale[key][1].parents.copy())
    return out


ref = api._sequence(At, scales, metrics, h, probes=(pg, pc, pi, pj))
sref, pref = snap(ref), ref.probes.copy()
print("reference call done; NaN probes:", int(np.isnan(pref).sum()), flush=True)
nflaky = 0
for rep in range(1, K):
    r = api._sequence(At, scales, metrics, h, probes=(pg, pc, pi, pj))
    sn = snap(r)
    bad = False
    for key in sref:
        a, b = sref[key], sn[key]
        g = metrics.index(key[0]) * len(scales) + scales.index(key[1])
        for c in range(len(a[0])):
            de, dr, dp = a[0][c] != b[0][c], a[1][c] != b[1][c], int(np.sum(a[2][c] != b[2][c]))
            sel = (pg == g) & (pc == c)
            dpr = np.nonzero(sel & (pref != r.probes))[0]
            if de or dr or dp or len(dpr):
                bad = True
                print(f"call {rep}: {key[0]} scale {key[1]} chunk {c}: eta {a[0][c]!r} -> {b[0][c]!r}  root {a[1][c]} -> {b[1][c]}"
                      f"  parents differing {dp}  probes differing {len(dpr)} of {int(sel.sum())}", flush=True)
                for q in dpr[:12]:
                    print(f"      D[{pi[q]},{pj[q]}]  {pref[q]!r} -> {r.probes[q]!r}", flush=True)
        if a[3] != b[3] or np.any(a[4] != b[4]):
            bad = True
            print(f"call {rep}: {key[0]} scale {key[1]} GROUP: eta {a[3]!r} -> {b[3]!r}  parents differing {int(np.sum(a[4] != b[4]))}",
                  flush=True)
    if r.eta != ref.eta or not np.array_equal(r.order, ref.order):
        bad = True
        print(f"call {rep}: FINAL eta {ref.eta!r} -> {r.eta!r}  order equal {np.array_equal(r.order, ref.order)}", flush=True)
    nflaky += bad
    if not bad:
        print(f"call {rep}: identical", flush=True)
print("env", {k: v for k, v in os.environ.items() if k.startswith("SEQJ_")}, "flaky calls:", nflaky, "of", K - 1, flush=True)
