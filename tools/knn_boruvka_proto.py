"""NumPy prototype of the kNN-filtered Boruvka used by the few-graph MST path (kernels_mst.cuh): validates the
list / bound / stuck logic against the oracle's Kruskal (exact edge sets, also under weight ties) and reports how
far the list-only rounds get on different data.  Test infrastructure (imports the oracle); not on the product path."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import seq_oracle as O  # noqa: E402

EPS = 1e-6
LANES, M_LOCAL, K = 32, 4, 16
NONE = np.uint64(0xFFFFFFFFFFFFFFFF)


def bits(x):
    return np.ascontiguousarray(x, dtype=np.float64).view(np.uint64)


def row_list(W, v, exclude, comp=None):
    """List + bound of row v over the vertices j with exclude[j] == False, as the device builds it: 32 lanes own
    interleaved 16-byte pairs and keep their M_LOCAL smallest (w, j); L = the smallest last entry among the lanes
    that left something unlisted; the list = the K smallest entries <= L.
    With `comp` (rescans, round 2): a lane keeps at most one entry per component -- its lightest edge into it -- and
    the merged list holds one entry per component: an edge into a component that is heavier than another edge of the
    row into the same component can never be the row's lightest outgoing edge (components only merge)."""
    N = W.shape[0]
    lane_of = (np.arange(N) // 2) % LANES
    cand, L = [], None
    for l in range(LANES):
        js = np.nonzero((lane_of == l) & ~exclude)[0]
        if len(js) == 0:
            continue
        o = np.lexsort((js, W[v, js]))
        if comp is not None:                       # first (lightest) entry of every component, in (w, j) order
            seen, keep = set(), []
            for t in o:
                c = int(comp[js[t]])
                if c not in seen:
                    seen.add(c); keep.append(t)
            o = np.array(keep, dtype=np.int64)
        o = o[:M_LOCAL]
        ent = [(int(W[v, js[t]]), int(js[t])) for t in o]
        cand += ent
        if len(ent) == M_LOCAL:       # full lane list: whatever the lane did not list is > its last entry (the
                                      # device does not know whether anything was left over, so a full list counts)
            b = ent[-1]
            L = b if L is None or b < L else L
    cand.sort()
    if L is not None:
        cand = [e for e in cand if e <= L]
    if comp is not None:                           # the warp merge skips what a lighter listed edge already covers
        seen, keep = set(), []
        for e in cand:
            c = int(comp[e[1]])
            if c not in seen:
                seen.add(c); keep.append(e)
        cand = keep
    if len(cand) > K:
        cand = cand[:K]
        B = cand[-1][0]
    elif len(cand) == K:
        B = cand[-1][0]
    else:
        B = L[0] if L is not None else int(NONE)
    return cand, B


def row_list_streaming(W, v, exclude, refresh=8):
    """Candidate for round 2 (lead 2): the same list, built the way the kernel streams the row
    (lane l sees pair t*32 + l at step t), with a warp-wide threshold tau = the K-th smallest weight currently held
    by any lane, refreshed every `refresh` steps.  An element is skipped when its weight is STRICTLY above tau
    (ties pass, so no skipped edge can tie with a listed one), otherwise it goes through the lane-local insertion.
    Whatever a lane did not keep is then > min(its last entry if its list is full, (tau, +inf)), so
    L = min(smallest last entry of the full lanes, (tau, +inf)).  Returns (list, bound, steps_with_an_insertion)."""
    N = W.shape[0]
    npairs = (N + 1) // 2
    steps = (npairs + LANES - 1) // LANES
    lane = [[] for _ in range(LANES)]
    tau = int(NONE)
    slow = 0
    for t in range(steps):
        hit = False
        for l in range(LANES):
            for e in (2 * (t * LANES + l), 2 * (t * LANES + l) + 1):
                if e >= N or exclude[e]:
                    continue
                x = (int(W[v, e]), e)
                if x[0] > tau:
                    continue
                if len(lane[l]) < M_LOCAL or x < lane[l][-1]:
                    hit = True
                    lane[l].append(x)
                    lane[l].sort()
                    del lane[l][M_LOCAL:]
        slow += hit
        if t % refresh == refresh - 1:
            held = sorted(w for ls in lane for w, _ in ls)
            if len(held) >= K:
                tau = min(tau, held[K - 1])
    L = None
    for ls in lane:
        if len(ls) == M_LOCAL and (L is None or ls[-1] < L):
            L = ls[-1]
    if tau != int(NONE):
        tb = (tau, 1 << 62)
        L = tb if L is None or tb < L else L
    cand = sorted(x for ls in lane for x in ls)
    if L is not None:
        cand = [x for x in cand if x <= L]
    if len(cand) > K:
        cand = cand[:K]
        B = cand[-1][0]
    else:
        B = L[0] if L is not None else int(NONE)
    return cand, B, slow


def knn_lists(D):
    """Per row v: entries sorted by (w, j), at most K of them, plus a bound weight B such that every edge of v that
    is not listed has weight >= B (exactly: (w, j) > (B, jB)); B = NONE when the list holds every edge."""
    N = D.shape[0]
    W = bits(np.maximum(D, EPS))
    lists, bound = [], np.empty(N, dtype=np.uint64)
    for v in range(N):
        ex = np.zeros(N, dtype=bool); ex[v] = True
        cand, B = row_list(W, v, ex)
        lists.append(cand)
        bound[v] = B
    return lists, bound


def threshold_lists(D, sample_div=8, sample_k=3, cap=96, block=128):
    """Round 2: the first list pass from the UPPER TRIANGLE (kernels_mst.cuh: knn_sample / knn_tiles / knn_merge
    kernels).  tau_v = the sample_k-th smallest of a contiguous sample of N / sample_div entries of row v; every entry
    (i, j), i < j, is read once and offered to both ends: it becomes a candidate of row i when it is <= tau_i and of
    row j when it is <= tau_j.  The list = the K smallest candidates in (weight, j) order, bound = the K-th weight or
    tau_v (every entry <= tau_v is a candidate).  A row without a threshold or with more than `cap` candidates gets an
    empty list with bound 0, i.e. it is rescanned in the first round.  Returns (lists, bound) like knn_lists."""
    N = D.shape[0]
    W = bits(np.maximum(D, EPS))
    m = max(64, (N // sample_div + 63) // 64 * 64)
    tau = np.empty(N, dtype=np.uint64)
    for v in range(N):
        s0 = (v // block) * block
        if s0 + m > N:
            s0 = max(0, ((N - m) + 1) & ~1)
        js = [j for j in range(s0, min(N, s0 + m)) if j != v]
        ws = sorted(int(W[v, j]) for j in js)
        tau[v] = ws[sample_k - 1] if len(ws) >= sample_k else NONE
        if len(ws) >= sample_k and ws[0] == ws[sample_k - 1]:
            tau[v] = 0                                      # the smallest sample entries tie: straight to the rescan
    cands = [[] for _ in range(N)]
    for i in range(N):                                     # one visit per unordered pair
        for j in range(i + 1, N):
            w = int(W[i, j])
            if w <= int(tau[i]):
                cands[i].append((w, j))
            if w <= int(tau[j]):
                cands[j].append((w, i))
    lists, bound = [], np.empty(N, dtype=np.uint64)
    for v in range(N):
        if tau[v] == NONE or tau[v] == 0 or len(cands[v]) > cap:
            lists.append([]); bound[v] = 0
            continue
        c = sorted(cands[v])
        if len(c) >= K:
            c = c[:K]
            bound[v] = c[-1][0]
        else:
            bound[v] = tau[v]
        lists.append(c)
    return lists, bound


def boruvka(D, lists, bound, verbose=False, rescan=True, dedupe=False, lazy=False):
    """dedupe: rescanned lists hold one entry per component; lazy: rescan only when no component can hook any more
    (the round-2 kernels do both; the round-1 behaviour is dedupe=False, lazy=False)."""
    N = D.shape[0]
    W = bits(np.maximum(D, EPS))
    nrescan = 0
    comp = np.arange(N)
    ptr = np.zeros(N, dtype=np.int64)
    edges = []
    rounds = 0
    while True:
        rounds += 1
        best = {}
        stuck = set()
        needy = []
        cand = [None] * N
        for v in range(N):
            c = comp[v]
            lv = lists[v]
            while ptr[v] < len(lv) and comp[lv[ptr[v]][1]] == c:
                ptr[v] += 1
            if ptr[v] < len(lv):
                w, j = lv[ptr[v]]
                key = (w, max(v, j), min(v, j))
                cand[v] = key
                if c not in best or key < best[c]:
                    best[c] = key
        for v in range(N):
            c = comp[v]
            if cand[v] is None and bound[v] != NONE:
                bw = best[c][0] if c in best else int(NONE)
                if int(bound[v]) <= bw:
                    stuck.add(c)
                    needy.append(v)
        hooks = {}
        for c, key in best.items():
            if c in stuck:
                continue
            w, hi, lo = key
            other = comp[hi] if comp[lo] == c else comp[lo]
            hooks[c] = (other, key)
        nh = 0
        parent = {c: c for c in set(comp.tolist())}
        for c, (other, key) in hooks.items():
            mutual = other in hooks and hooks[other][1] == key
            if mutual and c < other:
                continue
            parent[c] = other
            edges.append((key[2], key[1], key[0]))
            nh += 1
        if verbose:
            print(f"  round {rounds}: comps {len(parent)} stuck {len(stuck)} hooks {nh} needy {len(needy)}")
        if nh == 0 and not (rescan and needy):
            break

        def find(c):
            while parent[c] != c:
                c = parent[c]
            return c
        comp = np.array([find(c) for c in comp.tolist()])
        if len(set(comp.tolist())) == 1:
            break
        if rescan and not (lazy and nh > 0):
            for v in needy:                       # the rescan kernel runs after the round's relabelling
                lists[v], bound[v] = row_list(W, v, comp == comp[v], comp if dedupe else None)
                ptr[v] = 0
            nrescan += len(needy)
        if rounds > 80:
            break
    if verbose:
        print(f"  rescans: {nrescan} rows = {nrescan / N:.2f} passes")
    boruvka.last_rescans = nrescan
    return edges, len(set(comp.tolist())), rounds


def check(name, D):
    D = np.maximum(D, D.T) if not np.array_equal(D, D.T) else D
    lists, bound = knn_lists(D)
    edges, ncomp, rounds = boruvka(D, lists, bound, verbose=True)
    ei, ej, ew = O.kruskal_mst(D)
    ref = set(zip(ei.tolist(), ej.tolist()))
    got = set((a, b) for a, b, _ in edges)
    ok = got <= ref
    full = ncomp == 1 and got == ref
    avg = np.mean([len(x) for x in lists])
    print(f"{name}: N={D.shape[0]} rounds={rounds} comps_left={ncomp} edges_subset_of_mst={ok} complete={full} "
          f"avg_list={avg:.1f}")
    assert ok
    if ncomp == 1:
        assert got == ref
        wref = dict(zip(zip(ei.tolist(), ej.tolist()), bits(ew).tolist()))
        assert all(wref[(a, b)] == w for a, b, w in edges)


if __name__ == "__main__":
    rng = np.random.default_rng(5)
    N = int(sys.argv[1]) if len(sys.argv) > 1 else 400
    A = O.clamp_input(rng.random((64, N)))
    P = O.splitnorm(A, 1)[0]
    for m in (0, 1, 2, 3):
        check(f"random metric {m}", O.chunk_distance(m, P))
    x = np.arange(96)[:, None]
    mu = np.linspace(10, 86, N)[None, :]
    S = np.exp(-0.5 * ((x - mu) / 6.0) ** 2) + 1e-3
    S = S[:, rng.permutation(N)]
    Ps = O.splitnorm(O.clamp_input(S), 1)[0]
    for m in (0, 1):
        check(f"smooth metric {m}", O.chunk_distance(m, Ps))
    base = rng.random((64, N // 3))
    Dup = np.concatenate([base, base, base], axis=1)[:, rng.permutation(3 * (N // 3))]
    check("duplicates EMD", O.chunk_distance(1, O.splitnorm(O.clamp_input(Dup), 1)[0]))
    Wi = rng.integers(1, 4, (N, N)).astype(float)
    Wi = np.triu(Wi, 1); Wi = Wi + Wi.T
    check("integer ties", Wi)
    cl = np.concatenate([rng.normal(c, 0.01, (8, N // 5)) for c in range(5)], axis=1)
    Dc = np.sqrt(((cl[:, :, None] - cl[:, None, :]) ** 2).sum(0))
    check("5 tight clusters", Dc)
    for n in (2, 3, 5, 33, 70):
        Dn = rng.random((n, n)); Dn = np.triu(Dn, 1); Dn = Dn + Dn.T
        check(f"tiny N={n}", Dn)
    hb = os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "hummingbird_gray_u8.npy")
    if os.path.exists(hb):
        img = np.load(hb).astype(np.float64)
        img = img[:, :N] if img.shape[1] > N else img
        Ph = O.splitnorm(O.clamp_input(img), 1)[0]
        check("hummingbird L2", O.chunk_distance(0, Ph))
