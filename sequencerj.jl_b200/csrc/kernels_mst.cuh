// K5b: dense MST for FEW graphs per launch -- kNN-filtered Boruvka.  Replaces `_mst` (reference
// src/sequencer.jl:358-374, LightGraphs.kruskal_mst) when a launch holds too few graphs for the batched Prim
// (kernels_graph.cuh) to be bandwidth-bound: Prim is N dependent steps of one DRAM round trip each (~1.6 us), so 6
// group graphs at N = 20,000 cost 40 ms whatever the hardware could stream.  Here the matrix is read ONCE, by every SM:
//
//   1. knn_rows_kernel<false>: one warp per row keeps, per lane, the 4 smallest (weight, j) of its interleaved share;
//      the row's list is the 16 smallest entries that are <= L, L = the smallest "last entry" among the lanes that
//      overflowed, plus a bound B: every edge of the row that is NOT listed weighs >= B.  HBM-bound, 8 N^2 bytes.
//   2. boruvka_knn_kernel (one CTA per graph): Boruvka rounds over the lists.  The first list entry outside a vertex's
//      component IS its lightest outgoing edge (the list is sorted and everything unlisted is heavier).  A vertex whose
//      list is used up ("exhausted") can be ignored while B > the component's current candidate; otherwise the
//      component is stuck for this round (it does not hook; others may hook onto it) and the vertex goes on the needy
//      list.
//   3. knn_rows_kernel<true>: the needy vertices' rows are re-read by the whole GPU, now keeping only vertices outside
//      the component; then back to 2.  On unstructured data nothing is ever rescanned; on clustered data (tight
//      clusters larger than the list) one round rescans about one matrix worth of rows (tools/knn_boruvka_proto.py).
//   4. root_tree_kernel turns the edge list into the (parent, weight, order) arrays tree_kernel consumes.
//
// Edges are ordered like the reference's Kruskal visits them, (weight, max(u,v), min(u,v)); within one row that order
// is (weight, j).  Under a strict total order the MST is unique and every edge Boruvka adds belongs to it, so the
// result is Prim's / Kruskal's tree bit for bit, also with tied weights.  A graph that is not finished after the
// fixed number of rounds launched (or is disconnected) keeps done = 0 and is handed to the Prim kernel, which
// skips the finished ones.
#pragma once
#include "common.cuh"
#include "kernels_graph.cuh"

namespace seqj {

constexpr int kKnnK = 16;          // list entries per vertex
constexpr int kKnnWarps = 8;       // rows per CTA of the list kernels
constexpr unsigned long long kKnnNone = ~0ull;

struct KnnParams {
  const double* D;
  int64_t ldD, strideD;
  const int* mat_slot;
  int N;
  double eps_floor;
  int64_t ldv;
  unsigned long long* lk;      // [graph][ldv][K] weights (bit patterns), ascending by (weight, j)
  int* lj;                     // [graph][ldv][K] neighbours
  int* cnt;                    // [graph][ldv] entries in the list
  int* ptr;                    // [graph][ldv] first entry not known to be inside the component
  unsigned long long* bound;   // [graph][ldv] every unlisted edge weighs >= bound; ~0 = the list is complete
  const int* comp;             // rescan: component labels, needy vertices and their count, done flags
  const int* needy;
  const int* ncount;
  const int* done;
  int comp_in_smem;            // rescan: the CTA stages the graph's component labels in dynamic shared memory (4 N bytes)
};

// One warp per row.  Lane l owns the 16-byte pairs l, l + 32, ... of the row (coalesced 512-byte requests, four in
// flight per lane) and keeps its four smallest eligible entries sorted in registers; the test against the fourth
// is the only work on the common path.  Eligible = not the diagonal (first pass) / not in the row's component
// (rescan).  Scanning j in ascending order with a strict comparison keeps ties in ascending j.
//
// Rescan lists hold at most ONE entry per component (the lightest edge into it): an edge into a component that is
// heavier than another edge of the same row into the same component can never be the row's lightest outgoing edge --
// components only merge, so the two far ends stay together for ever.  Without this a tight cluster larger than the
// list used up the list again after every merge with a neighbouring cluster and had to be rescanned in every round
// (one matrix worth of rows per round, profiles/r01_knn_boruvka.md); with it a rescanned row knows its lightest edge
// into each of the 16 nearest components.
template <bool RESCAN>
__global__ void __launch_bounds__(kKnnWarps * 32) knn_rows_kernel(KnnParams p) {
  extern __shared__ int knn_comp_sm[];
  const int g = blockIdx.y, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int N = p.N;
  const int64_t gofs = (int64_t)g * p.ldv;
  if (RESCAN && p.done[g]) return;
  const int nrows = RESCAN ? p.ncount[g] : N;
  const int slot = p.mat_slot ? p.mat_slot[g] : g;
  const double* __restrict__ D = p.D + (int64_t)slot * p.strideD;
  const int* __restrict__ comp = RESCAN ? p.comp + gofs : nullptr;
  if (RESCAN) {
    // Every entry of the row that lies inside the row's own component is smaller than anything the list will keep
    // (that is why the row ran dry), so it passes the cheap test and needs its label: on clustered data that is a
    // dependent global load per element of a whole cluster.  The CTA therefore stages the labels once (CTA-uniform
    // conditions: done / row count depend on the graph and blockIdx only).
    if ((int)(blockIdx.x * kKnnWarps) >= nrows) return;
    if (p.comp_in_smem) {
      for (int i = threadIdx.x; i < N; i += kKnnWarps * 32) knn_comp_sm[i] = comp[i];
      __syncthreads();
      comp = knn_comp_sm;
    }
  }
  const unsigned long long eps_bits = (unsigned long long)__double_as_longlong(p.eps_floor);
  constexpr int U = 4;
  for (int r = blockIdx.x * kKnnWarps + warp; r < nrows; r += gridDim.x * kKnnWarps) {
    const int v = RESCAN ? p.needy[gofs + r] : r;
    const int cv = RESCAN ? comp[v] : 0;
    const double* __restrict__ row = D + (int64_t)v * p.ldD;
    unsigned long long k0 = kKnnNone, k1 = kKnnNone, k2 = kKnnNone, k3 = kKnnNone;
    int j0 = 0x7fffffff, j1 = 0x7fffffff, j2 = 0x7fffffff, j3 = 0x7fffffff;
    int c0 = -1, c1 = -1, c2 = -1, c3 = -1;               // rescan only: component of each kept entry
    auto consider = [&](unsigned long long x, int idx) {
      if (x < k3) {                                        // rare once the list has settled
        if (!RESCAN) {
          if (idx < N && idx != v) {
            x = max(x, eps_bits);
            if (x < k3) {
              if (x < k2) {
                k3 = k2; j3 = j2;
                if (x < k1) {
                  k2 = k1; j2 = j1;
                  if (x < k0) { k1 = k0; j1 = j0; k0 = x; j0 = idx; }
                  else { k1 = x; j1 = idx; }
                } else { k2 = x; j2 = idx; }
              } else { k3 = x; j3 = idx; }
            }
          }
        }
      }
    };
    // Rescan: the test against the fourth kept entry is no filter here -- with one entry per component the fourth entry
    // is the lightest edge into the FOURTH nearest component, so every vertex of the three nearest components (and of
    // the row's own) passes it, some lane of the warp takes the divergent insertion in nearly every step and the kernel
    // ran at 2.3 TB/s on clustered data (profiles/r02_mst.md).  The label is therefore looked up for every element
    // (shared memory) and a branch-free prefilter drops what cannot change the lane's list: own component, or a
    // component the lane already holds with an edge that is not heavier.
    auto consider_rescan = [&](unsigned long long x, int idx) {
      const int cc = comp[min(idx, N - 1)];
      x = max(x, eps_bits);
      const bool useful = (idx < N) & (x < k3) & (cc != cv) &
                          !(((cc == c0) & (x >= k0)) | ((cc == c1) & (x >= k1)) | ((cc == c2) & (x >= k2)));
      if (useful) {
        // position of an entry of the same component, if the lane keeps one (4 = none)
        const int q = (cc == c0) ? 0 : (cc == c1) ? 1 : (cc == c2) ? 2 : (cc == c3) ? 3 : 4;
        // remove entry q (or the last one) by shifting the tail up, then insert (x, idx, cc) in order
        if (q <= 0) { k0 = k1; j0 = j1; c0 = c1; }
        if (q <= 1) { k1 = k2; j1 = j2; c1 = c2; }
        if (q <= 2) { k2 = k3; j2 = j3; c2 = c3; }
        if (x < k2) {
          k3 = k2; j3 = j2; c3 = c2;
          if (x < k1) {
            k2 = k1; j2 = j1; c2 = c1;
            if (x < k0) { k1 = k0; j1 = j0; c1 = c0; k0 = x; j0 = idx; c0 = cc; }
            else { k1 = x; j1 = idx; c1 = cc; }
          } else { k2 = x; j2 = idx; c2 = cc; }
        } else { k3 = x; j3 = idx; c3 = cc; }
      }
    };
    for (int base = 2 * lane; base < N; base += 64 * U) {
      ulonglong2 dv[U];
#pragma unroll
      for (int q = 0; q < U; ++q) {
        const int idx = base + 64 * q;
        if (idx < N) dv[q] = ldg_stream_u64x2(row + idx);   // idx even, ldD % 16 == 0: aligned, idx + 1 < ldD
      }
#pragma unroll
      for (int q = 0; q < U; ++q) {
        const int idx = base + 64 * q;
        if (idx < N) {
          if (RESCAN) { consider_rescan(dv[q].x, idx); consider_rescan(dv[q].y, idx + 1); }
          else { consider(dv[q].x, idx); consider(dv[q].y, idx + 1); }
        }
      }
    }
    // L = the smallest (k3, j3) among the lanes whose list is full: whatever such a lane did not list is larger
    // (or, in a rescan, heavier than a listed edge into the same component)
    const bool full = k3 != kKnnNone;
    const unsigned fh = full ? (unsigned)(k3 >> 32) : 0xffffffffu;
    const unsigned mh = __reduce_min_sync(0xffffffffu, fh);
    const unsigned fl = (full && fh == mh) ? (unsigned)k3 : 0xffffffffu;
    const unsigned ml = __reduce_min_sync(0xffffffffu, fl);
    const unsigned fj = (full && fh == mh && fl == ml) ? (unsigned)j3 : 0x7fffffffu;
    const unsigned mj = __reduce_min_sync(0xffffffffu, fj);
    const bool anyfull = __any_sync(0xffffffffu, full);
    const unsigned long long Lk = anyfull ? (((unsigned long long)mh << 32) | ml) : kKnnNone;
    const int Lj = anyfull ? (int)mj : 0x7fffffff;
    // the K smallest entries <= L (rescan: of distinct components), by warp-argmins over the lanes' heads
    unsigned long long myk = kKnnNone, lastk = kKnnNone;
    int myj = -1, myc = -2, count = 0;
    for (int t = 0; t < (RESCAN ? 4 * 32 : kKnnK) && count < kKnnK; ++t) {
      const unsigned hh = (unsigned)(k0 >> 32);
      const unsigned mhh = __reduce_min_sync(0xffffffffu, hh);
      const unsigned hl = (hh == mhh) ? (unsigned)k0 : 0xffffffffu;
      const unsigned mhl = __reduce_min_sync(0xffffffffu, hl);
      const unsigned long long mk = ((unsigned long long)mhh << 32) | mhl;
      if (mk == kKnnNone) break;                            // every head is empty (uniform)
      const unsigned hj = (k0 == mk) ? (unsigned)j0 : 0x7fffffffu;
      const int mjj = (int)__reduce_min_sync(0xffffffffu, hj);
      if (mk > Lk || (mk == Lk && mjj > Lj)) break;         // beyond what the lanes can vouch for
      const bool mine = (k0 == mk && j0 == mjj);
      bool dup = false;
      int cstar = 0;
      if (RESCAN) {                                         // one entry per component: skip what a lighter edge covers
        const unsigned owner = __ballot_sync(0xffffffffu, mine);
        cstar = __shfl_sync(0xffffffffu, c0, __ffs(owner) - 1);
        dup = __any_sync(0xffffffffu, lane < count && myc == cstar);
      }
      if (!dup) {
        if (lane == count) { myk = mk; myj = mjj; myc = cstar; }
        lastk = mk;
        ++count;
      }
      if (mine) { k0 = k1; j0 = j1; c0 = c1; k1 = k2; j1 = j2; c1 = c2; k2 = k3; j2 = j3; c2 = c3; k3 = kKnnNone; j3 = 0x7fffffff; c3 = -1; }
    }
    const int64_t vo = gofs + v;
    if (lane < kKnnK) { p.lk[vo * kKnnK + lane] = myk; p.lj[vo * kKnnK + lane] = myj; }
    if (lane == 0) {
      p.cnt[vo] = count;
      p.ptr[vo] = 0;
      p.bound[vo] = (count == kKnnK) ? lastk : Lk;
    }
  }
}

struct BorKnnParams {
  int N;
  int64_t ldv;
  const unsigned long long* lk;
  const int* lj;
  const int* cnt;
  const unsigned long long* bound;
  int* ptr;
  int* comp;                   // [graph][ldv] labels between launches
  unsigned long long* best;    // [graph][ldv] scratch per component: lightest outgoing weight / hook target
  unsigned long long* bestp;   // [graph][ldv] scratch per component: (max << 32 | min) of that edge
  int* stuck;                  // [graph][ldv] scratch per component
  int* needy;                  // [graph][ldv] vertices to rescan
  int* ncount;                 // [graph]
  int* ecount;                 // [graph] MST edges found so far
  int* done;                   // [graph] 1 = all N-1 edges found, 2 = handed back to Prim (rescan budget used up)
  int* nresc;                  // [graph] rows rescanned so far
  int rescan_budget;           // rows a graph may have rescanned in total before it is handed to the batched Prim
  int rescan_once;             // ... and rows it may ask for in ONE request (blocked with most of its vertices needy)
  int* out_i;                  // [graph][ldv] MST edges (i < j)
  int* out_j;
  double* out_w;
  int use_smem;                // labels fit in dynamic shared memory (4 N bytes)
  int first;                   // first launch: initialise the state
};

constexpr int kBorKnnThreads = 1024;

__global__ void __launch_bounds__(kBorKnnThreads, 1) boruvka_knn_kernel(BorKnnParams p) {
  extern __shared__ int borknn_sm[];
  __shared__ int s_count, s_needy, s_hooks;
  const int tid = threadIdx.x, N = p.N, g = blockIdx.x;
  const int64_t gofs = (int64_t)g * p.ldv;
  if (!p.first && p.done[g]) return;
  int* compg = p.comp + gofs;
  int* comp = p.use_smem ? borknn_sm : compg;
  int* ptr = p.ptr + gofs;
  const int* __restrict__ cnt = p.cnt + gofs;
  const unsigned long long* __restrict__ bound = p.bound + gofs;
  const unsigned long long* __restrict__ lk = p.lk + gofs * kKnnK;
  const int* __restrict__ lj = p.lj + gofs * kKnnK;
  unsigned long long* best = p.best + gofs;
  unsigned long long* bestp = p.bestp + gofs;
  int* stuck = p.stuck + gofs;
  int* needy = p.needy + gofs;
  int* out_i = p.out_i + gofs;
  int* out_j = p.out_j + gofs;
  double* out_w = p.out_w + gofs;
  if (p.first) {
    for (int v = tid; v < N; v += kBorKnnThreads) comp[v] = v;
  } else if (p.use_smem) {
    for (int v = tid; v < N; v += kBorKnnThreads) comp[v] = compg[v];
  }
  if (tid == 0) s_count = p.first ? 0 : p.ecount[g];
  __syncthreads();
  int status = (N <= 1) ? 1 : 0;       // 1 = complete, 2 = rows to rescan, 3 = no progress (disconnected input)
  for (int it = 0; it < 64 && status == 0; ++it) {
    for (int v = tid; v < N; v += kBorKnnThreads) { best[v] = ~0ull; bestp[v] = ~0ull; stuck[v] = 0; }
    if (tid == 0) { s_needy = 0; s_hooks = 0; }
    __syncthreads();
    // 1. every vertex: skip the list entries that have joined its component; lightest outgoing weight per component
    for (int v = tid; v < N; v += kBorKnnThreads) {
      const int c = comp[v], n = cnt[v];
      int q = ptr[v];
      while (q < n && comp[lj[(int64_t)v * kKnnK + q]] == c) ++q;
      ptr[v] = q;
      if (q < n) atomicMin(&best[c], lk[(int64_t)v * kKnnK + q]);
    }
    __syncthreads();
    // 2. among the lightest, the first vertex pair in Kruskal's order; exhausted vertices that could hide a lighter
    //    edge make their component stuck and ask for a rescan
    for (int v = tid; v < N; v += kBorKnnThreads) {
      const int c = comp[v], n = cnt[v], q = ptr[v];
      if (q < n) {
        if (lk[(int64_t)v * kKnnK + q] == best[c]) {
          const int j = lj[(int64_t)v * kKnnK + q];
          atomicMin(&bestp[c], ((unsigned long long)(unsigned)max(v, j) << 32) | (unsigned)min(v, j));
        }
      } else {
        const unsigned long long b = bound[v];
        if (b != kKnnNone && b <= best[c]) {
          stuck[c] = 1;
          needy[atomicAdd(&s_needy, 1)] = v;
        }
      }
    }
    __syncthreads();
    // 3. hook every component that knows its lightest outgoing edge; a mutual pair keeps the smaller id as root
    for (int c = tid; c < N; c += kBorKnnThreads) {
      if (comp[c] != c || stuck[c]) continue;
      const unsigned long long pr = bestp[c];
      if (pr == ~0ull) continue;
      const int i = (int)(unsigned)(pr & 0xffffffffull), j = (int)(unsigned)(pr >> 32);
      const int ci = comp[i], cj = comp[j];
      const int other = (ci == c) ? cj : ci;
      const bool mutual = !stuck[other] && bestp[other] == pr;
      if (mutual && c < other) continue;                  // `other` hooks onto c and records the edge
      const int k = atomicAdd(&s_count, 1);
      out_i[k] = i;
      out_j[k] = j;
      out_w[k] = __longlong_as_double((long long)best[c]);
      best[c] = (1ull << 63) | (unsigned long long)(unsigned)other;   // only this thread reads best[c] from here on
      atomicAdd(&s_hooks, 1);
    }
    __syncthreads();
    for (int c = tid; c < N; c += kBorKnnThreads)
      if (comp[c] == c && !stuck[c] && bestp[c] != ~0ull && (best[c] >> 63)) comp[c] = (int)(best[c] & 0xffffffffull);
    __syncthreads();
    for (int jt = 0; jt < 32; ++jt) {                      // pointer jumping until every vertex points at its root
      int changed = 0;
      for (int v = tid; v < N; v += kBorKnnThreads) {
        const int c = comp[v], cc = comp[c];
        if (c != cc) { comp[v] = cc; changed = 1; }
      }
      if (!__syncthreads_or(changed)) break;
    }
    __syncthreads();
    const int ec = s_count, nd = s_needy, hk = s_hooks;
    __syncthreads();
    // Keep going while any component hooks: the stuck ones wait (others may hook onto them) and the launch only asks
    // for a rescan once nothing moves any more.  (Round 1 returned at the first needy vertex; on clustered data that
    // spent the fixed schedule of launches on a handful of rows each and rescanned clusters that were still merging.)
    if (ec == N - 1) status = 1;
    else if (hk == 0) status = (nd > 0) ? 2 : 3;
  }
  if (p.use_smem)
    for (int v = tid; v < N; v += kBorKnnThreads) compg[v] = comp[v];
  if (tid == 0) {
    // Rescan budget: every needy row is one more 8N-byte read by the whole GPU.  That is cheap for a launch of few
    // graphs and unbounded for a large batch of tightly clustered graphs (a cluster larger than the list is rescanned
    // in every round in which it runs dry), so a graph whose rescans would exceed the budget is handed to the
    // batched Prim instead (done = 2: the list kernels and later rounds skip it, Prim does not).
    int nd = (status == 2) ? s_needy : 0, dn = (status == 1) ? 1 : 0;
    const int used = p.first ? 0 : p.nresc[g];
    if (nd > 0 && (used + nd > p.rescan_budget || nd > p.rescan_once)) { nd = 0; dn = 2; }
    p.nresc[g] = used + nd;
    p.ecount[g] = s_count;
    p.ncount[g] = nd;
    p.done[g] = dn;
  }
}

}  // namespace seqj
