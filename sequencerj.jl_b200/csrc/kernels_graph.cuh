// K5-K8: graph layer.  Replaces `_measure_dm` (reference src/sequencer.jl:284-290):
//   _mst            (:358-374, LightGraphs.kruskal_mst)      -> batched dense Prim, one CTA (or cluster) per graph;
//                                                                launches of few graphs: kernels_mst.cuh
//   _leastcentralpt (:331, closeness_centrality)             -> O(N) rerooted tree-distance sums
//   elongation      (:334-354)                               -> hop depths by pointer jumping
//   bfs_tree        (:288, src/bfs.jl:6-9)                   -> re-rooted parent vector
//   unroll          (src/utils.jl:15-24)                     -> reverse DFS preorder, children ascending
// plus the eta-weighted combine (:231-247) and the proximity fuse (:310-328, :269).
//
// Edges are compared in the order the reference's Kruskal visits them, (weight, max(u,v), min(u,v)): a strict
// total order, under which the MST is unique, so Prim, Boruvka and Kruskal return the same tree also when weights tie.
#pragma once
#include <math_constants.h>
#include "common.cuh"

namespace seqj {

constexpr int kPrimThreads = 1024;
constexpr int kPrimUnroll = 8;      // 8 x 16 B loads per thread in flight: one batch covers N <= 16384
constexpr double kJuliaEps = 2.220446049250313e-16;       // eps(Float64)
constexpr double kFloatMax32 = 3.4028234663852886e38;     // floatmax(Float32)

struct PrimParams {
  const double* D;      // pool of dense symmetric N x N matrices (row r of matrix s at D + s*strideD + r*ldD)
  int64_t ldD, strideD;
  const int* mat_slot;  // optional: graph g reads matrix mat_slot[g] (else matrix g)
  int N;
  double eps_floor;     // _mst clamps to [eps, Inf] (src/sequencer.jl:360)
  int* parent;          // [graph][ldv]  Prim parent (start vertex 0 -> itself)
  double* weight;       // [graph][ldv]  weight of the edge (v, parent[v])
  int* order;           // [graph][ldv]  insertion order, order[0] = 0
  int64_t ldv;
  double* key_g;        // global scratch [graph][ldv] (used when the keys do not fit in smem)
  int* from_g;
  int use_smem;
  int prefetch;         // cluster kernel: prefetch the losing candidates' rows into L2 (few graphs in flight)
  const int* skip_done = nullptr;   // optional [graph]: graphs already solved by the kNN-Boruvka path (kernels_mst.cuh)
};

__device__ __forceinline__ void argmin_combine(double& k, int& v, double ok, int ov) {
  if (ok < k || (ok == k && ov < v)) { k = ok; v = ov; }
}

__device__ __forceinline__ double2 ldg_stream_f64x2(const double* p) {
  double2 v;
  asm volatile("ld.global.nc.L1::no_allocate.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "l"(p));
  return v;
}
__device__ __forceinline__ ulonglong2 ldg_stream_u64x2(const double* p) {
  ulonglong2 v;
  asm volatile("ld.global.nc.L1::no_allocate.v2.u64 {%0, %1}, [%2];" : "=l"(v.x), "=l"(v.y) : "l"(p));
  return v;
}

// Edge order of the reference's Kruskal: stable sort by weight of the upper-triangle edges enumerated in
// column-major order (LightGraphs.kruskal_mst over SimpleWeightedGraph), i.e. the strict total order
// (weight, max(u,v), min(u,v)).  The MST under a strict total order is unique, so Prim with the same order
// returns exactly Kruskal's tree even when weights tie (duplicate objects, integer data).
// Candidates are ordered by (key - kPrimDec): subtracting one from the HIGH word keeps the order of the real keys
// (weights >= eps have a high word >= 1) and sends the in-tree marker 0 to the top (high word 0xffffffff), above
// kPrimNone = "no candidate"; the decrement touches one register.  The relaxation test of the hot loop looks at high
// words only (`hi(d) <= hi(key)` is a superset of d <= key and false for almost every vertex); the clamp, the relax
// and the tie rule sit behind it.  (A high-word filter in front of the running argmin was measured slower: a
// thread's minimum over its 5-10 vertices changes too often for the branch to be rare at warp level.)
constexpr unsigned long long kPrimDec = 1ull << 32;
constexpr unsigned long long kPrimNone = 0xfffffffe00000000ull;

__device__ __forceinline__ bool edge_less(int v1, int f1, int v2, int f2) {      // weights equal
  const int h1 = max(v1, f1), h2 = max(v2, f2);
  return h1 < h2 || (h1 == h2 && min(v1, f1) < min(v2, f2));
}

// Warp argmin of candidates (key = weight bits - 1, v = vertex outside the tree) under that order.  The fast path is
// three REDUX.MIN.U32 (key high word, key low word, vertex) plus a vote that detects weight ties; only then (rare) the
// tied lanes fetch the tree endpoint of their candidate through `getf` and two more reductions order them by
// (max, min) of the endpoints.
struct NoHook { __device__ __forceinline__ void operator()() const {} };

// `pre` / `post` bracket the tie path (the cluster-level reduction fences its remote reads with them).
template <class F, class Pre = NoHook, class Post = NoHook>
__device__ __forceinline__ void warp_argmin_edge(unsigned long long& k, int& v, F getf, Pre pre = Pre(), Post post = Post()) {
  const unsigned hi = (unsigned)(k >> 32), lo = (unsigned)k;
  const unsigned mhi = __reduce_min_sync(0xffffffffu, hi);
  const unsigned mlo = __reduce_min_sync(0xffffffffu, hi == mhi ? lo : 0xffffffffu);
  const bool win = (hi == mhi) && (lo == mlo);
  const unsigned m = __ballot_sync(0xffffffffu, win);
  int mv = (int)__reduce_min_sync(0xffffffffu, win ? (unsigned)v : 0xffffffffu);
  if (__popc(m) > 1) {
    pre();
    const int f = (win && v != 0x7fffffff) ? getf(v) : 0x7fffffff;
    const unsigned eh = win ? (unsigned)max(v, f) : 0xffffffffu;
    const unsigned mh = __reduce_min_sync(0xffffffffu, eh);
    const unsigned el = (win && eh == mh) ? (unsigned)min(v, f) : 0xffffffffu;
    const unsigned ml = __reduce_min_sync(0xffffffffu, el);
    mv = __shfl_sync(0xffffffffu, v, __ffs(__ballot_sync(0xffffffffu, win && eh == mh && el == ml)) - 1);
    post();
  }
  v = mv;
  k = ((unsigned long long)mhi << 32) | mlo;
}

// Dense Prim, one CTA per graph (reference _mst, src/sequencer.jl:358-374).
// HBM-bound by construction: each of the N-1 steps streams one 8N-byte row exactly once (8*N^2 bytes
// per graph).  Keys live in shared memory as the bit patterns of non-negative doubles, which order
// like unsigned integers: the inner loop and both argmin levels are integer-only (no FP64-pipe
// compares).  key == 0 marks a vertex already in the tree: no weight (>= eps > 0) is smaller, so it is
// never relaxed, and (key - kPrimDec) wraps to the largest values, so it never wins the argmin.
// Thread t owns the vertex pairs {2t, 2t+1} + 2048 r: rows are fetched with 16-byte streaming loads,
// all R of them in flight at once.  R = 0 selects the generic path (any N, batches of 8 loads).
template <int R>
__global__ void __launch_bounds__(kPrimThreads, 1) prim_kernel(PrimParams p) {
  extern __shared__ unsigned long long prim_sm[];
  __shared__ unsigned long long rk[2][32];
  __shared__ int rv[2][32], rf[2][32];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int N = p.N;
  const int Npad = (N + 1) & ~1;
  if (p.skip_done && p.skip_done[blockIdx.x] == 1) return;   // 1 = solved by the Boruvka path (2 = handed back to Prim)
  const int64_t gofs = (int64_t)blockIdx.x * p.ldv;
  const int slot = p.mat_slot ? p.mat_slot[blockIdx.x] : (int)blockIdx.x;
  const double* __restrict__ D = p.D + (int64_t)slot * p.strideD;
  unsigned long long* key;
  int* from;
  if (p.use_smem) {
    key = prim_sm;
    from = reinterpret_cast<int*>(prim_sm + Npad);
  } else {
    key = reinterpret_cast<unsigned long long*>(p.key_g + gofs);
    from = p.from_g + gofs;
  }
  int* parent = p.parent + gofs;
  double* weight = p.weight + gofs;
  int* order = p.order + gofs;
  const unsigned long long kInf = 0x7ff0000000000000ull;
  const unsigned long long eps_bits = (unsigned long long)__double_as_longlong(p.eps_floor);

  for (int v = tid; v < Npad; v += kPrimThreads) { key[v] = (v < N) ? kInf : 0ull; from[v] = 0; }
  __syncthreads();
  if (tid == 0) { key[0] = 0ull; parent[0] = 0; weight[0] = 0.0; order[0] = 0; }
  __syncthreads();

  constexpr int U = (R > 0) ? R : kPrimUnroll;
  int u = 0, par = 0;
  for (int step = 1; step < N; ++step) {
    const double* __restrict__ row = D + (int64_t)u * p.ldD;
    unsigned long long bk = kPrimNone;
    int bv = 0x7fffffff, bl = 0x7fffffff;
    for (int base = 2 * tid; base < N; base += 2 * kPrimThreads * U) {
      ulonglong2 dv[U];
#pragma unroll
      for (int r = 0; r < U; ++r) {
        const int v = base + r * 2 * kPrimThreads;
        if (v < N) dv[r] = ldg_stream_u64x2(row + v);      // v even, ldD % 16 == 0: 16 B aligned; v+1 < ldD
      }
#pragma unroll
      for (int r = 0; r < U; ++r) {
        const int v = base + r * 2 * kPrimThreads;
        if (v < N) {
          ulonglong2 k2 = *reinterpret_cast<ulonglong2*>(key + v);
          if ((unsigned)(dv[r].x >> 32) <= (unsigned)(k2.x >> 32) || (unsigned)(dv[r].y >> 32) <= (unsigned)(k2.y >> 32)) {
            // Relaxation candidates: rare after the first steps, so everything careful lives in this branch -- the
            // eps clamp (keys are clamped values, so a clamped weight can only relax a key if the raw one passes the
            // high-word test above) and the tie rule: at equal weight the edge that comes first in Kruskal's order wins.
            const unsigned long long d0 = max(dv[r].x, eps_bits), d1 = max(dv[r].y, eps_bits);
            const bool u0 = k2.x != 0ull && (d0 < k2.x || (d0 == k2.x && edge_less(v, u, v, from[v])));
            const bool u1 = k2.y != 0ull && (d1 < k2.y || (d1 == k2.y && edge_less(v + 1, u, v + 1, from[v + 1])));
            if (u0) { k2.x = d0; from[v] = u; }
            if (u1) { k2.y = d1; from[v + 1] = u; }
            if (u0 || u1) *reinterpret_cast<ulonglong2*>(key + v) = k2;
          }
          // bv = first, bl = last of this thread's vertices at its minimum key: bv != bl flags a tie
          const unsigned long long c0 = k2.x - kPrimDec, c1 = k2.y - kPrimDec;
          if (c0 < bk) { bk = c0; bv = v; }
          if (c0 <= bk) bl = v;
          if (c1 < bk) { bk = c1; bv = v + 1; }
          if (c1 <= bk) bl = v + 1;
        }
      }
      if (R > 0) break;
    }
    if (bl != bv) {                                        // rare: order the tied candidates like Kruskal
      for (int base = 2 * tid; base < N; base += 2 * kPrimThreads) {
        const ulonglong2 k2 = *reinterpret_cast<ulonglong2*>(key + base);
        if (k2.x - kPrimDec == bk && base != bv && edge_less(base, from[base], bv, from[bv])) bv = base;
        if (k2.y - kPrimDec == bk && base + 1 != bv && edge_less(base + 1, from[base + 1], bv, from[bv])) bv = base + 1;
      }
    }
    auto getf = [&](int v) { return from[v]; };
    warp_argmin_edge(bk, bv, getf);
    // The tree endpoint of the warp's candidate travels with it: after the barrier a faster warp may already be relaxing
    // from[] for the next step while a slower one still orders tied candidates.
    if (lane == 0) { rk[par][warp] = bk; rv[par][warp] = bv; rf[par][warp] = (bv != 0x7fffffff) ? from[bv] : 0x7fffffff; }
    __syncthreads();
    bk = rk[par][lane];
    bv = rv[par][lane];
    const int bf = rf[par][lane];
    warp_argmin_edge(bk, bv, [&](int) { return bf; });
    u = bv;
    if (((u >> 1) & (kPrimThreads - 1)) == tid) {       // the owner of key[u] retires it
      order[step] = u;
      parent[u] = from[u];
      weight[u] = __longlong_as_double((long long)(bk + kPrimDec));
      key[u] = 0ull;
    }
    par ^= 1;
  }
}

// ---- cluster Prim: one thread-block CLUSTER per graph ------------------------------------------
// The single-CTA kernel above is bound by what one SM can pull from HBM plus the per-step argmin.
// Here a cluster of C CTAs shares one graph: CTA r owns the vertex slice [r*W, (r+1)*W), streams only
// its 8W-byte part of the row, and keeps its keys in ~12 KB of shared memory, so (a) a step costs one
// DRAM latency plus one DSMEM exchange, and (b) the CTAs are small enough (256 threads) to stay
// resident next to the FP64 distance tiles.  The cluster-wide argmin uses no barrier: every CTA sends
// its 16-byte candidate to all peers with st.async (distributed shared memory) that completes a
// transaction on the receiver's mbarrier; each CTA waits on its own mbarrier and reduces the C
// candidates locally.  Buffers and mbarriers are double-buffered by step parity.
// T threads per CTA: 512 (one pair per thread at N = 10k, shortest per-warp chain; 2 CTAs per SM) when all clusters
// are co-resident that way, else 256 (4 CTAs per SM).

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t mapa_u32(uint32_t addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
  return r;
}
__device__ __forceinline__ uint32_t ld_dsmem_u32(uint32_t addr) {
  uint32_t v;
  asm volatile("ld.shared::cluster.u32 %0, [%1];" : "=r"(v) : "r"(addr) : "memory");
  return v;
}
__device__ __forceinline__ void prefetch_l2_bulk(const void* src, uint32_t bytes) {     // bytes % 16 == 0, src 16 B aligned
  asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(src), "r"(bytes) : "memory");
}
__device__ __forceinline__ void st_async_u64x2(uint32_t dst, unsigned long long a, unsigned long long b, uint32_t mbar) {
  asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v2.b64 [%0], {%1, %2}, [%3];"
               ::"r"(dst), "l"(a), "l"(b), "r"(mbar) : "memory");
}

template <int C, int T>
__global__ void __cluster_dims__(C, 1, 1) __launch_bounds__(T, 1024 / T)
prim_cluster_kernel(PrimParams p, int W) {
  constexpr int kU = 1024 / T;                        // 16-byte loads in flight per thread: slices of <= 2048 vertices in one pass
  extern __shared__ unsigned long long primc_sm[];
  __shared__ __align__(16) ulonglong2 cand[2][C];
  __shared__ __align__(8) uint64_t mbar[2];
  __shared__ unsigned long long wk[T / 32];
  __shared__ int wv[T / 32];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int N = p.N;
  const int graph = blockIdx.x / C;
  if (p.skip_done && p.skip_done[graph] == 1) return; // the whole cluster leaves together, before any barrier
  const uint32_t rank = cluster_ctarank();
  const int v0 = (int)rank * W;                       // first vertex of this CTA's slice (W is even)
  const int Wn = max(0, min(W, N - v0));              // vertices in the slice
  const int Wp = (Wn + 1) & ~1;
  const int64_t gofs = (int64_t)graph * p.ldv;
  const int slot = p.mat_slot ? p.mat_slot[graph] : graph;
  const double* __restrict__ D = p.D + (int64_t)slot * p.strideD;
  unsigned long long* key = primc_sm;
  int* from = reinterpret_cast<int*>(primc_sm + ((W + 1) & ~1));
  int* parent = p.parent + gofs;
  double* weight = p.weight + gofs;
  int* order = p.order + gofs;
  const unsigned long long kInf = 0x7ff0000000000000ull;
  const unsigned long long eps_bits = (unsigned long long)__double_as_longlong(p.eps_floor);

  for (int v = tid; v < Wp; v += T) { key[v] = (v < Wn) ? kInf : 0ull; from[v] = 0; }
  if (tid == 0) {
    mbar_init(&mbar[0], 1);
    mbar_init(&mbar[1], 1);
    fence_barrier_init();
    if (rank == 0) { key[0] = 0ull; parent[0] = 0; weight[0] = 0.0; order[0] = 0; }
  }
  __syncthreads();
  cluster_sync_all();                                  // peers' mbarriers are initialised before any st.async

  const uint32_t cand_u32 = smem_u32(&cand[0][0]);
  const uint32_t mbar_u32 = smem_u32(&mbar[0]);
  const uint32_t from_u32 = smem_u32(from);
  int u = 0;
  for (int step = 1; step < N; ++step) {
    const int par = step & 1;
    const double* __restrict__ row = D + (int64_t)u * p.ldD + v0;
    unsigned long long bk = kPrimNone;
    int bv = 0x7fffffff, bl = 0x7fffffff;
    for (int base = 2 * tid; base < Wn; base += 2 * T * kU) {
      ulonglong2 dv[kU];
#pragma unroll
      for (int r = 0; r < kU; ++r) {
        const int v = base + r * 2 * T;
        if (v < Wn) dv[r] = ldg_stream_u64x2(row + v);
      }
#pragma unroll
      for (int r = 0; r < kU; ++r) {
        const int v = base + r * 2 * T;
        if (v < Wn) {
          ulonglong2 k2 = *reinterpret_cast<ulonglong2*>(key + v);
          const int gv = v0 + v;                                         // global vertex id
          if ((unsigned)(dv[r].x >> 32) <= (unsigned)(k2.x >> 32) || (unsigned)(dv[r].y >> 32) <= (unsigned)(k2.y >> 32)) {                      // rare: clamp + tie rule as in prim_kernel
            const unsigned long long d0 = max(dv[r].x, eps_bits), d1 = max(dv[r].y, eps_bits);
            const bool u0 = k2.x != 0ull && (d0 < k2.x || (d0 == k2.x && edge_less(gv, u, gv, from[v])));
            const bool u1 = k2.y != 0ull && (d1 < k2.y || (d1 == k2.y && edge_less(gv + 1, u, gv + 1, from[v + 1])));
            if (u0) { k2.x = d0; from[v] = u; }
            if (u1) { k2.y = d1; from[v + 1] = u; }
            if (u0 || u1) *reinterpret_cast<ulonglong2*>(key + v) = k2;
          }
          const unsigned long long c0 = k2.x - kPrimDec, c1 = k2.y - kPrimDec;
          if (c0 < bk) { bk = c0; bv = gv; }
          if (c0 <= bk) bl = gv;
          if (c1 < bk) { bk = c1; bv = gv + 1; }
          if (c1 <= bk) bl = gv + 1;
        }
      }
    }
    if (bl != bv) {                                        // rare: two of this thread's candidates tie at its minimum
      for (int v = 2 * tid; v < Wn; v += 2 * T) {
        const ulonglong2 k2 = *reinterpret_cast<ulonglong2*>(key + v);
        const int gv = v0 + v;
        if (k2.x - kPrimDec == bk && gv != bv && edge_less(gv, from[v], bv, from[bv - v0])) bv = gv;
        if (k2.y - kPrimDec == bk && gv + 1 != bv && edge_less(gv + 1, from[v + 1], bv, from[bv - v0])) bv = gv + 1;
      }
    }
    auto getf = [&](int v) { return from[v - v0]; };
    warp_argmin_edge(bk, bv, getf);
    if (lane == 0) { wk[warp] = bk; wv[warp] = bv; }
    __syncthreads();
    if (warp == 0) {
      bk = (lane < T / 32) ? wk[lane] : kPrimNone;
      bv = (lane < T / 32) ? wv[lane] : 0x7fffffff;
      warp_argmin_edge(bk, bv, getf);
      if (lane == 0) mbar_expect_tx(&mbar[par], C * 16);
      if (lane < C) {                                  // lane j delivers this CTA's candidate to CTA j
        const uint32_t dst = mapa_u32(cand_u32 + (uint32_t)((par * C + (int)rank) * 16), (uint32_t)lane);
        const uint32_t mb = mapa_u32(mbar_u32 + (uint32_t)(par * 8), (uint32_t)lane);
        st_async_u64x2(dst, bk, (unsigned long long)(unsigned)bv, mb);
      }
    }
    mbar_wait(&mbar[par], (uint32_t)(((step - 1) >> 1) & 1));   // k-th use of this barrier -> phase parity k & 1
    {
      // Every warp of every CTA reduces the same C candidates, so all of them take the tie path together: it may
      // therefore use cluster barriers.  The tied candidates' tree endpoints live in their owners' shared memory;
      // the barriers keep every owner from relaxing (rewriting from[]) until all readers are done.
      const ulonglong2 cnd = (lane < C) ? cand[par][lane] : make_ulonglong2(kPrimNone, 0x7fffffffull);
      bk = cnd.x;
      bv = (int)(unsigned)cnd.y;
      auto getf_remote = [&](int v) {
        const int owner = v / W;
        return (int)ld_dsmem_u32(mapa_u32(from_u32 + (uint32_t)((v - owner * W) * 4), (uint32_t)owner));
      };
      auto fence = [&]() { cluster_sync_all(); };
      const int cv = bv;                              // lane j < C: the candidate of CTA j
      warp_argmin_edge(bk, bv, getf_remote, fence, fence);
      // The C - 1 losers stay candidates for the next steps (their keys only compete with newly relaxed
      // vertices), so the next winner is usually one of them: pull this CTA's slice of their rows into L2 now and
      // the next step's row fetch costs an L2 hit instead of a DRAM round trip.  Rows already resident cost no
      // DRAM traffic, so the extra traffic is about one row per step.
      if (p.prefetch && warp == 0 && lane < C && cv != bv && cv != 0x7fffffff && Wn > 1)
        prefetch_l2_bulk(D + (int64_t)cv * p.ldD + v0, (uint32_t)((Wn * 8) & ~15));
    }
    u = bv;
    const int ul = u - v0;
    if (ul >= 0 && ul < Wn && ((ul >> 1) & (T - 1)) == tid) {   // the owner of key[u] retires it
      order[step] = u;
      parent[u] = from[ul];
      weight[u] = __longlong_as_double((long long)(bk + kPrimDec));
      key[ul] = 0ull;
    }
  }
  cluster_sync_all();                                  // no CTA exits while a peer may still write to it
}

// in-place exclusive scan of a[0..n) by one CTA of 1024 threads; returns the total
__device__ int block_excl_scan_i32(int* a, int n, int* s_part) {
  const int tid = threadIdx.x;
  const int per = (n + 1023) / 1024, lo = min(n, tid * per), hi = min(n, lo + per);
  int sum = 0;
  for (int v = lo; v < hi; ++v) sum += a[v];
  s_part[tid] = sum;
  __syncthreads();
  if (tid < 32) {                               // warp 0 scans the 1024 partial sums, 32 per lane
    int run = 0;
    for (int q = 0; q < 32; ++q) { const int x = s_part[tid * 32 + q]; s_part[tid * 32 + q] = run; run += x; }
    int incl = run;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int y = __shfl_up_sync(0xffffffffu, incl, o);
      if (tid >= o) incl += y;
    }
    const int base = incl - run;
    for (int q = 0; q < 32; ++q) s_part[tid * 32 + q] += base;
    if (tid == 31) s_part[1024] = incl;
  }
  __syncthreads();
  int run = s_part[tid];
  for (int v = lo; v < hi; ++v) { const int d = a[v]; a[v] = run; run += d; }
  const int total = s_part[1024];
  __syncthreads();
  return total;
}

// Euler tour of a rooted tree given by parent pointers, ranked by pointer jumping (O(log N) rounds, 1024 threads,
// integers only).  Children lists are built in CSR form (clist / cstart / qpos; ascending inside every list when
// `sorted`); child position q owns two arcs, 2q = parent -> child ("down") and 2q + 1 = back ("up").  Entering a
// vertex continues with its first child, returning from a child with the next sibling, else back to the parent; the
// tour ends when the root's last child returns.  Returns rank[]: the number of arcs after each arc, so that
// position = 2(N-1) - 1 - rank.  cstart has N entries (the end of the last list is N - 1).
__device__ int* euler_tour_rank(int N, int root, const int* par, bool sorted, int* cstart, int* clist, int* qpos,
                                int* fillc, int* tmp, int* b0, int* b1, int* b2, int* b3, int* s_scan) {
  const int tid = threadIdx.x, M = N - 1, A = 2 * M;
  auto cend = [&](int x) { return (x + 1 < N) ? cstart[x + 1] : M; };
  for (int v = tid; v < N; v += 1024) cstart[v] = 0;
  __syncthreads();
  for (int v = tid; v < N; v += 1024)
    if (v != root) atomicAdd(&cstart[par[v]], 1);
  __syncthreads();
  block_excl_scan_i32(cstart, N, s_scan);
  for (int v = tid; v < N; v += 1024) fillc[v] = cstart[v];
  __syncthreads();
  int* first = sorted ? tmp : clist;
  for (int v = tid; v < N; v += 1024)
    if (v != root) {
      const int q = atomicAdd(&fillc[par[v]], 1);
      first[q] = v;
      if (!sorted) qpos[v] = q;
    }
  __syncthreads();
  if (sorted) {                                         // rank sort inside the list (short lists: trees here are path-like)
    for (int q = tid; q < M; q += 1024) {
      const int c = tmp[q], x = par[c];
      const int s0 = cstart[x], s1 = cend(x);
      int r = s0;
      for (int k = s0; k < s1; ++k) r += (tmp[k] < c);
      clist[r] = c;
      qpos[c] = r;
    }
    __syncthreads();
  }
  int *succ_i = b0, *rank_i = b1, *succ_o = b2, *rank_o = b3;
  for (int q = tid; q < M; q += 1024) {
    const int c = clist[q], x = par[c];
    const int up = (q + 1 < cend(x)) ? 2 * (q + 1) : (x == root ? -1 : 2 * qpos[x] + 1);
    succ_i[2 * q] = (cstart[c] < cend(c)) ? 2 * cstart[c] : 2 * q + 1;
    succ_i[2 * q + 1] = up;
    rank_i[2 * q] = 1;
    rank_i[2 * q + 1] = (up < 0) ? 0 : 1;
  }
  __syncthreads();
  for (int span = 1; span < A; span <<= 1) {
    for (int a = tid; a < A; a += 1024) {
      const int sa = succ_i[a];
      if (sa >= 0) { rank_o[a] = rank_i[a] + rank_i[sa]; succ_o[a] = succ_i[sa]; }
      else { rank_o[a] = rank_i[a]; succ_o[a] = -1; }
    }
    __syncthreads();
    int* t = succ_i; succ_i = succ_o; succ_o = t;
    t = rank_i; rank_i = rank_o; rank_o = t;
  }
  return rank_i;
}

// ---- tree measurement -------------------------------------------------------------------------
struct TreeParams {
  int N;
  int64_t ldv;
  const int* parent;      // Prim outputs [graph][ldv]
  const double* weight;
  const int* order;
  // scratch, each [graph][ldv]
  int* size;
  double* wd;             // wd / wd2 and S / S2: ping-pong buffers of the pointer-jumping passes
  double* wd2;
  double* S;
  double* S2;
  int* anc0;
  int* anc1;
  int* dep0;
  int* dep1;
  int use_smem;           // parent + size fit in dynamic shared memory (8N bytes)
  // outputs (per graph g: index out0 + g)
  double* eta;            // [ngraphs]
  int* root;              // [ngraphs]
  int* newparent;         // [ngraphs][ldp]
  int64_t ldp;
  int* mst_i;             // optional [ngraphs][ldp]
  int* mst_j;
  double* mst_w;
  int* dfs_order;         // optional [ngraphs][ldp]
};

constexpr int kTreeThreads = 1024;

// One CTA per tree.  Replaces _leastcentralpt (N Dijkstra runs in the reference, src/sequencer.jl:331)
// by the O(N) rerooting identity  S(v) = S(parent) + w(v,parent) * (N - 2*size(v))  for the sums of
// weighted tree distances S(v) = sum_x d_w(v, x); closeness(v) = (N-1)/S(v).
//   1. subtree sizes and tour positions from the Euler tour of the tree rooted at the Prim start (euler_tour_rank);
//   2. wd(v) (weighted depth from the Prim start) and T(v) = S(v) - S(start): pointer jumping, log2 N rounds;
//   3. root = first argmin closeness; re-root (the path root -> start = the vertices whose tour interval contains
//      the root's); hop depths by pointer jumping; eta; optional DFS order from a second tour.
// No step is a sequential walk over the tree, so the cost does not depend on its depth.
__global__ void __launch_bounds__(kTreeThreads, 1) tree_kernel(TreeParams p) {
  extern __shared__ int tree_sm[];
  __shared__ double sred_d[32];
  __shared__ int sred_i[32];
  __shared__ long long sred_l[32];
  __shared__ int s_root;
  __shared__ double s_tot;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int N = p.N;
  const int64_t gofs = (int64_t)blockIdx.x * p.ldv;
  const int64_t oofs = (int64_t)blockIdx.x * p.ldp;
  const int* __restrict__ parent = p.parent + gofs;
  const double* __restrict__ weight = p.weight + gofs;
  const int* __restrict__ order = p.order + gofs;
  int* size = p.size + gofs;
  int* newparent = p.newparent + oofs;
  const int r0 = 0;   // Prim start vertex

  int* par_s = p.use_smem ? tree_sm : const_cast<int*>(parent);
  int* size_s = p.use_smem ? tree_sm + N : size;
  __shared__ int s_scan[1025];
  for (int v = tid; v < N; v += kTreeThreads)
    if (p.use_smem) par_s[v] = parent[v];
  if (p.mst_i) {
    for (int t = 1 + tid; t < N; t += kTreeThreads) {
      const int v = order[t], q = parent[v];
      p.mst_i[oofs + t - 1] = min(v, q);
      p.mst_j[oofs + t - 1] = max(v, q);
      p.mst_w[oofs + t - 1] = weight[v];
    }
  }
  __syncthreads();

  // 1. subtree sizes + position of every vertex's down arc in the tour from r0 (kept in newparent[] until the
  //    re-rooting, which is where newparent gets its real contents)
  int* pos_down = newparent;
  if (N > 1) {
    int* clist = p.anc1 + gofs;
    const int* rk = euler_tour_rank(N, r0, par_s, false, p.anc0 + gofs, clist, p.dep0 + gofs, p.dep1 + gofs, nullptr,
                                    reinterpret_cast<int*>(p.wd + gofs), reinterpret_cast<int*>(p.wd2 + gofs),
                                    reinterpret_cast<int*>(p.S + gofs), reinterpret_cast<int*>(p.S2 + gofs), s_scan);
    const int A = 2 * (N - 1);
    for (int q = tid; q < N - 1; q += kTreeThreads) {
      const int c = clist[q];
      const int pd = A - 1 - rk[2 * q], pu = A - 1 - rk[2 * q + 1];
      size_s[c] = (pu - pd + 1) >> 1;
      pos_down[c] = pd;
    }
  }
  if (tid == 0) { size_s[r0] = N; pos_down[r0] = -1; }
  __syncthreads();

  // 2. wd and T by pointer jumping: after round j, (a, b)[v] hold the sums over the first 2^j path edges
  double *wa = p.wd + gofs, *wb = p.wd2 + gofs, *ta = p.S + gofs, *tb = p.S2 + gofs;
  int *anc_a = p.anc0 + gofs, *anc_b = p.anc1 + gofs;
  for (int v = tid; v < N; v += kTreeThreads) {
    const double w = (v == r0) ? 0.0 : weight[v];
    wa[v] = w;
    ta[v] = (v == r0) ? 0.0 : w * (double)(N - 2 * size_s[v]);
    anc_a[v] = par_s[v];
  }
  __syncthreads();
  for (int span = 1; span < N; span <<= 1) {
    for (int v = tid; v < N; v += kTreeThreads) {
      const int a = anc_a[v];
      wb[v] = wa[v] + wa[a];
      tb[v] = ta[v] + ta[a];
      anc_b[v] = anc_a[a];
    }
    __syncthreads();
    double* td = wa; wa = wb; wb = td;
    td = ta; ta = tb; tb = td;
    int* ti = anc_a; anc_a = anc_b; anc_b = ti;
  }
  {                                                // S(r0) = sum_v wd(v)
    double acc = 0.0;
    for (int v = tid; v < N; v += kTreeThreads) acc += wa[v];
    acc = warp_sum(acc);
    if (lane == 0) sred_d[warp] = acc;
    __syncthreads();
    if (tid == 0) {
      double tot = 0.0;
      for (int w = 0; w < kTreeThreads / 32; ++w) tot += sred_d[w];
      s_tot = tot;
    }
    __syncthreads();
  }
  const double S0 = s_tot;
  const double* __restrict__ T = ta;

  // root = first argmin of closeness (N-1)/S(v)   (LightGraphs.closeness_centrality + findmin)
  {
    double bk = CUDART_INF;
    int bv = 0x7fffffff;
    for (int v = tid; v < N; v += kTreeThreads) {
      const double c = (N > 1) ? (double)(N - 1) / (S0 + T[v]) : 0.0;
      if (c < bk || bv == 0x7fffffff) { bk = c; bv = v; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const double ok = __shfl_xor_sync(0xffffffffu, bk, o);
      const int ov = __shfl_xor_sync(0xffffffffu, bv, o);
      argmin_combine(bk, bv, ok, ov);
    }
    if (lane == 0) { sred_d[warp] = bk; sred_i[warp] = bv; }
    __syncthreads();
    if (warp == 0) {
      bk = sred_d[lane];
      bv = sred_i[lane];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const double ok = __shfl_xor_sync(0xffffffffu, bk, o);
        const int ov = __shfl_xor_sync(0xffffffffu, bv, o);
        argmin_combine(bk, bv, ok, ov);
      }
      if (lane == 0) s_root = bv;
    }
    __syncthreads();
  }
  const int root = s_root;

  // re-root: reverse the parent pointers on the path root -> r0.  x lies on it iff the root's down arc falls inside
  // x's tour interval [pos_down(x), pos_down(x) + 2 size(x) - 1]  (r0 always does).
  {
    const int pr = (root == r0) ? -1 : pos_down[root];
    bool onpath[(16384 + kTreeThreads - 1) / kTreeThreads];
    __syncthreads();
    if (N <= 16384) {
      int k = 0;
      for (int v = tid; v < N; v += kTreeThreads, ++k) {
        const int pd = pos_down[v];
        onpath[k] = (v != r0) && (root != r0) && pd <= pr && pr <= pd + 2 * size_s[v] - 1;
      }
      __syncthreads();                                  // every pos_down has been read: newparent gets its contents
      for (int v = tid; v < N; v += kTreeThreads) newparent[v] = par_s[v];
      __syncthreads();
      k = 0;
      for (int v = tid; v < N; v += kTreeThreads, ++k)
        if (onpath[k]) newparent[par_s[v]] = v;
    } else {                                            // larger trees: the flags go through anc0 (free here)
      int* flag = p.anc0 + gofs;
      for (int v = tid; v < N; v += kTreeThreads) {
        const int pd = pos_down[v];
        flag[v] = (v != r0) && (root != r0) && pd <= pr && pr <= pd + 2 * size_s[v] - 1;
      }
      __syncthreads();
      for (int v = tid; v < N; v += kTreeThreads) newparent[v] = par_s[v];
      __syncthreads();
      for (int v = tid; v < N; v += kTreeThreads)
        if (flag[v]) newparent[par_s[v]] = v;
    }
    if (tid == 0) newparent[root] = root;
  }
  __syncthreads();

  // hop depths from the new root by pointer jumping (ceil(log2 N) rounds)
  anc_a = p.anc0 + gofs;
  anc_b = p.anc1 + gofs;
  int* dep_a = p.dep0 + gofs;
  int* dep_b = p.dep1 + gofs;
  for (int v = tid; v < N; v += kTreeThreads) {
    anc_a[v] = newparent[v];
    dep_a[v] = (v == root) ? 0 : 1;
  }
  __syncthreads();
  for (int span = 1; span < N; span <<= 1) {
    for (int v = tid; v < N; v += kTreeThreads) {
      const int a = anc_a[v];
      dep_b[v] = dep_a[v] + dep_a[a];
      anc_b[v] = anc_a[a];
    }
    __syncthreads();
    int* ti = anc_a; anc_a = anc_b; anc_b = ti;
    ti = dep_a; dep_a = dep_b; dep_b = ti;
  }

  // eta = mean(h) / (mean count per level / 2) + 1     (src/sequencer.jl:334-354)
  {
    long long sh = 0;
    int mx = 0;
    for (int v = tid; v < N; v += kTreeThreads) { sh += dep_a[v]; mx = max(mx, dep_a[v]); }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      sh += __shfl_xor_sync(0xffffffffu, sh, o);
      mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    }
    if (lane == 0) { sred_l[warp] = sh; sred_i[warp] = mx; }
    __syncthreads();
    if (tid == 0) {
      long long tot = 0;
      int m = 0;
      for (int w = 0; w < kTreeThreads / 32; ++w) { tot += sred_l[w]; m = max(m, sred_i[w]); }
      double hlen = (double)tot / (double)N;
      double hwid = ((double)N / (double)(m + 1)) / 2.0;
      hlen = fmin(fmax(hlen, kJuliaEps), kFloatMax32);
      hwid = fmin(fmax(hwid, kJuliaEps), kFloatMax32);
      p.eta[blockIdx.x] = hlen / hwid + 1.0;
      p.root[blockIdx.x] = root;
    }
  }

  // order = reverse(DFS preorder, children ascending)   (src/utils.jl:15-24): the preorder number of a vertex is the
  // number of down arcs up to the one that enters it in the tour of the re-rooted tree with ascending children.
  if (p.dfs_order) {
    __syncthreads();
    if (N == 1) {
      if (tid == 0) p.dfs_order[oofs] = root;
      return;
    }
    const int M = N - 1, A = 2 * M;
    int* clist = anc_b;
    int* b2 = reinterpret_cast<int*>(p.S + gofs);
    int* b3 = reinterpret_cast<int*>(p.S2 + gofs);
    const int* rk = euler_tour_rank(N, root, newparent, true, size, clist, dep_a, dep_b, anc_a,
                                    reinterpret_cast<int*>(p.wd + gofs), reinterpret_cast<int*>(p.wd2 + gofs), b2, b3, s_scan);
    int* flag = (rk == b3) ? b2 : b3;                   // a buffer the ranking no longer needs: 1 at the down-arc positions
    for (int a = tid; a < A; a += kTreeThreads) flag[A - 1 - rk[a]] = (a & 1) ? 0 : 1;
    __syncthreads();
    block_excl_scan_i32(flag, A, s_scan);               // flag[pos] = down arcs before pos = preorder number - 1
    for (int q = tid; q < M; q += kTreeThreads) {
      const int pre = flag[A - 1 - rk[2 * q]] + 1;      // 1 .. N-1; the root is 0
      p.dfs_order[oofs + (N - 1 - pre)] = clist[q];
    }
    if (tid == 0) p.dfs_order[oofs + N - 1] = root;
  }
}

// ---- eta-weighted combine ---------------------------------------------------------------------
// S = (first ? 0 : S) + sum_c clamp(eta_c * D_c, eps(), Inf)  in chunk order; if last: S /= sum(eta_all)
// (reference src/sequencer.jl:231-235,247; the sequential order keeps the rounding of the reference).
// HBM-bound: 8*(nb + 1 [+1]) bytes per entry.  One double2 per thread, nb independent loads in flight.
constexpr int kCombineMax = 32;
__global__ void __launch_bounds__(256) combine_kernel(double* __restrict__ S, const double* __restrict__ Db,
                                                      int64_t strideD, int nb, const double* __restrict__ eta_b,
                                                      int first, int last, const double* __restrict__ eta_all,
                                                      int n_all, int64_t total2 /* number of double2 */) {
  __shared__ double e[kCombineMax];
  __shared__ double s_tot;
  if (threadIdx.x < kCombineMax) e[threadIdx.x] = (threadIdx.x < nb) ? eta_b[threadIdx.x] : 0.0;
  if (threadIdx.x == 0) {
    double tot = 0.0;
    if (last)
      for (int c = 0; c < n_all; ++c) tot += eta_all[c];
    s_tot = tot;
  }
  __syncthreads();
  const double tot = s_tot;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total2;
       idx += (int64_t)gridDim.x * blockDim.x) {
    double2 s = first ? make_double2(0.0, 0.0) : reinterpret_cast<const double2*>(S)[idx];
    for (int c0 = 0; c0 < nb; c0 += 8) {
      double2 d[8];
#pragma unroll
      for (int c = 0; c < 8; ++c)
        if (c0 + c < nb) d[c] = ldg_stream_f64x2(Db + (int64_t)(c0 + c) * strideD + 2 * idx);
#pragma unroll
      for (int c = 0; c < 8; ++c)
        if (c0 + c < nb) {
          s.x += fmax(e[c0 + c] * d[c].x, kJuliaEps);
          s.y += fmax(e[c0 + c] * d[c].y, kJuliaEps);
        }
    }
    if (last) { s.x /= tot; s.y /= tot; }
    reinterpret_cast<double2*>(S)[idx] = s;
  }
}

// Same accumulation for the pipeline, where every chunk matrix is symmetric bit for bit (the distance, derive and
// fix-up kernels store each value and its mirror image): only the upper-triangle 64 x 64 tiles are read and
// accumulated, the lower ones are written as their transposes through shared memory.  HBM traffic per entry pair
// drops from 2*8*(nb + 1 [+1]) to 8*(nb + 2 [+1]) bytes, i.e. nearly half.  Grid (T, T), T = ceil(N / 64); the
// CTAs below the diagonal exit at once.
constexpr int kCombineTile = 64;
__global__ void __launch_bounds__(256) combine_sym_kernel(double* __restrict__ S, int64_t ldD, int N,
                                                          const double* __restrict__ Db, int64_t strideD, int nb,
                                                          const double* __restrict__ eta_b, int first, int last,
                                                          const double* __restrict__ eta_all, int n_all) {
  const int bi = blockIdx.y, bj = blockIdx.x;
  if (bi > bj) return;
  __shared__ double e[kCombineMax];
  __shared__ double s_tot;
  __shared__ double tile[kCombineTile][kCombineTile + 1];
  if (threadIdx.x < kCombineMax) e[threadIdx.x] = (threadIdx.x < nb) ? eta_b[threadIdx.x] : 0.0;
  if (threadIdx.x == 0) {
    double tot = 0.0;
    if (last)
      for (int c = 0; c < n_all; ++c) tot += eta_all[c];
    s_tot = tot;
  }
  __syncthreads();
  const double tot = s_tot;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int j = bj * kCombineTile + 2 * tx;                     // even; j + 1 < ldD (ldD is even and >= N)
#pragma unroll 1
  for (int pass = 0; pass < kCombineTile / 8; ++pass) {
    const int il = pass * 8 + ty, i = bi * kCombineTile + il;
    if (i < N && j < N) {
      const int64_t off = (int64_t)i * ldD + j;
      double2 s = first ? make_double2(0.0, 0.0) : *reinterpret_cast<const double2*>(S + off);
      for (int c0 = 0; c0 < nb; c0 += 8) {
        double2 d[8];
#pragma unroll
        for (int c = 0; c < 8; ++c)
          if (c0 + c < nb) d[c] = ldg_stream_f64x2(Db + (int64_t)(c0 + c) * strideD + off);
#pragma unroll
        for (int c = 0; c < 8; ++c)
          if (c0 + c < nb) {
            s.x += fmax(e[c0 + c] * d[c].x, kJuliaEps);
            s.y += fmax(e[c0 + c] * d[c].y, kJuliaEps);
          }
      }
      if (last) { s.x /= tot; s.y /= tot; }
      *reinterpret_cast<double2*>(S + off) = s;
      tile[il][2 * tx] = s.x;
      tile[il][2 * tx + 1] = s.y;
    }
  }
  if (bi == bj) return;                                         // a diagonal tile holds both halves already
  __syncthreads();
  const int ii = bi * kCombineTile + 2 * tx;
#pragma unroll 1
  for (int pass = 0; pass < kCombineTile / 8; ++pass) {
    const int jl = pass * 8 + ty, jj = bj * kCombineTile + jl;
    if (jj < N && ii < N) {
      double* dst = S + (int64_t)jj * ldD + ii;
      if (ii + 1 < N) *reinterpret_cast<double2*>(dst) = make_double2(tile[2 * tx][jl], tile[2 * tx + 1][jl]);
      else *dst = tile[2 * tx][jl];
    }
  }
}

// S /= tot, applied by the owner of a group after the partial sums have been reduced (chunk-major sharding)
__global__ void __launch_bounds__(256) scale_div_kernel(double* __restrict__ S, double tot, int64_t total) {
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (int64_t)gridDim.x * blockDim.x)
    S[idx] = S[idx] / tot;
}

// chunk etas from task (wave) order into (group, chunk) order, where the combine reads them: dst[pos[t]] = src[t]
__global__ void scatter_eta_kernel(const double* __restrict__ src, const int* __restrict__ pos, int n, double* __restrict__ dst) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t < n) dst[pos[t]] = src[t];
}

// Diagnostic sampling of matrix entries (seqj_set_probes): out[idx[q]] = base[off[q]]
__global__ void probe_gather_kernel(const double* __restrict__ base, const long long* __restrict__ off,
                                    const int* __restrict__ idx, int n, double* __restrict__ out) {
  const int q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q < n) out[idx[q]] = base[off[q]];
}

// Device-side exchange format of a group's result (seqj_sequence_shard / seqj_fuse_dev): one slot of 2 N doubles per
// group = { eta, w[N-1], (i, j) int32 pairs [N-1], pad }.  pack: slot q <- group arrays row perm[q]; unpack: group
// arrays row g <- slot slot_of[g].  Grid (ceil(N / 256), groups).
__global__ void pack_groups_kernel(double* __restrict__ pack, int64_t stride, const int* __restrict__ perm,
                                   const double* __restrict__ eta, const int* __restrict__ mi, const int* __restrict__ mj,
                                   const double* __restrict__ mw, int64_t ldp, int N) {
  const int q = blockIdx.y, e = blockIdx.x * blockDim.x + threadIdx.x;
  const int g = perm[q];
  double* slot = pack + (int64_t)q * stride;
  if (e == 0) { slot[0] = eta[g]; slot[2 * N - 1] = 0.0; }
  if (e < N - 1) {
    slot[1 + e] = mw[(int64_t)g * ldp + e];
    reinterpret_cast<int2*>(slot + N)[e] = make_int2(mi[(int64_t)g * ldp + e], mj[(int64_t)g * ldp + e]);
  }
}
__global__ void unpack_groups_kernel(const double* __restrict__ pack, int64_t stride, const int* __restrict__ slot_of,
                                     double* __restrict__ eta, int* __restrict__ mi, int* __restrict__ mj,
                                     double* __restrict__ mw, int64_t ldp, int N) {
  const int g = blockIdx.y, e = blockIdx.x * blockDim.x + threadIdx.x;
  const double* slot = pack + (int64_t)slot_of[g] * stride;
  if (e == 0) eta[g] = slot[0];
  if (e < N - 1) {
    mw[(int64_t)g * ldp + e] = slot[1 + e];
    const int2 ij = reinterpret_cast<const int2*>(slot + N)[e];
    mi[(int64_t)g * ldp + e] = ij.x;
    mj[(int64_t)g * ldp + e] = ij.y;
  }
}

// ---- proximity fuse ---------------------------------------------------------------------------
// P[i,j] = P[j,i] += eta * inv(w) for every MST edge of one group (src/sequencer.jl:313-322).
__global__ void fuse_scatter_kernel(double* __restrict__ P, int64_t ldD, const int* __restrict__ mi,
                                    const int* __restrict__ mj, const double* __restrict__ mw,
                                    const double* __restrict__ eta, int nedges) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= nedges) return;
  const int i = mi[e], j = mj[e];
  const double t = eta[0] * (1.0 / mw[e]);
  const double v = P[(int64_t)j * ldD + i] + t;
  P[(int64_t)j * ldD + i] = v;
  P[(int64_t)i * ldD + j] = v;
}

// D = inv.(P / sum(eta)) with +Inf diagonal proximity (=> 0 -> clamped to eps by _mst); non-edges stay +Inf
// (src/sequencer.jl:324-326,269,360).
__global__ void fuse_finalize_kernel(double* __restrict__ P, int64_t ldD, int N, const double* __restrict__ eta_g,
                                     const int* __restrict__ perm, int ngroups, double eps_floor) {
  double tot = 0.0;      // sum(eta_all) in the reference's (metric-major) group order
  for (int g = 0; g < ngroups; ++g) tot += eta_g[perm ? perm[g] : g];
  const int64_t total = (int64_t)N * ldD;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (int64_t)gridDim.x * blockDim.x) {
    const int i = (int)(idx / ldD), j = (int)(idx % ldD);
    if (j >= N) continue;
    const double pv = P[idx];
    double out;
    if (i == j) out = eps_floor;
    else if (pv != 0.0) out = fmax(1.0 / (pv / tot), eps_floor);
    else out = CUDART_INF;
    P[idx] = out;
  }
}

__global__ void gather_edges_kernel(const double* __restrict__ D, int64_t ldD, const int* __restrict__ mi,
                                    const int* __restrict__ mj, double* __restrict__ val, int nedges) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= nedges) return;
  val[e] = D[(int64_t)mi[e] * ldD + mj[e]];
}

// ---- sparse MST for the final graph ---------------------------------------------------------------
// The final distance matrix D = inv.(Pw) (reference src/sequencer.jl:269) is +Inf outside the union of the
// group MSTs, i.e. a sparse graph with at most G*(N-1) edges.  Kruskal over the dense N x N matrix (reference)
// or a dense Prim (N latency-bound steps) is wasteful here: Boruvka on the edge list needs O(log N) fully
// parallel rounds.  Edges are totally ordered by (weight, max(i,j), min(i,j)) -- the order Kruskal visits them in
// the reference -- so the result is the unique MST under that order (= the reference's tree even with tied
// weights) and the only hooking cycles are mutual pairs, broken by component id.
// One CTA; edge e = (g, k) lives at mi/mj/val[g*ldp + k].
struct BoruvkaParams {
  int N, G;
  int64_t ldp;
  const int* mi;
  const int* mj;
  const double* val;        // D[i][j] per edge (duplicates carry identical values)
  double eps_floor;
  int* comp;                // scratch [N]
  unsigned long long* best; // scratch [N]
  unsigned long long* bestp; // scratch [N]: (max(i,j) << 32 | min(i,j)) of the chosen edge
  int* out_i;               // MST edges [N-1]
  int* out_j;
  double* out_w;
  int* out_count;           // scratch [1]
};

constexpr int kBoruvkaThreads = 1024;

__global__ void __launch_bounds__(kBoruvkaThreads, 1) boruvka_sparse_kernel(BoruvkaParams p) {
  __shared__ int s_ncomp;
  const int tid = threadIdx.x;
  const int N = p.N, ne = N - 1;
  const long long E = (long long)p.G * ne;
  const unsigned long long eps_bits = (unsigned long long)__double_as_longlong(p.eps_floor);
  for (int v = tid; v < N; v += kBoruvkaThreads) p.comp[v] = v;
  if (tid == 0) { *p.out_count = 0; s_ncomp = N; }
  __syncthreads();
  for (int round = 0; round < 64 && s_ncomp > 1; ++round) {
    for (int v = tid; v < N; v += kBoruvkaThreads) { p.best[v] = ~0ull; p.bestp[v] = ~0ull; }
    __syncthreads();
    // 1. lightest crossing edge weight of every component
    for (long long e = tid; e < E; e += kBoruvkaThreads) {
      const int64_t a = (e / ne) * p.ldp + (e % ne);
      const int ci = p.comp[p.mi[a]], cj = p.comp[p.mj[a]];
      if (ci != cj) {
        const unsigned long long w = max((unsigned long long)__double_as_longlong(p.val[a]), eps_bits);
        atomicMin(&p.best[ci], w);
        atomicMin(&p.best[cj], w);
      }
    }
    __syncthreads();
    // 2. among those, the first vertex pair in Kruskal's order
    for (long long e = tid; e < E; e += kBoruvkaThreads) {
      const int64_t a = (e / ne) * p.ldp + (e % ne);
      const int ci = p.comp[p.mi[a]], cj = p.comp[p.mj[a]];
      if (ci != cj) {
        const unsigned long long w = max((unsigned long long)__double_as_longlong(p.val[a]), eps_bits);
        const int i = p.mi[a], j = p.mj[a];
        const unsigned long long pr = ((unsigned long long)(unsigned)max(i, j) << 32) | (unsigned)min(i, j);
        if (w == p.best[ci]) atomicMin(&p.bestp[ci], pr);
        if (w == p.best[cj]) atomicMin(&p.bestp[cj], pr);
      }
    }
    __syncthreads();
    // 3. hook every component root onto the other side of its edge; a mutual pair keeps the smaller id as root
    for (int c = tid; c < N; c += kBoruvkaThreads) {
      if (p.comp[c] != c) continue;
      const unsigned long long pr = p.bestp[c];
      if (pr == ~0ull) continue;                          // isolated component (disconnected input)
      const int i = (int)(unsigned)(pr & 0xffffffffull), j = (int)(unsigned)(pr >> 32);
      const int ci = p.comp[i], cj = p.comp[j];
      const int other = (ci == c) ? cj : ci;
      const bool mutual = (p.bestp[other] == pr);
      if (mutual && c < other) continue;                  // `other` hooks onto c and records the edge
      const int k = atomicAdd(p.out_count, 1);
      p.out_i[k] = i;
      p.out_j[k] = j;
      p.out_w[k] = __longlong_as_double((long long)p.best[c]);
      // bestp[] is read by other components' mutual test in this step, so only this component's own best[c] is
      // reused: bit 63 (never set in a finite weight) marks "hooked", the low word is the hook target
      p.best[c] = (1ull << 63) | (unsigned long long)(unsigned)other;
    }
    __syncthreads();
    for (int c = tid; c < N; c += kBoruvkaThreads)
      if (p.comp[c] == c && p.bestp[c] != ~0ull && (p.best[c] >> 63)) p.comp[c] = (int)(p.best[c] & 0xffffffffull);
    __syncthreads();
    // 4. pointer jumping until every vertex points at its root
    for (int it = 0; it < 32; ++it) {
      int changed = 0;
      for (int v = tid; v < N; v += kBoruvkaThreads) {
        const int c = p.comp[v], cc = p.comp[c];
        if (c != cc) { p.comp[v] = cc; changed = 1; }
      }
      if (!__syncthreads_or(changed)) break;
    }
    if (tid == 0) s_ncomp = N - *p.out_count;
    __syncthreads();
  }
}

// The same Boruvka on a thread-block CLUSTER of 8 CTAs (8 x 1024 threads on 8 SMs): the rounds are latency-bound scans
// of the <= G (N-1) edges with two label look-ups each, so eight times the threads in flight is nearly eight times the
// speed; all state lives in global memory (L2), CTA-wide barriers become cluster barriers (release / acquire at cluster
// scope) and the early exit of the pointer jumping reads a monotonic change counter instead of __syncthreads_or.
// Used for N >= 2048 (1.8 -> ~0.4 ms at N = 10,000, 20 groups); the single-CTA kernel stays for small graphs.
constexpr int kBoruvkaCluster = 8;
__global__ void __cluster_dims__(kBoruvkaCluster, 1, 1) __launch_bounds__(kBoruvkaThreads, 1)
boruvka_sparse_cluster_kernel(BoruvkaParams p) {
  const int tid = (int)cluster_ctarank() * kBoruvkaThreads + threadIdx.x;
  constexpr int NT = kBoruvkaCluster * kBoruvkaThreads;
  const int N = p.N, ne = N - 1;
  const unsigned long long eps_bits = (unsigned long long)__double_as_longlong(p.eps_floor);
  int* chg = p.out_count + 1;                            // monotonic "some label changed" counter
  // Labels and candidates are written by one CTA and read by the others between barriers: every such read bypasses the
  // (per-SM, non-coherent) L1 with ld.global.cg; the barrier is release / acquire at cluster scope after a fence.
  auto sync = [&]() { __threadfence(); cluster_sync_all(); };
  auto comp_of = [&](int v) { return __ldcg(p.comp + v); };
  for (int v = tid; v < N; v += NT) p.comp[v] = v;
  if (tid == 0) { p.out_count[0] = 0; *chg = 0; }
  sync();
  int seen = 0;
  for (int round = 0; round < 64; ++round) {
    // uniform: out_count is only written two barriers further down, so every thread of the cluster reads the same value
    if (N - __ldcg(p.out_count) <= 1) break;
    for (int v = tid; v < N; v += NT) { p.best[v] = ~0ull; p.bestp[v] = ~0ull; }
    sync();
    for (int g = 0; g < p.G; ++g)
      for (int k = tid; k < ne; k += NT) {
        const int64_t a = (int64_t)g * p.ldp + k;
        const int ci = comp_of(p.mi[a]), cj = comp_of(p.mj[a]);
        if (ci != cj) {
          const unsigned long long w = max((unsigned long long)__double_as_longlong(p.val[a]), eps_bits);
          atomicMin(&p.best[ci], w);
          atomicMin(&p.best[cj], w);
        }
      }
    sync();
    for (int g = 0; g < p.G; ++g)
      for (int k = tid; k < ne; k += NT) {
        const int64_t a = (int64_t)g * p.ldp + k;
        const int i = p.mi[a], j = p.mj[a];
        const int ci = comp_of(i), cj = comp_of(j);
        if (ci != cj) {
          const unsigned long long w = max((unsigned long long)__double_as_longlong(p.val[a]), eps_bits);
          const unsigned long long pr = ((unsigned long long)(unsigned)max(i, j) << 32) | (unsigned)min(i, j);
          if (w == __ldcg(p.best + ci)) atomicMin(&p.bestp[ci], pr);
          if (w == __ldcg(p.best + cj)) atomicMin(&p.bestp[cj], pr);
        }
      }
    sync();
    for (int c = tid; c < N; c += NT) {
      if (comp_of(c) != c) continue;
      const unsigned long long pr = __ldcg(p.bestp + c);
      if (pr == ~0ull) continue;
      const int i = (int)(unsigned)(pr & 0xffffffffull), j = (int)(unsigned)(pr >> 32);
      const int ci = comp_of(i), cj = comp_of(j);
      const int other = (ci == c) ? cj : ci;
      const bool mutual = (__ldcg(p.bestp + other) == pr);
      if (mutual && c < other) continue;
      const int k = atomicAdd(p.out_count, 1);
      p.out_i[k] = i;
      p.out_j[k] = j;
      p.out_w[k] = __longlong_as_double((long long)__ldcg(p.best + c));
      p.best[c] = (1ull << 63) | (unsigned long long)(unsigned)other;      // read back only by this thread (below)
    }
    sync();
    for (int c = tid; c < N; c += NT)
      if (comp_of(c) == c && __ldcg(p.bestp + c) != ~0ull && (__ldcg(p.best + c) >> 63))
        p.comp[c] = (int)(__ldcg(p.best + c) & 0xffffffffull);
    sync();
    for (int it = 0; it < 32; ++it) {
      int changed = 0;
      for (int v = tid; v < N; v += NT) {
        // a label may be shortened by another thread during this pass: both values lie on the same chain towards
        // the root, so the number of passes may vary but the fixed point (every vertex -> its root) does not
        const int c = comp_of(v), cc = comp_of(c);
        if (c != cc) { p.comp[v] = cc; changed = 1; }
      }
      if (changed) atomicAdd(chg, 1);
      sync();
      const int now = __ldcg(chg);
      sync();                                              // nobody bumps the counter again before everybody has read it
      if (now == seen) break;                              // uniform across the cluster
      seen = now;
    }
  }
}

// Root the MST edge list at vertex 0: produce the (parent, weight, order) arrays that tree_kernel consumes
// (the same contract as Prim's output: parents precede children in `order`).  The final tree is path-like
// (depth ~ N), so a BFS is N dependent steps of one thread (1.2 ms at N = 10k, 16 ms at N = 20k where its arrays
// no longer fit in shared memory).  Instead: Euler tour + list ranking, O(log N) parallel rounds whatever the
// shape of the tree.  Every edge becomes two arcs stored in CSR order; the successor of arc (u -> v) is the arc
// after (v -> u) in v's list (cyclically), which links all 2(N-1) arcs into one tour; cutting it at the root's first
// arc and ranking the list by pointer jumping gives every arc its position.  Arc (u -> v) is a tree edge
// parent -> child iff it comes before its twin; the child's preorder number is the count of parent -> child arcs
// up to it.  Integers only, so the result is exact.
struct RootTreeParams {
  int N;
  const int* ei;
  const int* ej;
  const double* ew;
  int* scratch;  // global scratch, 16 N + 64 ints
  int* parent;   // out [N]
  double* weight;
  int* order;
  // batched use (one CTA per graph, kernels_mst.cuh): strides of the edge lists, the scratch and the outputs, and an
  // optional per-graph flag (only graphs with done == 1 are rooted)
  int64_t ld_e = 0, ld_s = 0, ld_o = 0;
  const int* done = nullptr;
};

__global__ void __launch_bounds__(1024, 1) root_tree_kernel(RootTreeParams p) {
  __shared__ int s_part[1025];
  const int tid = threadIdx.x, N = p.N, M = N - 1, A = 2 * M;
  if (p.done && p.done[blockIdx.x] != 1) return;
  p.ei += blockIdx.x * p.ld_e; p.ej += blockIdx.x * p.ld_e; p.ew += blockIdx.x * p.ld_e;
  p.scratch += blockIdx.x * p.ld_s;
  p.parent += blockIdx.x * p.ld_o; p.weight += blockIdx.x * p.ld_o; p.order += blockIdx.x * p.ld_o;
  if (N == 1) {
    if (tid == 0) { p.parent[0] = 0; p.weight[0] = 0.0; p.order[0] = 0; }
    return;
  }
  int* off = p.scratch;              // [N + 1] CSR offsets
  int* fill = off + (N + 1);         // [N]     fill cursors
  int* adjv = fill + N;              // [A]     head vertex of every arc
  int* twin = adjv + A;              // [A]     index of the reverse arc
  int* eidx = twin + A;              // [A]     edge of the arc
  int* succ0 = eidx + A;             // [A] x 4 ping-pong buffers of the list ranking
  int* rank0 = succ0 + A;
  int* succ1 = rank0 + A;
  int* rank1 = succ1 + A;
  for (int v = tid; v <= N; v += 1024) off[v] = 0;
  __syncthreads();
  for (int e = tid; e < M; e += 1024) { atomicAdd(&off[p.ei[e]], 1); atomicAdd(&off[p.ej[e]], 1); }
  __syncthreads();
  block_excl_scan_i32(off, N, s_part);
  if (tid == 0) off[N] = A;
  for (int v = tid; v < N; v += 1024) fill[v] = off[v];
  __syncthreads();
  for (int e = tid; e < M; e += 1024) {
    const int i = p.ei[e], j = p.ej[e];
    const int pi = atomicAdd(&fill[i], 1), pj = atomicAdd(&fill[j], 1);
    adjv[pi] = j; adjv[pj] = i;
    twin[pi] = pj; twin[pj] = pi;
    eidx[pi] = e; eidx[pj] = e;
  }
  __syncthreads();
  // Euler tour: succ(u -> v) = the arc after (v -> u) in v's list; the tour starts at the root's first arc and ends
  // with the arc whose successor would wrap around to it
  const int last = twin[off[1] - 1];
  for (int a = tid; a < A; a += 1024) {
    const int v = adjv[a], t = twin[a];
    succ0[a] = (a == last) ? -1 : ((t + 1 < off[v + 1]) ? t + 1 : off[v]);
    rank0[a] = (a == last) ? 0 : 1;
  }
  __syncthreads();
  int *si = succ0, *ri = rank0, *so = succ1, *ro = rank1;
  for (int span = 1; span < A; span <<= 1) {     // rank = number of arcs after this one (Wyllie pointer jumping)
    for (int a = tid; a < A; a += 1024) {
      const int sa = si[a];
      if (sa >= 0) { ro[a] = ri[a] + ri[sa]; so[a] = si[sa]; }
      else { ro[a] = ri[a]; so[a] = -1; }
    }
    __syncthreads();
    int* t = si; si = so; so = t;
    t = ri; ri = ro; ro = t;
  }
  // position in the tour = A - 1 - rank; flag the parent -> child arcs by position and count them
  int* flag = so;                    // [A], free again
  for (int a = tid; a < A; a += 1024) {
    const int pa = A - 1 - ri[a], pt = A - 1 - ri[twin[a]];
    flag[pa] = (pa < pt) ? 1 : 0;
  }
  __syncthreads();
  block_excl_scan_i32(flag, A, s_part);          // flag[pos] = number of parent -> child arcs before position pos
  for (int a = tid; a < A; a += 1024) {
    const int pa = A - 1 - ri[a], pt = A - 1 - ri[twin[a]];
    if (pa < pt) {
      const int v = adjv[a], u = adjv[twin[a]];
      p.parent[v] = u;
      p.weight[v] = p.ew[eidx[a]];
      p.order[flag[pa] + 1] = v;                 // preorder: parents before children
    }
  }
  if (tid == 0) { p.parent[0] = 0; p.weight[0] = 0.0; p.order[0] = 0; }
}

// Unit-level: Symmetric(dm) view of a column-major upload (upper triangle wins), clamped to eps.
__global__ void symmetrize_kernel(const double* __restrict__ up, int N, double* __restrict__ out, int64_t ldD,
                                  double eps_floor) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (int64_t)N * N) return;
  const int r = (int)(idx / N), c = (int)(idx % N);
  const int lo = min(r, c), hi = max(r, c);
  out[(int64_t)r * ldD + c] = fmax(up[(int64_t)hi * N + lo], eps_floor);   // D[lo, hi] column-major
}

__global__ void copy_strided_kernel(const double* __restrict__ in, int64_t ld_in, double* __restrict__ out,
                                    int64_t ld_out, int rows, int cols) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (int64_t)rows * cols) return;
  const int r = (int)(idx / cols), c = (int)(idx % cols);
  out[(int64_t)r * ld_out + c] = in[(int64_t)r * ld_in + c];
}

}  // namespace seqj
