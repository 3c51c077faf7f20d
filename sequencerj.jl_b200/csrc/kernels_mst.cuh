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

// Tail shared by the list kernels: from the lanes' sorted (k0..k3, j0..j3) to the row's list.
// (Bk, Bj): an additional limit up to which the caller can vouch for its candidates (kKnnNone = none).
template <bool RESCAN>
__device__ __forceinline__ void knn_select_write(const KnnParams& p, int64_t vo, int lane, unsigned long long k0,
                                                 unsigned long long k1, unsigned long long k2, unsigned long long k3,
                                                 int j0, int j1, int j2, int j3, int c0, int c1, int c2, int c3,
                                                 unsigned long long Bk, int Bj) {
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
  unsigned long long Lk = anyfull ? (((unsigned long long)mh << 32) | ml) : kKnnNone;
  int Lj = anyfull ? (int)mj : 0x7fffffff;
  if (Bk < Lk || (Bk == Lk && Bj < Lj)) { Lk = Bk; Lj = Bj; }
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
  if (lane < kKnnK) { p.lk[vo * kKnnK + lane] = myk; p.lj[vo * kKnnK + lane] = myj; }
  if (lane == 0) {
    p.cnt[vo] = count;
    p.ptr[vo] = 0;
    p.bound[vo] = (count == kKnnK) ? lastk : Lk;
  }
}

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
    knn_select_write<RESCAN>(p, gofs + v, lane, k0, k1, k2, k3, j0, j1, j2, j3, c0, c1, c2, c3, kKnnNone, 0x7fffffff);
  }
}

// ---- first list pass from the UPPER TRIANGLE only (round 2) ------------------------------------------------------
// knn_rows_kernel<false> reads every row in full: 8 N^2 bytes per graph although the matrices are symmetric.  Here:
//   0. knn_sample_kernel: one warp per row reads N/8 contiguous entries of it and keeps the third smallest as the
//      row's threshold tau_v.  About 3 * 8 = 24 entries of the whole row are then <= tau_v (negative binomial).
//   1. knn_tiles_kernel: the 128 x 128 blocks (I, J), I <= J, are read ONCE and each entry (i, j) is offered to both
//      ends: it joins the candidates of row i when it is <= tau_i and those of row j when it is <= tau_j.  A lane owns
//      four columns (their thresholds stay in registers), the row's threshold is one shuffle; the common path per row
//      of 128 entries is 2 loads, a shuffle and 4 min/compare pairs per lane.  Candidates are appended with one atomic
//      per hit (a few dozen per row).
//   2. knn_merge_kernel: sorts a row's candidates by (weight, j) -- the order of the atomics does not matter -- and
//      writes the list format of the row kernel: the 16 smallest, bound = the 16th weight or tau_v (EVERY entry
//      <= tau_v is a candidate, so whatever is unlisted weighs at least that).  A row with more than kKtCap candidates
//      (ties, tight clusters: P = 3e-4 on unstructured data) gets an empty list with bound 0: the first Boruvka round
//      puts it on the needy list and the rescan kernel reads it in full, like any exhausted row.
// Bytes per graph: 4 N^2 (+ diagonal blocks) + N^2 for the samples = 0.64 of the row kernel's.
constexpr int kKtB = 128;                 // block edge
constexpr int kKtWarps = 8;               // blocks per CTA
constexpr int kKtCap = 96;                // candidates kept per row
constexpr int kKtQueue = 96;              // hits a warp parks before it sends them to memory
constexpr int kKtSampleDiv = 8;           // sample = N / 8 entries of the row
constexpr int kKtSampleK = 3;             // tau = third smallest of the sample
struct __align__(16) KnnCand {
  unsigned long long k;                   // weight (bit pattern, eps-clamped)
  int j, pad;
};

struct KnnTileParams {
  const double* D;
  int64_t ldD, strideD;
  const int* mat_slot;
  int N, nblk, ntiles;
  int g0;                                 // first graph of this launch (blockIdx.y = graph - g0)
  double eps_floor;
  unsigned long long* tau;                // [graph - g0][N]
  int* cnt;                               // [graph - g0][N]   candidates offered so far (may exceed kKtCap)
  KnnCand* cand;                          // [graph - g0][N][kKtCap]
};

__device__ __forceinline__ bool knn_less(unsigned long long ka, int ja, unsigned long long kb, int jb) {
  return ka < kb || (ka == kb && ja < jb);
}

// tau_v = the kKtSampleK-th smallest of N / kKtSampleDiv contiguous entries of row v (diagonal excluded), starting
// at the row's own block so that neighbouring rows share sectors.  kKnnNone when the sample holds fewer entries.
__global__ void __launch_bounds__(kKnnWarps * 32) knn_sample_kernel(KnnTileParams p) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int N = p.N;
  const int g = p.g0 + blockIdx.y;
  const int slot = p.mat_slot ? p.mat_slot[g] : g;
  const double* __restrict__ Dm = p.D + (int64_t)slot * p.strideD;
  const unsigned long long eps_bits = (unsigned long long)__double_as_longlong(p.eps_floor);
  const int m = max(64, (N / kKtSampleDiv + 63) & ~63);     // multiple of 64: whole 16-byte pairs per lane
  for (int v = blockIdx.x * kKnnWarps + warp; v < N; v += gridDim.x * kKnnWarps) {
    int s0 = (v / kKtB) * kKtB;
    if (s0 + m > N) s0 = max(0, ((N - m) + 1) & ~1);        // even start: 16-byte aligned pairs (ldD % 16 == 0)
    const double* __restrict__ row = Dm + (int64_t)v * p.ldD;
    unsigned long long k0 = kKnnNone, k1 = kKnnNone, k2 = kKnnNone;
    const int send = min(N, s0 + m);
    for (int b4 = s0 + 2 * lane; b4 < send; b4 += 4 * 64) {
      ulonglong2 d4[4];
#pragma unroll
      for (int q = 0; q < 4; ++q)
        if (b4 + 64 * q < send) d4[q] = ldg_stream_u64x2(row + b4 + 64 * q);
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int base = b4 + 64 * q;
        if (base < send) {
#pragma unroll
          for (int e = 0; e < 2; ++e) {
            const int j = base + e;
            unsigned long long x = e ? d4[q].y : d4[q].x;
            if (j < N && j != v && x < k2) {
              x = max(x, eps_bits);
              if (x < k2) {
                if (x < k1) { k2 = k1; if (x < k0) { k1 = k0; k0 = x; } else k1 = x; }
                else k2 = x;
              }
            }
          }
        }
      }
    }
    unsigned long long mk = kKnnNone, first = kKnnNone;
#pragma unroll
    for (int q = 0; q < kKtSampleK; ++q) {                  // pop the warp minimum kKtSampleK times
      const unsigned hh = (unsigned)(k0 >> 32);
      const unsigned mhh = __reduce_min_sync(0xffffffffu, hh);
      const unsigned hl = (hh == mhh) ? (unsigned)k0 : 0xffffffffu;
      const unsigned mhl = __reduce_min_sync(0xffffffffu, hl);
      mk = ((unsigned long long)mhh << 32) | mhl;
      if (q == 0) first = mk;
      const unsigned owner = __ballot_sync(0xffffffffu, k0 == mk);
      if (mk != kKnnNone && lane == __ffs(owner) - 1) { k0 = k1; k1 = k2; k2 = kKnnNone; }
    }
    // The smallest sample entries all tie (duplicate objects, constant chunks): the row has an unknown number of
    // entries at its threshold, possibly thousands.  tau = 0 keeps it out of the tile pass altogether -- nothing is
    // <= 0 after the eps clamp -- and the merge hands it to the rescan kernel, which lists one edge per component.
    if (mk != kKnnNone && first == mk) mk = 0ull;
    if (lane == 0) p.tau[(int64_t)blockIdx.y * N + v] = mk;
  }
}

template <int U, int MINB>
__global__ void __launch_bounds__(kKtWarps * 32, MINB) knn_tiles_kernel(KnnTileParams p) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int t = blockIdx.x * kKtWarps + warp;
  if (t >= p.ntiles) return;
  // Blocks in row-major order of the upper triangle, t = I nblk - I (I - 1) / 2 + (J - I): the eight warps of a CTA
  // read neighbouring column blocks of the same rows (8 KB contiguous per row and CTA, the same few pages).
  const int nb2 = 2 * p.nblk + 1;
  int I = (int)(((float)nb2 - sqrtf((float)nb2 * (float)nb2 - 8.0f * (float)t)) * 0.5f);
  I = max(0, min(I, p.nblk - 1));
  while (I > 0 && I * p.nblk - I * (I - 1) / 2 > t) --I;
  while ((I + 1) * p.nblk - (I + 1) * I / 2 <= t) ++I;
  const int J = I + (t - (I * p.nblk - I * (I - 1) / 2));
  const int N = p.N;
  const int g = p.g0 + blockIdx.y;
  const int slot = p.mat_slot ? p.mat_slot[g] : g;
  const int i0 = I * kKtB, j0 = J * kKtB;
  const int nr = min(kKtB, N - i0);
  const bool diag = (I == J);
  // The lane's four columns are adjacent (32 contiguous bytes per row).  Measured and rejected: the interleaved mapping
  // (16-byte pairs lane and lane + 32, so that each load instruction covers 512 contiguous bytes): MST stage 21.9
  // instead of 20.0 ms; and loads written as conditional expressions instead of under one branch per row (the
  // compiler then waits for each volatile load before the next): 22.9 ms.
  const int jc = j0 + 4 * lane;                            // first of the lane's four columns
  const bool ld_ok = jc < N;                               // jc + 3 < ldD: ldD is a multiple of 16
  const double* __restrict__ base = p.D + (int64_t)slot * p.strideD + (int64_t)i0 * p.ldD + jc;
  const unsigned long long* __restrict__ tau = p.tau + (int64_t)blockIdx.y * N;
  int* __restrict__ cnt = p.cnt + (int64_t)blockIdx.y * N;
  KnnCand* __restrict__ cand = p.cand + (int64_t)blockIdx.y * N * kKtCap;
  const unsigned long long eps_bits = (unsigned long long)__double_as_longlong(p.eps_floor);

  unsigned long long tj[4];                                // thresholds of the lane's columns (0 = column not there)
  unsigned tjh[4];
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    tj[c] = (jc + c < N) ? tau[jc + c] : 0ull;
    tjh[c] = (unsigned)(tj[c] >> 32);
  }
  // Hits are rare (~5 per 1,000 entries) and everything about them is kept out of the streaming loop, which only
  // records WHICH of the U rows hit (one bit each): a first version handled hits in place, and the unrolled copies of
  // that code made the kernel wait for instruction fetches (ncu: stall_no_instruction 35 per issue, 1.3-2.5 TB/s).
  // The hit rows are then re-read (one sector, from L2) in a rolled loop.  A hit costs a shared-memory queue entry;
  // the atomics on the rows' counters -- a round trip to L2 each -- are sent off together once 32 have gathered.
  __shared__ unsigned long long s_ti[kKtWarps][kKtB];       // thresholds of the block's rows (0 = row not there)
  __shared__ KnnCand s_q[kKtWarps][kKtQueue];               // .pad carries the row the candidate belongs to
  __shared__ int s_qn[kKtWarps];
  for (int r = lane; r < kKtB; r += 32) s_ti[warp][r] = (r < nr) ? tau[i0 + r] : 0ull;
  if (lane == 0) s_qn[warp] = 0;
  __syncwarp();
  const unsigned* s_tih = reinterpret_cast<const unsigned*>(&s_ti[warp][0]);   // high word of row r: s_tih[2 r + 1]
  auto offer = [&](int v, unsigned long long x, int other) {
    const int q = atomicAdd(&s_qn[warp], 1);
    if (q < kKtQueue) { KnnCand e; e.k = x; e.j = other; e.pad = v; s_q[warp][q] = e; }
    else atomicAdd(&cnt[v], kKtCap + 1);                    // queue full (pathological input): the row is rescanned
  };
  auto flush = [&](int at_least) {
    __syncwarp();
    const int n = min(s_qn[warp], kKtQueue);
    if (n >= at_least) {                                    // warp-uniform
#pragma unroll 1
      for (int e = lane; e < n; e += 32) {
        const KnnCand c = s_q[warp][e];
        const int pos = atomicAdd(&cnt[c.pad], 1);
        if (pos < kKtCap) {
          KnnCand o; o.k = c.k; o.j = c.j; o.pad = 0;
          cand[(int64_t)c.pad * kKtCap + pos] = o;
        }
      }
      __syncwarp();
      if (lane == 0) s_qn[warp] = 0;
    }
    __syncwarp();
  };

  // the loads of the next U rows are in flight while the current ones are tested (two register sets)
  ulonglong2 na[U], nb[U];
  auto load_rows = [&](int r0) {
#pragma unroll
    for (int u = 0; u < U; ++u)
      if (ld_ok && r0 + u < nr) {
        na[u] = ldg_stream_u64x2(base + (int64_t)(r0 + u) * p.ldD);
        nb[u] = ldg_stream_u64x2(base + (int64_t)(r0 + u) * p.ldD + 2);
      } else {
        na[u] = nb[u] = make_ulonglong2(kKnnNone, kKnnNone);
      }
  };
  load_rows(0);
#pragma unroll 1
  for (int r0 = 0; r0 < nr; r0 += U) {
    ulonglong2 va[U], vb[U];
#pragma unroll
    for (int u = 0; u < U; ++u) { va[u] = na[u]; vb[u] = nb[u]; }
    if (r0 + U < nr) load_rows(r0 + U);
    unsigned hm = 0;
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const unsigned tih = s_tih[2 * (r0 + u) + 1];
      bool hit = (unsigned)(va[u].x >> 32) <= max(tih, tjh[0]);
      hit |= (unsigned)(va[u].y >> 32) <= max(tih, tjh[1]);
      hit |= (unsigned)(vb[u].x >> 32) <= max(tih, tjh[2]);
      hit |= (unsigned)(vb[u].y >> 32) <= max(tih, tjh[3]);
      hm |= (hit ? 1u : 0u) << u;
    }
    if (hm) {                                               // a few lanes per group of rows
#pragma unroll 1
      for (; hm; hm &= hm - 1) {
        const int r = r0 + __ffs(hm) - 1;
        const int i = i0 + r;
        if (r >= nr || !ld_ok) continue;
        const unsigned long long tir = s_ti[warp][r];
        const unsigned long long* src = reinterpret_cast<const unsigned long long*>(base + (int64_t)r * p.ldD);
#pragma unroll 1
        for (int c = 0; c < 4; ++c) {
          const int j = jc + c;
          if (j >= N || j == i) continue;
          const unsigned long long xc = __ldg(src + c);
          const unsigned long long xv = max(xc, eps_bits);
          // the diagonal block holds (i, j) and (j, i): each entry is offered to its row only
          if (xv <= tir) offer(i, xv, j);
          const unsigned long long tjc = (c == 0) ? tj[0] : (c == 1) ? tj[1] : (c == 2) ? tj[2] : tj[3];
          if (!diag && xv <= tjc) offer(j, xv, i);
        }
      }
    }
    flush(32);
  }
  flush(1);
}

// One warp per object: its candidates (any order) -> the list format of knn_rows_kernel<false>.
__global__ void __launch_bounds__(kKnnWarps * 32) knn_merge_kernel(KnnParams p, KnnTileParams tp) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int N = p.N;
  const int g = tp.g0 + blockIdx.y;
  const int64_t gofs = (int64_t)g * p.ldv;
  for (int v = blockIdx.x * kKnnWarps + warp; v < N; v += gridDim.x * kKnnWarps) {
    const int n = tp.cnt[(int64_t)blockIdx.y * N + v];
    const unsigned long long tv = tp.tau[(int64_t)blockIdx.y * N + v];
    if (n > kKtCap || tv == kKnnNone || tv == 0ull) {       // too many candidates / no threshold / tied sample: rescan the row
      if (lane < kKnnK) { p.lk[(gofs + v) * kKnnK + lane] = kKnnNone; p.lj[(gofs + v) * kKnnK + lane] = -1; }
      if (lane == 0) { p.cnt[gofs + v] = 0; p.ptr[gofs + v] = 0; p.bound[gofs + v] = 0ull; }
      continue;
    }
    const KnnCand* __restrict__ C = tp.cand + ((int64_t)blockIdx.y * N + v) * kKtCap;
    unsigned long long k0 = kKnnNone, k1 = kKnnNone, k2 = kKnnNone, k3 = kKnnNone;
    int j0 = 0x7fffffff, j1 = 0x7fffffff, j2 = 0x7fffffff, j3 = 0x7fffffff;
    for (int q = lane; q < n; q += 32) {                    // at most 3 per lane: sorted by (weight, j) in registers
      const KnnCand e = C[q];
      if (knn_less(e.k, e.j, k2, j2)) {
        k3 = k2; j3 = j2;
        if (knn_less(e.k, e.j, k1, j1)) {
          k2 = k1; j2 = j1;
          if (knn_less(e.k, e.j, k0, j0)) { k1 = k0; j1 = j0; k0 = e.k; j0 = e.j; }
          else { k1 = e.k; j1 = e.j; }
        } else { k2 = e.k; j2 = e.j; }
      } else if (knn_less(e.k, e.j, k3, j3)) { k3 = e.k; j3 = e.j; }
    }
    // every entry <= tau_v is a candidate: the lists may vouch up to (tau_v, any j); kKtCap <= 3 * 32 keeps k3 empty,
    // so no lane lowers the limit
    knn_select_write<false>(p, gofs + v, lane, k0, k1, k2, k3, j0, j1, j2, j3, -1, -1, -1, -1, tv, 0x7fffffff);
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
