// General pair distance between two weighted empirical CDFs on their OWN grids:
// `EMD(u, v, uw, vw)` / `Energy(u, v, uw, vw)` -> `_cdf_distance(ecdf(u; weights = uw), ecdf(v; weights = vw), p)`
// (reference src/distancemetrics.jl:82-87, 143-148, 191-223; StatsBase.ecdf).  This is the per-pair form the
// reference's known-answer tests use (test/algotests.jl:36-107); `sequence` itself always takes the shared-grid
// tile kernels.  One CTA does the whole pair with global scratch:
//   1. stable sort of (grid value, weight) of u and of v (bitonic, ties by original index);
//   2. inclusive prefix sums of the sorted weights (sequential below 8192 entries = the order StatsBase uses);
//   3. identical grids  -> CDFs on the grid itself, dx = non-zero differences of sort([0; uv; vv])   (:206-209)
//                          (nu of them, or a single one that broadcasts)
//      otherwise        -> CDFs of both on A = sort([0; uv; vv])[1:end-1], dx = diff                 (:210-214)
//   4. p = 1: sum |U - V| dx          p = 2: sqrt(sum (U - V)^2 dx)   (Energy multiplies by sqrt 2).
#pragma once
#include <math_constants.h>
#include "common.cuh"

namespace seqj {

constexpr int kPairThreads = 1024;
constexpr int kPairSeqScan = 8192;

struct CdfPairParams {
  const double *u, *uw, *v, *vw;   // device copies of grid values and weights
  int nu, nv, p;
  int pu, pv, pa;                  // power-of-two padded lengths of the sort buffers (pa >= nu + nv + 1)
  double *su, *sv, *A;             // sorted grid values [pu], [pv]; merged grid [pa]
  int *iu, *iv;                    // sort permutations [pu], [pv]
  double *cu, *cv;                 // inclusive prefix sums of the sorted weights [nu], [nv]
  double *dx;                      // non-zero grid steps of the same-grid branch [nu + nv]
  int* pos;                        // compaction scratch [nu + nv + 1]
  double* out;                     // [1]
  int* status;                     // 0 ok, 1 = non-finite input, 2 = same-grid branch with repeated / zero grid values
};

// Bitonic sort of n (power of two) keys in global memory by one CTA; ties ordered by idx (= stable).
__device__ void pair_bitonic(double* key, int* idx, int n) {
  for (int k = 2; k <= n; k <<= 1)
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int t = threadIdx.x; t < n; t += blockDim.x) {
        const int l = t ^ j;
        if (l > t) {
          const double a = key[t], b = key[l];
          const int ia = idx ? idx[t] : 0, ib = idx ? idx[l] : 0;
          const bool gt = a > b || (a == b && ia > ib);
          if (gt == ((t & k) == 0)) {
            key[t] = b; key[l] = a;
            if (idx) { idx[t] = ib; idx[l] = ia; }
          }
        }
      }
      __syncthreads();
    }
}

// c[i] = w[perm[i]] summed inclusively, i < n.
__device__ void pair_scan_weights(const double* w, const int* perm, double* c, int n, double* s_part) {
  const int tid = threadIdx.x;
  if (n <= kPairSeqScan) {
    if (tid == 0) {
      double run = 0.0;
      for (int i = 0; i < n; ++i) { run += w[perm[i]]; c[i] = run; }
    }
    __syncthreads();
    return;
  }
  const int per = (n + kPairThreads - 1) / kPairThreads;
  const int lo = min(n, tid * per), hi = min(n, lo + per);
  double run = 0.0;
  for (int i = lo; i < hi; ++i) { run += w[perm[i]]; c[i] = run; }
  s_part[tid] = run;
  __syncthreads();
  if (tid == 0) {
    double acc = 0.0;
    for (int t = 0; t < kPairThreads; ++t) { const double x = s_part[t]; s_part[t] = acc; acc += x; }
  }
  __syncthreads();
  const double off = s_part[tid];
  for (int i = lo; i < hi; ++i) c[i] += off;
  __syncthreads();
}

// ecdf(x): total weight of the sorted values <= x over the weight sum; exactly 1 at and beyond the last value.
__device__ __forceinline__ double pair_ecdf(const double* sv, const double* c, int n, double x) {
  int lo = 0, hi = n;                           // upper bound: first index with sv > x
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if (sv[mid] <= x) lo = mid + 1; else hi = mid;
  }
  if (lo == n) return 1.0;
  return lo == 0 ? 0.0 : c[lo - 1] / c[n - 1];
}

__global__ void __launch_bounds__(kPairThreads, 1) cdf_pair_kernel(CdfPairParams p) {
  __shared__ double s_part[kPairThreads];
  __shared__ int s_ipart[kPairThreads];
  const int tid = threadIdx.x;
  const int nu = p.nu, nv = p.nv, na = nu + nv + 1;

  // 0. load + validate
  int bad = 0;
  for (int i = tid; i < p.pu; i += kPairThreads) {
    const double x = i < nu ? p.u[i] : CUDART_INF;
    if (i < nu && !(isfinite(x) && isfinite(p.uw[i]))) bad = 1;
    p.su[i] = x; p.iu[i] = i;
  }
  for (int i = tid; i < p.pv; i += kPairThreads) {
    const double x = i < nv ? p.v[i] : CUDART_INF;
    if (i < nv && !(isfinite(x) && isfinite(p.vw[i]))) bad = 1;
    p.sv[i] = x; p.iv[i] = i;
  }
  if (__syncthreads_or(bad)) {
    if (tid == 0) { *p.status = 1; *p.out = CUDART_NAN; }
    return;
  }
  // 1. sorted grids, 2. cumulative weights
  pair_bitonic(p.su, p.iu, p.pu);
  pair_bitonic(p.sv, p.iv, p.pv);
  pair_scan_weights(p.uw, p.iu, p.cu, nu, s_part);
  pair_scan_weights(p.vw, p.iv, p.cv, nv, s_part);
  // merged grid with the leading zero
  for (int i = tid; i < p.pa; i += kPairThreads)
    p.A[i] = (i == 0) ? 0.0 : (i <= nu ? p.su[i - 1] : (i < na ? p.sv[i - 1 - nu] : CUDART_INF));
  __syncthreads();
  pair_bitonic(p.A, nullptr, p.pa);
  // 3. identical grids?
  int differ = (nu != nv);
  if (!differ)
    for (int i = tid; i < nu; i += kPairThreads) differ |= (p.su[i] != p.sv[i]);
  const bool same = !__syncthreads_or(differ);

  double acc = 0.0;
  if (same) {
    // dx = filter(!iszero, diff(A)): compact the non-zero steps in order; there must be exactly nu of them
    const int nd = na - 1;
    const int per = (nd + kPairThreads - 1) / kPairThreads;
    const int lo = min(nd, tid * per), hi = min(nd, lo + per);
    int cnt = 0;
    for (int k = lo; k < hi; ++k) cnt += (p.A[k + 1] - p.A[k]) != 0.0;
    s_ipart[tid] = cnt;
    __syncthreads();
    if (tid == 0) {
      int run = 0;
      for (int t = 0; t < kPairThreads; ++t) { const int x = s_ipart[t]; s_ipart[t] = run; run += x; }
      p.pos[0] = run;
    }
    __syncthreads();
    // `(UA .- VA) .* dx` broadcasts: dx must have nu entries, or exactly one (e.g. the grid [0, 8] of
    // test/algotests.jl:74-78, whose only non-zero step is 8); an empty dx against one CDF value sums to 0.
    const int ndx = p.pos[0];
    if (ndx == 0 && nu == 1) {
      if (tid == 0) { *p.status = 0; *p.out = 0.0; }
      return;
    }
    if (ndx != nu && ndx != 1) {               // the reference fails here with a DimensionMismatch
      if (tid == 0) { *p.status = 2; *p.out = CUDART_NAN; }
      return;
    }
    int w = s_ipart[tid];
    for (int k = lo; k < hi; ++k) {
      const double d = p.A[k + 1] - p.A[k];
      if (d != 0.0) p.dx[w++] = d;
    }
    __syncthreads();
    for (int k = tid; k < nu; k += kPairThreads) {
      const double x = p.su[k];
      const double d = pair_ecdf(p.su, p.cu, nu, x) - pair_ecdf(p.sv, p.cv, nv, x);
      acc += (p.p == 1 ? fabs(d) : d * d) * p.dx[ndx == 1 ? 0 : k];
    }
  } else {
    for (int k = tid; k < na - 1; k += kPairThreads) {
      const double x = p.A[k];
      const double d = pair_ecdf(p.su, p.cu, nu, x) - pair_ecdf(p.sv, p.cv, nv, x);
      acc += (p.p == 1 ? fabs(d) : d * d) * (p.A[k + 1] - x);
    }
  }
  // 4. block sum
  s_part[tid] = acc;
  __syncthreads();
  for (int o = kPairThreads / 2; o > 0; o >>= 1) {
    if (tid < o) s_part[tid] += s_part[tid + o];
    __syncthreads();
  }
  if (tid == 0) {
    *p.status = 0;
    *p.out = (p.p == 1) ? s_part[0] : sqrt(s_part[0]);
  }
}

}  // namespace seqj
