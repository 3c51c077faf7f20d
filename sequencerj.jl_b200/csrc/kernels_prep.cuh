// K1: per-chunk normalisation + prefix scan.
// Replaces _splitnorm (reference src/sequencer.jl:383-407) and the weighted ECDF evaluated on the
// shared unit grid inside EMD/Energy (src/distancemetrics.jl:86,147,206-209; StatsBase.ecdf).
//
// One warp per (object j, chunk c).  Input A is object-major [N][ldA] (= Julia column-major M x N).
// Outputs, all K-contiguous [N][ldX] with chunk c at columns [c*cl_pad, c*cl_pad + m_c) and zero
// padding up to cl_pad (a multiple of 16 doubles so TMA boxes never straddle chunks):
//   P    = a / sum(a)                    (PDF; Euclidean operand, KL left operand)
//   logP = log(P)                        (KL right operand)
//   Uc   = cumsum(P)/sum(P) - (k+1)/m    (CDF centred on the uniform ramp; EMD / Energy operand.
//                                         Differences U_i - U_j are unchanged by the centring, and
//                                         |Uc|^2 << |U|^2 removes the Gram-form cancellation.)
//   U    = cumsum(P)/sum(P), last = 1    (optional, uncentred; unit-level API only)
// and per-chunk vectors [nchunks][ldn]:  sP = sum P^2, sU = sum Uc^2, H = sum P log P.
// HBM-bound: 8*M*N bytes read, up to 3 * 8*M*N written per scale.
#pragma once
#include <math_constants.h>
#include "common.cuh"

namespace seqj {

// clamp!(A, eps(), typemax) of the reference's `sequence` (src/sequencer.jl:130) is applied on load, so the
// device results equal those of the clamped input whether or not the host copy has been clamped yet.
constexpr double kClampEps = 2.220446049250313e-16;

struct InputStats {
  double minv, maxv;
  int has_nan;
  int pad;
};

// Input conditioning check of `sequence` (src/sequencer.jl:111): min / max / NaN over the whole input.
__global__ void __launch_bounds__(256) input_stats_kernel(const double* __restrict__ A, int64_t n, InputStats* out) {
  double mn = CUDART_INF, mx = -CUDART_INF;
  int nan = 0;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const double a = A[i];
    nan |= (a != a);
    mn = fmin(mn, a);
    mx = fmax(mx, a);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    mn = fmin(mn, __shfl_xor_sync(0xffffffffu, mn, o));
    mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    nan |= __shfl_xor_sync(0xffffffffu, nan, o);
  }
  if ((threadIdx.x & 31) == 0) {
    // non-negative finite doubles order like their bit patterns; negatives / NaN are caught by the flags
    if (nan) atomicOr(&out->has_nan, 1);
    if (mn < 0.0) atomicOr(&out->pad, 1);
    atomicMin(reinterpret_cast<unsigned long long*>(&out->minv), (unsigned long long)__double_as_longlong(fmax(mn, 0.0)));
    atomicMax(reinterpret_cast<unsigned long long*>(&out->maxv), (unsigned long long)__double_as_longlong(fmax(mx, 0.0)));
  }
}

// Float32 input (reference quirk Q2: the image tests pass Float32 matrices): widen to the FP64 working copy.  The
// reference's `clamp!(A, eps(), typemax(Float32))` (src/sequencer.jl:130) is the same clamp-on-load as for Float64:
// eps() = 2^-52 is exactly representable in Float32.
__global__ void __launch_bounds__(256) widen_f32_kernel(const float* __restrict__ A, int64_t n, double* __restrict__ out) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    out[i] = (double)A[i];
}

struct PrepParams {
  const double* A;
  int64_t ldA;
  int M, N;
  int cl, cl_pad, nchunks;
  int64_t ldX, ldn;
  double* P;
  double* logP;
  double* Uc;
  double* U;
  double* sP;
  double* sU;
  double* H;
  int normalize;
  int need_log;           // log P and the entropy term (KL only): an FP64 log per element otherwise wasted
  int need_cdf;           // the CDF pass (EMD / Energy operands and their per-chunk vectors)
  int f32_pdf;            // Float32 input (quirk Q2): the PDF is the reference's Float32 quotient Float32(a) / Float32(sum)
  // explicit grid (reference: metrics passed as the TYPES EMD / Energy, src/sequencer.jl:217-219): per-row
  // deltas g_k - g_{k-1} of the chunk-local grid with g_0 = 0 (src/distancemetrics.jl:199-209).  Since the
  // deltas are positive, |U-V|*d = |d*U - d*V| and (U-V)^2*d = (sqrt(d)*U - sqrt(d)*V)^2: the grid only scales
  // the operands.
  const double* grid;     // [M] grid values, or nullptr
  double* UwE;            // Uc * delta        (EMD on the grid)
  double* UwG;            // Uc * sqrt(delta)  (Energy on the grid)
  double* sUG;            // sum (Uc*sqrt(delta))^2
  // per-chunk vectors used to derive coarser scales from the finest one (kernels_derive.cuh)
  double* mass;           // sum a          (the chunk mass the column is normalised by)
  double* sR;             // sum Uc * (k+1)/m
  double* sT;             // sum Uc
};

__global__ void __launch_bounds__(256) prep_kernel(PrepParams p) {
  const int lane = threadIdx.x & 31;
  const int64_t wid = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (wid >= (int64_t)p.N * p.nchunks) return;
  const int j = (int)(wid / p.nchunks);
  const int c = (int)(wid % p.nchunks);
  const int r0 = c * p.cl;
  const int m = min(p.cl, p.M - r0);
  const double* __restrict__ src = p.A + (int64_t)j * p.ldA + r0;
  const int64_t dst = (int64_t)j * p.ldX + (int64_t)c * p.cl_pad;

  // pass 1: chunk mass
  double s = 0.0;
  for (int k = lane; k < m; k += 32) s += fmax(src[k], kClampEps);
  s = warp_sum(s);

  // pass 2: PDF, log PDF, moments
  double w = 0.0, sq = 0.0, h = 0.0;
  for (int k = lane; k < p.cl_pad; k += 32) {
    double pv = 0.0, lp = 0.0;
    if (k < m) {
      const double a = fmax(src[k], kClampEps);
      pv = p.normalize ? (p.f32_pdf ? (double)((float)a / (float)s) : a / s) : a;
      w += pv;
      sq += pv * pv;
      if (p.need_log) {
        lp = log(pv);
        h += pv * lp;
      }
    }
    if (p.P) p.P[dst + k] = pv;
    if (p.logP) p.logP[dst + k] = lp;
  }
  w = warp_sum(w);
  sq = warp_sum(sq);
  h = warp_sum(h);

  // pass 3: CDF by warp-shuffle inclusive scan with a running carry
  double su = 0.0, sug = 0.0, sr = 0.0, st = 0.0;
  if (p.need_cdf) {
    double carry = 0.0;
    for (int base = 0; base < p.cl_pad; base += 32) {
      const int k = base + lane;
      double x = 0.0;
      if (k < m) {
        const double a = fmax(src[k], kClampEps);
        x = p.normalize ? (p.f32_pdf ? (double)((float)a / (float)s) : a / s) : a;
      }
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const double y = __shfl_up_sync(0xffffffffu, x, o);
        if (lane >= o) x += y;
      }
      const double cum = carry + x;
      carry = __shfl_sync(0xffffffffu, cum, 31);
      double u = 0.0, uc = 0.0;
      if (k < m) {
        u = (k == m - 1) ? 1.0 : cum / w;
        const double ramp = (double)(k + 1) / (double)m;
        uc = u - ramp;
        su += uc * uc;
        sr += uc * ramp;
        st += uc;
      }
      double ue = 0.0, ug = 0.0;
      if (p.grid && k < m) {
        const double dlt = p.grid[r0 + k] - (k == 0 ? 0.0 : p.grid[r0 + k - 1]);
        ue = uc * dlt;
        ug = uc * sqrt(dlt);
        sug += ug * ug;
      }
      if (k < p.cl_pad) {            // cl_pad is a multiple of 16, not of the 32-wide scan step
        if (p.Uc) p.Uc[dst + k] = uc;
        if (p.U) p.U[dst + k] = u;
        if (p.UwE) p.UwE[dst + k] = ue;
        if (p.UwG) p.UwG[dst + k] = ug;
      }
    }
    su = warp_sum(su);
    sug = warp_sum(sug);
    sr = warp_sum(sr);
    st = warp_sum(st);
  }
  if (lane == 0) {
    const int64_t o = (int64_t)c * p.ldn + j;
    if (p.sP) p.sP[o] = sq;
    if (p.sU) p.sU[o] = su;
    if (p.H) p.H[o] = h;
    if (p.sUG) p.sUG[o] = sug;
    if (p.mass) p.mass[o] = s;
    if (p.sR) p.sR[o] = sr;
    if (p.sT) p.sT[o] = st;
  }
}

// Unit-level copy-out: padded K-contiguous [N][ldX] chunk layout -> dense M x N column-major.
__global__ void unpad_kernel(const double* __restrict__ X, int64_t ldX, int cl, int cl_pad, int M, int N,
                             double* __restrict__ out) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (int64_t)M * N) return;
  const int j = (int)(idx / M);
  const int r = (int)(idx % M);
  const int c = r / cl;
  out[idx] = X[(int64_t)j * ldX + (int64_t)c * cl_pad + (r - c * cl)];
}

}  // namespace seqj
