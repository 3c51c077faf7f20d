// Coarse scales derived from the finest one.
//
// The reference evaluates `pairwise` independently for every (metric, scale, chunk) (src/sequencer.jl:194-226).
// When the chunk boundaries of a coarse scale are a subset of those of the finest scale (e.g. scales 1,2,4,8,16
// of a length-4096 grid), the coarse chunk of object i is the concatenation of fine chunks f, rescaled:
//      p^c_ik = w_i^f p^f_ik ,     U^c_ik = w_i^f U^f_ik + b_i^f      (k in f),
// with w_i^f = mass_i^f / mass_i^c (fraction of the coarse mass in f) and b_i^f = sum of w_i over the fine chunks
// of c that precede f.  Expanding the three contraction metrics over the fine chunks gives them in terms of the
// FINISHED fine distance matrices plus per-object vectors, with no further O(N^2 m) work:
//   Euclidean^2   = sum_f [ w_i w_j v_ij^f + (w_i - w_j)(w_i g_i^f - w_j g_j^f) ]        v^f = (D^f - eps)^2, g = sum p^2
//   KL(i||j)      = sum_f   w_i [ KL_ij^f + log w_i - log w_j ]                          KL^f = D^f - eps  (i < j)
//   Energy^2 / 2  = sum_f [ w_i w_j e_ij^f + (w_i - w_j)(w_i q_i - w_j q_j) + 2 (w_i - w_j)(w_i R_i - w_j R_j)
//                           + 2 (b_i - b_j)(w_i T_i - w_j T_j) + (w_i - w_j)^2 P2 + 2 (w_i - w_j)(b_i - b_j) P1
//                           + m_f (b_i - b_j)^2 ]      e^f = ((D^f - eps)/sqrt 2)^2, q = sum Uc^2, R = sum Uc*rho,
//                                                      T = sum Uc, rho_t = t/m_f, P1 = sum rho, P2 = sum rho^2
// (Uc = fine CDF centred on the ramp rho).  The leading terms use the accurate fine distances (already fixed up
// where the Gram form cancelled) and vanish for identical objects; the rest involves only O(N) vectors.
// This replaces 4 of the 5 DMMA launches per metric at config 3 by an HBM-bound pass: read the n_f fine
// matrices, write one coarse matrix -> 8 (n_f + 1) bytes per entry of the upper triangle.
// Entries whose result is small against the self terms are flagged for the same direct fix-up as the tiles.
// EMD (L1 between CDFs) is not derivable this way and always runs the tile kernel.
#pragma once
#include "common.cuh"
#include "kernels_dist.cuh"

namespace seqj {

constexpr int kDeriveMaxFine = 32;     // 2 x 2 x 32 x 32 doubles of weights = 32 KB of static shared memory

struct DeriveParams {
  int mode;               // MODE_L2 / MODE_KL / MODE_ENERGY
  int N;
  int64_t ldD, strideD;
  const double* Dfine;    // fine chunk matrices: chunk f at Dfine + (f - 0) * strideD  (f = fine chunk index in the group)
  double* Dout;           // coarse chunk matrix
  int f0, nf;             // fine chunks [f0, f0 + nf) make up this coarse chunk
  int64_t ldn;
  const double* fmass;    // fine-scale vectors [nfine_total][ldn]
  const double* fg;       // sum p^2        (L2)  or  sum Uc^2 (Energy)
  const double* fR;       // Energy only
  const double* fT;
  const int* fm;          // length of every fine chunk
  const double* cmass;    // coarse-scale vectors for THIS chunk: [ldn]
  const double* cself;    // coarse self term for the flag: sum p^2 / sum Uc^2 / H
  double eps_floor, tau;
  int chunk_local;        // index of the coarse chunk inside the current fix-up batch
  int4* fix_list;
  unsigned int* fix_count;
  unsigned int fix_cap;
  int* fix_overflow;
};

// 32 x 32 tiles of the upper triangle; block = 32 x 8 threads; result mirrored through shared memory.
__global__ void __launch_bounds__(256) derive_kernel(DeriveParams p) {
  __shared__ double tile[32][33];
  __shared__ double s_w[2][kDeriveMaxFine][32];   // w_i^f for the 32 rows (0) and 32 cols (1) of this tile
  __shared__ double s_b[2][kDeriveMaxFine][32];   // b_i^f
  const int bi = blockIdx.y, bj = blockIdx.x;
  if (bi > bj) return;
  const int tx = threadIdx.x, ty = threadIdx.y;
  const int N = p.N;
  const int i0 = bi * 32, j0 = bj * 32;

  // per-object weights of the 64 objects touched by this tile
  for (int t = ty * 32 + tx; t < 64; t += 256) {
    const int side = t >> 5, o = (side ? j0 : i0) + (t & 31);
    double run = 0.0;
    const double cm = (o < N) ? p.cmass[o] : 1.0;
    for (int q = 0; q < p.nf; ++q) {
      const double fmv = (o < N) ? p.fmass[(int64_t)(p.f0 + q) * p.ldn + o] : 0.0;
      s_w[side][q][t & 31] = fmv / cm;
      s_b[side][q][t & 31] = run / cm;
      run += fmv;
    }
  }
  __syncthreads();

  const int j = j0 + tx;
  for (int r = ty; r < 32; r += 8) {
    const int i = i0 + r;
    double out = 0.0;
    if (i < N && j < N) {
      if (i == j) {
        out = p.eps_floor;
      } else {
        // (lo, hi) ordering matters for KL only: the stored value is KL(p_lo || p_hi)
        const bool swap = i > j;           // only inside diagonal tiles
        const int lo = swap ? j : i, hi = swap ? i : j;
        const int slo = swap ? 1 : 0, shi = swap ? 0 : 1;
        const int tlo = swap ? tx : r, thi = swap ? r : tx;
        double acc = 0.0;
        for (int q = 0; q < p.nf; ++q) {
          const int f = p.f0 + q;
          const double dv = p.Dfine[(int64_t)f * p.strideD + (int64_t)i * p.ldD + j] - p.eps_floor;
          const double wi = s_w[slo][q][tlo], wj = s_w[shi][q][thi];
          if (p.mode == MODE_KL) {
            acc += wi * (dv + log(wi) - log(wj));
          } else if (p.mode == MODE_L2) {
            const double gi = p.fg[(int64_t)f * p.ldn + lo], gj = p.fg[(int64_t)f * p.ldn + hi];
            acc += wi * wj * dv * dv + (wi - wj) * (wi * gi - wj * gj);
          } else {
            const double e = 0.5 * dv * dv;
            const double qi = p.fg[(int64_t)f * p.ldn + lo], qj = p.fg[(int64_t)f * p.ldn + hi];
            const double Ri = p.fR[(int64_t)f * p.ldn + lo], Rj = p.fR[(int64_t)f * p.ldn + hi];
            const double Ti = p.fT[(int64_t)f * p.ldn + lo], Tj = p.fT[(int64_t)f * p.ldn + hi];
            const double bi_ = s_b[slo][q][tlo], bj_ = s_b[shi][q][thi];
            const double m = (double)p.fm[f];
            const double P1 = 0.5 * (m + 1.0);                               // sum_{t=1..m} t/m
            const double P2 = (m + 1.0) * (2.0 * m + 1.0) / (6.0 * m);       // sum (t/m)^2
            const double dw = wi - wj, db = bi_ - bj_;
            acc += wi * wj * e + dw * (wi * qi - wj * qj) + 2.0 * dw * (wi * Ri - wj * Rj) +
                   2.0 * db * (wi * Ti - wj * Tj) + dw * dw * P2 + 2.0 * dw * db * P1 + m * db * db;
          }
        }
        double d;
        bool flag;
        if (p.mode == MODE_KL) {
          const double hh = p.cself[lo];
          d = acc;
          flag = d < p.tau * fabs(hh);
        } else {
          const double selfterms = p.cself[lo] + p.cself[hi];
          flag = acc < p.tau * selfterms;
          d = fast_sqrt(acc);
          if (p.mode == MODE_ENERGY) d = 1.4142135623730951 * d;
        }
        out = fabs(d) + p.eps_floor;
        if (flag && i < j) {      // (j, i) of a diagonal tile is rewritten by the mirrored store of the fix-up
          const unsigned int idx = atomicAdd(p.fix_count, 1u);
          if (idx < p.fix_cap) p.fix_list[idx] = make_int4(p.chunk_local, i, j, 0);
          else { *p.fix_overflow = 1; out = -1.0; }
        }
      }
    }
    if (i < N && j < N) p.Dout[(int64_t)i * p.ldD + j] = out;
    tile[r][tx] = out;
  }
  if (bi == bj) return;          // diagonal tiles computed both halves directly
  __syncthreads();
  // mirror: element (ii, jj) of the tile is also written at (jj, ii), coalesced along ii
  for (int r = ty; r < 32; r += 8) {
    const int jj = j0 + r, ii = i0 + tx;
    if (ii < N && jj < N) p.Dout[(int64_t)jj * p.ldD + ii] = tile[tx][r];
  }
}

}  // namespace seqj
