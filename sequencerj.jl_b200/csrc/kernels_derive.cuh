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

constexpr int kDeriveMaxFine = 64;     // dynamic shared memory: 2 KB per fine chunk

struct DeriveParams {
  int mode;               // MODE_L2 / MODE_KL / MODE_ENERGY
  int N;
  int64_t ldD, strideD;
  const double* Dfine;    // fine chunk matrices: chunk f at Dfine + (f - 0) * strideD  (f = fine chunk index in the group)
  double* Dout;           // coarse chunk matrix
  int ratio, nfine;       // coarse chunk z = blockIdx.z covers fine chunks [z*ratio + f_base, ...) (at most `ratio`, nfine in total)
  int c_first;            // first coarse chunk of this launch
  int f0, nf;             // (derived per block)
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
  // dense blocks: more than 1/32 of a 32 x 32 block flagged -> flag the enclosing 64 x 128 tile for the direct
  // tile kernels instead of listing the entries one by one
  unsigned char* tile_flags;   // [coarse chunks of the launch][tile_flag_stride]; may be nullptr
  int tile_flag_stride;
  const int* tile_col_offset;  // index of the first tile of every 128-column block in the tile list
};

// 32 x 32 tiles of the upper triangle; block = 32 x 8 threads; result mirrored through shared memory.
// Per-object factors of the 64 objects of a tile are precomputed once per block in (dynamic) shared memory,
// 4 doubles per (side, fine chunk, object):  L2: {w, w*g}   KL: {w, log w}   Energy: {w, b, w*(q + 2R), w*T},
// so the inner loop is 6-14 FP64 operations per (entry, fine chunk) and the pass stays HBM-bound.
__global__ void __launch_bounds__(256, 4) derive_kernel(DeriveParams p) {
  extern __shared__ double dsm[];                 // [2][nf][4][32]: consecutive lanes read consecutive doubles
  __shared__ double tile[32][33];
  const int bi = blockIdx.y, bj = blockIdx.x;
  if (bi > bj) return;
  const int tx = threadIdx.x, ty = threadIdx.y;
  {                                   // one launch covers all coarse chunks of the group: z selects the chunk
    const int z = blockIdx.z, cc = p.c_first + z;
    p.f0 = cc * p.ratio;
    p.nf = min(p.ratio, p.nfine - p.f0);
    p.Dout += (int64_t)z * p.strideD;
    p.cmass += (int64_t)cc * p.ldn;
    p.cself += (int64_t)cc * p.ldn;
    p.chunk_local = z;
  }
  const int N = p.N, nf = p.nf;
  const int i0 = bi * 32, j0 = bj * 32;
  auto fac = [&](int side, int q, int o) -> double* { return dsm + ((size_t)side * nf + q) * 128 + o; };   // + 32 * component

  // per-object factors, all 256 threads: item = (side, q, object)
  for (int t = ty * 32 + tx; t < 64 * nf; t += 256) {
    const int oo = t & 31, side = (t >> 5) & 1, q = t >> 6;
    const int o = (side ? j0 : i0) + oo;
    const int64_t fo = (int64_t)(p.f0 + q) * p.ldn + o;
    const bool ok = o < N;
    const double cm = ok ? p.cmass[o] : 1.0;
    const double fmv = ok ? p.fmass[fo] : 1.0;
    const double w = fmv / cm;
    double* d = fac(side, q, oo);
    d[0] = w;
    if (p.mode == MODE_L2) {
      d[32] = ok ? w * p.fg[fo] : 0.0;
    } else if (p.mode == MODE_KL) {
      d[32] = log(w);
    } else {
      d[32] = fmv;                                   // turned into b = (mass of the preceding fine chunks) / cm below
      d[64] = ok ? w * (p.fg[fo] + 2.0 * p.fR[fo]) : 0.0;
      d[96] = ok ? w * p.fT[fo] : 0.0;
    }
  }
  __syncthreads();
  if (p.mode == MODE_ENERGY) {
    const int t = ty * 32 + tx;
    if (t < 64) {
      const int oo = t & 31, side = t >> 5, o = (side ? j0 : i0) + oo;
      const double cm = (o < N) ? p.cmass[o] : 1.0;
      double run = 0.0;
      for (int q = 0; q < nf; ++q) {
        double* d = fac(side, q, oo);
        const double fmv = d[32];
        d[32] = run / cm;
        run += fmv;
      }
    }
    __syncthreads();
  }

  // each thread: column j, rows ty + 8 rr (rr = 0..3): 4 independent loads per fine chunk in flight
  const int j = j0 + tx;
  double acc[4] = {0.0, 0.0, 0.0, 0.0};
  bool live[4], swp[4];
  const double* src[4];
#pragma unroll
  for (int rr = 0; rr < 4; ++rr) {
    const int i = i0 + ty + 8 * rr;
    live[rr] = (i < N) && (j < N) && (i != j);
    swp[rr] = i > j;                                  // only inside diagonal tiles: stored value is for (lo, hi) = (j, i)
    src[rr] = p.Dfine + (int64_t)p.f0 * p.strideD + (int64_t)min(i, N - 1) * p.ldD + min(j, N - 1);
  }
#pragma unroll 2
  for (int q = 0; q < nf; ++q) {
    double dv[4];
#pragma unroll
    for (int rr = 0; rr < 4; ++rr) dv[rr] = __ldg(src[rr] + (int64_t)q * p.strideD) - p.eps_floor;
    const double* fc = fac(1, q, tx);                 // factors of the column object j
    const double c0 = fc[0], c1 = fc[32];
    double c2 = 0.0, c3 = 0.0, m = 0.0, P1 = 0.0, P2 = 0.0;
    if (p.mode == MODE_ENERGY) {
      c2 = fc[64]; c3 = fc[96];
      m = (double)p.fm[p.f0 + q];
      P1 = 0.5 * (m + 1.0);                                  // sum_{t=1..m} t/m
      P2 = (m + 1.0) * (2.0 * m + 1.0) / (6.0 * m);          // sum (t/m)^2
    }
#pragma unroll
    for (int rr = 0; rr < 4; ++rr) {
      const double* fr = fac(0, q, ty + 8 * rr);      // factors of the row object i
      // (lo, hi) = (row, col) except below the diagonal of a diagonal tile
      const double l0 = swp[rr] ? c0 : fr[0], h0 = swp[rr] ? fr[0] : c0;
      const double l1 = swp[rr] ? c1 : fr[32], h1 = swp[rr] ? fr[32] : c1;
      if (p.mode == MODE_KL) {
        acc[rr] += l0 * (dv[rr] + l1 - h1);
      } else if (p.mode == MODE_L2) {
        acc[rr] += l0 * h0 * (dv[rr] * dv[rr]) + (l0 - h0) * (l1 - h1);
      } else {
        const double r2 = fr[64], r3 = fr[96];
        const double dw = l0 - h0, db = l1 - h1;
        const double dA = swp[rr] ? (c2 - r2) : (r2 - c2), dB = swp[rr] ? (c3 - r3) : (r3 - c3);
        acc[rr] += l0 * h0 * (0.5 * dv[rr] * dv[rr]) + dw * dA + 2.0 * db * dB + dw * (dw * P2 + 2.0 * db * P1) + m * db * db;
      }
    }
  }
  // phase A: results + flags
  double outv[4];
  bool flg[4];
  int nflag = 0;
#pragma unroll
  for (int rr = 0; rr < 4; ++rr) {
    const int i = i0 + ty + 8 * rr;
    outv[rr] = 0.0;
    flg[rr] = false;
    if (i < N && j < N) {
      if (i == j) {
        outv[rr] = p.eps_floor;
      } else {
        const int lo = min(i, j), hi = max(i, j);
        double d;
        bool flag;
        if (p.mode == MODE_KL) {
          const double hh = p.cself[lo];
          d = acc[rr];
          flag = d < p.tau * fabs(hh);
        } else {
          const double selfterms = p.cself[lo] + p.cself[hi];
          flag = acc[rr] < p.tau * selfterms;
          d = fast_sqrt(acc[rr]);
          if (p.mode == MODE_ENERGY) d = 1.4142135623730951 * d;
        }
        outv[rr] = fabs(d) + p.eps_floor;
        flg[rr] = flag && i < j;      // (j, i) of a diagonal tile is rewritten by the mirrored store of the fix-up
        nflag += flg[rr] ? 1 : 0;
      }
    }
  }
  // block-level decision
  bool dense = false;
  if (p.tile_flags != nullptr) {
    __shared__ int s_cnt;
    if (tx == 0 && ty == 0) s_cnt = 0;
    __syncthreads();
    if (nflag) atomicAdd(&s_cnt, nflag);
    __syncthreads();
    dense = s_cnt > 32;
    if (dense && tx == 0 && ty == 0) {
      const int tidx = p.tile_col_offset[j0 / kTileN] + i0 / kTileM;
      p.tile_flags[(size_t)p.chunk_local * p.tile_flag_stride + tidx] = 1;
    }
  }
  // phase B: stores (+ per-entry list when the block is not dense)
#pragma unroll
  for (int rr = 0; rr < 4; ++rr) {
    const int r = ty + 8 * rr, i = i0 + r;
    double out = outv[rr];
    if (i < N && j < N) {
      if (flg[rr] && !dense) {
        const unsigned int idx = atomicAdd(p.fix_count, 1u);
        if (idx < p.fix_cap) p.fix_list[idx] = make_int4(p.chunk_local, i, j, 0);
        else { *p.fix_overflow = 1; out = -1.0; }
      }
      p.Dout[(int64_t)i * p.ldD + j] = out;
    }
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
