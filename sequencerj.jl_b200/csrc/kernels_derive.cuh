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

constexpr int kDeriveMaxFine = 32;     // dynamic shared memory: 2 KB (L2, KL) / 4 KB (Energy) per fine chunk

struct DeriveParams {
  int mode;               // MODE_L2 / MODE_KL / MODE_ENERGY
  int N;
  int64_t ldD, strideD;
  const double* Dfine;    // fine chunk matrices: chunk f of the fine scale at Dfine + (f - f_base) * strideD
  int f_base;             // first fine chunk resident at Dfine (a wave may hold only a block of the fine scale)
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
  // per-object factors of every fine chunk, precomputed by derive_factors_kernel: fac[(f * NF + comp) * ldn + object]
  // and, for Energy, the ramp moments {m, P1 = sum rho, P2 = sum rho^2} of every fine chunk: pm[3 f + 0..2]
  double* fac;
  double* pm;
};

// 64 x 64 tiles of the upper triangle; block = 32 x 8 threads, each thread owns 2 adjacent columns (one 16-byte
// chunk) of 8 rows.  The fine tiles are staged in shared memory with per-thread cp.async (16 bytes each, L2 only):
// a thread copies exactly the chunks it reads back, so the staging needs no barrier, costs no registers, and all of
// a CTA's loads (64 KB for two fine chunks) are in flight at once; three CTAs per SM overlap one block's arithmetic
// with the others' loads.  (The first version loaded into registers, 4 x 16 bytes per thread at a time with 16
// warps per SM: 49 % of the HBM rate, stalled on the loads.)  Fine chunks are consumed two at a time.
// Results are written over stage 0, re-read by their owner for the direct store (after the block-level dense
// decision) and re-read transposed for the mirror image.  Stage rows are XOR-swizzled by 16-byte chunk
// (slot = chunk ^ (row >> 1)) so that the row-wise and the transposed accesses are both (nearly) conflict-free
// without padding.
// Per-object factors of the 128 objects of a tile are precomputed once per block in (dynamic) shared memory,
// NF doubles per (side, fine chunk, object):  L2: {w, w*g}   KL: {w, log w}   Energy: {w, b, w*(q + 2R), w*T},
// so the inner loop is 6-14 FP64 operations per (entry, fine chunk) and the pass stays HBM-bound.
constexpr int kDeriveTile = 64;
constexpr int kDeriveStage = 2;                       // fine chunks staged at a time
constexpr int kDeriveStageDoubles = kDeriveTile * kDeriveTile;
__host__ __device__ constexpr int derive_nf(int mode) { return mode == MODE_ENERGY ? 4 : 2; }
inline size_t derive_smem_bytes(int mode, int nf) {
  return ((size_t)kDeriveStage * kDeriveStageDoubles + (size_t)2 * nf * derive_nf(mode) * kDeriveTile) * sizeof(double);
}

__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() {
  asm volatile("cp.async.commit_group;\n cp.async.wait_group 0;" ::: "memory");
}

// Per-object factors for one (metric, coarse scale): one thread per (fine chunk, object).  Kept out of derive_kernel
// so that its prologue is nothing but asynchronous copies (the division / log / dependent global loads used to sit
// at the head of every block's critical path).
template <int MODE>
__global__ void __launch_bounds__(256) derive_factors_kernel(DeriveParams p) {
  constexpr int NF = derive_nf(MODE);
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t total = (int64_t)p.nfine * p.ldn;
  if (idx >= total) return;
  const int f = (int)(idx / p.ldn), o = (int)(idx % p.ldn);
  const int cc = f / p.ratio, fbeg = cc * p.ratio;
  double* out = p.fac + (int64_t)f * NF * p.ldn + o;
  const bool ok = o < p.N;
  const int64_t fo = (int64_t)f * p.ldn + o;
  const double cm = ok ? p.cmass[(int64_t)cc * p.ldn + o] : 1.0;
  const double fmv = ok ? p.fmass[fo] : 1.0;
  const double w = fmv / cm;
  out[0] = w;
  if (MODE == MODE_L2) {
    out[p.ldn] = ok ? w * p.fg[fo] : 0.0;
  } else if (MODE == MODE_KL) {
    out[p.ldn] = log(w);
  } else {
    double run = 0.0;                                 // b = (mass of the preceding fine chunks of the coarse chunk) / cm
    for (int q = fbeg; q < f; ++q) run += ok ? p.fmass[(int64_t)q * p.ldn + o] : 0.0;
    out[p.ldn] = run / cm;
    out[2 * p.ldn] = ok ? w * (p.fg[fo] + 2.0 * p.fR[fo]) : 0.0;
    out[3 * p.ldn] = ok ? w * p.fT[fo] : 0.0;
    if (o == 0) {
      const double m = (double)p.fm[f];
      p.pm[3 * f] = m;
      p.pm[3 * f + 1] = 0.5 * (m + 1.0);                          // sum_{t=1..m} t/m
      p.pm[3 * f + 2] = (m + 1.0) * (2.0 * m + 1.0) / (6.0 * m);  // sum (t/m)^2
    }
  }
}

template <int MODE>
__global__ void __launch_bounds__(256, 3) derive_kernel(DeriveParams p) {
  constexpr int NF = derive_nf(MODE);
  constexpr int TS = kDeriveTile;
  extern __shared__ __align__(16) double dsm[];   // [kDeriveStage][64][64] staged fine tiles, then [2][nf][NF][64] factors
  __shared__ int s_cnt;
  const int bi = blockIdx.y, bj = blockIdx.x;
  if (bi > bj) return;
  const int tx = threadIdx.x, ty = threadIdx.y, tid = ty * 32 + tx;
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
  const int i0 = bi * TS, j0 = bj * TS;
  const bool diag = (bi == bj);
  double* stage = dsm;
  double* facs = dsm + kDeriveStage * kDeriveStageDoubles;
  auto fac = [&](int side, int q, int o) -> double* { return facs + ((size_t)side * nf + q) * (NF * TS) + o; };   // + TS * component
  // element (r, 2*chunk + e) of stage s
  auto at = [&](int s_, int r, int chunk) -> double* { return stage + s_ * kDeriveStageDoubles + r * TS + ((chunk ^ ((r >> 1) & 31)) << 1); };

  const int j = j0 + 2 * tx;                          // this thread's columns: j, j + 1 (j even, j + 1 < ldD)
  const int jc = min(j, (N - 1) & ~1);                // clamped load column (results of columns >= N are dropped)
  const double* gbase = p.Dfine + (int64_t)(p.f0 - p.f_base) * p.strideD + jc;
  auto issue = [&](int qb) {                          // stage fine chunks qb, qb + 1 (this thread's 16 chunks)
#pragma unroll
    for (int s_ = 0; s_ < kDeriveStage; ++s_)
      if (qb + s_ < nf) {
        const double* g = gbase + (int64_t)(qb + s_) * p.strideD;
#pragma unroll
        for (int k = 0; k < 8; ++k)
          cp_async16(smem_u32(at(s_, ty + 8 * k, tx)), g + (int64_t)min(i0 + ty + 8 * k, N - 1) * p.ldD);
      }
  };
  issue(0);

  // per-object factors: asynchronous copies of the precomputed rows, one 16-byte chunk (2 objects) per item.
  // Objects beyond ldn (last tile) read into the next row / the buffer's tail padding; their results are dropped.
  for (int t = tid; t < 2 * nf * NF * (TS / 2); t += 256) {
    const int ch = t & (TS / 2 - 1), row = t / (TS / 2);          // row = (side * nf + q) * NF + comp
    const int comp = row % NF, sq = row / NF, q = sq % nf, side = sq / nf;
    const double* src = p.fac + ((int64_t)(p.f0 + q) * NF + comp) * p.ldn + (side ? j0 : i0) + 2 * ch;
    cp_async16(smem_u32(facs + (size_t)row * TS + 2 * ch), src);
  }
  if (tid == 0) s_cnt = 0;

  double acc[8][2];
#pragma unroll
  for (int k = 0; k < 8; ++k) acc[k][0] = acc[k][1] = 0.0;
  for (int qb = 0; qb < nf; qb += kDeriveStage) {
    cp_async_wait_all();
    if (qb == 0) __syncthreads();                         // the factor rows were copied by other threads
#pragma unroll
    for (int s_ = 0; s_ < kDeriveStage; ++s_) {
      const int q = qb + s_;
      if (q >= nf) break;
      const double* fc = fac(1, q, 2 * tx);           // factors of the column objects j, j + 1
      const double2 c0 = *reinterpret_cast<const double2*>(fc), c1 = *reinterpret_cast<const double2*>(fc + TS);
      double2 c2 = make_double2(0.0, 0.0), c3 = c2;
      double m = 0.0, P1 = 0.0, P2 = 0.0;
      if (MODE == MODE_ENERGY) {
        c2 = *reinterpret_cast<const double2*>(fc + 2 * TS);
        c3 = *reinterpret_cast<const double2*>(fc + 3 * TS);
        m = p.pm[3 * (p.f0 + q)];
        P1 = p.pm[3 * (p.f0 + q) + 1];                         // sum_{t=1..m} t/m
        P2 = p.pm[3 * (p.f0 + q) + 2];                         // sum (t/m)^2
      }
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const int r = ty + 8 * k, i = i0 + r;
        const double* fr = fac(0, q, r);              // factors of the row object i (warp-wide broadcast)
        const double r0 = fr[0], r1 = fr[TS];
        double2 dv = *reinterpret_cast<const double2*>(at(s_, r, tx));
        dv.x -= p.eps_floor;
        dv.y -= p.eps_floor;
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const double cc0 = e ? c0.y : c0.x, cc1 = e ? c1.y : c1.x, d = e ? dv.y : dv.x;
          if (MODE == MODE_KL) {
            // (lo, hi) = (row, col) except below the diagonal of a diagonal tile, where the stored value is for (j, i)
            const bool swp = diag && i > j + e;
            const double l0 = swp ? cc0 : r0, l1 = swp ? cc1 : r1, h1 = swp ? r1 : cc1;
            acc[k][e] += l0 * (d + l1 - h1);
          } else if (MODE == MODE_L2) {               // symmetric in (row, col): no swap needed
            acc[k][e] += r0 * cc0 * (d * d) + (r0 - cc0) * (r1 - cc1);
          } else {
            const double r2 = fr[2 * TS], r3 = fr[3 * TS];
            const double cc2 = e ? c2.y : c2.x, cc3 = e ? c3.y : c3.x;
            const double dw = r0 - cc0, db = r1 - cc1;
            acc[k][e] += r0 * cc0 * (0.5 * d * d) + dw * (r2 - cc2) + 2.0 * db * (r3 - cc3) +
                         dw * (dw * P2 + 2.0 * db * P1) + m * db * db;
          }
        }
      }
    }
    if (qb + kDeriveStage < nf) issue(qb + kDeriveStage);     // the thread overwrites only chunks it has consumed
  }
  // results + flags, written over stage 0 (each thread's own chunks)
  unsigned flagbits = 0;                              // bit (2k + e) = entry flagged for the fix-up
  int nflag = 0;
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const int r = ty + 8 * k, i = i0 + r;
    double2 o2;
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      const int jj = j + e;
      double out = 0.0;
      if (i < N && jj < N) {
        if (i == jj) {
          out = p.eps_floor;
        } else {
          const int lo = min(i, jj), hi = max(i, jj);
          double d;
          bool flag;
          if (MODE == MODE_KL) {
            d = acc[k][e];
            flag = d < p.tau * fabs(p.cself[lo]);
          } else {
            flag = acc[k][e] < p.tau * (p.cself[lo] + p.cself[hi]);
            d = fast_sqrt(acc[k][e]);
            if (MODE == MODE_ENERGY) d = 1.4142135623730951 * d;
          }
          out = fabs(d) + p.eps_floor;
          if (flag && i < jj) {       // (j, i) of a diagonal tile is rewritten by the mirrored store of the fix-up
            flagbits |= 1u << (2 * k + e);
            ++nflag;
          }
        }
      }
      if (e) o2.y = out; else o2.x = out;
    }
    *reinterpret_cast<double2*>(at(0, r, tx)) = o2;
  }
  // block-level decision: more than 1/32 of the block flagged -> the enclosing 64 x 128 tile goes to the direct kernels
  bool dense = false;
  if (p.tile_flags != nullptr) {
    if (nflag) atomicAdd(&s_cnt, nflag);
    __syncthreads();
    dense = s_cnt > TS * TS / 32;
    if (dense && tid == 0) {
      const int tidx = p.tile_col_offset[j0 / kTileN] + i0 / kTileM;
      p.tile_flags[(size_t)p.chunk_local * p.tile_flag_stride + tidx] = 1;
    }
  }
  // direct stores of the thread's own entries (+ per-entry list when the block is not dense)
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const int r = ty + 8 * k, i = i0 + r;
    if (i < N && j < N) {
      double2* own = reinterpret_cast<double2*>(at(0, r, tx));
      double2 out = *own;
      if (!dense && ((flagbits >> (2 * k)) & 3u)) {
#pragma unroll
        for (int e = 0; e < 2; ++e)
          if ((flagbits >> (2 * k + e)) & 1u) {
            const unsigned int idx = atomicAdd(p.fix_count, 1u);
            if (idx < p.fix_cap) p.fix_list[idx] = make_int4(p.chunk_local, i, j + e, 0);
            else {
              *p.fix_overflow = 1;
              if (e) out.y = -1.0; else out.x = -1.0;
              *own = out;
            }
          }
      }
      *reinterpret_cast<double2*>(p.Dout + (int64_t)i * p.ldD + j) = out;      // column j + 1 == N lands in the row padding
    }
  }
  if (diag) return;              // diagonal tiles computed both halves directly
  __syncthreads();
  // mirror: element (ii, jj) of the tile is also written at (jj, ii), coalesced along ii
  const int ii = i0 + 2 * tx;
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const int r = ty + 8 * k, jj = j0 + r;
    if (jj < N && ii < N) {
      const double a = at(0, 2 * tx, r >> 1)[r & 1], b = at(0, 2 * tx + 1, r >> 1)[r & 1];
      double* dst = p.Dout + (int64_t)jj * p.ldD + ii;
      if (ii + 1 < N) *reinterpret_cast<double2*>(dst) = make_double2(a, b);
      else *dst = a;
    }
  }
}

}  // namespace seqj
