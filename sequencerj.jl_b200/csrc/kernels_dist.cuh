// K2/K3/K4: pairwise distance tile kernels.
// Replace `abs.(pairwise(alg, m; dims=2)) .+ eps` (reference src/sequencer.jl:226) for the four
// ALL_METRICS instances (src/distancemetrics.jl:234-257):
//   Euclidean / Energy / KL : dense contractions -> FP64 tensor cores (DMMA.8x8x4), TMA-staged
//                             operands (SWIZZLE_128B), fused norm / sqrt / entropy epilogue.
//   EMD                     : L1 between CDFs, not a contraction -> CUDA-core FP64 (DADD x2 per
//                             element), same TMA pipeline, register-tiled 8x8 per thread.
// Both exploit symmetry: only tiles touching the upper triangle are computed, every value is
// stored at (i,j) and mirrored to (j,i) (the reference evaluates i>j and mirrors for SemiMetrics;
// for KL only the upper triangle is ever consumed, src/sequencer.jl:365).
//
// Tile = 64 (i) x 128 (j) x 16 (k) doubles, 4 warps (thread 0 also issues the TMA loads), 4 stages,
// 2 CTAs / SM so that one CTA's epilogue overlaps the other's main loop.
//
// Accuracy: Gram forms cancel when two objects are close.  Entries whose Gram value is small
// relative to the norms are recomputed directly (sum of squared differences / sum p log(p/q)) by a
// fix-up kernel, exactly the Distances.jl `Euclidean(thresh)` recipe (src/distancemetrics.jl:6,234)
// with a safer relative threshold, so results stay within 1e-9 of the reference's direct formulas
// (Energy: src/distancemetrics.jl:219; KL: Distances.KLDivergence).
#pragma once
#include "common.cuh"

namespace seqj {

enum DistMode { MODE_L2 = 0, MODE_EMD = 1, MODE_KL = 2, MODE_ENERGY = 3 };

struct DistParams {
  int N;
  int64_t ldD;          // leading dimension of every output matrix (doubles)
  int64_t strideD;      // elements between consecutive chunk matrices in the batch
  double* D;            // output of the first chunk in this launch
  int nk;               // k-stages per chunk (= cl_pad / 16)
  int cl_pad;
  int chunk0;           // first chunk of this launch (blockIdx.y = chunk - chunk0)
  const double* normA;  // [nchunks][ldn]: row norms (L2: sum p^2; Energy: sum Uc^2; KL: H)
  int64_t ldn;
  const int2* tiles;    // (i0, j0) of each CTA tile
  double eps_floor;
  double tau;           // relative fix-up threshold
  int mirror;           // 1: write (i,j) and (j,i) for i<j only;  0: write every (i,j) (KL full)
  // dense tiles: a tile whose epilogue flags more than `dense_thresh` entries is recomputed as a whole by the
  // direct tile kernel (direct_dist_kernel / kl_direct_kernel) instead of entry by entry
  unsigned char* tile_flags;   // [nb][tile_flag_stride], indexed by blockIdx.y, blockIdx.x; may be nullptr
  int tile_flag_stride;
  int dense_thresh;
  // direct kernels: with a dense list, a small persistent grid loops over the listed (chunk_local, tile) entries;
  // without one (EMD) every CTA computes the tile (blockIdx.x, blockIdx.y)
  const int2* dense_list;
  const unsigned int* dense_count;
  int kl_swap;          // KL full mode: operand A = log P, B = P, so that D[r][c] = KL(p_c || p_r), i.e. the
                        // row-major buffer is the column-major pairwise matrix of the C ABI
  // fix-up list
  int4* fix_list;       // (chunk, i, j, 0)
  unsigned int* fix_count;
  unsigned int fix_cap;
  int* fix_overflow;
  const double* period; // OP_PERIODIC: period of every row of the (single) chunk, padded to cl_pad with 1.0
  double out_scale = 1.0;  // OP_L1: factor on the sum (a power of two)
};

// rho: logical fragment row g -> physical row inside an 8-row group, chosen so that the two rows
// read by one quarter-warp differ in bit 2 (=> their SWIZZLE_128B chunks never collide).
__device__ __forceinline__ int rho(int g) { return (g >> 1) + ((g & 1) << 2); }

// sqrt(v) for the epilogue.  The FP64 units are shared between DMMA and scalar FP64 math
// (profiles/r01_fp64_mix.json), so every epilogue instruction is taken from the tile main loops; the CUDA
// library sqrt costs ~30 FP64-pipe instructions, this one 9: MUFU.RSQ64H seed (2^-22), one Newton step on
// 1/sqrt (-> 7e-14), one residual correction of the root (error squared again).  Within 1 ulp; v <= 0 gives 0.
__device__ __forceinline__ double fast_sqrt(double v) {
  double y;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(v));
  const double hv = 0.5 * v;
  y = y * fma(-hv * y, y, 1.5);
  double sq = v * y;
  sq = fma(fma(-sq, sq, v), 0.5 * y, sq);
  return v > 0.0 ? sq : 0.0;
}

template <int MODE, bool INTERIOR>
__device__ __forceinline__ void gemm_epilogue(const DistParams& p, double (&acc)[8][4][2], int i0, int j0, int chunk,
                                              int warp, int rg, int t, int* s_cnt) {
  const double* __restrict__ nrm = p.normA + (int64_t)chunk * p.ldn;
  double* __restrict__ D = p.D + (int64_t)blockIdx.y * p.strideD;
  const int N = p.N;
  const int jbase = j0 + 32 * warp + t;
  double nj[4][2];
#pragma unroll
  for (int b = 0; b < 4; ++b)
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      const int j = jbase + 8 * b + 4 * e;
      nj[b][e] = (INTERIOR || j < N) ? nrm[j] : 0.0;
    }
  // One pass: every entry is finished and stored (direct + mirrored) as soon as it is computed, so that the
  // accumulators die row by row and the epilogue needs no registers beyond the main loop's (168 per thread = three
  // CTAs per SM for every metric; the two-phase version kept all 64 distances alive across the tile-level decision and
  // spilled ~1 KB per thread for L2 / Energy there).  Only the flag bits survive; flagged entries are listed afterwards.
  unsigned long long flags = 0ull;
#pragma unroll
  for (int a = 0; a < 8; ++a) {
    const int i = i0 + 8 * a + rg;
    const bool irow = INTERIOR || i < N;
    const double ni = irow ? nrm[i] : 0.0;
    double* __restrict__ drow = D + (int64_t)i * p.ldD + jbase;          // direct stores: drow[8b + 4e]
    double* __restrict__ dcol = D + (int64_t)jbase * p.ldD + i;          // mirrored stores: dcol[(8b + 4e) * ldD]
#pragma unroll
    for (int b = 0; b < 4; ++b) {
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int j = jbase + 8 * b + 4 * e;
        const bool live = INTERIOR || (i < N && j < N && i != j && !(p.mirror && i > j));
        const double c = acc[a][b][e];
        bool flag;
        double d;
        if (MODE == MODE_KL) {
          const double hh = p.kl_swap ? nj[b][e] : ni;
          d = hh - c;                                   // H_i - sum_k p_ik log p_jk
          flag = d < p.tau * fabs(hh);
        } else {
          const double selfterms = ni + nj[b][e];
          const double v = selfterms - 2.0 * c;
          flag = v < p.tau * selfterms;
          d = fast_sqrt(v);
          if (MODE == MODE_ENERGY) d = 1.4142135623730951 * d;   // sqrt(2) * sqrt(sum)
        }
        if (live && flag) flags |= 1ull << (a * 8 + b * 2 + e);
        if (!INTERIOR) {
          if (!irow || j >= N) continue;
          if (p.mirror && i > j) continue;
        }
        const double out = (!INTERIOR && i == j) ? p.eps_floor : fabs(d) + p.eps_floor;   // pairwise sets the diagonal to 0, then .+ eps
        drow[8 * b + 4 * e] = out;
        if (INTERIOR || (p.mirror && i != j)) dcol[(int64_t)(8 * b + 4 * e) * p.ldD] = out;
      }
    }
  }
  // tile-level decision: many flagged entries (near-duplicate objects) -> the whole tile goes to the direct kernel
  bool dense = false;
  if (p.tile_flags != nullptr) {
    if (flags) atomicAdd(s_cnt, __popcll(flags));
    __syncthreads();
    dense = *s_cnt > p.dense_thresh;
    if (dense && threadIdx.x == 0) p.tile_flags[(size_t)blockIdx.y * p.tile_flag_stride + blockIdx.x] = 1;
  }
  // per-entry fix-up list when the tile is not dense (rare: nothing is flagged on unstructured data)
  if (!dense && flags) {
#pragma unroll 1
    for (unsigned long long f = flags; f; f &= f - 1) {
      const int bit = __ffsll((long long)f) - 1;
      const int a = bit >> 3, b = (bit >> 1) & 3, e = bit & 1;
      const int i = i0 + 8 * a + rg, j = jbase + 8 * b + 4 * e;
      const unsigned int idx = atomicAdd(p.fix_count, 1u);
      if (idx < p.fix_cap) {
        p.fix_list[idx] = make_int4((int)blockIdx.y, i, j, p.kl_swap);
      } else {
        *p.fix_overflow = 1;                                // marker for the dense fix-up pass
        D[(int64_t)i * p.ldD + j] = -1.0;
        if (INTERIOR || (p.mirror && i != j)) D[(int64_t)j * p.ldD + i] = -1.0;
      }
    }
  }
}

template <int MODE>
// KL runs 3 CTAs per SM (168 registers, the few spills land outside the DMMA loop: measured 13.6 -> 12.4 ms at config 3);
// the sqrt epilogue of L2 / Energy spills more at 168 and is faster at 2 CTAs (14.2 vs 14.9 ms).
__global__ void __launch_bounds__(kTileThreads, 3)
gemm_dist_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, DistParams p) {
  __shared__ int s_cnt;
  extern __shared__ uint8_t smem_raw[];
  const TileSmem s = carve_tile_smem(smem_raw);
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int2 tile = p.tiles[blockIdx.x];
  const int i0 = tile.x, j0 = tile.y;
  const int chunk = p.chunk0 + blockIdx.y;
  const int kbase = chunk * p.cl_pad;

  if (threadIdx.x == 0) s_cnt = 0;
  if (threadIdx.x == 0) {
    tile_barriers_init(s);
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
  }
  __syncthreads();

  if (threadIdx.x == 0)
    for (int it = 0; it < min(kStages, p.nk); ++it) tile_issue(&tmA, &tmB, s, it, kbase, i0, j0);

  // ---- consumers: warp w owns columns [32w, 32w+32) of the 64 x 128 tile ----
  const int g = lane >> 2, t = lane & 3;
  const int rg = rho(g);
  // byte offset of this lane's 16 B chunk for k-half p (chunks 4p+t), swizzled by the row
  uint32_t off[2];
  off[0] = rg * 128 + (((0 + t) ^ rg) << 4);
  off[1] = rg * 128 + (((4 + t) ^ rg) << 4);

  double acc[8][4][2];
#pragma unroll
  for (int a = 0; a < 8; ++a)
#pragma unroll
    for (int b = 0; b < 4; ++b) acc[a][b][0] = acc[a][b][1] = 0.0;

  const uint32_t stages_u32 = smem_u32(s.stages);
  for (int it = 0; it < p.nk; ++it) {
    const int st = it % kStages;
    const uint32_t ph = (it / kStages) & 1;
    mbar_wait(&s.full[st], ph);
    tile_release_refill(&tmA, &tmB, s, it, p.nk, kbase, i0, j0);
    const uint32_t sa = stages_u32 + st * kStageBytes;
    const uint32_t sb = sa + kStageBytesA + warp * (32 * 128);
#pragma unroll
    for (int half = 0; half < 2; ++half) {
      double2 fa[8], fb[4];
#pragma unroll
      for (int a = 0; a < 8; ++a) fa[a] = lds128(sa + a * 1024 + off[half]);
#pragma unroll
      for (int b = 0; b < 4; ++b) fb[b] = lds128(sb + b * 1024 + off[half]);
#pragma unroll
      for (int a = 0; a < 8; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) dmma884(acc[a][b][0], acc[a][b][1], fa[a].x, fb[b].x);
#pragma unroll
      for (int a = 0; a < 8; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) dmma884(acc[a][b][0], acc[a][b][1], fa[a].y, fb[b].y);
    }
  }

  // ---- fused epilogue ----
  // acc[a][b][e] = C[i][j] with i = i0 + 8a + rho(g), j = j0 + 32*warp + 8b + t + 4e
  // Interior tiles (entirely above the diagonal and inside the matrix: ~95 % of them) skip every per-entry
  // bounds / triangle test; the general path handles the diagonal band and the ragged edge.
  const bool interior = p.mirror && (i0 + kTileM <= j0) && (j0 + kTileN <= p.N);
  if (interior) gemm_epilogue<MODE, true>(p, acc, i0, j0, chunk, warp, rg, t, &s_cnt);
  else gemm_epilogue<MODE, false>(p, acc, i0, j0, chunk, warp, rg, t, &s_cnt);
}

// flags -> list of (chunk_local, tile) entries for the persistent direct kernels
__global__ void compact_flags_kernel(const unsigned char* __restrict__ flags, int total, int stride, int2* list,
                                     unsigned int* count) {
  const int q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q < total && flags[q]) list[atomicAdd(count, 1u)] = make_int2(q / stride, q % stride);
}

// ---- EMD: sum_k |U_ik - U_jk| on CUDA cores -------------------------------------------------
// thread T of the 128 consumers: ti = T>>4 (rows 8ti..8ti+7), tj = T&15 (cols (tj&7)+8(tj>>3)+16bb).
// OP_L1: EMD.  OP_SQ / OP_SQ_ENERGY: sum_k (a-b)^2 -> Euclidean / Energy computed WITHOUT the Gram form; used
// as the fix-up of whole tiles that the DMMA epilogue (or derive_kernel) found full of cancelling entries.
enum DirectOp { OP_L1 = 0, OP_SQ = 1, OP_SQ_ENERGY = 2, OP_PERIODIC = 3, OP_MAX = 4, OP_SQ_RAW = 5 };
// OP_MAX: max_k |a - b| (Chebyshev); OP_SQ_RAW: sum (a - b)^2 without the root (SqEuclidean); OP_L1 with out_scale = 0.5
// is TotalVariation, OP_L1 on the PDFs Cityblock.

// |a - b| mod p folded onto [0, p/2] (Distances.PeriodicEuclidean eval_op): the mod is exact (fmod) and only
// taken when the difference reaches the period.
__device__ __noinline__ double periodic_wrap_slow(double s, double per) { return fmod(s, per); }
__device__ __forceinline__ double periodic_fold(double d, double per) {
  double s = fabs(d);
  if (s >= per) s = periodic_wrap_slow(s, per);
  return fmin(s, per - s);
}

template <int OP>
__global__ void __launch_bounds__(kTileThreads, 2)
direct_dist_kernel(const __grid_constant__ CUtensorMap tmU, DistParams p) {
  extern __shared__ uint8_t smem_raw[];
  const TileSmem s = carve_tile_smem(smem_raw);
  const int lane = threadIdx.x & 31;
  const bool listed = p.dense_list != nullptr;
  const int ntodo = listed ? (int)*p.dense_count : 1;
  const int T = threadIdx.x;
  const int ti = T >> 4, tj = T & 15;
  const int jrow0 = (tj & 7) + 8 * (tj >> 3);   // + 16*bb ; (row & 7) == (tj & 7)
  const int jsw = tj & 7;
  const uint32_t stages_u32 = smem_u32(s.stages);
  const int N = p.N;

  for (int e = listed ? (int)blockIdx.x : 0; e < ntodo; e += listed ? (int)gridDim.x : 1) {
    int tile_idx = blockIdx.x, chunk_local = blockIdx.y;
    if (listed) { const int2 ent = p.dense_list[e]; chunk_local = ent.x; tile_idx = ent.y; }
    const int2 tile = p.tiles[tile_idx];
    const int i0 = tile.x, j0 = tile.y;
    const int kbase = (p.chunk0 + chunk_local) * p.cl_pad;

    if (threadIdx.x == 0) {
      tile_barriers_init(s);
      tma_prefetch_desc(&tmU);
    }
    __syncthreads();
    if (threadIdx.x == 0)
      for (int it = 0; it < min(kStages, p.nk); ++it) tile_issue(&tmU, &tmU, s, it, kbase, i0, j0);

    double acc[8][8];
#pragma unroll
    for (int a = 0; a < 8; ++a)
#pragma unroll
      for (int b = 0; b < 8; ++b) acc[a][b] = 0.0;

    for (int it = 0; it < p.nk; ++it) {
      const int st = it % kStages;
      const uint32_t ph = (it / kStages) & 1;
      mbar_wait(&s.full[st], ph);
      tile_release_refill(&tmU, &tmU, s, it, p.nk, kbase, i0, j0);
      const uint32_t sa = stages_u32 + st * kStageBytes + (8 * ti) * 128;
      const uint32_t sb = stages_u32 + st * kStageBytes + kStageBytesA + jrow0 * 128;
      // The j-side fragments of chunk c + 1 are loaded while chunk c is being accumulated (two register sets,
      // fully unrolled so that the set index is static): without this every chunk started with nine back-to-back
      // LDS.128 and ~60 cycles in which neither of the SMSP's two warps had a DADD to issue (ncu, end of r01).
      double2 fb[2][8];
#pragma unroll
      for (int b = 0; b < 8; ++b) fb[0][b] = lds128(sb + b * (16 * 128) + ((0 ^ jsw) << 4));
#pragma unroll
      for (int c = 0; c < 8; ++c) {    // 16 B chunk = 2 consecutive k
        const int cur = c & 1;
        if (c + 1 < 8) {
#pragma unroll
          for (int b = 0; b < 8; ++b) fb[cur ^ 1][b] = lds128(sb + b * (16 * 128) + (((c + 1) ^ jsw) << 4));
        }
        double2 per = make_double2(1.0, 1.0);
        if (OP == OP_PERIODIC) per = __ldg(reinterpret_cast<const double2*>(p.period + it * kTileK + 2 * c));
#pragma unroll
        for (int a = 0; a < 8; ++a) {
          const double2 fa = lds128(sa + a * 128 + ((c ^ a) << 4));
#pragma unroll
          for (int b = 0; b < 8; ++b) {
            if (OP == OP_L1) {
              acc[a][b] += fabs(fa.x - fb[cur][b].x);
              acc[a][b] += fabs(fa.y - fb[cur][b].y);
            } else if (OP == OP_MAX) {
              acc[a][b] = fmax(acc[a][b], fmax(fabs(fa.x - fb[cur][b].x), fabs(fa.y - fb[cur][b].y)));
            } else if (OP == OP_PERIODIC) {
              const double sx = periodic_fold(fa.x - fb[cur][b].x, per.x), sy = periodic_fold(fa.y - fb[cur][b].y, per.y);
              acc[a][b] = fma(sx, sx, acc[a][b]);
              acc[a][b] = fma(sy, sy, acc[a][b]);
            } else {
              const double dx = fa.x - fb[cur][b].x, dy = fa.y - fb[cur][b].y;
              acc[a][b] = fma(dx, dx, acc[a][b]);
              acc[a][b] = fma(dy, dy, acc[a][b]);
            }
          }
        }
      }
    }

    double* __restrict__ D = p.D + (int64_t)chunk_local * p.strideD;
#pragma unroll
    for (int a = 0; a < 8; ++a) {
      const int i = i0 + 8 * ti + a;
      if (i >= N) continue;
#pragma unroll
      for (int b = 0; b < 8; ++b) {
        const int j = j0 + jrow0 + 16 * b;
        if (j >= N || i > j) continue;
        double d = acc[a][b];
        if (OP != OP_L1 && OP != OP_MAX && OP != OP_SQ_RAW) d = fast_sqrt(d);
        if (OP == OP_SQ_ENERGY) d = 1.4142135623730951 * d;
        if (OP == OP_L1) d *= p.out_scale;                               // 1 (EMD, Cityblock) or 1/2 (TotalVariation): exact
        const double out = (i == j) ? p.eps_floor : fabs(d) + p.eps_floor;
        D[(int64_t)i * p.ldD + j] = out;
        if (i != j) D[(int64_t)j * p.ldD + i] = out;
      }
    }
    if (listed) __syncthreads();      // every warp is done with the pipeline before the barriers are re-initialised
  }
}

// ---- KL on whole tiles without the entropy-minus-cross-term form -----------------------------------
// KL(p_i || p_j) = sum_k p_ik (log p_ik - log p_jk): the per-term difference is small when the objects are close,
// so there is no cancellation of large sums.  Operands per stage: P and log P for the 64 rows i, log P for the
// 128 rows j (32 KB, 3 stages, 2 CTAs/SM).  Same 8 x 8 register tile as the EMD kernel; 2 FP64 instructions per
// element (DADD + DFMA).  Used only for tiles flagged dense by the DMMA epilogue / derive_kernel.
constexpr int kKlStages = 3;
constexpr int kKlStageBytes = 2 * kStageBytesA + kStageBytesB;          // P_i, logP_i (8 KB each) + logP_j (16 KB)
constexpr int kKlSmemBytes = kKlStages * kKlStageBytes + 1024 + 256;

__device__ __forceinline__ void kl_issue(const CUtensorMap* tmP, const CUtensorMap* tmL, uint8_t* stages, uint64_t* full,
                                         int it, int kbase, int i0, int j0) {
  const int st = it % kKlStages;
  mbar_expect_tx(&full[st], kKlStageBytes);
  uint8_t* base = stages + st * kKlStageBytes;
  const int k0 = kbase + it * kTileK;
  tma_load_2d(base, tmP, &full[st], k0, i0);
  tma_load_2d(base + kStageBytesA, tmL, &full[st], k0, i0);
  tma_load_2d(base + 2 * kStageBytesA, tmL, &full[st], k0, j0);
  tma_load_2d(base + 3 * kStageBytesA, tmL, &full[st], k0, j0 + kTileM);
}

__global__ void __launch_bounds__(kTileThreads, 2)
kl_direct_kernel(const __grid_constant__ CUtensorMap tmP, const __grid_constant__ CUtensorMap tmL, DistParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* stages = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full = reinterpret_cast<uint64_t*>(stages + kKlStages * kKlStageBytes);
  uint64_t* empty = full + kKlStages;
  const int lane = threadIdx.x & 31;
  const int ntodo = (int)*p.dense_count;
  const int T = threadIdx.x;
  const int ti = T >> 4, tj = T & 15;
  const int jrow0 = (tj & 7) + 8 * (tj >> 3);
  const int jsw = tj & 7;
  const uint32_t stages_u32 = smem_u32(stages);
  const int N = p.N;

  for (int e = blockIdx.x; e < ntodo; e += gridDim.x) {
    const int2 ent = p.dense_list[e];
    const int chunk_local = ent.x;
    const int2 tile = p.tiles[ent.y];
    const int i0 = tile.x, j0 = tile.y;
    const int kbase = (p.chunk0 + chunk_local) * p.cl_pad;
    if (threadIdx.x == 0) {
      for (int q = 0; q < kKlStages; ++q) { mbar_init(&full[q], 1); mbar_init(&empty[q], kConsumerWarps); }
      fence_barrier_init();
      tma_prefetch_desc(&tmP);
      tma_prefetch_desc(&tmL);
    }
    __syncthreads();
    if (threadIdx.x == 0)
      for (int it = 0; it < min(kKlStages, p.nk); ++it) kl_issue(&tmP, &tmL, stages, full, it, kbase, i0, j0);

    double acc[8][8];
#pragma unroll
    for (int a = 0; a < 8; ++a)
#pragma unroll
      for (int b = 0; b < 8; ++b) acc[a][b] = 0.0;

    for (int it = 0; it < p.nk; ++it) {
      const int st = it % kKlStages;
      const uint32_t ph = (it / kKlStages) & 1;
      mbar_wait(&full[st], ph);
      {                                            // release of slot it-1 trails its reads by one wait: see tile_release_refill
        const int nxt = it - 1 + kKlStages;
        if (it >= 1 && nxt < p.nk) {
          __syncwarp();
          if (lane == 0) mbar_arrive(&empty[(it - 1) % kKlStages]);
          if (threadIdx.x == 0) {
            mbar_wait(&empty[(it - 1) % kKlStages], ((it - 1) / kKlStages) & 1);
            kl_issue(&tmP, &tmL, stages, full, nxt, kbase, i0, j0);
          }
        }
      }
      const uint32_t sp = stages_u32 + st * kKlStageBytes + (8 * ti) * 128;       // P rows i
      const uint32_t sl = sp + kStageBytesA;                                        // log P rows i
      const uint32_t sb = stages_u32 + st * kKlStageBytes + 2 * kStageBytesA + jrow0 * 128;   // log P rows j
#pragma unroll 2
      for (int c = 0; c < 8; ++c) {
        double2 fb[8];
#pragma unroll
        for (int b = 0; b < 8; ++b) fb[b] = lds128(sb + b * (16 * 128) + ((c ^ jsw) << 4));
#pragma unroll
        for (int a = 0; a < 8; ++a) {
          const uint32_t off = a * 128 + ((c ^ a) << 4);
          const double2 pa = lds128(sp + off);
          const double2 la = lds128(sl + off);
#pragma unroll
          for (int b = 0; b < 8; ++b) {
            acc[a][b] = fma(pa.x, la.x - fb[b].x, acc[a][b]);
            acc[a][b] = fma(pa.y, la.y - fb[b].y, acc[a][b]);
          }
        }
      }
    }

    double* __restrict__ D = p.D + (int64_t)chunk_local * p.strideD;
#pragma unroll
    for (int a = 0; a < 8; ++a) {
      const int i = i0 + 8 * ti + a;
      if (i >= N) continue;
#pragma unroll
      for (int b = 0; b < 8; ++b) {
        const int j = j0 + jrow0 + 16 * b;
        if (j >= N || i > j) continue;           // mirrored storage: the stored value is KL(p_i || p_j), i < j
        const double out = (i == j) ? p.eps_floor : fabs(acc[a][b]) + p.eps_floor;
        D[(int64_t)i * p.ldD + j] = out;
        if (i != j) D[(int64_t)j * p.ldD + i] = out;
      }
    }
    __syncthreads();      // every warp is done with the pipeline before the barriers are re-initialised
  }
}

// ---- fix-up: direct recomputation of flagged entries (one warp per entry) ---------------------
struct FixParams {
  int mode;
  int N;
  int64_t ldD, strideD;
  double* D;
  const double* X;      // L2: P ; Energy: Uc ; KL: P
  int64_t ldX;
  int cl_pad, chunk0;
  const int* m_of_chunk;   // valid length of every chunk
  double eps_floor;
  int mirror;
  const int4* fix_list;
  const unsigned int* fix_count;
  unsigned int fix_cap;
  const int* fix_overflow;
  int kl_swap;
};

__device__ __forceinline__ double direct_pair(int mode, const double* __restrict__ xi, const double* __restrict__ xj,
                                              int m, int lane) {
  double sacc = 0.0;
  if (mode == MODE_KL) {
    for (int k = lane; k < m; k += 32) {
      const double a = xi[k], b = xj[k];
      sacc += a * log(a / b);               // Distances.KLDivergence term
    }
    return warp_sum(sacc);
  }
  for (int k = lane; k < m; k += 32) {
    const double d = xi[k] - xj[k];
    sacc += d * d;
  }
  sacc = warp_sum(sacc);
  double d = sqrt(sacc);
  if (mode == MODE_ENERGY) d = 1.4142135623730951 * d;
  return d;
}

__global__ void __launch_bounds__(256) fixup_list_kernel(FixParams p) {
  const int lane = threadIdx.x & 31;
  const unsigned int nwarps = gridDim.x * (blockDim.x >> 5);
  const unsigned int count = min(*p.fix_count, p.fix_cap);
  for (unsigned int e = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); e < count; e += nwarps) {
    const int4 f = p.fix_list[e];
    const int chunk = p.chunk0 + f.x;
    const int m = p.m_of_chunk[chunk];
    const double* xr = p.X + (int64_t)f.y * p.ldX + (int64_t)chunk * p.cl_pad;
    const double* xc = p.X + (int64_t)f.z * p.ldX + (int64_t)chunk * p.cl_pad;
    const double d = f.w ? direct_pair(p.mode, xc, xr, m, lane) : direct_pair(p.mode, xr, xc, m, lane);
    if (lane == 0) {
      const double out = fabs(d) + p.eps_floor;
      double* D = p.D + (int64_t)f.x * p.strideD;
      D[(int64_t)f.y * p.ldD + f.z] = out;
      if (p.mirror) D[(int64_t)f.z * p.ldD + f.y] = out;
    }
  }
}

// Dense fallback, only does work when the list overflowed: every entry marked -1 is recomputed.
__global__ void __launch_bounds__(256) fixup_dense_kernel(FixParams p, int nchunks_batch) {
  if (*p.fix_overflow == 0) return;
  const int lane = threadIdx.x & 31;
  const int64_t nwarps = (int64_t)gridDim.x * (blockDim.x >> 5);
  const int64_t total = (int64_t)nchunks_batch * p.N;
  for (int64_t r = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); r < total; r += nwarps) {
    const int cb = (int)(r / p.N);
    const int i = (int)(r % p.N);
    const int chunk = p.chunk0 + cb;
    const int m = p.m_of_chunk[chunk];
    double* Drow = p.D + (int64_t)cb * p.strideD + (int64_t)i * p.ldD;
    const double* xi = p.X + (int64_t)i * p.ldX + (int64_t)chunk * p.cl_pad;
    for (int jb = 0; jb < p.N; jb += 32) {
      const int j = jb + lane;
      const bool marked = (j < p.N) && (Drow[j] < 0.0) && (!p.mirror || i < j);
      unsigned int mask = __ballot_sync(0xffffffffu, marked);
      while (mask) {
        const int l = __ffs(mask) - 1;
        mask &= mask - 1;
        const int jj = jb + l;
        const double* xj = p.X + (int64_t)jj * p.ldX + (int64_t)chunk * p.cl_pad;
        const double d = p.kl_swap ? direct_pair(p.mode, xj, xi, m, lane) : direct_pair(p.mode, xi, xj, m, lane);
        if (lane == 0) {
          const double out = fabs(d) + p.eps_floor;
          Drow[jj] = out;
          if (p.mirror) p.D[(int64_t)cb * p.strideD + (int64_t)jj * p.ldD + i] = out;
        }
      }
      __syncwarp();
    }
  }
}

}  // namespace seqj
