// Shared device helpers: mbarrier / TMA (cp.async.bulk.tensor) wrappers, DMMA, warp reductions.
// sm_100a only.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace seqj {

constexpr int kTileM = 64;        // rows (objects i) per CTA tile
constexpr int kTileN = 128;       // cols (objects j) per CTA tile
constexpr int kTileK = 16;        // doubles per smem row = 128 B (SWIZZLE_128B atom)
constexpr int kStages = 3;     // 3 x 24 KB + barriers = 75 KB per CTA: three CTAs fit in an SM's 228 KB (KL tiles); 4 stages measured no faster
constexpr int kStageBytesA = kTileM * kTileK * 8;   // 8 KB
constexpr int kStageBytesB = kTileN * kTileK * 8;   // 16 KB
constexpr int kStageBytes = kStageBytesA + kStageBytesB;
constexpr int kConsumerWarps = 4;
constexpr int kTileThreads = kConsumerWarps * 32;   // 4 warps; thread 0 also issues the TMA loads (a separate
                                                    // producer warp would put 3 warps on one SM sub-partition at
                                                    // 2 CTAs/SM and cap registers at 168/thread -> spills)
constexpr int kTileSmemBytes = kStages * kStageBytes + 1024 /*align slack*/ + 256 /*barriers*/;

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_barrier_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}\n"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
// Bounded wait: a pipeline bug traps (kernel error) instead of hanging the GPU box.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  long long t0 = clock64();
  while (!mbar_try_wait(bar, parity)) {
    if (clock64() - t0 > 8000000000LL) __trap();
  }
}
__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* tmap, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"(tmap), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* tmap) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(tmap) : "memory");
}

// FP64 tensor-core primitive: DMMA.8x8x4 (the only FP64 MMA shape in sm_100a SASS; m16n8k{4,8,16}
// decompose into it, see profiles/r01_fp64_peaks.json).
// A frag: lane (g = lane>>2, t = lane&3) holds A[g][t]; B frag: B[t][g]; C frag: C[g][2t], C[g][2t+1].
__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(c0), "+d"(c1)
               : "d"(a), "d"(b));
}

__device__ __forceinline__ double2 lds128(uint32_t addr) {
  double2 v;
  asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"(addr));
  return v;
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Carve the dynamic smem of a tile kernel: stages (1024 B aligned for SWIZZLE_128B) then barriers.
struct TileSmem {
  uint8_t* stages;
  uint64_t* full;
  uint64_t* empty;
};
__device__ __forceinline__ TileSmem carve_tile_smem(uint8_t* raw) {
  TileSmem s;
  uintptr_t p = (reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023);
  s.stages = reinterpret_cast<uint8_t*>(p);
  s.full = reinterpret_cast<uint64_t*>(s.stages + kStages * kStageBytes);
  s.empty = s.full + kStages;
  return s;
}

// Producer side of the TMA pipeline shared by the DMMA and the EMD tile kernels (one elected thread).
// Loads k-stage `it`: the 64 x 16 A tile (rows i0..) and the 128 x 16 B tile (rows j0..) of the
// K-contiguous operand matrices into pipeline slot it % kStages.
__device__ __forceinline__ void tile_issue(const CUtensorMap* tmA, const CUtensorMap* tmB, const TileSmem& s,
                                           int it, int kbase, int i0, int j0) {
  const int st = it % kStages;
  mbar_expect_tx(&s.full[st], kStageBytes);
  uint8_t* sa = s.stages + st * kStageBytes;
  uint8_t* sb = sa + kStageBytesA;
  const int k0 = kbase + it * kTileK;
  tma_load_2d(sa, tmA, &s.full[st], k0, i0);
  tma_load_2d(sb, tmB, &s.full[st], k0, j0);
  tma_load_2d(sb + kStageBytesA, tmB, &s.full[st], k0, j0 + kTileM);
}
// Consumer release + producer refill, called by EVERY consumer thread (warp-converged) of iteration `it` right after
// its wait on full[it % kStages]: the warp releases the slot it consumed in iteration it-1 and the elected thread
// refills it with k-stage it-1+kStages.
// Why the release trails by one wait: ptxas schedules `mbarrier.arrive` directly behind the last `ld.shared` of the
// iteration, ahead of the math instructions that consume the loaded fragments, and attaches no scoreboard wait to it
// (SASS: LDS.128 x3, SYNCS.ARRIVE, DMMA ...).  The arrive can then be performed while those loads are still queued,
// and once the last warp has arrived the refill's TMA write may land in the slot before they are served: the last
// fragment rows of a k-stage were occasionally read from the NEXT k-stage (seen on B200 in the first call of a
// process, GPU clocks still ramping: a few 24 x 32 blocks of a 10k x 10k Euclidean matrix off by ~2 %;
// tools/determinism_stress.py, profiles/r02_pipeline_race.md).  Placed after the next full-wait, the arrive follows a
// spin loop that itself follows every math instruction of iteration it-1 in the instruction stream, and those were
// only issued once their operands had arrived from shared memory.
__device__ __forceinline__ void tile_release_refill(const CUtensorMap* tmA, const CUtensorMap* tmB, const TileSmem& s,
                                                    int it, int nk, int kbase, int i0, int j0) {
  const int nxt = it - 1 + kStages;
  if (it >= 1 && nxt < nk) {                    // slots that are never refilled need no release
    __syncwarp();
    if ((threadIdx.x & 31) == 0) mbar_arrive(&s.empty[(it - 1) % kStages]);
    if (threadIdx.x == 0) {
      mbar_wait(&s.empty[(it - 1) % kStages], ((it - 1) / kStages) & 1);
      tile_issue(tmA, tmB, s, nxt, kbase, i0, j0);
    }
  }
}

__device__ __forceinline__ void tile_barriers_init(const TileSmem& s) {
  for (int i = 0; i < kStages; ++i) {
    mbar_init(&s.full[i], 1);
    mbar_init(&s.empty[i], kConsumerWarps);
  }
  fence_barrier_init();
}

}  // namespace seqj
