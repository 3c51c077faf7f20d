// C ABI + host orchestration of the B200-native Sequencer hot path (see include/seqj.h).
// Replaces the reference's `_sequence` driver loop (src/sequencer.jl:184-276): scale-major
// iteration, per (metric, scale) group: distance tile kernels over a batch of chunks -> batched
// Prim -> tree measurement (eta) -> eta-weighted combine -> group Prim/tree; then the proximity
// fuse and the final MST / order.  Everything is asynchronous on one stream: eta values never
// leave the device between kernels, and the host only synchronises to read the final results.
#include <algorithm>
#include <cmath>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <condition_variable>
#include <functional>
#include <map>
#include <memory>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "../../include/seqj.h"
#include "common.cuh"
#include "kernels_dist.cuh"
#include "kernels_derive.cuh"
#include "kernels_graph.cuh"
#include "kernels_mst.cuh"
#include "kernels_pair.cuh"
#include "kernels_prep.cuh"

using namespace seqj;

#include "host_upload.cuh"
#include "host_core.cuh"
#include "host_plan.cuh"
#include "host_shard.cuh"
#include "host_stages.cuh"


// --------------------------------------------------------------------------------------------
extern "C" {

int seqj_version(void) { return SEQJ_VERSION; }

static int create_one(seqj_handle** out, const int* devices, int ndev);

int seqj_create(seqj_handle** out, const int* devices, int ndev) {
  // devices[0] is the primary device (final fuse, results); devices[1..] become peer handles, each driven by
  // its own host thread inside seqj_sequence (single-process multi-GPU for the Julia host)
  int st = create_one(out, devices, ndev);
  if (st != SEQJ_OK || !devices || ndev <= 1) return st;
  for (int d = 1; d < ndev; ++d) {
    seqj_handle* p = nullptr;
    st = create_one(&p, devices + d, 1);
    if (st != SEQJ_OK) {
      (*out)->err = "peer device " + std::to_string(devices[d]) + ": " + (p ? p->err : std::string("create failed"));
      if (p) seqj_destroy(p);
      return st;
    }
    (*out)->peers.push_back(p);
  }
  // peer access between every pair of the handle's devices (NVLink loads / stores and direct peer copies); a pair
  // that cannot be mapped still works, cudaMemcpyPeerAsync then stages through the host
  for (int a = 0; a < ndev; ++a)
    for (int b = 0; b < ndev; ++b) {
      if (a == b) continue;
      int can = 0;
      if (cudaDeviceCanAccessPeer(&can, devices[a], devices[b]) == cudaSuccess && can) {
        cudaSetDevice(devices[a]);
        if (cudaDeviceEnablePeerAccess(devices[b], 0) != cudaSuccess) cudaGetLastError();   // already enabled: fine
      }
    }
  cudaSetDevice((*out)->dev);
  return SEQJ_OK;
}

static int create_one(seqj_handle** out, const int* devices, int ndev) {
  if (!out) return SEQJ_E_BADARG;
  *out = nullptr;
  seqj_handle* h = new seqj_handle();
  int count = 0;
  cudaError_t e = cudaGetDeviceCount(&count);
  if (e != cudaSuccess || count == 0) {
    // keep the handle so the caller can read the message; every compute call will fail loudly
    h->err = std::string("no CUDA device: ") + cudaGetErrorString(e);
    *out = h;
    return SEQJ_E_CUDA;
  }
  if (devices && ndev > 0) h->dev = devices[0];
  else cudaGetDevice(&h->dev);
  cudaDeviceProp prop;
  if (cudaSetDevice(h->dev) != cudaSuccess || cudaGetDeviceProperties(&prop, h->dev) != cudaSuccess) {
    h->err = "cudaSetDevice failed";
    *out = h;
    return SEQJ_E_CUDA;
  }
  if (prop.major != 10) {
    h->err = std::string("device is not sm_100 (Blackwell B200): ") + prop.name;
    *out = h;
    return SEQJ_E_CUDA;
  }
  h->sm_count = prop.multiProcessorCount;
  h->max_smem = prop.sharedMemPerBlockOptin;
  int prio_lo = 0, prio_hi = 0;
  cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi);
  const char* pr = getenv("SEQJ_PRIO");
  const bool use_prio = pr ? atoi(pr) != 0 : false;
  cudaStreamCreateWithPriority(&h->st, cudaStreamNonBlocking, use_prio ? prio_lo : 0);
  cudaStreamCreateWithPriority(&h->st2, cudaStreamNonBlocking, use_prio ? prio_hi : 0);
  cudaHostAlloc(&h->stats_pinned, 64, cudaHostAllocDefault);
  cudaEventCreateWithFlags(&h->stats_ev, cudaEventDisableTiming);
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult q;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q) != cudaSuccess || !fn) {
    h->err = "cuTensorMapEncodeTiled not available in this driver";
    *out = h;
    return SEQJ_E_CUDA;
  }
  h->encode = reinterpret_cast<EncodeTiledFn>(fn);
  *out = h;
  return set_kernel_attrs(h);
}

void seqj_destroy(seqj_handle* h) {
  if (!h) return;
  for (DeviceWorker* w : h->workers) delete w;          // joins the worker threads before their devices go away
  h->workers.clear();
  for (seqj_handle* p : h->peers) seqj_destroy(p);
  h->peers.clear();
  cudaSetDevice(h->dev);
  if (h->st) cudaStreamSynchronize(h->st);
  if (h->st2) cudaStreamSynchronize(h->st2);
  release_cache(h);
  if (h->st) cudaStreamDestroy(h->st);
  if (h->st2) cudaStreamDestroy(h->st2);
  upload_stage_free(h->upload);
  if (h->stats_pinned) cudaFreeHost(h->stats_pinned);
  if (h->stats_ev) cudaEventDestroy(h->stats_ev);
  delete h;
}

const char* seqj_last_error(const seqj_handle* h) { return h ? h->err.c_str() : "null handle"; }

int seqj_release_workspace(seqj_handle* h) {
  if (!h) return SEQJ_E_BADARG;
  std::lock_guard<std::recursive_mutex> lock__(h->mu);      // one call at a time per handle
  if (!h) return SEQJ_E_BADARG;
  if (h->st) cudaStreamSynchronize(h->st);
  release_cache(h);
  return SEQJ_OK;
}

int seqj_set_grid(seqj_handle* h, const double* grid, int64_t M) {
  if (!h) return SEQJ_E_BADARG;
  std::lock_guard<std::recursive_mutex> lock__(h->mu);      // one call at a time per handle
  if (!h) return SEQJ_E_BADARG;
  h->grid.clear();
  if (!grid || M <= 0) return SEQJ_OK;
  double prev = 0.0;
  for (int64_t i = 0; i < M; ++i) {
    if (!(grid[i] > prev) || !std::isfinite(grid[i]))
      return fail(h, SEQJ_E_BADARG, "grid must be finite, positive and strictly increasing (the same-grid shortcut of "
                                    "_cdf_distance, src/distancemetrics.jl:206-209, needs non-zero deltas)");
    prev = grid[i];
  }
  h->grid.assign(grid, grid + M);
  return SEQJ_OK;
}

int seqj_set_periods(seqj_handle* h, const double* periods, int64_t M) {
  if (!h) return SEQJ_E_BADARG;
  std::lock_guard<std::recursive_mutex> lock__(h->mu);      // one call at a time per handle
  if (!h) return SEQJ_E_BADARG;
  h->periods.clear();
  if (!periods || M <= 0) return SEQJ_OK;
  for (int64_t i = 0; i < M; ++i)
    if (!(periods[i] > 0.0) || !std::isfinite(periods[i])) return fail(h, SEQJ_E_BADARG, "periods must be finite and positive");
  h->periods.assign(periods, periods + M);
  return SEQJ_OK;
}

int seqj_set_probes(seqj_handle* h, const int32_t* group, const int32_t* chunk, const int32_t* i, const int32_t* j, int64_t n) {
  if (!h) return SEQJ_E_BADARG;
  std::lock_guard<std::recursive_mutex> lock__(h->mu);      // one call at a time per handle
  if (!h) return SEQJ_E_BADARG;
  h->probe_g.clear(); h->probe_c.clear(); h->probe_i.clear(); h->probe_j.clear();
  if (n <= 0 || !group) return SEQJ_OK;
  if (!chunk || !i || !j || n > (1 << 24)) return fail(h, SEQJ_E_BADARG, "bad probe arguments");
  h->probe_g.assign(group, group + n); h->probe_c.assign(chunk, chunk + n);
  h->probe_i.assign(i, i + n); h->probe_j.assign(j, j + n);
  return SEQJ_OK;
}

int seqj_plan_describe(int64_t M, const int32_t* metric_ids, int nmetrics, const int32_t* scales, int nscales,
                       const uint8_t* group_mask, int64_t mats, int hier, char* buf, int64_t buflen) {
  if (!metric_ids || !scales || nmetrics < 1 || nscales < 1 || M < 1 || !buf || buflen < 64) return SEQJ_E_BADARG;
  std::vector<ScaleLayout> layouts;
  for (int l = 0; l < nscales; ++l) {
    if (scales[l] < 1 || scales[l] > M) return SEQJ_E_BADARG;
    layouts.push_back(make_layout((int)M, scales[l]));
  }
  PlanInput pin;
  pin.M = (int)M; pin.nmetrics = nmetrics; pin.nscales = nscales; pin.metric_ids = metric_ids; pin.layouts = &layouts;
  pin.mask = [&](int k, int l) { return !group_mask || group_mask[k * nscales + l] != 0; };
  pin.c_lo = [](int, int) { return 0; };
  pin.c_hi = [&](int, int l) { return layouts[l].nchunks; };
  pin.partial = false; pin.hier = hier != 0; pin.mats = mats; pin.max_fine = kDeriveMaxFine;
  Plan plan = make_plan(pin);
  std::string o;
  if (!plan.error.empty()) o = "error " + plan.error + "\n";
  o += "plan pool_slots " + std::to_string(plan.pool_slots) + " waves " + std::to_string(plan.waves.size()) + " total_chunks " +
       std::to_string(plan.total_chunks) + " derived_groups " + std::to_string(plan.n_derived_groups) + "\n";
  for (size_t wi = 0; wi < plan.waves.size(); ++wi) {
    const PlanWave& w = plan.waves[wi];
    o += "wave " + std::to_string(wi) + " tasks " + std::to_string(w.ntasks) + " out0 " + std::to_string(w.out0) + " done " +
         std::to_string(w.gp_done0) + " " + std::to_string(w.gp_done1) + "\n";
    for (const PlanSeg& sg : w.segs)
      o += "seg g " + std::to_string(sg.g) + " l " + std::to_string(sg.l) + " c0 " + std::to_string(sg.c0) + " c1 " +
           std::to_string(sg.c1) + " slot0 " + std::to_string(sg.slot0) + " sslot " + std::to_string(sg.sslot) + " src_slot0 " +
           std::to_string(sg.src_slot0) + " src_l " + std::to_string(sg.src_l) + " fbase " + std::to_string(sg.fbase) + " first " +
           std::to_string(sg.first) + " last " + std::to_string(sg.last) + " gp " + std::to_string(sg.gp) + "\n";
  }
  if ((int64_t)o.size() + 1 > buflen) return SEQJ_E_BADARG;
  std::memcpy(buf, o.c_str(), o.size() + 1);
  return SEQJ_OK;
}

int seqj_plan_shards(int64_t M, const int32_t* metric_ids, int nmetrics, const int32_t* scales, int nscales, int world,
                     int32_t* owner_out, double* load_out) {
  if (!metric_ids || !scales || !owner_out || nmetrics < 1 || nscales < 1 || M < 1 || world < 1) return SEQJ_E_BADARG;
  for (int l = 0; l < nscales; ++l) if (scales[l] < 1 || scales[l] > M) return SEQJ_E_BADARG;
  std::vector<int> owner;
  std::vector<double> load;
  plan_shards(M, metric_ids, nmetrics, scales, nscales, world, owner, &load);
  for (int g = 0; g < nmetrics * nscales; ++g) owner_out[g] = owner[g];
  if (load_out) for (int r = 0; r < world; ++r) load_out[r] = load[r];
  return SEQJ_OK;
}

int seqj_set_workspace_limit(seqj_handle* h, uint64_t bytes) {
  if (!h) return SEQJ_E_BADARG;
  std::lock_guard<std::recursive_mutex> lock__(h->mu);      // one call at a time per handle
  if (!h) return SEQJ_E_BADARG;
  h->ws_limit = bytes;
  return SEQJ_OK;
}

static int check_ready(seqj_handle* h) {
  if (!h) return SEQJ_E_BADARG;
  if (!h->encode || !h->st || !h->st2) return fail(h, SEQJ_E_CUDA, "handle has no usable sm_100 device (no CPU fallback): " + h->err);
  if (cudaSetDevice(h->dev) != cudaSuccess) return fail(h, SEQJ_E_CUDA, "cudaSetDevice failed");
  return SEQJ_OK;
}

// Shared implementation of seqj_sequence_dev (whole groups, optional group mask) and seqj_sequence_partial
// (chunk-major sharding: this rank computes chunks [chunk_lo[g], chunk_hi[g]) of every group g and leaves the
// eta-weighted partial sums  sum_c eta_c * D_c  in S_out; no division, no group-level measure, no fuse).
static int sequence_impl(seqj_handle* h, const double* A_dev, int64_t M64, int64_t N64, const int32_t* metric_ids,
                         int nmetrics, const int32_t* scales, int nscales, double eps_floor, double l2_thresh,
                         const uint8_t* group_mask, const int32_t* chunk_lo, const int32_t* chunk_hi, double* S_out,
                         int64_t S_stride, seqj_result** out, double* pack_out = nullptr, int f32_pdf = 0) {
  ST_TRY(check_ready(h));
  if (!A_dev || !metric_ids || !scales || !out || nmetrics < 1 || nscales < 1 || M64 < 1 || N64 < 2 ||
      M64 > (1 << 30) || N64 > (1 << 30))
    return fail(h, SEQJ_E_BADARG, "bad argument (need M >= 1, N >= 2, at least one metric and one scale)");
  const int M = (int)M64, N = (int)N64;
  for (int k = 0; k < nmetrics; ++k)
    if (metric_ids[k] < 0 || metric_ids[k] > SEQJ_METRIC_MAX) return fail(h, SEQJ_E_BADARG, "unknown metric id");
  for (int l = 0; l < nscales; ++l)
    if (scales[l] < 1 || scales[l] > M)
      return fail(h, SEQJ_E_BADARG, "Scale (" + std::to_string(scales[l]) +
                                        ") cannot be larger than the number of data elements (" + std::to_string(M) + ").");
  *out = nullptr;
  Ctx c(h);
  const int G = nmetrics * nscales;
  std::vector<ScaleLayout> layouts;
  int64_t max_ldX = 0;
  int max_chunks = 0;
  for (int l = 0; l < nscales; ++l) {
    layouts.push_back(make_layout(M, scales[l]));
    max_ldX = std::max(max_ldX, layouts.back().ldX);
    max_chunks = std::max(max_chunks, layouts.back().nchunks);
  }
  const bool partial = (chunk_lo != nullptr);
  if (partial && (!chunk_hi || !S_out)) return fail(h, SEQJ_E_BADARG, "partial mode needs chunk_hi and S_out");
  auto c_lo = [&](int k, int l) { return partial ? std::max(0, (int)chunk_lo[k * nscales + l]) : 0; };
  auto c_hi = [&](int k, int l) { return partial ? std::min(layouts[l].nchunks, (int)chunk_hi[k * nscales + l]) : layouts[l].nchunks; };
  auto mask = [&](int k, int l) {
    if (partial) return c_lo(k, l) < c_hi(k, l);
    return !group_mask || group_mask[k * nscales + l] != 0;
  };
  const bool do_final = (group_mask == nullptr) && !partial;
  bool needP = false, needLog = false, needUc = false, needWE = false, needWG = false, needPer = false;
  for (int k = 0; k < nmetrics; ++k) {
    bool used = false;                            // operands only for the metrics this call (this rank) computes
    for (int l = 0; l < nscales; ++l) used |= mask(k, l);
    if (!used) continue;
    needP |= metric_ids[k] == SEQJ_L2 || metric_ids[k] == SEQJ_KLD || metric_ids[k] == SEQJ_PERIODIC ||
             (metric_ids[k] >= SEQJ_CITYBLOCK && metric_ids[k] <= SEQJ_SQEUCLIDEAN);
    needPer |= metric_ids[k] == SEQJ_PERIODIC;
    needLog |= metric_ids[k] == SEQJ_KLD;
    needUc |= metric_ids[k] == SEQJ_EMD || metric_ids[k] == SEQJ_ENERGY;
    needWE |= metric_ids[k] == SEQJ_EMD_GRID;
    needWG |= metric_ids[k] == SEQJ_ENERGY_GRID;
  }
  if ((needWE || needWG) && (int64_t)h->grid.size() != M64)
    return fail(h, SEQJ_E_BADARG, "Grid length must match dims 1 (row count) of A. Got grid:" +
                                      std::to_string(h->grid.size()) + " and A:" + std::to_string(M64));
  if (needPer) {                  // PeriodicEuclidean(periods) only evaluates vectors of length(periods)
    if ((int64_t)h->periods.size() != M64)
      return fail(h, SEQJ_E_BADARG, "DimensionMismatch: PeriodicEuclidean needs one period per row. Got periods:" +
                                        std::to_string(h->periods.size()) + " and A:" + std::to_string(M64));
    for (int l = 0; l < nscales; ++l)
      if (layouts[l].nchunks != 1)
        return fail(h, SEQJ_E_BADARG, "DimensionMismatch: PeriodicEuclidean(periods) applies to whole columns only; scale " +
                                          std::to_string(scales[l]) + " splits them into chunks of " + std::to_string(layouts[l].cl) + " rows");
  }
  if (partial) {                                   // S_out slot = rank of the group among the active ones (metric-major)
    const size_t need = (size_t)N64 * round_up(N64, 16);
    if (S_stride < (int64_t)need || (S_stride & 1)) return fail(h, SEQJ_E_BADARG, "S_stride too small (need N*roundup(N,16), even)");
  }

  int tt = c.timer.begin(SEQJ_T_TOTAL);
  const bool htrace = getenv("SEQJ_TRACE") != nullptr;
  const auto h0 = std::chrono::steady_clock::now();
  auto hmark = [&](const char* what) {
    if (htrace) fprintf(stderr, "[seqj] host %-14s at %8.3f ms\n", what, std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - h0).count());
  };

  // `@assert all(A .>= 0) && no NaN/Inf` (src/sequencer.jl:111), checked on the device copy
  InputStats* stats_dev;
  CU_TRY(c.pool.alloc(&stats_dev, 1));
  if (!h->stats_pinned || !h->stats_ev) return fail(h, SEQJ_E_CUDA, "pinned stats buffer missing");
  InputStats* st0p = static_cast<InputStats*>(h->stats_pinned);
  CU_TRY(cudaMemsetAsync(stats_dev, 0, sizeof(InputStats), c.st));
  CU_TRY(cudaMemsetAsync(&stats_dev->minv, 0xff, sizeof(double), c.st));     // atomicMin on the bit pattern starts at the top
  input_stats_kernel<<<h->sm_count * 8, 256, 0, c.st>>>(A_dev, (int64_t)M * N, stats_dev);
  c.launches++;
  CU_TRY(cudaMemcpyAsync(st0p, stats_dev, sizeof(InputStats), cudaMemcpyDeviceToHost, c.st));
  CU_TRY(cudaEventRecord(h->stats_ev, c.st));
  bool stats_checked = false;
  // Looked at once the distance kernels of the first wave are in the queue (they are memory-safe on any input: values
  // are clamped on load, NaN / Inf only produce NaN distances).  A bad input then costs the queued work, not a hang.
  auto check_input = [&]() -> int {
    if (stats_checked) return SEQJ_OK;
    stats_checked = true;
    CU_TRY(cudaEventSynchronize(h->stats_ev));
    if (st0p->has_nan || st0p->pad || !std::isfinite(st0p->maxv)) {
      cudaStreamSynchronize(c.st_compute);
      cudaStreamSynchronize(c.st_graph);
      return fail(h, SEQJ_E_DOMAIN, "Input data cannot contain negative, NaN or infinite values.");
    }
    return SEQJ_OK;
  };

  hmark("stats queued");
  PrepBufs pb;
  pb.f32_pdf = f32_pdf;
  ST_TRY(alloc_prep(c, pb, N, max_ldX, max_chunks, needP, needLog, needUc, false, needWE, needWG));
  hmark("alloc_prep");
  if (needWE || needWG) {
    CU_TRY(c.pool.alloc(&pb.grid, (size_t)M));
    CU_TRY(cudaMemcpyAsync(pb.grid, h->grid.data(), sizeof(double) * M, cudaMemcpyHostToDevice, c.st));
  }
  if (needPer) ST_TRY(upload_periods(c, pb, h->periods, layouts[0].cl_pad));
  FixBufs fb;
  std::vector<int2> tiles = make_tiles(N, true);
  ST_TRY(alloc_fix(c, fb, 1u << 24, (size_t)max_chunks * tiles.size()));     // 16M flagged entries (256 MB) in the list
  int2* tiles_dev;
  CU_TRY(c.pool.alloc(&tiles_dev, tiles.size()));
  CU_TRY(cudaMemcpyAsync(tiles_dev, tiles.data(), tiles.size() * sizeof(int2), cudaMemcpyHostToDevice, c.st));
  // first tile of every 128-column block (the derive kernel maps its 32 x 32 blocks onto the tile list with it)
  std::vector<int> col_off((N + kTileN - 1) / kTileN + 1, 0);
  for (size_t q = 0; q < tiles.size(); ++q)
    if (tiles[q].x == 0) col_off[tiles[q].y / kTileN] = (int)q;
  int* col_off_dev;
  CU_TRY(c.pool.alloc(&col_off_dev, col_off.size()));
  CU_TRY(cudaMemcpyAsync(col_off_dev, col_off.data(), col_off.size() * sizeof(int), cudaMemcpyHostToDevice, c.st));

  // ---- static schedule (host_plan.cuh) -------------------------------------------------------
  const int64_t ldp = N;
  const int64_t ldD = round_up(N, 16);
  const size_t mat_elems = (size_t)N * ldD;
  hmark("fix+tiles");
  size_t free_b = 0, total_b = 0;
  CU_TRY(cudaMemGetInfo(&free_b, &total_b));
  hmark("memgetinfo");
  free_b += h->cached_bytes;                    // blocks cached from the previous call are reusable
  // 5 GB stay outside the matrix pool: the graph workspace (lists, candidate scratch of the upper-triangle list pass:
  // up to 0.6 GB), result staging and whatever the host framework allocates next to the library (the sharded upload of
  // config 4 needs 0.55 GB of torch tensors per rank; with 4 GB it no longer fitted once the candidate scratch existed)
  const size_t reserve_b = size_t(5) << 30;
  size_t budget = free_b > reserve_b ? free_b - reserve_b : free_b / 2;
  if (h->ws_limit && h->ws_limit < budget) budget = h->ws_limit;
  const char* henv = getenv("SEQJ_HIER");
  PlanInput pin;
  pin.M = M; pin.nmetrics = nmetrics; pin.nscales = nscales; pin.metric_ids = metric_ids; pin.layouts = &layouts;
  pin.mask = mask; pin.c_lo = c_lo; pin.c_hi = c_hi; pin.partial = partial;
  // Float32 PDFs are rounded per scale (p = Float32(a) / Float32(sum)), so a coarse chunk is no longer an exact
  // rescaling of its fine chunks: every scale runs the tile kernels in that mode
  pin.hier = !(henv && atoi(henv) == 0) && !f32_pdf;
  pin.mats = (int64_t)(budget / (mat_elems * 8));
  pin.max_fine = kDeriveMaxFine;
  {
    int total = 0;
    for (int k = 0; k < nmetrics; ++k)
      for (int l = 0; l < nscales; ++l) if (mask(k, l)) total += c_hi(k, l) - c_lo(k, l);
    const char* wenv = getenv("SEQJ_WAVES");
    if (wenv && atoi(wenv) > 1) pin.wave_target = (total + atoi(wenv) - 1) / atoi(wenv);
    if (const char* wf = getenv("SEQJ_WAVE_FRAC")) pin.wave_target = std::max(1, (int)(atof(wf) * total));
  }
  Plan plan = make_plan(pin);
  if (!plan.error.empty()) return fail(h, SEQJ_E_OOM, plan.error);
  const int total_chunks = plan.total_chunks, ngp = plan.ngp;
  const std::vector<int>& gp_of = plan.gp_of;
  const std::vector<int>& chunk_off = plan.chunk_off;
  if (total_chunks == 0) {
    // nothing selected (an all-zero group mask, or a rank of a chunk-major run that was assigned no chunk): a valid
    // call with an empty result, so that the rank can still join every collective of the sharded driver
    ST_TRY(check_input());
    seqj_result* r = new seqj_result();
    r->N = N;
    r->groups.resize(G);
    for (int k = 0; k < nmetrics; ++k)
      for (int l = 0; l < nscales; ++l) {
        GroupRes& gr = r->groups[k * nscales + l];
        gr.metric = metric_ids[k]; gr.scale = scales[l]; gr.nchunks = layouts[l].nchunks; gr.computed = 0;
      }
    c.timer.end(tt);
    CU_TRY(cudaStreamSynchronize(c.st));
    c.timer.collect(r->timers);
    r->launches = c.launches;
    r->in_min = st0p->minv; r->in_max = st0p->maxv;
    r->probe_vals.assign(h->probe_g.size(), NAN);
    *out = r;
    return SEQJ_OK;
  }
  if (plan.pool_slots < 1 || plan.pool_slots > pin.mats)
    return fail(h, SEQJ_E_OOM, "not enough device memory for two N x N chunk matrices plus the accumulator");
  hmark("waves planned");
  const int lf = std::max_element(layouts.begin(), layouts.end(), [](const ScaleLayout& a, const ScaleLayout& b) { return a.nchunks < b.nchunks; }) - layouts.begin();
  FineVecs fv;
  if (plan.n_derived_groups) {
    const size_t nn = (size_t)layouts[lf].nchunks * pb.ldn;
    CU_TRY(c.pool.alloc(&fv.mass, nn)); CU_TRY(c.pool.alloc(&fv.sP, nn)); CU_TRY(c.pool.alloc(&fv.sU, nn));
    CU_TRY(c.pool.alloc(&fv.sR, nn)); CU_TRY(c.pool.alloc(&fv.sT, nn));
    CU_TRY(c.pool.alloc(&fv.m_of_chunk, (size_t)layouts[lf].nchunks));
    CU_TRY(c.pool.alloc(&fv.fac, nn * 4 + 128)); CU_TRY(c.pool.alloc(&fv.pm, (size_t)layouts[lf].nchunks * 3));
  }

  double *eta_c, *eta_bg, *eta_g, *mstw_g, *eta_f, *fmst_w, *edge_val;
  int *root_c, *par_c, *root_g, *par_g, *msti_g, *mstj_g, *root_f, *par_f, *fmst_i, *fmst_j, *order_f, *sslot_dev, *perm_dev,
      *task_slot_dev, *task_pos_dev;
  CU_TRY(c.pool.alloc(&eta_c, (size_t)total_chunks)); CU_TRY(c.pool.alloc(&root_c, (size_t)total_chunks));
  CU_TRY(c.pool.alloc(&eta_bg, (size_t)total_chunks));          // chunk etas in (group, chunk) order for the combine
  CU_TRY(c.pool.alloc(&par_c, (size_t)total_chunks * ldp));
  CU_TRY(c.pool.alloc(&eta_g, (size_t)G)); CU_TRY(c.pool.alloc(&root_g, (size_t)G));
  CU_TRY(c.pool.alloc(&par_g, (size_t)G * ldp)); CU_TRY(c.pool.alloc(&msti_g, (size_t)G * ldp));
  CU_TRY(c.pool.alloc(&mstj_g, (size_t)G * ldp)); CU_TRY(c.pool.alloc(&mstw_g, (size_t)G * ldp));
  CU_TRY(c.pool.alloc(&eta_f, 1)); CU_TRY(c.pool.alloc(&root_f, 1)); CU_TRY(c.pool.alloc(&par_f, (size_t)ldp));
  CU_TRY(c.pool.alloc(&fmst_i, (size_t)ldp)); CU_TRY(c.pool.alloc(&fmst_j, (size_t)ldp));
  CU_TRY(c.pool.alloc(&fmst_w, (size_t)ldp)); CU_TRY(c.pool.alloc(&order_f, (size_t)ldp));
  CU_TRY(c.pool.alloc(&edge_val, (size_t)G * ldp));
  CU_TRY(c.pool.alloc(&sslot_dev, (size_t)G)); CU_TRY(c.pool.alloc(&perm_dev, (size_t)G));
  CU_TRY(c.pool.alloc(&task_slot_dev, (size_t)total_chunks)); CU_TRY(c.pool.alloc(&task_pos_dev, (size_t)total_chunks));
  std::vector<int> sslot_host(G, 0), perm_host;
  for (int gp = 0; gp < ngp; ++gp) sslot_host[gp] = std::max(0, plan.sslot_of_gp[gp]);
  for (int g = 0; g < G; ++g) if (gp_of[g] >= 0) perm_host.push_back(gp_of[g]);
  CU_TRY(cudaMemcpyAsync(sslot_dev, sslot_host.data(), sizeof(int) * G, cudaMemcpyHostToDevice, c.st));
  CU_TRY(cudaMemcpyAsync(perm_dev, perm_host.data(), sizeof(int) * perm_host.size(), cudaMemcpyHostToDevice, c.st));
  CU_TRY(cudaMemcpyAsync(task_slot_dev, plan.task_slot.data(), sizeof(int) * total_chunks, cudaMemcpyHostToDevice, c.st));
  CU_TRY(cudaMemcpyAsync(task_pos_dev, plan.task_pos.data(), sizeof(int) * total_chunks, cudaMemcpyHostToDevice, c.st));

  // active groups in metric-major order -> slot in S_out (partial mode)
  std::vector<int> out_slot(G, -1);
  {
    int sidx = 0;
    for (int g = 0; g < G; ++g) if (gp_of[g] >= 0) out_slot[g] = sidx++;
  }
  // ONE pool of N x N matrices: every phase of the plan keeps its accumulators, wave slots and carry slots in it
  double* Mpool;
  CU_TRY(c.pool.alloc(&Mpool, mat_elems * (size_t)plan.pool_slots, /*larger_ok=*/true));
  GraphWS gws;
  ST_TRY(graphws_alloc(c, gws, std::max(plan.max_tasks, std::max(1, plan.max_done)), N));

  // diagnostic probes (seqj_set_probes): values land in probe_out, NaN where this call did not produce the entry
  const int nprobe = (int)h->probe_g.size();
  double* probe_out = nullptr;
  std::vector<std::vector<long long>> probe_offs;      // kept alive until the final synchronisation
  std::vector<std::vector<int>> probe_idxs;
  if (nprobe) {
    CU_TRY(c.pool.alloc(&probe_out, (size_t)nprobe));
    CU_TRY(cudaMemsetAsync(probe_out, 0xff, sizeof(double) * nprobe, c.st));
    probe_offs.reserve(2 * plan.waves.size() + 2); probe_idxs.reserve(2 * plan.waves.size() + 2);
  }
  // gathers the probes selected by `pick(q, &offset)` from the matrix pool on the current stream
  auto run_probes = [&](const std::function<bool(int, long long*)>& pick) -> int {
    std::vector<long long> offs; std::vector<int> idxs;
    for (int q = 0; q < nprobe; ++q) {
      long long o;
      if (pick(q, &o)) { offs.push_back(o); idxs.push_back(q); }
    }
    if (offs.empty()) return SEQJ_OK;
    probe_offs.push_back(std::move(offs)); probe_idxs.push_back(std::move(idxs));
    const std::vector<long long>& ov = probe_offs.back(); const std::vector<int>& iv = probe_idxs.back();
    long long* off_d; int* idx_d;
    CU_TRY(c.pool.alloc(&off_d, ov.size())); CU_TRY(c.pool.alloc(&idx_d, iv.size()));
    CU_TRY(cudaMemcpyAsync(off_d, ov.data(), ov.size() * sizeof(long long), cudaMemcpyHostToDevice, c.st));
    CU_TRY(cudaMemcpyAsync(idx_d, iv.data(), iv.size() * sizeof(int), cudaMemcpyHostToDevice, c.st));
    probe_gather_kernel<<<(unsigned)((ov.size() + 255) / 256), 256, 0, c.st>>>(Mpool, off_d, idx_d, (int)ov.size(), probe_out);
    c.launches++;
    CU_TRY(cudaGetLastError());
    return SEQJ_OK;
  };

  // prep state: `cur_scale` = scale whose operands / vectors are in pb, `fv_scale` = scale whose vectors are in fv
  int cur_scale = -1, fv_scale = -1;
  auto prep_scale = [&](int l) -> int {
    if (cur_scale == l) return SEQJ_OK;
    hmark("before prep");
    ST_TRY(run_prep(c, A_dev, M, M, N, layouts[l], pb, 1));
    hmark("after prep");
    cur_scale = l;
    return SEQJ_OK;
  };
  auto stash_fine = [&]() -> int {                 // the vectors of the scale in pb become the derivation source
    const ScaleLayout& Lp = layouts[cur_scale];
    const size_t nb8 = (size_t)Lp.nchunks * pb.ldn * 8;
    CU_TRY(cudaMemcpyAsync(fv.mass, pb.mass, nb8, cudaMemcpyDeviceToDevice, c.st));
    CU_TRY(cudaMemcpyAsync(fv.sP, pb.sP, nb8, cudaMemcpyDeviceToDevice, c.st));
    CU_TRY(cudaMemcpyAsync(fv.sU, pb.sU, nb8, cudaMemcpyDeviceToDevice, c.st));
    CU_TRY(cudaMemcpyAsync(fv.sR, pb.sR, nb8, cudaMemcpyDeviceToDevice, c.st));
    CU_TRY(cudaMemcpyAsync(fv.sT, pb.sT, nb8, cudaMemcpyDeviceToDevice, c.st));
    CU_TRY(cudaMemcpyAsync(fv.m_of_chunk, pb.m_of_chunk, sizeof(int) * Lp.nchunks, cudaMemcpyDeviceToDevice, c.st));
    fv.nchunks = Lp.nchunks; fv.cl = Lp.cl; fv_scale = cur_scale;
    return SEQJ_OK;
  };
  std::vector<uint8_t> group_derived(G, 0);
  hmark("setup done");
  for (size_t wi = 0; wi < plan.waves.size(); ++wi) {
    PlanWave& w = plan.waves[wi];
    c.timer.wave = (int)wi;
    for (auto& sg : w.segs) {
      const ScaleLayout& L = layouts[sg.l];
      double* Dseg = Mpool + (size_t)sg.slot0 * mat_elems;
      c.timer.group = sg.g;                          // distance / derive spans are charged to the segment's group
      if (sg.src_slot0 >= 0) {
        // derived from the chunks of the next finer scale (resident in this wave or in the carry slots)
        if (fv_scale != sg.src_l) {
          if (cur_scale != sg.src_l) ST_TRY(prep_scale(sg.src_l));
          ST_TRY(stash_fine());
        }
        ST_TRY(prep_scale(sg.l));
        ST_TRY(run_derive(c, metric_ids[sg.k], L, pb, fv, Mpool + (size_t)sg.src_slot0 * mat_elems, sg.fbase, N, Dseg, ldD,
                          (int64_t)mat_elems, sg.c0, sg.c1, eps_floor, l2_thresh, fb, tiles_dev, (int)tiles.size(), col_off_dev));
        group_derived[sg.g] = 1;
        continue;
      }
      ST_TRY(prep_scale(sg.l));
      ST_TRY(run_distance(c, metric_ids[sg.k], L, pb, tiles_dev, (int)tiles.size(), N, Dseg, ldD, (int64_t)mat_elems, sg.c0,
                          sg.c1 - sg.c0, eps_floor, l2_thresh, 0, fb));
    }
    c.timer.group = -1;                              // measure / combine spans: charged to the wave, split by task count
    hmark("dist queued");
    ST_TRY(check_input());
    hmark("input checked");
    if (nprobe)
      ST_TRY(run_probes([&](int q, long long* o) {
        const int pc = h->probe_c[q], pi = h->probe_i[q], pj = h->probe_j[q];
        if (pc < 0 || pi < 0 || pj < 0 || pi >= N || pj >= N) return false;
        for (auto& sg : w.segs)
          if (sg.g == h->probe_g[q] && pc >= sg.c0 && pc < sg.c1) {
            *o = (long long)(sg.slot0 + pc - sg.c0) * (long long)mat_elems + (long long)pi * ldD + pj;
            return true;
          }
        return false;
      }));
    ST_TRY(run_measure(c, gws, Mpool, ldD, (int64_t)mat_elems, w.ntasks, eps_floor, eta_c + w.out0, root_c + w.out0,
                       par_c + (int64_t)w.out0 * ldp, ldp, nullptr, nullptr, nullptr, nullptr, task_slot_dev + w.out0));
    int t = c.timer.begin(SEQJ_T_COMBINE);
    scatter_eta_kernel<<<(w.ntasks + 255) / 256, 256, 0, c.st>>>(eta_c + w.out0, task_pos_dev + w.out0, w.ntasks, eta_bg);
    c.launches++;
    for (auto& sg : w.segs) {
      const int g = sg.g;
      const int nch = layouts[sg.l].nchunks;
      const int lo = c_lo(sg.k, sg.l);
      double* Sg = partial ? S_out + (size_t)out_slot[g] * (size_t)S_stride : Mpool + (size_t)sg.sslot * mat_elems;
      for (int b0 = sg.c0; b0 < sg.c1; b0 += kCombineMax) {
        const int nb = std::min(kCombineMax, sg.c1 - b0);
        const unsigned T = (unsigned)((N + kCombineTile - 1) / kCombineTile);
        combine_sym_kernel<<<dim3(T, T), 256, 0, c.st>>>(
            Sg, ldD, N, Mpool + (size_t)(sg.slot0 + b0 - sg.c0) * mat_elems, (int64_t)mat_elems, nb,
            eta_bg + chunk_off[g] + (b0 - lo), b0 == lo, !partial && b0 + nb >= nch, eta_bg + chunk_off[g], nch);
        c.launches++;
      }
    }
    c.timer.end(t);
    CU_TRY(cudaGetLastError());
    if (nprobe && !partial && w.gp_done1 > w.gp_done0)
      ST_TRY(run_probes([&](int q, long long* o) {
        const int pg = h->probe_g[q], pi = h->probe_i[q], pj = h->probe_j[q];
        if (h->probe_c[q] >= 0 || pg < 0 || pg >= G || gp_of[pg] < w.gp_done0 || gp_of[pg] >= w.gp_done1 || pi < 0 ||
            pj < 0 || pi >= N || pj >= N) return false;
        *o = (long long)plan.sslot_of_gp[gp_of[pg]] * (long long)mat_elems + (long long)pi * ldD + pj;
        return true;
      }));
    if (w.gp_done1 > w.gp_done0) {
      const int a = w.gp_done0, nd = w.gp_done1 - w.gp_done0;
      ST_TRY(run_measure(c, gws, Mpool, ldD, (int64_t)mat_elems, nd, eps_floor, eta_g + a, root_g + a, par_g + (int64_t)a * ldp,
                         ldp, msti_g + (int64_t)a * ldp, mstj_g + (int64_t)a * ldp, mstw_g + (int64_t)a * ldp, nullptr,
                         sslot_dev + a));
    }
  }
  c.timer.wave = -1;
  int n_derived_done = 0;
  for (int g = 0; g < G; ++g) n_derived_done += group_derived[g];
  if (do_final)
    ST_TRY(run_fuse(c, gws, Mpool, ldD, N, G, eta_g, msti_g, mstj_g, mstw_g, ldp, eps_floor, eta_f, root_f, par_f, fmst_i,
                    fmst_j, fmst_w, order_f, edge_val, perm_host, perm_dev));

  // device-side exchange of the group results (group-major sharding): slot q = q-th computed group, metric-major
  if (pack_out && !partial && !perm_host.empty()) {
    pack_groups_kernel<<<dim3((unsigned)((N + 255) / 256), (unsigned)perm_host.size()), 256, 0, c.st>>>(
        pack_out, 2 * (int64_t)N, perm_dev, eta_g, msti_g, mstj_g, mstw_g, ldp, N);
    c.launches++;
    CU_TRY(cudaGetLastError());
  }

  // ---- copy-out ----
  seqj_result* r = new seqj_result();
  r->N = N;
  r->groups.resize(G);
  int td = c.timer.begin(SEQJ_T_D2H);
  std::vector<double> h_eta_c(total_chunks), h_eta_g(G), h_mstw((size_t)G * ldp), h_val;
  std::vector<int> h_root_c(total_chunks), h_par_c((size_t)total_chunks * ldp), h_root_g(G), h_par_g((size_t)G * ldp),
      h_msti((size_t)G * ldp), h_mstj((size_t)G * ldp);
  cudaError_t e = cudaSuccess;
  auto d2h = [&](void* dst, const void* src, size_t bytes) {
    if (e == cudaSuccess && bytes) e = cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, c.st);
  };
  d2h(h_eta_c.data(), eta_c, sizeof(double) * total_chunks);
  d2h(h_root_c.data(), root_c, sizeof(int) * total_chunks);
  d2h(h_par_c.data(), par_c, sizeof(int) * (size_t)total_chunks * ldp);
  d2h(h_eta_g.data(), eta_g, sizeof(double) * G);
  d2h(h_root_g.data(), root_g, sizeof(int) * G);
  d2h(h_par_g.data(), par_g, sizeof(int) * (size_t)G * ldp);
  d2h(h_msti.data(), msti_g, sizeof(int) * (size_t)G * ldp);
  d2h(h_mstj.data(), mstj_g, sizeof(int) * (size_t)G * ldp);
  d2h(h_mstw.data(), mstw_g, sizeof(double) * (size_t)G * ldp);
  if (do_final) {
    r->has_final = 1;
    r->order.resize(N); r->parents.resize(N); r->mst_i.resize(N - 1); r->mst_j.resize(N - 1); r->mst_w.resize(N - 1);
    h_val.resize((size_t)G * ldp);
    d2h(&r->eta, eta_f, sizeof(double));
    d2h(&r->root, root_f, sizeof(int));
    d2h(r->order.data(), order_f, sizeof(int) * N);
    d2h(r->parents.data(), par_f, sizeof(int) * N);
    d2h(r->mst_i.data(), fmst_i, sizeof(int) * (N - 1));
    d2h(r->mst_j.data(), fmst_j, sizeof(int) * (N - 1));
    d2h(r->mst_w.data(), fmst_w, sizeof(double) * (N - 1));
    d2h(h_val.data(), edge_val, sizeof(double) * (size_t)G * ldp);
  }
  if (nprobe) {
    r->probe_vals.resize(nprobe);
    d2h(r->probe_vals.data(), probe_out, sizeof(double) * nprobe);
  }
  c.timer.end(td);
  c.timer.end(tt);
  if (e == cudaSuccess) e = cudaStreamSynchronize(c.st);
  if (e != cudaSuccess) {
    cudaStreamSynchronize(c.st);
    delete r;
    return fail(h, SEQJ_E_CUDA, std::string("sequence failed: ") + cudaGetErrorString(e));
  }
  for (int k = 0; k < nmetrics; ++k)
    for (int l = 0; l < nscales; ++l) {
      const int g = k * nscales + l;
      GroupRes& gr = r->groups[g];
      gr.metric = metric_ids[k]; gr.scale = scales[l]; gr.nchunks = layouts[l].nchunks; gr.computed = mask(k, l);
      if (!gr.computed) continue;
      gr.chunk_lo = c_lo(k, l); gr.chunk_hi = c_hi(k, l);
      const int nloc = gr.chunk_hi - gr.chunk_lo;
      const size_t gp = (size_t)gp_of[g];        // row of this group in the device arrays (finishing order)
      gr.eta_chunk.resize(nloc); gr.root_chunk.resize(nloc); gr.parents_chunk.resize((size_t)nloc * ldp);
      if (partial) continue;                     // group-level results come from seqj_groups_finish on the owner rank
      gr.eta = h_eta_g[gp]; gr.root = h_root_g[gp];
      gr.parents.assign(h_par_g.begin() + gp * ldp, h_par_g.begin() + gp * ldp + N);
      gr.mst_i.assign(h_msti.begin() + gp * ldp, h_msti.begin() + gp * ldp + N - 1);
      gr.mst_j.assign(h_mstj.begin() + gp * ldp, h_mstj.begin() + gp * ldp + N - 1);
      gr.mst_w.assign(h_mstw.begin() + gp * ldp, h_mstw.begin() + gp * ldp + N - 1);
    }
  // chunk-level results arrive in task (wave) order: file them under (group, chunk)
  for (auto& w : plan.waves)
    for (auto& sg : w.segs) {
      GroupRes& gr = r->groups[sg.g];
      for (int cc = sg.c0; cc < sg.c1; ++cc) {
        const size_t tsk = (size_t)sg.out0 + (cc - sg.c0), dst = (size_t)(cc - gr.chunk_lo);
        gr.eta_chunk[dst] = h_eta_c[tsk];
        gr.root_chunk[dst] = h_root_c[tsk];
        std::memcpy(gr.parents_chunk.data() + dst * ldp, h_par_c.data() + tsk * ldp, sizeof(int) * (size_t)ldp);
      }
    }
  if (do_final) stash_triplets(r, G, ldp, h_msti, h_mstj, h_val);
  c.timer.collect(r->timers);
  // per-group elapsed time (the reference prints `tt = @elapsed` of the group's chunk loop, src/sequencer.jl:210,260):
  // the group's own distance / derive kernels plus its share (by chunk count) of its waves' batched measure + combine
  for (int g = 0; g < G; ++g)
    if ((int)c.timer.group_ms.size() > g) r->groups[g].elapsed_ms = c.timer.group_ms[g];
  for (size_t wi = 0; wi < plan.waves.size(); ++wi) {
    const PlanWave& w = plan.waves[wi];
    if (c.timer.wave_ms.size() <= wi || w.ntasks == 0) continue;
    for (auto& sg : w.segs) r->groups[sg.g].elapsed_ms += c.timer.wave_ms[wi] * (double)(sg.c1 - sg.c0) / (double)w.ntasks;
  }
  r->launches = c.launches;
  r->in_min = st0p->minv; r->in_max = st0p->maxv;       // check_input() ran before the first Prim launch
  r->derived_groups = n_derived_done;
  r->waves = (int)plan.waves.size();
  *out = r;
  return SEQJ_OK;
}

int seqj_sequence_dev(seqj_handle* h, const double* A_dev, int64_t M, int64_t N, const int32_t* metric_ids,
                      int nmetrics, const int32_t* scales, int nscales, double eps_floor, double l2_thresh,
                      const uint8_t* group_mask, seqj_result** out) {
  if (!h) return SEQJ_E_BADARG;
  std::lock_guard<std::recursive_mutex> lock__(h->mu);      // one call at a time per handle
  return sequence_impl(h, A_dev, M, N, metric_ids, nmetrics, scales, nscales, eps_floor, l2_thresh, group_mask, nullptr,
                       nullptr, nullptr, 0, out);
}

int seqj_sequence_partial(seqj_handle* h, const double* A_dev, int64_t M, int64_t N, const int32_t* metric_ids,
                          int nmetrics, const int32_t* scales, int nscales, double eps_floor, double l2_thresh,
                          const int32_t* chunk_lo, const int32_t* chunk_hi, double* S_out_dev, int64_t S_stride,
                          seqj_result** out) {
  if (!h) return SEQJ_E_BADARG;
  std::lock_guard<std::recursive_mutex> lock__(h->mu);      // one call at a time per handle
  if (!chunk_lo || !chunk_hi || !S_out_dev) return fail(h, SEQJ_E_BADARG, "bad argument");
  return sequence_impl(h, A_dev, M, N, metric_ids, nmetrics, scales, nscales, eps_floor, l2_thresh, nullptr, chunk_lo,
                       chunk_hi, S_out_dev, S_stride, out);
}

int seqj_groups_finish(seqj_handle* h, double* S_dev, int64_t S_stride, int ngroups, const int32_t* slots, int64_t N64,
                       const double* eta_sum, double eps_floor, double* eta, int32_t* root, int32_t* parents,
                       int32_t* mst_i, int32_t* mst_j, double* mst_w) {
  if (!h) return SEQJ_E_BADARG;
  std::lock_guard<std::recursive_mutex> lock__(h->mu);      // one call at a time per handle
  ST_TRY(check_ready(h));
  if (!S_dev || !eta_sum || ngroups < 1 || N64 < 2) return fail(h, SEQJ_E_BADARG, "bad argument");
  const int N = (int)N64, ng = ngroups;
  Ctx c(h);
  const int64_t ldD = round_up(N, 16), ldp = N;
  double *eta_d, *w_d;
  int *root_d, *par_d, *mi_d, *mj_d, *slot_d;
  CU_TRY(c.pool.alloc(&eta_d, (size_t)ng)); CU_TRY(c.pool.alloc(&root_d, (size_t)ng));
  CU_TRY(c.pool.alloc(&par_d, (size_t)ng * ldp)); CU_TRY(c.pool.alloc(&mi_d, (size_t)ng * ldp));
  CU_TRY(c.pool.alloc(&mj_d, (size_t)ng * ldp)); CU_TRY(c.pool.alloc(&w_d, (size_t)ng * ldp));
  CU_TRY(c.pool.alloc(&slot_d, (size_t)ng));
  std::vector<int> slot_h(ng);
  for (int g = 0; g < ng; ++g) slot_h[g] = slots ? slots[g] : g;
  CU_TRY(cudaMemcpyAsync(slot_d, slot_h.data(), sizeof(int) * ng, cudaMemcpyHostToDevice, c.st));
  for (int g = 0; g < ng; ++g) {                  // Dkl = Dklms / sum(etas)  (reference src/sequencer.jl:247)
    scale_div_kernel<<<h->sm_count * 8, 256, 0, c.st>>>(S_dev + (size_t)slot_h[g] * (size_t)S_stride, eta_sum[g],
                                                       (int64_t)N * ldD);
    c.launches++;
  }
  CU_TRY(cudaGetLastError());
  GraphWS gws;
  ST_TRY(graphws_alloc(c, gws, ng, N));
  ST_TRY(run_measure(c, gws, S_dev, ldD, S_stride, ng, eps_floor, eta_d, root_d, par_d, ldp, mi_d, mj_d, w_d, nullptr, slot_d));
  if (eta) CU_TRY(cudaMemcpyAsync(eta, eta_d, 8 * (size_t)ng, cudaMemcpyDeviceToHost, c.st));
  if (root) CU_TRY(cudaMemcpyAsync(root, root_d, 4 * (size_t)ng, cudaMemcpyDeviceToHost, c.st));
  if (parents) CU_TRY(cudaMemcpyAsync(parents, par_d, 4 * (size_t)ng * N, cudaMemcpyDeviceToHost, c.st));
  for (int g = 0; g < ng; ++g) {
    if (mst_i) CU_TRY(cudaMemcpyAsync(mst_i + (size_t)g * (N - 1), mi_d + (size_t)g * ldp, 4 * (size_t)(N - 1), cudaMemcpyDeviceToHost, c.st));
    if (mst_j) CU_TRY(cudaMemcpyAsync(mst_j + (size_t)g * (N - 1), mj_d + (size_t)g * ldp, 4 * (size_t)(N - 1), cudaMemcpyDeviceToHost, c.st));
    if (mst_w) CU_TRY(cudaMemcpyAsync(mst_w + (size_t)g * (N - 1), w_d + (size_t)g * ldp, 8 * (size_t)(N - 1), cudaMemcpyDeviceToHost, c.st));
  }
  CU_TRY(cudaStreamSynchronize(c.st));
  return SEQJ_OK;
}

static int sequence_host_one(seqj_handle* h, const void* A, int f32, int64_t M, int64_t N, const int32_t* metric_ids,
                             int nmetrics, const int32_t* scales, int nscales, double eps_floor, double l2_thresh,
                             const uint8_t* group_mask, seqj_result** out);
static int fuse_impl(seqj_handle* h, int64_t N64, int ngroups, const double* eta_group, const int32_t* mst_i,
                     const int32_t* mst_j, const double* mst_w, const double* pack_dev, const int32_t* slot_of_group,
                     double eps_floor, seqj_result** out);

// Single-process multi-GPU: whole (metric, scale) groups are assigned to the devices of the handle by an LPT
// greedy on estimated cost, one host thread per device runs the masked sequence, the per-group (eta, MST) are
// merged on the host (~160 KB per group) and the primary device runs the fuse.  No N x N data leaves a GPU.
static int sequence_multi(seqj_handle* h, const double* A, int64_t M, int64_t N, const int32_t* metric_ids, int nmetrics,
                          const int32_t* scales, int nscales, double eps_floor, double l2_thresh, seqj_result** out) {
  const int ndev = 1 + (int)h->peers.size(), G = nmetrics * nscales;
  std::vector<seqj_handle*> hs;
  hs.push_back(h);
  for (auto* p : h->peers) hs.push_back(p);
  // which device owns which group: cost-model LPT over EMD groups and (sub-)chains of the contraction metrics (host_shard.cuh)
  std::vector<int> owner;
  plan_shards(M, metric_ids, nmetrics, scales, nscales, ndev, owner, nullptr);
  std::vector<std::vector<uint8_t>> masks(ndev, std::vector<uint8_t>(G, 0));
  for (int g = 0; g < G; ++g) masks[owner[g]][g] = 1;
  std::vector<seqj_result*> parts(ndev, nullptr);
  std::vector<int> status(ndev, SEQJ_OK);
  while ((int)h->workers.size() < ndev - 1) h->workers.push_back(new DeviceWorker());
  for (int d = 0; d < ndev; ++d) {
    hs[d]->grid = h->grid;
    hs[d]->periods = h->periods;
  }
  // Every device uploads 1/ndev of the objects over its own PCIe link, then pulls the other slabs from its peers over
  // NVLink; the group results are exchanged in the device format (pack slots) and the primary device fuses from them.
  const int64_t rows = (N + ndev - 1) / ndev;                        // objects per slab
  auto slab_lo = [&](int d) { return std::min<int64_t>(N, (int64_t)d * rows); };
  auto slab_hi = [&](int d) { return std::min<int64_t>(N, (int64_t)(d + 1) * rows); };
  int cap = 1;
  for (int d = 0; d < ndev; ++d) {
    int n = 0;
    for (uint8_t b : masks[d]) n += b != 0;
    cap = std::max(cap, n);
  }
  const int64_t pstride = 2 * N;
  std::vector<std::unique_ptr<DevPool>> pools(ndev);
  std::vector<double*> A_dev(ndev, nullptr), pack(ndev, nullptr);
  std::vector<float> h2d_ms(ndev, 0.f);
  auto fail_dev = [&](int d, cudaError_t e, const char* what) {
    hs[d]->err = std::string(what) + ": " + cudaGetErrorString(e);
    status[d] = (e == cudaErrorMemoryAllocation) ? SEQJ_E_OOM : SEQJ_E_CUDA;
  };
  auto phase_upload = [&](int d) {
    seqj_handle* hd = hs[d];
    if (check_ready(hd) != SEQJ_OK) { status[d] = SEQJ_E_CUDA; return; }
    pools[d].reset(new DevPool(hd));
    cudaError_t e = pools[d]->alloc(&A_dev[d], (size_t)M * N);
    if (e == cudaSuccess) e = pools[d]->alloc(&pack[d], (size_t)cap * pstride);
    if (e != cudaSuccess) { fail_dev(d, e, "device allocation"); return; }
    const int64_t lo = slab_lo(d), hi = slab_hi(d);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0, hd->st);
    if (hi > lo) {
      const double* src = A + lo * M;
      double* dst = A_dev[d] + lo * M;
      const size_t bytes = (size_t)(hi - lo) * M * 8;
      const bool staged = bytes >= (size_t(4) << 20) && !host_pointer_is_pinned(src);
      e = staged ? upload_pageable(hd->upload, hd->dev, dst, src, bytes, hd->st)
                 : cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, hd->st);
    }
    cudaEventRecord(e1, hd->st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(hd->st);
    if (e == cudaSuccess) cudaEventElapsedTime(&h2d_ms[d], e0, e1);
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    if (e != cudaSuccess) fail_dev(d, e, "H2D copy");
  };
  auto phase_compute = [&](int d) {
    seqj_handle* hd = hs[d];
    bool any = false;
    for (uint8_t b : masks[d]) any |= b != 0;
    if (!any || status[d] != SEQJ_OK) return;
    cudaSetDevice(hd->dev);
    for (int e = 0; e < ndev; ++e) {                                // all-gather of the slabs over NVLink (peer copies)
      if (e == d || slab_hi(e) <= slab_lo(e)) continue;
      cudaError_t ce = cudaMemcpyPeerAsync(A_dev[d] + slab_lo(e) * M, hd->dev, A_dev[e] + slab_lo(e) * M, hs[e]->dev,
                                           (size_t)(slab_hi(e) - slab_lo(e)) * M * 8, hd->st);
      if (ce != cudaSuccess) { fail_dev(d, ce, "peer copy of A"); return; }
    }
    status[d] = sequence_impl(hd, A_dev[d], M, N, metric_ids, nmetrics, scales, nscales, eps_floor, l2_thresh, masks[d].data(),
                              nullptr, nullptr, nullptr, 0, &parts[d], pack[d]);
  };
  for (int d = 1; d < ndev; ++d) h->workers[d - 1]->post([&, d]() { phase_upload(d); });
  phase_upload(0);                                                  // the primary device runs on the calling thread
  for (int d = 1; d < ndev; ++d) h->workers[d - 1]->wait();
  bool upload_ok = true;
  for (int d = 0; d < ndev; ++d) upload_ok = upload_ok && status[d] == SEQJ_OK;
  if (upload_ok) {
    for (int d = 1; d < ndev; ++d) h->workers[d - 1]->post([&, d]() { phase_compute(d); });
    phase_compute(0);
    for (int d = 1; d < ndev; ++d) h->workers[d - 1]->wait();
  }
  cudaSetDevice(h->dev);
  for (int d = 0; d < ndev; ++d)
    if (status[d] != SEQJ_OK) {
      const int st = status[d];
      h->err = "device " + std::to_string(hs[d]->dev) + ": " + hs[d]->err;
      for (auto* r : parts) delete r;
      return st;
    }
  // gather the pack slots on the primary device (peer copies over NVLink), fuse from the device buffer
  double* gathered = nullptr;
  {
    cudaError_t e = pools[0]->alloc(&gathered, (size_t)ndev * cap * pstride);
    for (int d = 0; d < ndev && e == cudaSuccess; ++d)
      e = cudaMemcpyPeerAsync(gathered + (size_t)d * cap * pstride, h->dev, pack[d], hs[d]->dev, (size_t)cap * pstride * 8, h->st);
    if (e != cudaSuccess) {
      for (auto* r : parts) delete r;
      return fail(h, SEQJ_E_CUDA, std::string("gather of the group results: ") + cudaGetErrorString(e));
    }
  }
  std::vector<int32_t> slot_of(G, 0);
  {
    std::vector<int> seen(ndev, 0);
    for (int g = 0; g < G; ++g) slot_of[g] = owner[g] * cap + seen[owner[g]]++;
  }
  seqj_result* fin = nullptr;
  int st = fuse_impl(h, N, G, nullptr, nullptr, nullptr, nullptr, gathered, slot_of.data(), eps_floor, &fin);
  if (st != SEQJ_OK) {
    for (auto* r : parts) delete r;
    return st;
  }
  fin->groups.resize(G);
  double fuse_total = fin->timers[SEQJ_T_TOTAL];
  for (int t = 0; t < SEQJ_NTIMERS; ++t) fin->timers[t] = (t == SEQJ_T_FUSE) ? fuse_total : 0.0;
  double slowest = 0.0;
  fin->in_min = INFINITY; fin->in_max = 0.0;
  for (int d = 0; d < ndev; ++d) {
    if (!parts[d]) continue;
    for (int g = 0; g < G; ++g)
      if (owner[g] == d) fin->groups[g] = std::move(parts[d]->groups[g]);
    for (int t = 1; t < SEQJ_NTIMERS; ++t)
      if (t != SEQJ_T_FUSE) fin->timers[t] = std::max(fin->timers[t], parts[d]->timers[t]);   // max over devices
    slowest = std::max(slowest, parts[d]->timers[SEQJ_T_TOTAL]);
    fin->launches += parts[d]->launches;
    fin->in_min = std::min(fin->in_min, parts[d]->in_min);
    fin->in_max = std::max(fin->in_max, parts[d]->in_max);
    delete parts[d];
  }
  float h2d = 0.f;
  for (float x : h2d_ms) h2d = std::max(h2d, x);
  fin->timers[SEQJ_T_H2D] = h2d;
  fin->timers[SEQJ_T_TOTAL] = h2d + slowest + fuse_total;
  *out = fin;
  return SEQJ_OK;
}

int seqj_sequence(seqj_handle* h, const double* A, int64_t M, int64_t N, const int32_t* metric_ids, int nmetrics,
                  const int32_t* scales, int nscales, double eps_floor, double l2_thresh, const uint8_t* group_mask,
                  seqj_result** out) {
  if (!h) return SEQJ_E_BADARG;
  std::lock_guard<std::recursive_mutex> lock__(h->mu);      // one call at a time per handle
  ST_TRY(check_ready(h));
  if (!A || !metric_ids || !scales || !out || nmetrics < 1 || nscales < 1 || M < 1 || N < 2)
    return fail(h, SEQJ_E_BADARG, "bad argument");
  if (!h->peers.empty() && !group_mask && nmetrics * nscales > 1)
    return sequence_multi(h, A, M, N, metric_ids, nmetrics, scales, nscales, eps_floor, l2_thresh, out);
  return sequence_host_one(h, A, 0, M, N, metric_ids, nmetrics, scales, nscales, eps_floor, l2_thresh, group_mask, out);
}

int seqj_sequence_f32(seqj_handle* h, const float* A, int64_t M, int64_t N, const int32_t* metric_ids, int nmetrics,
                      const int32_t* scales, int nscales, double eps_floor, double l2_thresh, const uint8_t* group_mask,
                      seqj_result** out) {
  if (!h) return SEQJ_E_BADARG;
  std::lock_guard<std::recursive_mutex> lock__(h->mu);      // one call at a time per handle
  ST_TRY(check_ready(h));
  if (!A || !metric_ids || !scales || !out || nmetrics < 1 || nscales < 1 || M < 1 || N < 2)
    return fail(h, SEQJ_E_BADARG, "bad argument");
  if (!h->peers.empty())
    return fail(h, SEQJ_E_BADARG, "Float32 input is supported on single-device handles (convert to Float64 for a multi-device handle)");
  return sequence_host_one(h, A, 1, M, N, metric_ids, nmetrics, scales, nscales, eps_floor, l2_thresh, group_mask, out);
}

static int sequence_host_one(seqj_handle* h, const void* A, int f32, int64_t M, int64_t N, const int32_t* metric_ids,
                             int nmetrics, const int32_t* scales, int nscales, double eps_floor, double l2_thresh,
                             const uint8_t* group_mask, seqj_result** out) {
  ST_TRY(check_ready(h));
  if (!A || M < 1 || N < 2) return fail(h, SEQJ_E_BADARG, "bad argument");
  DevPool pool(h);
  double* A_dev;
  float* A32_dev = nullptr;
  CU_TRY(pool.alloc(&A_dev, (size_t)M * N));
  if (f32) CU_TRY(pool.alloc(&A32_dev, (size_t)M * N));
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0, h->st);
  // pinned / registered callers: one DMA; pageable callers (a Julia Matrix, a NumPy array): staged upload
  const size_t abytes = (size_t)M * N * (f32 ? 4 : 8);
  void* dst = f32 ? static_cast<void*>(A32_dev) : static_cast<void*>(A_dev);
  const char* up = getenv("SEQJ_UPLOAD");
  const bool staged = !(up && !strcmp(up, "direct")) && abytes >= (size_t(4) << 20) && !host_pointer_is_pinned(A);
  cudaError_t e = staged ? upload_pageable(h->upload, h->dev, dst, A, abytes, h->st)
                         : cudaMemcpyAsync(dst, A, abytes, cudaMemcpyHostToDevice, h->st);
  cudaEventRecord(e1, h->st);
  if (e != cudaSuccess) {
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    return fail(h, SEQJ_E_CUDA, std::string("H2D copy failed: ") + cudaGetErrorString(e));
  }
  if (f32) widen_f32_kernel<<<h->sm_count * 8, 256, 0, h->st>>>(A32_dev, (int64_t)M * N, A_dev);
  int s = sequence_impl(h, A_dev, M, N, metric_ids, nmetrics, scales, nscales, eps_floor, l2_thresh, group_mask, nullptr,
                        nullptr, nullptr, 0, out, nullptr, f32);
  if (s == SEQJ_OK) {
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    (*out)->timers[SEQJ_T_H2D] = ms;
    (*out)->timers[SEQJ_T_TOTAL] += ms;
  }
  cudaEventDestroy(e0); cudaEventDestroy(e1);
  return s;
}

// Shared body of seqj_fuse (per-group results in host arrays) and seqj_fuse_dev (in the device exchange format).
static int fuse_impl(seqj_handle* h, int64_t N64, int ngroups, const double* eta_group, const int32_t* mst_i,
                     const int32_t* mst_j, const double* mst_w, const double* pack_dev, const int32_t* slot_of_group,
                     double eps_floor, seqj_result** out) {
  ST_TRY(check_ready(h));
  const int N = (int)N64, G = ngroups;
  Ctx c(h);
  const int64_t ldp = N, ldD = round_up(N, 16);
  int tt = c.timer.begin(SEQJ_T_TOTAL);
  double *eta_g, *mstw_g, *eta_f, *fmst_w, *edge_val, *Sbuf;
  int *msti_g, *mstj_g, *root_f, *par_f, *fmst_i, *fmst_j, *order_f;
  CU_TRY(c.pool.alloc(&eta_g, (size_t)G)); CU_TRY(c.pool.alloc(&msti_g, (size_t)G * ldp));
  CU_TRY(c.pool.alloc(&mstj_g, (size_t)G * ldp)); CU_TRY(c.pool.alloc(&mstw_g, (size_t)G * ldp));
  CU_TRY(c.pool.alloc(&eta_f, 1)); CU_TRY(c.pool.alloc(&root_f, 1)); CU_TRY(c.pool.alloc(&par_f, (size_t)ldp));
  CU_TRY(c.pool.alloc(&fmst_i, (size_t)ldp)); CU_TRY(c.pool.alloc(&fmst_j, (size_t)ldp));
  CU_TRY(c.pool.alloc(&fmst_w, (size_t)ldp)); CU_TRY(c.pool.alloc(&order_f, (size_t)ldp));
  CU_TRY(c.pool.alloc(&edge_val, (size_t)G * ldp)); CU_TRY(c.pool.alloc(&Sbuf, (size_t)N * ldD, /*larger_ok=*/true));
  if (pack_dev) {
    int* slot_d;
    CU_TRY(c.pool.alloc(&slot_d, (size_t)G));
    CU_TRY(cudaMemcpyAsync(slot_d, slot_of_group, sizeof(int) * G, cudaMemcpyHostToDevice, c.st));
    unpack_groups_kernel<<<dim3((unsigned)((N + 255) / 256), (unsigned)G), 256, 0, c.st>>>(pack_dev, 2 * (int64_t)N, slot_d, eta_g,
                                                                                         msti_g, mstj_g, mstw_g, ldp, N);
    c.launches++;
    CU_TRY(cudaGetLastError());
  } else {
    CU_TRY(cudaMemcpyAsync(eta_g, eta_group, sizeof(double) * G, cudaMemcpyHostToDevice, c.st));
    CU_TRY(cudaMemcpy2DAsync(msti_g, sizeof(int) * ldp, mst_i, sizeof(int) * (N - 1), sizeof(int) * (N - 1), G, cudaMemcpyHostToDevice, c.st));
    CU_TRY(cudaMemcpy2DAsync(mstj_g, sizeof(int) * ldp, mst_j, sizeof(int) * (N - 1), sizeof(int) * (N - 1), G, cudaMemcpyHostToDevice, c.st));
    CU_TRY(cudaMemcpy2DAsync(mstw_g, sizeof(double) * ldp, mst_w, sizeof(double) * (N - 1), sizeof(double) * (N - 1), G, cudaMemcpyHostToDevice, c.st));
  }
  GraphWS gws;
  ST_TRY(graphws_alloc(c, gws, 1, N));
  std::vector<int> perm(G);
  for (int g = 0; g < G; ++g) perm[g] = g;
  ST_TRY(run_fuse(c, gws, Sbuf, ldD, N, G, eta_g, msti_g, mstj_g, mstw_g, ldp, eps_floor, eta_f, root_f, par_f, fmst_i,
                  fmst_j, fmst_w, order_f, edge_val, perm, nullptr));
  seqj_result* r = new seqj_result();
  r->N = N; r->has_final = 1;
  r->order.resize(N); r->parents.resize(N); r->mst_i.resize(N - 1); r->mst_j.resize(N - 1); r->mst_w.resize(N - 1);
  std::vector<double> h_val((size_t)G * ldp);
  std::vector<int> h_mi((size_t)G * ldp), h_mj((size_t)G * ldp);
  cudaMemcpyAsync(&r->eta, eta_f, sizeof(double), cudaMemcpyDeviceToHost, c.st);
  cudaMemcpyAsync(&r->root, root_f, sizeof(int), cudaMemcpyDeviceToHost, c.st);
  cudaMemcpyAsync(r->order.data(), order_f, sizeof(int) * N, cudaMemcpyDeviceToHost, c.st);
  cudaMemcpyAsync(r->parents.data(), par_f, sizeof(int) * N, cudaMemcpyDeviceToHost, c.st);
  cudaMemcpyAsync(r->mst_i.data(), fmst_i, sizeof(int) * (N - 1), cudaMemcpyDeviceToHost, c.st);
  cudaMemcpyAsync(r->mst_j.data(), fmst_j, sizeof(int) * (N - 1), cudaMemcpyDeviceToHost, c.st);
  cudaMemcpyAsync(r->mst_w.data(), fmst_w, sizeof(double) * (N - 1), cudaMemcpyDeviceToHost, c.st);
  cudaMemcpyAsync(h_val.data(), edge_val, sizeof(double) * (size_t)G * ldp, cudaMemcpyDeviceToHost, c.st);
  cudaMemcpyAsync(h_mi.data(), msti_g, sizeof(int) * (size_t)G * ldp, cudaMemcpyDeviceToHost, c.st);
  cudaMemcpyAsync(h_mj.data(), mstj_g, sizeof(int) * (size_t)G * ldp, cudaMemcpyDeviceToHost, c.st);
  c.timer.end(tt);
  cudaError_t e = cudaStreamSynchronize(c.st);
  if (e != cudaSuccess) {
    delete r;
    return fail(h, SEQJ_E_CUDA, std::string("fuse failed: ") + cudaGetErrorString(e));
  }
  stash_triplets(r, G, ldp, h_mi, h_mj, h_val);
  c.timer.collect(r->timers);
  r->launches = c.launches;
  *out = r;
  return SEQJ_OK;
}

int seqj_fuse(seqj_handle* h, int64_t N64, int ngroups, const double* eta_group, const int32_t* mst_i,
              const int32_t* mst_j, const double* mst_w, double eps_floor, seqj_result** out) {
  if (!h) return SEQJ_E_BADARG;
  std::lock_guard<std::recursive_mutex> lock__(h->mu);      // one call at a time per handle
  if (!h) return SEQJ_E_BADARG;
  if (!eta_group || !mst_i || !mst_j || !mst_w || !out || N64 < 2 || ngroups < 1) return fail(h, SEQJ_E_BADARG, "bad argument");
  return fuse_impl(h, N64, ngroups, eta_group, mst_i, mst_j, mst_w, nullptr, nullptr, eps_floor, out);
}

int64_t seqj_pack_stride(int64_t N) { return 2 * N; }

int seqj_sequence_shard(seqj_handle* h, const double* A_dev, int64_t M, int64_t N, const int32_t* metric_ids, int nmetrics,
                        const int32_t* scales, int nscales, double eps_floor, double l2_thresh, const uint8_t* group_mask,
                        double* pack_dev, seqj_result** out) {
  if (!h) return SEQJ_E_BADARG;
  std::lock_guard<std::recursive_mutex> lock__(h->mu);      // one call at a time per handle
  if (!h) return SEQJ_E_BADARG;
  if (!group_mask || !pack_dev) return fail(h, SEQJ_E_BADARG, "seqj_sequence_shard needs a group mask and a pack buffer");
  return sequence_impl(h, A_dev, M, N, metric_ids, nmetrics, scales, nscales, eps_floor, l2_thresh, group_mask, nullptr,
                       nullptr, nullptr, 0, out, pack_dev);
}

int seqj_fuse_dev(seqj_handle* h, int64_t N64, int ngroups, const double* pack_dev, const int32_t* slot_of_group,
                  double eps_floor, seqj_result** out) {
  if (!h) return SEQJ_E_BADARG;
  std::lock_guard<std::recursive_mutex> lock__(h->mu);      // one call at a time per handle
  if (!h) return SEQJ_E_BADARG;
  if (!pack_dev || !slot_of_group || !out || N64 < 2 || ngroups < 1) return fail(h, SEQJ_E_BADARG, "bad argument");
  return fuse_impl(h, N64, ngroups, nullptr, nullptr, nullptr, nullptr, pack_dev, slot_of_group, eps_floor, out);
}

// ---- result accessors -------------------------------------------------------------------------
int seqj_result_dims(const seqj_result* r, int64_t* N, int* ngroups) {
  if (!r) return SEQJ_E_BADARG;
  if (N) *N = r->N;
  if (ngroups) *ngroups = (int)r->groups.size();
  return SEQJ_OK;
}
int seqj_result_group_info(const seqj_result* r, int g, int* metric, int* scale, int* nchunks, int* computed) {
  if (!r || g < 0 || g >= (int)r->groups.size()) return SEQJ_E_BADARG;
  const GroupRes& gr = r->groups[g];
  if (metric) *metric = gr.metric;
  if (scale) *scale = gr.scale;
  if (nchunks) *nchunks = gr.nchunks;
  if (computed) *computed = gr.computed;
  return SEQJ_OK;
}
int seqj_result_group_elapsed(const seqj_result* r, int g, double* ms) {
  if (!r || !ms || g < 0 || g >= (int)r->groups.size()) return SEQJ_E_BADARG;
  *ms = r->groups[g].elapsed_ms;
  return SEQJ_OK;
}
int seqj_result_group_range(const seqj_result* r, int g, int* chunk_lo, int* chunk_hi) {
  if (!r || g < 0 || g >= (int)r->groups.size()) return SEQJ_E_BADARG;
  if (chunk_lo) *chunk_lo = r->groups[g].chunk_lo;
  if (chunk_hi) *chunk_hi = r->groups[g].chunk_hi;
  return SEQJ_OK;
}
int seqj_result_group_chunks(const seqj_result* r, int g, double* eta_chunk, int32_t* root_chunk, int32_t* parents_chunk) {
  if (!r || g < 0 || g >= (int)r->groups.size() || !r->groups[g].computed) return SEQJ_E_BADARG;
  const GroupRes& gr = r->groups[g];
  copy_out(eta_chunk, gr.eta_chunk); copy_out(root_chunk, gr.root_chunk); copy_out(parents_chunk, gr.parents_chunk);
  return SEQJ_OK;
}
int seqj_result_group(const seqj_result* r, int g, double* eta, int32_t* root, int32_t* parents, int32_t* mst_i,
                      int32_t* mst_j, double* mst_w) {
  if (!r || g < 0 || g >= (int)r->groups.size() || !r->groups[g].computed) return SEQJ_E_BADARG;
  const GroupRes& gr = r->groups[g];
  if (eta) *eta = gr.eta;
  if (root) *root = gr.root;
  copy_out(parents, gr.parents); copy_out(mst_i, gr.mst_i); copy_out(mst_j, gr.mst_j); copy_out(mst_w, gr.mst_w);
  return SEQJ_OK;
}
int seqj_result_final(const seqj_result* r, double* eta, int32_t* root, int32_t* order, int32_t* parents,
                      int32_t* mst_i, int32_t* mst_j, double* mst_w) {
  if (!r || !r->has_final) return SEQJ_E_BADARG;
  if (eta) *eta = r->eta;
  if (root) *root = r->root;
  copy_out(order, r->order); copy_out(parents, r->parents); copy_out(mst_i, r->mst_i); copy_out(mst_j, r->mst_j);
  copy_out(mst_w, r->mst_w);
  return SEQJ_OK;
}
int seqj_result_D_nnz(const seqj_result* r, int64_t* nnz) {
  if (!r || !nnz) return SEQJ_E_BADARG;
  ensure_triplets(r);
  *nnz = (int64_t)r->Dv.size();
  return SEQJ_OK;
}
int seqj_result_D_triplets(const seqj_result* r, int32_t* i, int32_t* j, double* val) {
  if (!r) return SEQJ_E_BADARG;
  ensure_triplets(r);
  copy_out(i, r->Di); copy_out(j, r->Dj); copy_out(val, r->Dv);
  return SEQJ_OK;
}
int seqj_result_timings(const seqj_result* r, double* ms) {
  if (!r || !ms) return SEQJ_E_BADARG;
  std::memcpy(ms, r->timers, sizeof(r->timers));
  return SEQJ_OK;
}
int seqj_result_probes(const seqj_result* r, double* vals, int64_t n) {
  if (!r || !vals || n != (int64_t)r->probe_vals.size()) return SEQJ_E_BADARG;
  copy_out(vals, r->probe_vals);
  return SEQJ_OK;
}
int seqj_result_derived_groups(const seqj_result* r, int* n) {
  if (!r || !n) return SEQJ_E_BADARG;
  *n = r->derived_groups;
  return SEQJ_OK;
}
int seqj_result_waves(const seqj_result* r, int* n) {
  if (!r || !n) return SEQJ_E_BADARG;
  *n = r->waves;
  return SEQJ_OK;
}
int seqj_result_launches(const seqj_result* r, int64_t* launches) {
  if (!r || !launches) return SEQJ_E_BADARG;
  *launches = r->launches;
  return SEQJ_OK;
}
int seqj_result_input_range(const seqj_result* r, double* minv, double* maxv) {
  if (!r) return SEQJ_E_BADARG;
  if (minv) *minv = r->in_min;
  if (maxv) *maxv = r->in_max;
  return SEQJ_OK;
}
void seqj_result_free(seqj_result* r) { delete r; }

#include "api_units.inl"

}  // extern "C"
