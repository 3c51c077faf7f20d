// Host-side core of the C ABI library: handle, result container, device block pool, CUDA-event phase timers,
// chunk layouts, TMA tensor maps and the per-call context.  Included once, by seqj_api.cu.
#pragma once
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

// One persistent host thread per peer device (single-process multi-GPU): a thread that has used CUDA is expensive to
// create and to tear down, so the workers live as long as the handle and are handed one job per sequence() call.
struct DeviceWorker {
  std::thread th;
  std::mutex m;
  std::condition_variable cv;
  std::function<void()> job;
  bool has_job = false, done = false, stop = false;
  DeviceWorker() {
    th = std::thread([this]() {
      std::unique_lock<std::mutex> lk(m);
      while (true) {
        cv.wait(lk, [this]() { return has_job || stop; });
        if (stop) return;
        std::function<void()> j = std::move(job);
        has_job = false;
        lk.unlock();
        j();
        lk.lock();
        done = true;
        cv.notify_all();
      }
    });
  }
  void post(std::function<void()> j) {
    std::lock_guard<std::mutex> lk(m);
    job = std::move(j); has_job = true; done = false;
    cv.notify_all();
  }
  void wait() {
    std::unique_lock<std::mutex> lk(m);
    cv.wait(lk, [this]() { return done; });
  }
  ~DeviceWorker() {
    { std::lock_guard<std::mutex> lk(m); stop = true; }
    cv.notify_all();
    if (th.joinable()) th.join();
  }
};

struct seqj_handle {
  std::recursive_mutex mu;       // serialises the public entry points on this handle: its workspace cache, streams, pinned
                                 // buffers, grid / periods / probes and error string are per-handle state
  int dev = 0;
  cudaStream_t st = nullptr;     // compute stream: prep + distance tiles
  cudaStream_t st2 = nullptr;    // graph stream (higher priority): Prim, tree, combine, fuse
  std::string err;
  uint64_t ws_limit = 0;
  EncodeTiledFn encode = nullptr;
  int sm_count = 0;
  size_t max_smem = 0;
  // Device blocks kept between calls (exact-size reuse): repeated sequence() calls of the same shape do not
  // pay cudaMalloc/cudaFree of ~100 GB of workspace each time.  Released by seqj_release_workspace/destroy.
  std::vector<seqj_handle*> peers;   // additional devices driven by this handle (seqj_create with ndev > 1)
  std::vector<DeviceWorker*> workers; // one per peer, created on first use
  std::vector<double> grid;      // explicit grid for the EMD / Energy TYPES (seqj_set_grid), empty = none
  std::vector<double> periods;   // periods of SEQJ_PERIODIC (seqj_set_periods), empty = none
  // entries of the chunk / group matrices to sample during the next sequence calls (seqj_set_probes): group index
  // (metric-major), chunk (-1 = the combined group matrix Dkl), row, column
  std::vector<int32_t> probe_g, probe_c, probe_i, probe_j;
  std::multimap<size_t, void*> cache;
  size_t cached_bytes = 0;
  // pinned landing buffer + event of the input check: its result is read back asynchronously and only looked at
  // after the distance work has been enqueued, so the GPU never waits for the host at the start of a call
  void* stats_pinned = nullptr;
  cudaEvent_t stats_ev = nullptr;
  UploadStage upload;            // pinned staging of pageable inputs (host_upload.cuh), allocated on first use
};

struct GroupRes {
  int metric = 0, scale = 0, nchunks = 0, computed = 0;
  int chunk_lo = 0, chunk_hi = 0;     // chunks of this group held by this result (partial / sharded runs)
  std::vector<double> eta_chunk;
  std::vector<int32_t> root_chunk, parents_chunk;
  double eta = 0;
  int32_t root = 0;
  std::vector<int32_t> parents, mst_i, mst_j;
  std::vector<double> mst_w;
  double elapsed_ms = 0;              // device time attributed to this group (the reference's `tt = @elapsed`, src/sequencer.jl:210)
};

struct seqj_result {
  int64_t N = 0;
  std::vector<GroupRes> groups;
  int has_final = 0;
  double eta = 0;
  int32_t root = 0;
  std::vector<int32_t> order, parents, mst_i, mst_j, Di, Dj;
  std::vector<double> mst_w, Dv;
  double timers[SEQJ_NTIMERS] = {0};
  int64_t launches = 0;
  double in_min = 0, in_max = 0;
  int derived_groups = 0;
  int waves = 0;                      // HBM-fulls the chunk matrices were streamed in
  std::vector<double> probe_vals;     // one value per probe of the handle (NaN: the entry was not produced by this call)
  // raw group edge lists + D values; the deduplicated triplets (Di, Dj, Dv) are built on first request
  mutable std::vector<int32_t> raw_i, raw_j;
  mutable std::vector<double> raw_v;
  mutable int raw_G = 0;
  mutable int64_t raw_ldp = 0;
  mutable bool triplets_ready = false;
};

namespace {

int fail(seqj_handle* h, int code, const std::string& msg) {
  if (h) h->err = msg;
  return code;
}

#define CU_TRY(call)                                                                              \
  do {                                                                                            \
    cudaError_t e__ = (call);                                                                     \
    if (e__ != cudaSuccess)                                                                       \
      return fail(h, e__ == cudaErrorMemoryAllocation ? SEQJ_E_OOM : SEQJ_E_CUDA,                 \
                  std::string(#call) + ": " + cudaGetErrorString(e__));                           \
  } while (0)

#define ST_TRY(call)          \
  do {                        \
    int s__ = (call);         \
    if (s__ != SEQJ_OK) return s__; \
  } while (0)

inline int64_t round_up(int64_t x, int64_t m) { return (x + m - 1) / m * m; }

void release_cache(seqj_handle* h) {
  for (auto& kv : h->cache) cudaFree(kv.second);
  h->cache.clear();
  h->cached_bytes = 0;
}

// RAII device allocations; blocks go back to the handle's cache when the owning call returns.
struct DevPool {
  seqj_handle* h;
  std::vector<std::pair<size_t, void*>> blocks;
  explicit DevPool(seqj_handle* hh) : h(hh) {}
  ~DevPool() {
    // kernels of a call that failed half-way may still be queued on the handle's streams: the blocks must not be
    // handed to the next call under them (a completed call has synchronised already, so this costs nothing then)
    if (!blocks.empty()) {
      if (h->st) cudaStreamSynchronize(h->st);
      if (h->st2) cudaStreamSynchronize(h->st2);
    }
    for (auto& b : blocks) {
      h->cache.emplace(b.first, b.second);
      h->cached_bytes += b.first;
    }
  }
  // `larger_ok`: the request may be served by a larger cached block.  Used for the N x N buffers that dominate a call
  // -- the matrix pool of seqj_sequence, the proximity matrix of seqj_fuse / seqj_fuse_dev -- so that they hand the
  // same block (nearly all of the device memory) back and forth.  Without it the fuse's request went to cudaMalloc,
  // failed, released the whole cache, and the next seqj_sequence call did the same in the other direction: two full
  // free / malloc cycles of ~170 GB per step (seen at config 4 on 2 GPUs once the sharded upload had forced a
  // re-plan).  The block keeps its own size in the cache.
  template <typename T>
  cudaError_t alloc(T** out, size_t count, bool larger_ok = false) {
    const size_t bytes = (std::max<size_t>(count, 1) * sizeof(T) + 255) & ~size_t(255);
    void* p = nullptr;
    auto it = h->cache.find(bytes);
    if (it == h->cache.end() && larger_ok) it = h->cache.lower_bound(bytes);
    if (it != h->cache.end()) {
      p = it->second;
      const size_t have = it->first;
      h->cache.erase(it);
      h->cached_bytes -= have;
      blocks.emplace_back(have, p);
      *out = static_cast<T*>(p);
      return cudaSuccess;
    } else {
      cudaError_t e = cudaMalloc(&p, bytes);
      if (e == cudaErrorMemoryAllocation && !h->cache.empty()) {
        cudaGetLastError();
        release_cache(h);
        e = cudaMalloc(&p, bytes);
      }
      if (e != cudaSuccess) { *out = nullptr; return e; }
    }
    blocks.emplace_back(bytes, p);
    *out = static_cast<T*>(p);
    return cudaSuccess;
  }
};

struct PhaseTimer {
  struct Span { int phase, group, wave; cudaEvent_t a, b; };
  std::vector<Span> spans;
  cudaStream_t* cur;             // the context's current stream
  int group = -1, wave = -1;     // tags of the spans opened from now on: (metric, scale) group / wave they belong to
  std::vector<double> group_ms, wave_ms;   // filled by collect(): time of the spans tagged with a group / only a wave
  explicit PhaseTimer(cudaStream_t* s) : cur(s) {}
  ~PhaseTimer() {
    for (auto& s : spans) { cudaEventDestroy(s.a); cudaEventDestroy(s.b); }
  }
  int begin(int phase) {
    Span s; s.phase = phase; s.group = group; s.wave = wave;
    cudaEventCreate(&s.a); cudaEventCreate(&s.b);
    cudaEventRecord(s.a, *cur);
    spans.push_back(s);
    return (int)spans.size() - 1;
  }
  void end(int id) { cudaEventRecord(spans[id].b, *cur); }
  void collect(double* timers) {
    const bool trace = getenv("SEQJ_TRACE") != nullptr;      // per-span listing on stderr (debugging)
    for (auto& s : spans) {
      float ms = 0;
      if (cudaEventElapsedTime(&ms, s.a, s.b) == cudaSuccess) {
        timers[s.phase] += ms;
        if (s.phase != SEQJ_T_TOTAL && s.phase != SEQJ_T_D2H) {
          if (s.group >= 0) { if ((int)group_ms.size() <= s.group) group_ms.resize(s.group + 1, 0.0); group_ms[s.group] += ms; }
          else if (s.wave >= 0) { if ((int)wave_ms.size() <= s.wave) wave_ms.resize(s.wave + 1, 0.0); wave_ms[s.wave] += ms; }
        }
        if (trace) {
          float at = 0;
          cudaEventElapsedTime(&at, spans[0].a, s.a);
          fprintf(stderr, "[seqj] phase %2d  %9.3f ms  (starts at %9.3f ms)\n", s.phase, ms, at);
        }
      }
    }
  }
};

struct ScaleLayout {
  int scale = 1, cl = 0, cl_pad = 0, nchunks = 0;
  int64_t ldX = 0;
  std::vector<int> m_of_chunk;
};

// _splitnorm chunking (reference src/sequencer.jl:389-397): chunklen = cld(M, scale); starts 1:chunklen:M
ScaleLayout make_layout(int M, int scale) {
  ScaleLayout L;
  L.scale = scale;
  L.cl = (M + scale - 1) / scale;
  L.cl_pad = (int)round_up(L.cl, kTileK);
  for (int r0 = 0; r0 < M; r0 += L.cl) L.m_of_chunk.push_back(std::min(L.cl, M - r0));
  L.nchunks = (int)L.m_of_chunk.size();
  L.ldX = (int64_t)L.nchunks * L.cl_pad;
  return L;
}

std::vector<int2> make_tiles(int N, bool upper) {
  std::vector<int2> t;
  for (int j0 = 0; j0 < N; j0 += kTileN)
    for (int i0 = 0; i0 < N; i0 += kTileM)
      if (!upper || i0 < j0 + kTileN) t.push_back(make_int2(i0, j0));
  return t;
}

int make_tmap(seqj_handle* h, CUtensorMap* tm, const double* base, int64_t ldX, int64_t rows) {
  cuuint64_t dims[2] = {(cuuint64_t)ldX, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)ldX * 8};
  cuuint32_t box[2] = {(cuuint32_t)kTileK, (cuuint32_t)kTileM};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = h->encode(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, const_cast<double*>(base), dims, strides, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(h, SEQJ_E_CUDA, "cuTensorMapEncodeTiled failed: " + std::to_string((int)r));
  return SEQJ_OK;
}

// Everything the graph layer needs for up to `cap` graphs of N vertices.
struct GraphWS {
  int cap = 0, N = 0;
  int64_t ldv = 0;
  int use_smem = 1;
  size_t prim_smem = 0;
  int *parent = nullptr, *order = nullptr, *size = nullptr, *anc0 = nullptr, *anc1 = nullptr, *dep0 = nullptr,
      *dep1 = nullptr, *from_g = nullptr;
  double *weight = nullptr, *wd = nullptr, *wd2 = nullptr, *S = nullptr, *S2 = nullptr, *key_g = nullptr;
  int prim_R = 0;
  int* fuse_i0 = nullptr;         // scratch of the final tree's rooting (root_tree_kernel, 16 N + 64 ints)
  int cluster = 0;                // cluster Prim available (N large enough, slices fit in shared memory)
  int tree_use_smem = 1;
  size_t tree_smem = 0;
  // kNN-Boruvka MST for launches of few graphs (kernels_mst.cuh): its own buffers for up to knn_cap graphs; the
  // per-vertex scratch of the rounds lives in the tree kernel's buffers above (free until tree_kernel runs)
  int knn_cap = 0, knn_use_smem = 1;
  size_t knn_smem = 0;
  unsigned long long* knn_lk = nullptr;
  int *knn_lj = nullptr, *knn_ei = nullptr, *knn_ej = nullptr, *knn_root = nullptr, *knn_flags = nullptr;
  // first list pass from the upper triangle (knn_sample / knn_tiles / knn_merge kernels): candidate scratch for knn_tile_graphs
  // graphs at a time, nullptr = the row kernel reads the full matrices
  KnnCand* knn_cand = nullptr;
  unsigned long long* knn_tau = nullptr;
  int* knn_cnt = nullptr;
  int knn_tile_graphs = 0;
};

constexpr int kKnnRounds = 12;      // Boruvka launches per measurement (each runs rounds until rows need a rescan)

constexpr int kPrimCluster = 8;     // CTAs per graph; 2 and 4 measured slower at every graph count (profiles/r01_prim_variants.md)

// Slice width and dynamic shared memory of a C-CTA cluster Prim over N vertices.
void prim_cluster_geometry(int N, int C, int* W_out, size_t* smem_out) {
  const int W = (int)round_up((N + C - 1) / C, 2);
  if (W_out) *W_out = W;
  if (smem_out) *smem_out = (size_t)((W + 1) & ~1) * 8 + (size_t)W * 4 + 16;
}

typedef void (*PrimFn)(PrimParams);
PrimFn prim_fn(int R) {
  switch (R) {
    case 1: return prim_kernel<1>;
    case 2: return prim_kernel<2>;
    case 3: return prim_kernel<3>;
    case 4: return prim_kernel<4>;
    case 5: return prim_kernel<5>;
    case 6: return prim_kernel<6>;
    case 7: return prim_kernel<7>;
    case 8: return prim_kernel<8>;
    default: return prim_kernel<0>;
  }
}

struct Ctx {
  seqj_handle* h;
  cudaStream_t st;               // current stream: every run_* helper enqueues on it
  cudaStream_t st_compute, st_graph;
  DevPool pool;
  PhaseTimer timer;
  int64_t launches = 0;
  bool overlap = false;
  std::vector<cudaEvent_t> events;
  explicit Ctx(seqj_handle* hh) : h(hh), st(hh->st), st_compute(hh->st), st_graph(hh->st2), pool(hh), timer(&st) {
    // Default: one stream.  SEQJ_OVERLAP=1 runs the graph layer on a second stream so that it overlaps the
    // distance tiles of the next wave; measured on B200 (profiles/r01_overlap_experiments.md) the
    // interference costs more than it hides at config 3, so it is off until the graph kernels shrink.
    const char* ov = getenv("SEQJ_OVERLAP");
    overlap = ov && atoi(ov) != 0;
    if (!overlap) st_graph = st_compute;
  }
  ~Ctx() {
    for (auto e : events) cudaEventDestroy(e);
  }
  void on_compute() { st = st_compute; }
  void on_graph() { st = st_graph; }
  cudaEvent_t new_event() {
    cudaEvent_t e;
    cudaEventCreateWithFlags(&e, cudaEventDisableTiming);
    events.push_back(e);
    return e;
  }
};

}  // namespace
