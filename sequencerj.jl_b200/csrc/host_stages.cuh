// Host-side stages of one sequence() call: each run_* function enqueues the kernels of one stage on the call's
// stream (nothing here synchronises).  Included once, by seqj_api.cu, after host_core.cuh.
#pragma once

namespace {

int graphws_alloc(Ctx& c, GraphWS& g, int cap, int N) {
  seqj_handle* h = c.h;
  g.cap = cap; g.N = N; g.ldv = round_up(N, 32);
  const size_t n = (size_t)cap * g.ldv + 64;
  CU_TRY(c.pool.alloc(&g.parent, n)); CU_TRY(c.pool.alloc(&g.order, n)); CU_TRY(c.pool.alloc(&g.size, n));
  CU_TRY(c.pool.alloc(&g.anc0, n)); CU_TRY(c.pool.alloc(&g.anc1, n)); CU_TRY(c.pool.alloc(&g.dep0, n));
  CU_TRY(c.pool.alloc(&g.dep1, n)); CU_TRY(c.pool.alloc(&g.weight, n)); CU_TRY(c.pool.alloc(&g.wd, n));
  CU_TRY(c.pool.alloc(&g.S, n)); CU_TRY(c.pool.alloc(&g.wd2, n)); CU_TRY(c.pool.alloc(&g.S2, n));
  CU_TRY(c.pool.alloc(&g.fuse_i0, (size_t)16 * N + 64));
  g.prim_smem = (size_t)((N + 1) & ~1) * 8 + (size_t)N * 4;
  g.use_smem = g.prim_smem + 1024 <= h->max_smem;
  if (!g.use_smem) {
    g.prim_smem = 0;
    CU_TRY(c.pool.alloc(&g.key_g, n)); CU_TRY(c.pool.alloc(&g.from_g, n));
  }
  const int R = (N + 2 * kPrimThreads - 1) / (2 * kPrimThreads);
  g.prim_R = (g.use_smem && R <= 8) ? R : 0;
  CU_TRY(cudaFuncSetAttribute(prim_fn(g.prim_R), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)g.prim_smem));
  // cluster Prim for graphs large enough that one SM's load rate / argmin latency is the limit
  const char* force = getenv("SEQJ_PRIM");
  g.cluster = (N >= 1024) ? 1 : 0;
  if (force && !strcmp(force, "single")) g.cluster = 0;
  if (g.cluster) {
    size_t smem2 = 0;
    prim_cluster_geometry(N, kPrimCluster, nullptr, &smem2);
    if (smem2 + 4096 > h->max_smem) g.cluster = 0;
    else {
      CU_TRY(cudaFuncSetAttribute(prim_cluster_kernel<kPrimCluster, 512>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem2));
      CU_TRY(cudaFuncSetAttribute(prim_cluster_kernel<kPrimCluster, 256>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem2));
    }
  }
  g.tree_smem = (size_t)N * 8;
  g.tree_use_smem = g.tree_smem + 2048 <= h->max_smem;
  if (!g.tree_use_smem) g.tree_smem = 0;
  CU_TRY(cudaFuncSetAttribute(tree_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)g.tree_smem));
  // kNN-Boruvka for every launch whose lists fit 2 GB (config 3: 124 graphs = 0.33 GB); large batches run under a
  // rescan budget and hand graphs that exceed it to the batched Prim (run_knn_mst).  SEQJ_MST=prim turns the path
  // off, SEQJ_MST=small restores the round-1 rule (only launches of up to SMs/3 graphs).
  const char* menv = getenv("SEQJ_MST");
  const bool knn_small = menv && !strcmp(menv, "small");
  const size_t per_graph = (size_t)g.ldv * (kKnnK * 12 + 8) + ((size_t)16 * N + 64) * 4;
  const size_t knn_budget = size_t(2) << 30;
  g.knn_cap = (menv && !strcmp(menv, "prim")) ? 0
            : (int)std::min<size_t>(std::min<size_t>((size_t)cap, knn_small ? (size_t)std::max(1, h->sm_count / 3) : (size_t)cap),
                                    knn_budget / per_graph);
  if (N < 2) g.knn_cap = 0;
  if (g.knn_cap > 0) {
    const size_t nk = (size_t)g.knn_cap * g.ldv;
    CU_TRY(c.pool.alloc(&g.knn_lk, nk * kKnnK)); CU_TRY(c.pool.alloc(&g.knn_lj, nk * kKnnK));
    CU_TRY(c.pool.alloc(&g.knn_ei, nk)); CU_TRY(c.pool.alloc(&g.knn_ej, nk));
    CU_TRY(c.pool.alloc(&g.knn_root, (size_t)g.knn_cap * ((size_t)16 * N + 64)));
    CU_TRY(c.pool.alloc(&g.knn_flags, (size_t)4 * g.knn_cap + 16));
    g.knn_smem = (size_t)N * 4;
    g.knn_use_smem = g.knn_smem + 1024 <= h->max_smem;
    if (!g.knn_use_smem) g.knn_smem = 0;
    CU_TRY(cudaFuncSetAttribute(boruvka_knn_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)g.knn_smem));
    // Upper-triangle list pass (0.64 of the row kernel's bytes): thresholds from an N/8 sample of every row, so it
    // is used where that sample is large enough to be representative, N >= 6144.  SEQJ_KNN=rows / tiles forces either.
    const char* kenv = getenv("SEQJ_KNN");
    const bool tiles = kenv ? !strcmp(kenv, "tiles") : N >= 6144;
    if (tiles) {
      const size_t per = (size_t)N * kKtCap * sizeof(KnnCand);
      g.knn_tile_graphs = (int)std::max<size_t>(1, std::min<size_t>((size_t)g.knn_cap, (size_t(512) << 20) / per));
      CU_TRY(c.pool.alloc(&g.knn_cand, (size_t)N * kKtCap * g.knn_tile_graphs));
      CU_TRY(c.pool.alloc(&g.knn_tau, (size_t)N * g.knn_tile_graphs));
      CU_TRY(c.pool.alloc(&g.knn_cnt, (size_t)N * g.knn_tile_graphs));
    }
  }
  return SEQJ_OK;
}

// MST of `nb` dense matrices by kNN-filtered Boruvka (kernels_mst.cuh): lists, rounds with rescans, rooting.  Leaves
// (parent, weight, order) of every finished graph in the workspace and done[graph] = 1; the caller runs Prim with
// skip_done for whatever is left.
int run_knn_mst(Ctx& c, GraphWS& g, const double* D, int64_t ldD, int64_t strideD, int nb, double eps_floor,
                const int* mat_slot) {
  seqj_handle* h = c.h;
  int* ncount = g.knn_flags;
  int* ecount = g.knn_flags + g.knn_cap;
  int* done = g.knn_flags + 2 * g.knn_cap;
  KnnParams kp{};                                  // every field starts at 0 / nullptr
  kp.D = D; kp.ldD = ldD; kp.strideD = strideD; kp.mat_slot = mat_slot; kp.N = g.N; kp.eps_floor = eps_floor; kp.ldv = g.ldv;
  kp.lk = g.knn_lk; kp.lj = g.knn_lj; kp.cnt = g.size; kp.ptr = g.dep1;
  kp.bound = reinterpret_cast<unsigned long long*>(g.S);
  kp.comp = g.anc0; kp.needy = g.dep0; kp.ncount = ncount; kp.done = done;
  BorKnnParams bp{};
  bp.N = g.N; bp.ldv = g.ldv; bp.lk = g.knn_lk; bp.lj = g.knn_lj; bp.cnt = g.size; bp.bound = kp.bound;
  bp.ptr = g.dep1; bp.comp = g.anc0; bp.best = reinterpret_cast<unsigned long long*>(g.wd);
  bp.bestp = reinterpret_cast<unsigned long long*>(g.wd2); bp.stuck = g.anc1; bp.needy = g.dep0;
  bp.ncount = ncount; bp.ecount = ecount; bp.done = done; bp.nresc = g.knn_flags + 3 * g.knn_cap;
  // launches of few graphs may rescan freely (a row costs the same 8N bytes however many graphs share the launch,
  // and Prim would be latency-bound there).  In a large batch a graph may re-read two matrices worth of rows before it
  // is handed to the batched Prim, whose N dependent steps cost 25-65 ms whatever the number of graphs: with rescans
  // only when a graph is completely blocked and one list entry per component (kernels_mst.cuh), structured inputs
  // stay well below that (profiles/r02_mst.md)
  const char* benv = getenv("SEQJ_KNN_BUDGET");
  bp.rescan_budget = benv ? atoi(benv) : ((nb * 3 <= h->sm_count) ? 0x7fffffff : 2 * g.N);
  // A cap on a single request was tried and dropped: with rescans only when a graph is blocked, the smooth family
  // also asks for more than N/2 rows at once (80 ms instead of 44 with the cap, profiles/r02_mst.md)
  const char* oenv = getenv("SEQJ_KNN_ONCE");
  bp.rescan_once = oenv ? atoi(oenv) : 0x7fffffff;
  bp.out_i = g.knn_ei; bp.out_j = g.knn_ej; bp.out_w = g.S2; bp.use_smem = g.knn_use_smem; bp.first = 1;
  const dim3 grid_all((unsigned)((g.N + kKnnWarps - 1) / kKnnWarps), (unsigned)nb);
  // Rescan grid: at least 64 CTAs per graph.  The needy rows are spread very unevenly over the graphs of a batch (on
  // the clustered input a third of the 124 graphs asks for all N rows, the rest for none), and with SMs*8/nb = 9 CTAs
  // per graph the busy graphs kept 2-3 CTAs per SM in flight: 45 ms for one pass that the list kernel does in 17
  // (profiles/r02_mst.md).  CTAs without rows return at once.
  const int gx = std::max(1, std::min((g.N + kKnnWarps - 1) / kKnnWarps, std::max(h->sm_count * 8 / nb, 64)));
  const dim3 grid_rescan((unsigned)gx, (unsigned)nb);
  // rescans keep the graph's component labels in shared memory when they fit next to a few resident CTAs
  const size_t rescan_smem = ((size_t)g.N * 4 + 1024 <= h->max_smem) ? (size_t)g.N * 4 : 0;
  kp.comp_in_smem = rescan_smem != 0;
  if (rescan_smem) CU_TRY(cudaFuncSetAttribute(knn_rows_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)rescan_smem));
  int list_launches = 1;
  if (g.knn_cand) {
    KnnTileParams tp{};
    tp.D = D; tp.ldD = ldD; tp.strideD = strideD; tp.mat_slot = mat_slot; tp.N = g.N; tp.eps_floor = eps_floor;
    tp.nblk = (g.N + kKtB - 1) / kKtB; tp.ntiles = tp.nblk * (tp.nblk + 1) / 2;
    tp.tau = g.knn_tau; tp.cnt = g.knn_cnt; tp.cand = g.knn_cand;
    list_launches = 0;
    const unsigned rows_grid = (unsigned)((g.N + kKnnWarps - 1) / kKnnWarps);
    for (int g0 = 0; g0 < nb; g0 += g.knn_tile_graphs) {
      const int gb = std::min(g.knn_tile_graphs, nb - g0);
      tp.g0 = g0;
      CU_TRY(cudaMemsetAsync(g.knn_cnt, 0, sizeof(int) * (size_t)g.N * gb, c.st));
      knn_sample_kernel<<<dim3(rows_grid, (unsigned)gb), kKnnWarps * 32, 0, c.st>>>(tp);
      const dim3 tgrid((unsigned)((tp.ntiles + kKtWarps - 1) / kKtWarps), (unsigned)gb);
      // 4 rows per register set, 4 CTAs per SM: 19.99 ms for the MST stage of config 3; 8 rows / 2 CTAs: 22.8 ms
      knn_tiles_kernel<4, 4><<<tgrid, kKtWarps * 32, 0, c.st>>>(tp);
      knn_merge_kernel<<<dim3(rows_grid, (unsigned)gb), kKnnWarps * 32, 0, c.st>>>(kp, tp);
      list_launches += 3;
    }
  } else {
    knn_rows_kernel<false><<<grid_all, kKnnWarps * 32, 0, c.st>>>(kp);
  }
  const int rounds = getenv("SEQJ_KNN_ROUNDS") ? std::max(1, atoi(getenv("SEQJ_KNN_ROUNDS"))) : kKnnRounds;
  for (int r = 0; r < rounds; ++r) {
    if (r > 0) knn_rows_kernel<true><<<grid_rescan, kKnnWarps * 32, rescan_smem, c.st>>>(kp);
    bp.first = (r == 0);
    boruvka_knn_kernel<<<nb, kBorKnnThreads, g.knn_smem, c.st>>>(bp);
  }
  RootTreeParams rp;
  rp.N = g.N; rp.ei = g.knn_ei; rp.ej = g.knn_ej; rp.ew = g.S2; rp.scratch = g.knn_root;
  rp.parent = g.parent; rp.weight = g.weight; rp.order = g.order;
  rp.ld_e = g.ldv; rp.ld_s = (int64_t)16 * g.N + 64; rp.ld_o = g.ldv; rp.done = done;
  root_tree_kernel<<<nb, 1024, 0, c.st>>>(rp);
  c.launches += 2 * rounds + list_launches;
  CU_TRY(cudaGetLastError());
  return SEQJ_OK;
}

// _measure_dm for `nb` dense matrices: Prim + tree kernels.  Outputs go to device arrays.
int run_measure(Ctx& c, GraphWS& g, const double* D, int64_t ldD, int64_t strideD, int nb, double eps_floor,
                double* eta_out, int* root_out, int* parents_out, int64_t ldp, int* mst_i, int* mst_j, double* mst_w,
                int* dfs_order, const int* mat_slot = nullptr) {
  seqj_handle* h = c.h;
  if (nb > g.cap) return fail(h, SEQJ_E_INTERNAL, "graph workspace too small");
  PrimParams pp;
  pp.D = D; pp.ldD = ldD; pp.strideD = strideD; pp.N = g.N; pp.eps_floor = eps_floor; pp.mat_slot = mat_slot;
  pp.parent = g.parent; pp.weight = g.weight; pp.order = g.order; pp.ldv = g.ldv;
  pp.key_g = g.key_g; pp.from_g = g.from_g; pp.use_smem = g.use_smem;
  pp.prefetch = getenv("SEQJ_PRIM_PREFETCH") ? atoi(getenv("SEQJ_PRIM_PREFETCH")) : 1;
  int t = c.timer.begin(SEQJ_T_PRIM);
  if (g.knn_cap > 0 && nb <= g.knn_cap) {
    ST_TRY(run_knn_mst(c, g, D, ldD, strideD, nb, eps_floor, mat_slot));
    pp.skip_done = g.knn_flags + 2 * g.knn_cap;
  }
  // Measured on B200 at N = 10k (profiles/r01_prim_variants.md): with >= ~50 graphs in flight the single-CTA
  // kernel streams 4.3 TB/s (66 % of HBM) and beats the cluster kernel, whose 8x more CTAs only add
  // contention; with few graphs every step is latency-bound and the cluster kernel's shorter step wins.
  const bool force_cluster = getenv("SEQJ_PRIM") && !strcmp(getenv("SEQJ_PRIM"), "cluster");
  // cluster kernel while every cluster is resident at once: 512-thread CTAs (2 per SM) up to 2*SMs/8 graphs,
  // 256-thread CTAs up to SMs/3 graphs (profiles/r01_prim_variants.md); beyond that one CTA per graph
  int W = 0;
  size_t smem = 0;
  prim_cluster_geometry(g.N, kPrimCluster, &W, &smem);
  if (g.cluster && nb * kPrimCluster <= 2 * h->sm_count)
    prim_cluster_kernel<kPrimCluster, 512><<<nb * kPrimCluster, 512, smem, c.st>>>(pp, W);
  else if (g.cluster && (force_cluster || nb * 3 <= h->sm_count))
    prim_cluster_kernel<kPrimCluster, 256><<<nb * kPrimCluster, 256, smem, c.st>>>(pp, W);
  else
    prim_fn(g.prim_R)<<<nb, kPrimThreads, g.prim_smem, c.st>>>(pp);
  c.timer.end(t); c.launches++;
  CU_TRY(cudaGetLastError());
  TreeParams tp;
  tp.N = g.N; tp.ldv = g.ldv; tp.parent = g.parent; tp.weight = g.weight; tp.order = g.order;
  tp.size = g.size; tp.wd = g.wd; tp.wd2 = g.wd2; tp.S = g.S; tp.S2 = g.S2; tp.use_smem = g.tree_use_smem; tp.anc0 = g.anc0; tp.anc1 = g.anc1; tp.dep0 = g.dep0; tp.dep1 = g.dep1;
  tp.eta = eta_out; tp.root = root_out; tp.newparent = parents_out; tp.ldp = ldp;
  tp.mst_i = mst_i; tp.mst_j = mst_j; tp.mst_w = mst_w; tp.dfs_order = dfs_order;
  t = c.timer.begin(SEQJ_T_TREE);
  tree_kernel<<<nb, kTreeThreads, g.tree_smem, c.st>>>(tp);
  c.timer.end(t); c.launches++;
  CU_TRY(cudaGetLastError());
  return SEQJ_OK;
}

struct PrepBufs {
  double *P = nullptr, *logP = nullptr, *Uc = nullptr, *U = nullptr, *sP = nullptr, *sU = nullptr, *H = nullptr;
  double *UwE = nullptr, *UwG = nullptr, *sUG = nullptr, *grid = nullptr, *period = nullptr;
  double *mass = nullptr, *sR = nullptr, *sT = nullptr;
  int64_t ldn = 0, rows_pad = 0;
  int* m_of_chunk = nullptr;
  int f32_pdf = 0;               // Float32 input: PDFs rounded through Float32 as the reference computes them (Q2)
};

int run_prep(Ctx& c, const double* A_dev, int64_t ldA, int M, int N, const ScaleLayout& L, PrepBufs& b,
             int normalize) {
  seqj_handle* h = c.h;
  PrepParams pp;
  pp.A = A_dev; pp.ldA = ldA; pp.M = M; pp.N = N; pp.cl = L.cl; pp.cl_pad = L.cl_pad; pp.nchunks = L.nchunks;
  pp.ldX = L.ldX; pp.ldn = b.ldn; pp.P = b.P; pp.logP = b.logP; pp.Uc = b.Uc; pp.U = b.U;
  pp.sP = b.sP; pp.sU = b.sU; pp.H = b.H; pp.normalize = normalize;
  pp.grid = b.grid; pp.UwE = b.UwE; pp.UwG = b.UwG; pp.sUG = b.sUG;
  pp.mass = b.mass; pp.sR = b.sR; pp.sT = b.sT;
  pp.need_log = b.logP != nullptr;
  pp.need_cdf = b.Uc != nullptr || b.U != nullptr || b.UwE != nullptr || b.UwG != nullptr;
  pp.f32_pdf = b.f32_pdf;
  int t = c.timer.begin(SEQJ_T_PREP);
  // rows [N, rows_pad) are read by TMA boxes of the last tiles: keep them zero
  const size_t tail = (size_t)(b.rows_pad - N) * L.ldX * 8;
  if (tail) {
    if (b.P) CU_TRY(cudaMemsetAsync(b.P + (size_t)N * L.ldX, 0, tail, c.st));
    if (b.logP) CU_TRY(cudaMemsetAsync(b.logP + (size_t)N * L.ldX, 0, tail, c.st));
    if (b.Uc) CU_TRY(cudaMemsetAsync(b.Uc + (size_t)N * L.ldX, 0, tail, c.st));
    if (b.UwE) CU_TRY(cudaMemsetAsync(b.UwE + (size_t)N * L.ldX, 0, tail, c.st));
    if (b.UwG) CU_TRY(cudaMemsetAsync(b.UwG + (size_t)N * L.ldX, 0, tail, c.st));
  }
  CU_TRY(cudaMemcpyAsync(b.m_of_chunk, L.m_of_chunk.data(), sizeof(int) * L.nchunks, cudaMemcpyHostToDevice, c.st));
  const int64_t warps = (int64_t)N * L.nchunks;
  const int wpb = 8;
  prep_kernel<<<(unsigned)((warps + wpb - 1) / wpb), wpb * 32, 0, c.st>>>(pp);
  c.timer.end(t); c.launches++;
  CU_TRY(cudaGetLastError());
  return SEQJ_OK;
}

struct FixBufs {
  int2* dense_list = nullptr;            // compacted (chunk_local, tile) entries of the flagged tiles
  unsigned int* dense_count = nullptr;
  unsigned char* tile_flags = nullptr;   // [max chunks per launch][ntiles] dense-tile flags
  size_t tile_flags_bytes = 0;
  int4* list = nullptr;
  unsigned int* count = nullptr;
  int* overflow = nullptr;
  unsigned int cap = 0;
};

// Relative fix-up thresholds (see kernels_dist.cuh).  Rounding in the DMMA accumulation over m terms behaves like
// a random walk: |dv| ~ sqrt(m/4) * 2^-53 * (s_i + s_j) ~ 4e-15 * selfterms at m = 4096 (worst case m/4 times
// that).  A distance is within 1e-9 relative when dv / v <= 2e-9, i.e. v >= 2e-6 * selfterms typically;
// 1e-4 leaves a 50x margin and keeps the flagged set small on smooth data (neighbouring objects that differ
// by < 1 %).  KL: |d(H - C)| ~ 3e-14 against results compared as d + 1e-6, flagged below 1e-4 * |H|.
constexpr double kTauGram = 1e-4;
constexpr double kTauKL = 1e-4;

// Dense fix-up: compact the tile flags into a list, then a small persistent grid of the direct tile kernel
// recomputes the listed tiles from the operands (no Gram form).  With nothing flagged this costs two tiny launches.
int run_dense_tiles(Ctx& c, int metric, DistParams dd, const CUtensorMap& tm1, const CUtensorMap& tm2, FixBufs& fb,
                    int ntiles, int nb) {
  seqj_handle* h = c.h;
  CU_TRY(cudaMemsetAsync(fb.dense_count, 0, sizeof(unsigned int), c.st));
  const int total = ntiles * nb;
  compact_flags_kernel<<<(total + 255) / 256, 256, 0, c.st>>>(fb.tile_flags, total, ntiles, fb.dense_list, fb.dense_count);
  dd.dense_list = fb.dense_list; dd.dense_count = fb.dense_count;
  const int grid = h->sm_count * 2 * 4;
  if (metric == SEQJ_L2) direct_dist_kernel<OP_SQ><<<grid, kTileThreads, kTileSmemBytes, c.st>>>(tm1, dd);
  else if (metric == SEQJ_ENERGY) direct_dist_kernel<OP_SQ_ENERGY><<<grid, kTileThreads, kTileSmemBytes, c.st>>>(tm1, dd);
  else kl_direct_kernel<<<grid, kTileThreads, kKlSmemBytes, c.st>>>(tm1, tm2, dd);
  c.launches += 2;
  CU_TRY(cudaGetLastError());
  return SEQJ_OK;
}

// Device copy of the PeriodicEuclidean periods, padded to the chunk's padded length with 1.0 (the operands are 0 there)
int upload_periods(Ctx& c, PrepBufs& pb, const std::vector<double>& periods, int cl_pad) {
  seqj_handle* h = c.h;
  std::vector<double> padded((size_t)cl_pad, 1.0);
  std::copy(periods.begin(), periods.end(), padded.begin());
  CU_TRY(c.pool.alloc(&pb.period, (size_t)cl_pad));
  CU_TRY(cudaMemcpy(pb.period, padded.data(), sizeof(double) * cl_pad, cudaMemcpyHostToDevice));
  return SEQJ_OK;
}

// distance matrices of chunks [chunk0, chunk0+nb) of one scale into D (stride strideD)
int run_distance(Ctx& c, int metric, const ScaleLayout& L, const PrepBufs& pb, const int2* tiles_dev, int ntiles,
                 int N, double* D, int64_t ldD, int64_t strideD, int chunk0, int nb, double eps_floor,
                 double l2_thresh, int kl_full, FixBufs& fb) {
  seqj_handle* h = c.h;
  DistParams dp;
  dp.N = N; dp.ldD = ldD; dp.strideD = strideD; dp.D = D; dp.nk = L.cl_pad / kTileK; dp.cl_pad = L.cl_pad;
  dp.chunk0 = chunk0; dp.ldn = pb.ldn; dp.tiles = tiles_dev; dp.eps_floor = eps_floor;
  dp.mirror = 1; dp.kl_swap = 0;
  dp.tile_flags = nullptr; dp.tile_flag_stride = ntiles; dp.dense_thresh = 0x7fffffff; dp.dense_list = nullptr; dp.dense_count = nullptr;
  dp.fix_list = fb.list; dp.fix_count = fb.count; dp.fix_cap = fb.cap; dp.fix_overflow = fb.overflow;
  dp.out_scale = 1.0;
  if (metric >= SEQJ_CITYBLOCK && metric <= SEQJ_SQEUCLIDEAN) {
    // element-wise Distances.jl metrics on the PDFs: direct tiles (same TMA pipeline and register tile as EMD), timed
    // with the Euclidean phase
    CUtensorMap tmP;
    ST_TRY(make_tmap(h, &tmP, pb.P, L.ldX, pb.rows_pad));
    dp.normA = nullptr; dp.tau = 0; dp.period = nullptr;
    int t0 = c.timer.begin(SEQJ_T_DIST_L2);
    dim3 grid0(ntiles, nb);
    if (metric == SEQJ_CITYBLOCK) direct_dist_kernel<OP_L1><<<grid0, kTileThreads, kTileSmemBytes, c.st>>>(tmP, dp);
    else if (metric == SEQJ_TOTALVARIATION) { dp.out_scale = 0.5; direct_dist_kernel<OP_L1><<<grid0, kTileThreads, kTileSmemBytes, c.st>>>(tmP, dp); }
    else if (metric == SEQJ_CHEBYSHEV) direct_dist_kernel<OP_MAX><<<grid0, kTileThreads, kTileSmemBytes, c.st>>>(tmP, dp);
    else direct_dist_kernel<OP_SQ_RAW><<<grid0, kTileThreads, kTileSmemBytes, c.st>>>(tmP, dp);
    c.launches++;
    CU_TRY(cudaGetLastError());
    c.timer.end(t0);
    return SEQJ_OK;
  }
  const bool periodic = metric == SEQJ_PERIODIC;
  if (periodic) metric = SEQJ_L2;                      // timed with the Euclidean phase
  const bool on_grid = metric >= SEQJ_EMD_GRID;        // EMD / Energy on the explicit grid: pre-scaled operands
  if (on_grid) metric = (metric == SEQJ_EMD_GRID) ? SEQJ_EMD : SEQJ_ENERGY;
  dp.period = nullptr;
  const int phase = SEQJ_T_DIST_L2 + metric;
  int t = c.timer.begin(phase);
  dim3 grid(ntiles, nb);
  if (periodic) {                 // Distances.PeriodicEuclidean: element-wise fold, direct tiles on the PDFs
    CUtensorMap tmP;
    ST_TRY(make_tmap(h, &tmP, pb.P, L.ldX, pb.rows_pad));
    dp.normA = nullptr; dp.tau = 0; dp.period = pb.period;
    direct_dist_kernel<OP_PERIODIC><<<grid, kTileThreads, kTileSmemBytes, c.st>>>(tmP, dp);
    c.launches++;
    CU_TRY(cudaGetLastError());
    c.timer.end(t);
    return SEQJ_OK;
  }
  if (metric == SEQJ_EMD) {
    CUtensorMap tmU;
    ST_TRY(make_tmap(h, &tmU, on_grid ? pb.UwE : pb.Uc, L.ldX, pb.rows_pad));
    dp.normA = nullptr; dp.tau = 0; dp.tile_flags = nullptr; dp.dense_list = nullptr; dp.dense_count = nullptr;
    direct_dist_kernel<OP_L1><<<grid, kTileThreads, kTileSmemBytes, c.st>>>(tmU, dp);
    c.launches++;
    CU_TRY(cudaGetLastError());
    c.timer.end(t);
    return SEQJ_OK;
  }
  CU_TRY(cudaMemsetAsync(fb.count, 0, sizeof(unsigned int), c.st));
  CU_TRY(cudaMemsetAsync(fb.overflow, 0, sizeof(int), c.st));
  CUtensorMap tmA, tmB;
  // dense-tile path (L2 / Energy with mirrored output): tiles with > 1/32 of their entries flagged
  const bool dense_ok = (metric == SEQJ_L2 || metric == SEQJ_ENERGY || (metric == SEQJ_KLD && !kl_full)) && fb.tile_flags &&
                        (size_t)nb * ntiles <= fb.tile_flags_bytes && !(getenv("SEQJ_DENSE_TILES") && atoi(getenv("SEQJ_DENSE_TILES")) == 0);
  if (dense_ok) {
    CU_TRY(cudaMemsetAsync(fb.tile_flags, 0, (size_t)nb * ntiles, c.st));
    dp.tile_flags = fb.tile_flags; dp.dense_thresh = kTileM * kTileN / 32;
  }
  FixParams fp;
  fp.mode = metric; fp.N = N; fp.ldD = ldD; fp.strideD = strideD; fp.D = D; fp.ldX = L.ldX; fp.cl_pad = L.cl_pad;
  fp.chunk0 = chunk0; fp.m_of_chunk = pb.m_of_chunk; fp.eps_floor = eps_floor; fp.mirror = 1; fp.kl_swap = 0;
  fp.fix_list = fb.list; fp.fix_count = fb.count; fp.fix_cap = fb.cap; fp.fix_overflow = fb.overflow;
  if (metric == SEQJ_L2) {
    ST_TRY(make_tmap(h, &tmA, pb.P, L.ldX, pb.rows_pad));
    dp.normA = pb.sP; dp.tau = std::max(kTauGram, l2_thresh); fp.X = pb.P;
    gemm_dist_kernel<MODE_L2><<<grid, kTileThreads, kTileSmemBytes, c.st>>>(tmA, tmA, dp);
  } else if (metric == SEQJ_ENERGY) {
    ST_TRY(make_tmap(h, &tmA, on_grid ? pb.UwG : pb.Uc, L.ldX, pb.rows_pad));
    dp.normA = on_grid ? pb.sUG : pb.sU; dp.tau = kTauGram; fp.X = on_grid ? pb.UwG : pb.Uc;
    gemm_dist_kernel<MODE_ENERGY><<<grid, kTileThreads, kTileSmemBytes, c.st>>>(tmA, tmA, dp);
  } else {  // KL
    ST_TRY(make_tmap(h, &tmA, pb.P, L.ldX, pb.rows_pad));
    ST_TRY(make_tmap(h, &tmB, pb.logP, L.ldX, pb.rows_pad));
    dp.normA = pb.H; dp.tau = kTauKL; fp.X = pb.P;
    if (kl_full) {
      dp.mirror = 0; dp.kl_swap = 1; fp.mirror = 0; fp.kl_swap = 1;
      gemm_dist_kernel<MODE_KL><<<grid, kTileThreads, kTileSmemBytes, c.st>>>(tmB, tmA, dp);
    } else {
      gemm_dist_kernel<MODE_KL><<<grid, kTileThreads, kTileSmemBytes, c.st>>>(tmA, tmB, dp);
    }
  }
  c.launches++;
  CU_TRY(cudaGetLastError());
  fixup_list_kernel<<<h->sm_count * 4, 256, 0, c.st>>>(fp);
  fixup_dense_kernel<<<h->sm_count * 4, 256, 0, c.st>>>(fp, nb);
  c.launches += 2;
  if (dense_ok)              // whole tiles full of cancelling entries: recompute them without the Gram form
    ST_TRY(run_dense_tiles(c, metric, dp, tmA, metric == SEQJ_KLD ? tmB : tmA, fb, ntiles, nb));
  CU_TRY(cudaGetLastError());
  c.timer.end(t);
  return SEQJ_OK;
}

// Fine-scale per-chunk vectors kept aside while the prep buffers are reused for the coarser scales.
struct FineVecs {
  double *mass = nullptr, *sP = nullptr, *sU = nullptr, *sR = nullptr, *sT = nullptr;
  double *fac = nullptr, *pm = nullptr;     // scratch of derive_factors_kernel: [nchunks][4][ldn] + tail padding, [nchunks][3]
  int* m_of_chunk = nullptr;
  int nchunks = 0, cl = 0;
};

// chunk matrices [c0, c1) of a coarse (metric, scale) group derived from the finished fine matrices
int run_derive(Ctx& c, int metric, const ScaleLayout& L, const PrepBufs& pb, const FineVecs& fv, const double* Dfine,
               int fbase, int N, double* D, int64_t ldD, int64_t strideD, int c0, int c1, double eps_floor, double l2_thresh,
               FixBufs& fb, const int2* tiles_dev, int ntiles, const int* tile_col_offset) {
  seqj_handle* h = c.h;
  int t = c.timer.begin(SEQJ_T_DERIVE);
  CU_TRY(cudaMemsetAsync(fb.count, 0, sizeof(unsigned int), c.st));
  CU_TRY(cudaMemsetAsync(fb.overflow, 0, sizeof(int), c.st));
  const int ratio = L.cl / fv.cl;
  DeriveParams dp;
  dp.mode = metric; dp.N = N; dp.ldD = ldD; dp.strideD = strideD; dp.Dfine = Dfine; dp.f_base = fbase; dp.ldn = pb.ldn;
  dp.fmass = fv.mass; dp.fg = (metric == SEQJ_ENERGY) ? fv.sU : fv.sP; dp.fR = fv.sR; dp.fT = fv.sT; dp.fm = fv.m_of_chunk;
  dp.eps_floor = eps_floor;
  dp.tau = (metric == SEQJ_L2) ? std::max(kTauGram, l2_thresh) : (metric == SEQJ_KLD ? kTauKL : kTauGram);
  dp.fix_list = fb.list; dp.fix_count = fb.count; dp.fix_cap = fb.cap; dp.fix_overflow = fb.overflow;
  const bool dense_ok = fb.tile_flags && tile_col_offset && (size_t)(c1 - c0) * ntiles <= fb.tile_flags_bytes &&
                        !(getenv("SEQJ_DENSE_TILES") && atoi(getenv("SEQJ_DENSE_TILES")) == 0);
  dp.tile_flags = dense_ok ? fb.tile_flags : nullptr; dp.tile_flag_stride = ntiles; dp.tile_col_offset = tile_col_offset;
  if (dense_ok) CU_TRY(cudaMemsetAsync(fb.tile_flags, 0, (size_t)(c1 - c0) * ntiles, c.st));
  const unsigned T = (unsigned)((N + kDeriveTile - 1) / kDeriveTile);
  dp.ratio = ratio; dp.nfine = fv.nchunks; dp.c_first = c0; dp.f0 = 0; dp.nf = 0; dp.chunk_local = 0;
  dp.Dout = D;
  dp.cmass = pb.mass;
  dp.cself = (metric == SEQJ_L2) ? pb.sP : (metric == SEQJ_KLD ? pb.H : pb.sU);
  const size_t dsmem = derive_smem_bytes(metric, std::min(ratio, fv.nchunks));
  const dim3 dgrid(T, T, (unsigned)(c1 - c0)), dblock(32, 8);
  dp.fac = fv.fac; dp.pm = fv.pm;
  const unsigned fgrid = (unsigned)(((int64_t)fv.nchunks * pb.ldn + 255) / 256);
  if (metric == SEQJ_L2) {
    derive_factors_kernel<MODE_L2><<<fgrid, 256, 0, c.st>>>(dp);
    derive_kernel<MODE_L2><<<dgrid, dblock, dsmem, c.st>>>(dp);
  } else if (metric == SEQJ_KLD) {
    derive_factors_kernel<MODE_KL><<<fgrid, 256, 0, c.st>>>(dp);
    derive_kernel<MODE_KL><<<dgrid, dblock, dsmem, c.st>>>(dp);
  } else {
    derive_factors_kernel<MODE_ENERGY><<<fgrid, 256, 0, c.st>>>(dp);
    derive_kernel<MODE_ENERGY><<<dgrid, dblock, dsmem, c.st>>>(dp);
  }
  c.launches++;
  c.launches++;
  CU_TRY(cudaGetLastError());
  FixParams fp;
  fp.mode = metric; fp.N = N; fp.ldD = ldD; fp.strideD = strideD; fp.D = D; fp.ldX = L.ldX; fp.cl_pad = L.cl_pad;
  fp.chunk0 = c0; fp.m_of_chunk = pb.m_of_chunk; fp.eps_floor = eps_floor; fp.mirror = 1; fp.kl_swap = 0;
  fp.fix_list = fb.list; fp.fix_count = fb.count; fp.fix_cap = fb.cap; fp.fix_overflow = fb.overflow;
  fp.X = (metric == SEQJ_ENERGY) ? pb.Uc : pb.P;
  fixup_list_kernel<<<h->sm_count * 4, 256, 0, c.st>>>(fp);
  fixup_dense_kernel<<<h->sm_count * 4, 256, 0, c.st>>>(fp, c1 - c0);
  c.launches += 2;
  if (dense_ok) {            // tiles full of near-duplicate pairs: recompute them directly from the coarse operands
    DistParams dd;
    dd.N = N; dd.ldD = ldD; dd.strideD = strideD; dd.D = D; dd.nk = L.cl_pad / kTileK; dd.cl_pad = L.cl_pad;
    dd.chunk0 = c0; dd.normA = nullptr; dd.ldn = pb.ldn; dd.tiles = tiles_dev; dd.eps_floor = eps_floor; dd.tau = 0;
    dd.mirror = 1; dd.kl_swap = 0; dd.tile_flags = fb.tile_flags; dd.tile_flag_stride = ntiles; dd.dense_thresh = 0;
    dd.dense_list = nullptr; dd.dense_count = nullptr;
    dd.fix_list = fb.list; dd.fix_count = fb.count; dd.fix_cap = fb.cap; dd.fix_overflow = fb.overflow;
    CUtensorMap tm1, tm2;
    if (metric == SEQJ_ENERGY) {
      ST_TRY(make_tmap(h, &tm1, pb.Uc, L.ldX, pb.rows_pad));
      tm2 = tm1;
    } else {
      ST_TRY(make_tmap(h, &tm1, pb.P, L.ldX, pb.rows_pad));
      if (metric == SEQJ_KLD) ST_TRY(make_tmap(h, &tm2, pb.logP, L.ldX, pb.rows_pad));
      else tm2 = tm1;
    }
    ST_TRY(run_dense_tiles(c, metric, dd, tm1, tm2, fb, ntiles, c1 - c0));
  }
  CU_TRY(cudaGetLastError());
  c.timer.end(t);
  return SEQJ_OK;
}

int set_kernel_attrs(seqj_handle* h) {
  CU_TRY(cudaFuncSetAttribute(gemm_dist_kernel<MODE_L2>, cudaFuncAttributeMaxDynamicSharedMemorySize, kTileSmemBytes));
  CU_TRY(cudaFuncSetAttribute(gemm_dist_kernel<MODE_ENERGY>, cudaFuncAttributeMaxDynamicSharedMemorySize, kTileSmemBytes));
  CU_TRY(cudaFuncSetAttribute(gemm_dist_kernel<MODE_KL>, cudaFuncAttributeMaxDynamicSharedMemorySize, kTileSmemBytes));
  CU_TRY(cudaFuncSetAttribute(direct_dist_kernel<OP_L1>, cudaFuncAttributeMaxDynamicSharedMemorySize, kTileSmemBytes));
  CU_TRY(cudaFuncSetAttribute(direct_dist_kernel<OP_SQ>, cudaFuncAttributeMaxDynamicSharedMemorySize, kTileSmemBytes));
  CU_TRY(cudaFuncSetAttribute(direct_dist_kernel<OP_SQ_ENERGY>, cudaFuncAttributeMaxDynamicSharedMemorySize, kTileSmemBytes));
  CU_TRY(cudaFuncSetAttribute(direct_dist_kernel<OP_PERIODIC>, cudaFuncAttributeMaxDynamicSharedMemorySize, kTileSmemBytes));
  CU_TRY(cudaFuncSetAttribute(direct_dist_kernel<OP_MAX>, cudaFuncAttributeMaxDynamicSharedMemorySize, kTileSmemBytes));
  CU_TRY(cudaFuncSetAttribute(direct_dist_kernel<OP_SQ_RAW>, cudaFuncAttributeMaxDynamicSharedMemorySize, kTileSmemBytes));
  CU_TRY(cudaFuncSetAttribute(kl_direct_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kKlSmemBytes));
  CU_TRY(cudaFuncSetAttribute(derive_kernel<MODE_L2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)derive_smem_bytes(MODE_L2, kDeriveMaxFine)));
  CU_TRY(cudaFuncSetAttribute(derive_kernel<MODE_KL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)derive_smem_bytes(MODE_KL, kDeriveMaxFine)));
  CU_TRY(cudaFuncSetAttribute(derive_kernel<MODE_ENERGY>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)derive_smem_bytes(MODE_ENERGY, kDeriveMaxFine)));
  return SEQJ_OK;
}

int alloc_prep(Ctx& c, PrepBufs& pb, int N, int64_t max_ldX, int max_chunks, bool needP, bool needLog, bool needUc,
               bool needU, bool needWE = false, bool needWG = false) {
  seqj_handle* h = c.h;
  pb.rows_pad = round_up(N, kTileN);
  pb.ldn = round_up(N, 8);
  const size_t n = (size_t)pb.rows_pad * max_ldX;
  if (needP) CU_TRY(c.pool.alloc(&pb.P, n));
  if (needLog) CU_TRY(c.pool.alloc(&pb.logP, n));
  if (needUc) CU_TRY(c.pool.alloc(&pb.Uc, n));
  if (needU) CU_TRY(c.pool.alloc(&pb.U, n));
  if (needWE) CU_TRY(c.pool.alloc(&pb.UwE, n));
  if (needWG) CU_TRY(c.pool.alloc(&pb.UwG, n));
  const size_t nn = (size_t)max_chunks * pb.ldn;
  CU_TRY(c.pool.alloc(&pb.sP, nn)); CU_TRY(c.pool.alloc(&pb.sU, nn)); CU_TRY(c.pool.alloc(&pb.H, nn));
  CU_TRY(c.pool.alloc(&pb.sUG, nn));
  CU_TRY(c.pool.alloc(&pb.mass, nn)); CU_TRY(c.pool.alloc(&pb.sR, nn)); CU_TRY(c.pool.alloc(&pb.sT, nn));
  CU_TRY(c.pool.alloc(&pb.m_of_chunk, (size_t)max_chunks));
  return SEQJ_OK;
}

int alloc_fix(Ctx& c, FixBufs& fb, unsigned int cap, size_t tile_flags_bytes = 0) {
  seqj_handle* h = c.h;
  fb.cap = cap;
  fb.tile_flags_bytes = tile_flags_bytes;
  if (tile_flags_bytes) {
    CU_TRY(c.pool.alloc(&fb.tile_flags, tile_flags_bytes));
    CU_TRY(c.pool.alloc(&fb.dense_list, tile_flags_bytes));
    CU_TRY(c.pool.alloc(&fb.dense_count, 1));
  }
  CU_TRY(c.pool.alloc(&fb.list, (size_t)cap));
  CU_TRY(c.pool.alloc(&fb.count, 1));
  CU_TRY(c.pool.alloc(&fb.overflow, 1));
  return SEQJ_OK;
}

// final fuse on device: group MSTs (device arrays, [G][ldp]) -> D (in Sbuf) -> final measure
// `perm[g]` = index (in the device arrays) of the g-th group in the reference's metric-major order;
// the scatter and the eta sum follow that order so the rounding matches the reference's loop.
int run_fuse(Ctx& c, GraphWS& gws, double* Sbuf, int64_t ldD, int N, int G, const double* eta_g, const int* mst_i,
             const int* mst_j, const double* mst_w, int64_t ldp, double eps_floor, double* eta_f, int* root_f,
             int* parents_f, int* fmst_i, int* fmst_j, double* fmst_w, int* order_f, double* edge_val,
             const std::vector<int>& perm, const int* perm_dev) {
  seqj_handle* h = c.h;
  int t = c.timer.begin(SEQJ_T_FUSE);
  CU_TRY(cudaMemsetAsync(Sbuf, 0, (size_t)N * ldD * 8, c.st));
  const int ne = N - 1;
  for (int gg = 0; gg < G; ++gg) {
    const int g = perm[gg];
    fuse_scatter_kernel<<<(ne + 255) / 256, 256, 0, c.st>>>(Sbuf, ldD, mst_i + (int64_t)g * ldp, mst_j + (int64_t)g * ldp,
                                                          mst_w + (int64_t)g * ldp, eta_g + g, ne);
    c.launches++;
  }
  fuse_finalize_kernel<<<h->sm_count * 8, 256, 0, c.st>>>(Sbuf, ldD, N, eta_g, perm_dev, G, eps_floor);
  c.launches++;
  for (int g = 0; g < G; ++g) {
    gather_edges_kernel<<<(ne + 255) / 256, 256, 0, c.st>>>(Sbuf, ldD, mst_i + (int64_t)g * ldp, mst_j + (int64_t)g * ldp,
                                                          edge_val + (int64_t)g * ldp, ne);
    c.launches++;
  }
  CU_TRY(cudaGetLastError());
  // final MST on the sparse edge union (Boruvka), rooted at vertex 0, then the usual tree measurement
  {
    BoruvkaParams bp;
    bp.N = N; bp.G = G; bp.ldp = ldp; bp.mi = mst_i; bp.mj = mst_j; bp.val = edge_val; bp.eps_floor = eps_floor;
    bp.comp = gws.anc0; bp.best = reinterpret_cast<unsigned long long*>(gws.wd); bp.bestp = reinterpret_cast<unsigned long long*>(gws.wd2);
    bp.out_i = gws.dep0; bp.out_j = gws.dep1; bp.out_w = gws.S; bp.out_count = gws.size;
    const char* fenv = getenv("SEQJ_FUSE_MST");
    if (N >= 2048 && !(fenv && !strcmp(fenv, "single")))
      boruvka_sparse_cluster_kernel<<<kBoruvkaCluster, kBoruvkaThreads, 0, c.st>>>(bp);     // one cluster of 8 CTAs
    else
      boruvka_sparse_kernel<<<1, kBoruvkaThreads, 0, c.st>>>(bp);
    RootTreeParams rp;
    rp.N = N; rp.ei = gws.dep0; rp.ej = gws.dep1; rp.ew = gws.S;
    rp.scratch = gws.fuse_i0;
    rp.parent = gws.parent; rp.weight = gws.weight; rp.order = gws.order;
    root_tree_kernel<<<1, 1024, 0, c.st>>>(rp);
    c.launches += 2;
    CU_TRY(cudaGetLastError());
    TreeParams tp;
    tp.N = N; tp.ldv = gws.ldv; tp.parent = gws.parent; tp.weight = gws.weight; tp.order = gws.order;
    tp.size = gws.size; tp.wd = gws.wd; tp.wd2 = gws.wd2; tp.S = gws.S; tp.S2 = gws.S2; tp.use_smem = gws.tree_use_smem;
    tp.anc0 = gws.anc0; tp.anc1 = gws.anc1; tp.dep0 = gws.dep0; tp.dep1 = gws.dep1;
    tp.eta = eta_f; tp.root = root_f; tp.newparent = parents_f; tp.ldp = ldp;
    tp.mst_i = fmst_i; tp.mst_j = fmst_j; tp.mst_w = fmst_w; tp.dfs_order = order_f;
    tree_kernel<<<1, kTreeThreads, gws.tree_smem, c.st>>>(tp);
    c.launches++;
    CU_TRY(cudaGetLastError());
  }
  c.timer.end(t);
  return SEQJ_OK;
}

void stash_triplets(seqj_result* r, int G, int64_t ldp, std::vector<int>& mi, std::vector<int>& mj, std::vector<double>& val) {
  r->raw_i.swap(mi); r->raw_j.swap(mj); r->raw_v.swap(val);
  r->raw_G = G; r->raw_ldp = ldp; r->triplets_ready = false;
}

void dedup_triplets_impl(seqj_result* r, int G, int N, int64_t ldp, const std::vector<int>& mi, const std::vector<int>& mj,
                         const std::vector<double>& val);

// D triplets are rarely read (the wrapper materialises D(r) on demand): deduplicate on first access
void ensure_triplets(const seqj_result* cr) {
  if (cr->triplets_ready) return;
  seqj_result* r = const_cast<seqj_result*>(cr);
  if (r->raw_G > 0) dedup_triplets_impl(r, r->raw_G, (int)r->N, r->raw_ldp, r->raw_i, r->raw_j, r->raw_v);
  r->raw_i.clear(); r->raw_i.shrink_to_fit(); r->raw_j.clear(); r->raw_j.shrink_to_fit(); r->raw_v.clear(); r->raw_v.shrink_to_fit();
  r->triplets_ready = true;
}

void dedup_triplets_impl(seqj_result* r, int G, int N, int64_t ldp, const std::vector<int>& mi, const std::vector<int>& mj,
                         const std::vector<double>& val) {
  // first occurrence of every (i, j) in group-major edge order, via an open-addressing table
  const size_t total = (size_t)G * (N - 1);
  size_t cap = 16;
  while (cap < 2 * total) cap <<= 1;
  std::vector<uint64_t> table(cap, ~uint64_t(0));
  r->Di.reserve(total); r->Dj.reserve(total); r->Dv.reserve(total);
  for (int g = 0; g < G; ++g)
    for (int e = 0; e < N - 1; ++e) {
      const int64_t k = (int64_t)g * ldp + e;
      const uint64_t key = (uint64_t)mi[k] * (uint64_t)N + (uint64_t)mj[k];
      size_t hpos = (key * 0x9E3779B97F4A7C15ull) & (cap - 1);
      bool seen = false;
      while (table[hpos] != ~uint64_t(0)) {
        if (table[hpos] == key) { seen = true; break; }
        hpos = (hpos + 1) & (cap - 1);
      }
      if (seen) continue;
      table[hpos] = key;
      r->Di.push_back(mi[k]); r->Dj.push_back(mj[k]); r->Dv.push_back(val[k]);
    }
}

template <typename T>
void copy_out(T* dst, const std::vector<T>& v) {
  if (dst && !v.empty()) std::memcpy(dst, v.data(), v.size() * sizeof(T));
}

}  // namespace
