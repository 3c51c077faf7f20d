// Unit-level entry points of the C ABI (one reference function each; used by the parity tests and by callers that
// want a single stage).  Included inside the extern "C" block of seqj_api.cu.
// ---- unit-level entry points ------------------------------------------------------------------
int seqj_splitnorm(seqj_handle* h, const double* A, int64_t M64, int64_t N64, int32_t scale, double* P_out,
                   double* U_out, int32_t* nchunks_out) {
  if (!h) return SEQJ_E_BADARG;
  std::lock_guard<std::recursive_mutex> lock__(h->mu);      // one call at a time per handle
  ST_TRY(check_ready(h));
  if (!A || M64 < 1 || N64 < 1 || scale < 1) return fail(h, SEQJ_E_BADARG, "bad argument");
  if (scale > M64)
    return fail(h, SEQJ_E_BADARG, "Scale (" + std::to_string(scale) + ") cannot be larger than the number of data elements (" +
                                      std::to_string(M64) + ").");
  const int M = (int)M64, N = (int)N64;
  Ctx c(h);
  ScaleLayout L = make_layout(M, scale);
  PrepBufs pb;
  ST_TRY(alloc_prep(c, pb, N, L.ldX, L.nchunks, true, false, false, true));
  double *A_dev, *out_dev;
  CU_TRY(c.pool.alloc(&A_dev, (size_t)M * N)); CU_TRY(c.pool.alloc(&out_dev, (size_t)M * N));
  CU_TRY(cudaMemcpyAsync(A_dev, A, (size_t)M * N * 8, cudaMemcpyHostToDevice, c.st));
  ST_TRY(run_prep(c, A_dev, M, M, N, L, pb, 1));
  const int64_t tot = (int64_t)M * N;
  if (P_out) {
    unpad_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, c.st>>>(pb.P, L.ldX, L.cl, L.cl_pad, M, N, out_dev);
    CU_TRY(cudaMemcpyAsync(P_out, out_dev, tot * 8, cudaMemcpyDeviceToHost, c.st));
  }
  if (U_out) {
    unpad_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, c.st>>>(pb.U, L.ldX, L.cl, L.cl_pad, M, N, out_dev);
    CU_TRY(cudaMemcpyAsync(U_out, out_dev, tot * 8, cudaMemcpyDeviceToHost, c.st));
  }
  CU_TRY(cudaStreamSynchronize(c.st));
  if (nchunks_out) *nchunks_out = L.nchunks;
  return SEQJ_OK;
}

int seqj_pairwise(seqj_handle* h, int32_t metric, const double* chunk, int64_t m64, int64_t N64, int normalize,
                  int kl_full, double eps_floor, double l2_thresh, double* D_out) {
  if (!h) return SEQJ_E_BADARG;
  std::lock_guard<std::recursive_mutex> lock__(h->mu);      // one call at a time per handle
  ST_TRY(check_ready(h));
  if (!chunk || !D_out || m64 < 1 || N64 < 1 || metric < 0 || metric > SEQJ_METRIC_MAX) return fail(h, SEQJ_E_BADARG, "bad argument");
  const bool grid_metric = metric == SEQJ_EMD_GRID || metric == SEQJ_ENERGY_GRID;
  if (grid_metric && (int64_t)h->grid.size() != m64)
    return fail(h, SEQJ_E_BADARG, "Grid length must match dims 1 (row count) of the chunk");
  if (metric == SEQJ_PERIODIC && (int64_t)h->periods.size() != m64)
    return fail(h, SEQJ_E_BADARG, "DimensionMismatch: PeriodicEuclidean needs one period per row of the chunk");
  const int m = (int)m64, N = (int)N64;
  Ctx c(h);
  ScaleLayout L = make_layout(m, 1);
  PrepBufs pb;
  ST_TRY(alloc_prep(c, pb, N, L.ldX, 1, metric == SEQJ_L2 || metric == SEQJ_KLD || metric == SEQJ_PERIODIC || metric >= SEQJ_CITYBLOCK, metric == SEQJ_KLD,
                    metric == SEQJ_EMD || metric == SEQJ_ENERGY, false, metric == SEQJ_EMD_GRID, metric == SEQJ_ENERGY_GRID));
  if (metric == SEQJ_PERIODIC) ST_TRY(upload_periods(c, pb, h->periods, L.cl_pad));
  if (grid_metric) {
    CU_TRY(c.pool.alloc(&pb.grid, (size_t)m));
    CU_TRY(cudaMemcpyAsync(pb.grid, h->grid.data(), sizeof(double) * m, cudaMemcpyHostToDevice, c.st));
  }
  FixBufs fb;
  ST_TRY(alloc_fix(c, fb, 1u << 20));
  const bool full = (metric == SEQJ_KLD) && kl_full;
  std::vector<int2> tiles = make_tiles(N, !full);
  int2* tiles_dev;
  CU_TRY(c.pool.alloc(&tiles_dev, tiles.size()));
  CU_TRY(cudaMemcpyAsync(tiles_dev, tiles.data(), tiles.size() * sizeof(int2), cudaMemcpyHostToDevice, c.st));
  const int64_t ldD = round_up(N, 16);
  double *A_dev, *D_dev;
  CU_TRY(c.pool.alloc(&A_dev, (size_t)m * N)); CU_TRY(c.pool.alloc(&D_dev, (size_t)N * ldD));
  CU_TRY(cudaMemcpyAsync(A_dev, chunk, (size_t)m * N * 8, cudaMemcpyHostToDevice, c.st));
  ST_TRY(run_prep(c, A_dev, m, m, N, L, pb, normalize));
  ST_TRY(run_distance(c, metric, L, pb, tiles_dev, (int)tiles.size(), N, D_dev, ldD, 0, 0, 1, eps_floor, l2_thresh,
                      full ? 1 : 0, fb));
  CU_TRY(cudaMemcpy2DAsync(D_out, (size_t)N * 8, D_dev, (size_t)ldD * 8, (size_t)N * 8, N, cudaMemcpyDeviceToHost, c.st));
  CU_TRY(cudaStreamSynchronize(c.st));
  return SEQJ_OK;
}

int seqj_measure_dm(seqj_handle* h, const double* D, int64_t N64, double eps_floor, double* eta, int32_t* root,
                    int32_t* parents, int32_t* mst_i, int32_t* mst_j, double* mst_w, int32_t* order) {
  if (!h) return SEQJ_E_BADARG;
  std::lock_guard<std::recursive_mutex> lock__(h->mu);      // one call at a time per handle
  ST_TRY(check_ready(h));
  if (!D || N64 < 2) return fail(h, SEQJ_E_BADARG, "bad argument");
  const int N = (int)N64;
  Ctx c(h);
  const int64_t ldD = round_up(N, 16), ldp = N;
  double *up, *Dsym, *eta_d, *w_d;
  int *root_d, *par_d, *mi_d, *mj_d, *ord_d;
  CU_TRY(c.pool.alloc(&up, (size_t)N * N)); CU_TRY(c.pool.alloc(&Dsym, (size_t)N * ldD));
  CU_TRY(c.pool.alloc(&eta_d, 1)); CU_TRY(c.pool.alloc(&root_d, 1)); CU_TRY(c.pool.alloc(&par_d, (size_t)ldp));
  CU_TRY(c.pool.alloc(&mi_d, (size_t)ldp)); CU_TRY(c.pool.alloc(&mj_d, (size_t)ldp)); CU_TRY(c.pool.alloc(&w_d, (size_t)ldp));
  CU_TRY(c.pool.alloc(&ord_d, (size_t)ldp));
  CU_TRY(cudaMemcpyAsync(up, D, (size_t)N * N * 8, cudaMemcpyHostToDevice, c.st));
  const int64_t tot = (int64_t)N * N;
  symmetrize_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, c.st>>>(up, N, Dsym, ldD, eps_floor);
  GraphWS gws;
  ST_TRY(graphws_alloc(c, gws, 1, N));
  ST_TRY(run_measure(c, gws, Dsym, ldD, 0, 1, eps_floor, eta_d, root_d, par_d, ldp, mi_d, mj_d, w_d, order ? ord_d : nullptr));
  if (eta) CU_TRY(cudaMemcpyAsync(eta, eta_d, 8, cudaMemcpyDeviceToHost, c.st));
  if (root) CU_TRY(cudaMemcpyAsync(root, root_d, 4, cudaMemcpyDeviceToHost, c.st));
  if (parents) CU_TRY(cudaMemcpyAsync(parents, par_d, 4 * (size_t)N, cudaMemcpyDeviceToHost, c.st));
  if (mst_i) CU_TRY(cudaMemcpyAsync(mst_i, mi_d, 4 * (size_t)(N - 1), cudaMemcpyDeviceToHost, c.st));
  if (mst_j) CU_TRY(cudaMemcpyAsync(mst_j, mj_d, 4 * (size_t)(N - 1), cudaMemcpyDeviceToHost, c.st));
  if (mst_w) CU_TRY(cudaMemcpyAsync(mst_w, w_d, 8 * (size_t)(N - 1), cudaMemcpyDeviceToHost, c.st));
  if (order) CU_TRY(cudaMemcpyAsync(order, ord_d, 4 * (size_t)N, cudaMemcpyDeviceToHost, c.st));
  CU_TRY(cudaStreamSynchronize(c.st));
  return SEQJ_OK;
}

int seqj_accumulate(seqj_handle* h, const double* Dchunks, const double* etas, int n, int64_t N64, double* Dkl_out) {
  if (!h) return SEQJ_E_BADARG;
  std::lock_guard<std::recursive_mutex> lock__(h->mu);      // one call at a time per handle
  ST_TRY(check_ready(h));
  if (!Dchunks || !etas || !Dkl_out || n < 1 || N64 < 1) return fail(h, SEQJ_E_BADARG, "bad argument");
  const int N = (int)N64;
  Ctx c(h);
  const size_t mat = (size_t)N * N, matp = (mat + 1) & ~size_t(1);     // 16 B aligned matrices for double2 access
  double *D_dev, *S_dev, *eta_dev;
  const int B = std::min(n, kCombineMax);
  CU_TRY(c.pool.alloc(&D_dev, matp * B)); CU_TRY(c.pool.alloc(&S_dev, matp)); CU_TRY(c.pool.alloc(&eta_dev, (size_t)n));
  CU_TRY(cudaMemsetAsync(D_dev, 0, matp * B * 8, c.st));
  CU_TRY(cudaMemcpyAsync(eta_dev, etas, sizeof(double) * n, cudaMemcpyHostToDevice, c.st));
  for (int c0 = 0; c0 < n; c0 += B) {
    const int nb = std::min(B, n - c0);
    for (int q = 0; q < nb; ++q)
      CU_TRY(cudaMemcpyAsync(D_dev + (size_t)q * matp, Dchunks + (size_t)(c0 + q) * mat, mat * 8, cudaMemcpyHostToDevice, c.st));
    combine_kernel<<<h->sm_count * 16, 256, 0, c.st>>>(S_dev, D_dev, (int64_t)matp, nb, eta_dev + c0, c0 == 0, c0 + nb >= n,
                                                      eta_dev, n, (int64_t)(matp / 2));
    CU_TRY(cudaGetLastError());
  }
  CU_TRY(cudaMemcpyAsync(Dkl_out, S_dev, mat * 8, cudaMemcpyDeviceToHost, c.st));
  CU_TRY(cudaStreamSynchronize(c.st));
  return SEQJ_OK;
}

int seqj_cdf_distance(seqj_handle* h, int32_t metric, const double* u, const double* uw, int64_t nu64, const double* v,
                      const double* vw, int64_t nv64, double* out) {
  if (!h) return SEQJ_E_BADARG;
  std::lock_guard<std::recursive_mutex> lock__(h->mu);      // one call at a time per handle
  ST_TRY(check_ready(h));
  if (!u || !uw || !v || !vw || !out || nu64 < 1 || nv64 < 1 || nu64 + nv64 >= (1ll << 26))
    return fail(h, SEQJ_E_BADARG, "bad argument");
  if (metric != SEQJ_EMD && metric != SEQJ_ENERGY) return fail(h, SEQJ_E_BADARG, "metric must be SEQJ_EMD or SEQJ_ENERGY");
  const int nu = (int)nu64, nv = (int)nv64;
  auto pow2 = [](int n) { int q = 1; while (q < n) q <<= 1; return q; };
  Ctx c(h);
  CdfPairParams pp;
  pp.nu = nu; pp.nv = nv; pp.p = (metric == SEQJ_EMD) ? 1 : 2;
  pp.pu = pow2(nu); pp.pv = pow2(nv); pp.pa = pow2(nu + nv + 1);
  double *u_d, *uw_d, *v_d, *vw_d, *out_d;
  int* status_d;
  CU_TRY(c.pool.alloc(&u_d, (size_t)nu)); CU_TRY(c.pool.alloc(&uw_d, (size_t)nu));
  CU_TRY(c.pool.alloc(&v_d, (size_t)nv)); CU_TRY(c.pool.alloc(&vw_d, (size_t)nv));
  CU_TRY(c.pool.alloc(&pp.su, (size_t)pp.pu)); CU_TRY(c.pool.alloc(&pp.sv, (size_t)pp.pv)); CU_TRY(c.pool.alloc(&pp.A, (size_t)pp.pa));
  CU_TRY(c.pool.alloc(&pp.iu, (size_t)pp.pu)); CU_TRY(c.pool.alloc(&pp.iv, (size_t)pp.pv));
  CU_TRY(c.pool.alloc(&pp.cu, (size_t)nu)); CU_TRY(c.pool.alloc(&pp.cv, (size_t)nv));
  CU_TRY(c.pool.alloc(&pp.dx, (size_t)nu + nv)); CU_TRY(c.pool.alloc(&pp.pos, (size_t)nu + nv + 1));
  CU_TRY(c.pool.alloc(&out_d, 1)); CU_TRY(c.pool.alloc(&status_d, 1));
  CU_TRY(cudaMemcpyAsync(u_d, u, 8 * (size_t)nu, cudaMemcpyHostToDevice, c.st));
  CU_TRY(cudaMemcpyAsync(uw_d, uw, 8 * (size_t)nu, cudaMemcpyHostToDevice, c.st));
  CU_TRY(cudaMemcpyAsync(v_d, v, 8 * (size_t)nv, cudaMemcpyHostToDevice, c.st));
  CU_TRY(cudaMemcpyAsync(vw_d, vw, 8 * (size_t)nv, cudaMemcpyHostToDevice, c.st));
  pp.u = u_d; pp.uw = uw_d; pp.v = v_d; pp.vw = vw_d; pp.out = out_d; pp.status = status_d;
  cdf_pair_kernel<<<1, kPairThreads, 0, c.st>>>(pp);
  CU_TRY(cudaGetLastError());
  double res = 0.0;
  int status = 0;
  CU_TRY(cudaMemcpyAsync(&res, out_d, 8, cudaMemcpyDeviceToHost, c.st));
  CU_TRY(cudaMemcpyAsync(&status, status_d, 4, cudaMemcpyDeviceToHost, c.st));
  CU_TRY(cudaStreamSynchronize(c.st));
  if (status == 1) return fail(h, SEQJ_E_DOMAIN, "grid values and weights must be finite");
  if (status == 2) return fail(h, SEQJ_E_DOMAIN, "identical grids: the non-zero grid steps do not match the grid length (the reference fails with a DimensionMismatch)");
  *out = (metric == SEQJ_ENERGY) ? 1.4142135623730951 * res : res;
  return SEQJ_OK;
}
