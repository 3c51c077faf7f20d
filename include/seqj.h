/*
 * seqj.h -- C ABI of the B200-native Sequencer hot path (libseqj_b200.so).
 *
 * The reference (turingtest37/SequencerJ.jl) is pure Julia and has no FFI of its own; this ABI
 * is what a Julia `ccall` (see INTEGRATION.md and sequencerj.jl_b200/julia/SequencerJB200.jl) or
 * a Python ctypes binding (sequencerj.jl_b200/_lib.py) binds to replace the reference's hot
 * path.  Each entry point cites the reference code it replaces (paths relative to the
 * reference repository root).
 *
 * Conventions
 *   - every function returns an int status: 0 = ok, <0 = error class (SEQJ_E_*);
 *     seqj_last_error(handle) returns a human-readable message for the last failure.
 *   - the library owns handle/result objects; the caller owns every array buffer.
 *   - matrices are passed exactly as Julia stores them: `A` is M x N column-major, i.e. object
 *     (column) j is the contiguous run A[j*M .. j*M+M-1].  N x N matrices are column-major too
 *     (for symmetric matrices this equals row-major).
 *   - all vertex / column indices are 0-based on the C side; the Julia wrapper adds 1.
 *   - calls are blocking; calls on one handle are serialised by the library (a per-handle lock), different handles
 *     run concurrently; there is no global state.
 *   - there is NO CPU fallback: every compute entry point fails with SEQJ_E_CUDA when no
 *     sm_100 device is usable.
 */
#ifndef SEQJ_H
#define SEQJ_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SEQJ_VERSION 200 /* 0.2.0 */

/* status codes */
#define SEQJ_OK 0
#define SEQJ_E_BADARG (-1)
#define SEQJ_E_CUDA (-2)
/* -3 is unused: the library issues no collective itself; the multi-process exchange (one all-gather of packed group
 * results) belongs to the caller, see seqj_sequence_shard / seqj_fuse_dev, and the single-process path uses peer copies */
#define SEQJ_E_OOM (-4)
#define SEQJ_E_INTERNAL (-5)
#define SEQJ_E_DOMAIN (-6) /* input has negative, NaN or infinite values (AssertionError in the reference) */

/* metric ids, in the order of the reference's ALL_METRICS (src/distancemetrics.jl:234-257) */
#define SEQJ_L2 0     /* L2 = Distances.Euclidean(1e-7)   src/distancemetrics.jl:234 */
#define SEQJ_EMD 1    /* WASS1D = EMD()                   src/distancemetrics.jl:247 */
#define SEQJ_KLD 2    /* KLD = Distances.KLDivergence()   src/distancemetrics.jl:239 */
#define SEQJ_ENERGY 3 /* ENERGY = Energy()                src/distancemetrics.jl:254 */
/* the TYPES EMD / Energy passed as metrics: the reference then builds EMD((lgrid, lgrid)) / Energy((lgrid, lgrid))
 * from the chunk-local slice of `grid` (src/sequencer.jl:217-219), i.e. delta-weighted CDF distances */
#define SEQJ_EMD_GRID 4
#define SEQJ_ENERGY_GRID 5
#define SEQJ_PERIODIC 6 /* Distances.PeriodicEuclidean(periods), set with seqj_set_periods (test/runtests.jl:189-236) */
/* further element-wise Distances.jl (semi)metrics a caller can pass in `metrics` (any Distances.PreMetric is accepted
 * by the reference's `pairwise` call, src/sequencer.jl:226): evaluated on the chunk PDFs
 * by the direct tile kernel, no Gram form */
#define SEQJ_CITYBLOCK 7      /* Distances.Cityblock():      sum |a - b|       */
#define SEQJ_TOTALVARIATION 8 /* Distances.TotalVariation(): sum |a - b| / 2   */
#define SEQJ_CHEBYSHEV 9      /* Distances.Chebyshev():      max |a - b|       */
#define SEQJ_SQEUCLIDEAN 10   /* Distances.SqEuclidean():    sum (a - b)^2     */
#define SEQJ_METRIC_MAX 10

/* phase timers returned by seqj_result_timings (milliseconds, CUDA events on the library's streams) */
#define SEQJ_T_TOTAL 0    /* whole seqj_sequence call, device side (H2D .. results ready) */
#define SEQJ_T_H2D 1      /* host -> device copy of A */
#define SEQJ_T_PREP 2     /* normalise + prefix-scan kernels */
#define SEQJ_T_DIST_L2 3  /* distance kernels per metric */
#define SEQJ_T_DIST_EMD 4
#define SEQJ_T_DIST_KLD 5
#define SEQJ_T_DIST_ENERGY 6
#define SEQJ_T_PRIM 7     /* MST stage: batched dense Prim / kNN-filtered Boruvka */
#define SEQJ_T_TREE 8     /* root / elongation / BFS-tree kernels */
#define SEQJ_T_COMBINE 9  /* eta-weighted accumulation */
#define SEQJ_T_FUSE 10    /* proximity fuse + final MST / order */
#define SEQJ_T_D2H 11     /* result copy-out */
#define SEQJ_T_DERIVE 12  /* coarse-scale chunk matrices derived from the finest scale (no DMMA) */
#define SEQJ_NTIMERS 13

typedef struct seqj_handle seqj_handle;
typedef struct seqj_result seqj_result;

/* ---- lifecycle ----------------------------------------------------------------------- */

int seqj_version(void);

/* Create a handle driving `ndev` CUDA devices (devices == NULL or ndev <= 0: current device).
 * With ndev > 1, seqj_sequence (host input, no group mask) shards the (metric, scale) groups over the devices
 * from this ONE process (one host thread per device; devices[0] runs the final fuse).  The one-process-per-GPU
 * layout (torch.distributed above this ABI) uses single-device handles with group masks or chunk ranges. */
int seqj_create(seqj_handle** out, const int* devices, int ndev);
void seqj_destroy(seqj_handle* h);
const char* seqj_last_error(const seqj_handle* h);

/* The handle keeps its device workspace between calls (repeated calls of one shape reuse it);
 * this frees it without destroying the handle. */
int seqj_release_workspace(seqj_handle* h);

/* Grid used by SEQJ_EMD_GRID / SEQJ_ENERGY_GRID: M finite, positive, strictly increasing values (the `grid`
 * keyword of `sequence`, src/sequencer.jl:105-109,296-304).  NULL clears it. */
int seqj_set_grid(seqj_handle* h, const double* grid, int64_t M);

/* Periods of SEQJ_PERIODIC = `Distances.PeriodicEuclidean(periods)`: d(a,b) = sqrt(sum_k s_k^2) with
 * s = |a_k - b_k| mod p_k, s_k = min(s, p_k - s).  M positive finite values, one per row of a chunk; as in
 * Distances.jl the metric only applies to vectors of exactly that length, i.e. to scales whose single chunk is
 * the whole column (the reference's tests use scales = (1,), test/runtests.jl:189-236); other scales are
 * rejected with SEQJ_E_BADARG where Julia throws a DimensionMismatch.  periods == NULL clears them. */
int seqj_set_periods(seqj_handle* h, const double* periods, int64_t M);

/* Diagnostic (pure host code, needs no GPU): the static schedule seqj_sequence uses for an M-row input with these
 * metrics / scales / group mask when `mats` N x N matrices fit the device -- which chunks of the loop
 * src/sequencer.jl:194-258 are resident together (waves), which are computed by the tile kernels and which derived from
 * the next finer scale, and where the accumulators live -- as text lines ("plan ...", "wave ...", "seg ...") in buf. */
int seqj_plan_describe(int64_t M, const int32_t* metric_ids, int nmetrics, const int32_t* scales, int nscales,
                       const uint8_t* group_mask, int64_t mats, int hier, char* buf, int64_t buflen);

/* Group-major sharding plan (pure host code): owner_out[k * nscales + l] = rank in [0, world) that computes the
 * (metric, scale) group of the loop src/sequencer.jl:194-196.  EMD-type groups are units of their own; a contraction
 * metric over nested scales stays on one rank (its coarse scales are derived from the finest there) unless cutting
 * the chain balances the ranks better.  load_out (may be NULL): modelled cost per rank, in picoseconds per N^2 matrix
 * entry.  The one-process-per-GPU driver (sharding.py) and the single-process multi-device handle use the same plan. */
int seqj_plan_shards(int64_t M, const int32_t* metric_ids, int nmetrics, const int32_t* scales, int nscales, int world,
                     int32_t* owner_out, double* load_out);

/* Limit the device memory the handle may use for chunk matrices (bytes; 0 = auto from free memory). */
int seqj_set_workspace_limit(seqj_handle* h, uint64_t bytes);

/* Diagnostic probes (parity tests at sizes where a whole N x N matrix is too large to copy out): during the following
 * seqj_sequence / seqj_sequence_dev calls on this single-device handle, entry (i[q], j[q]) of chunk matrix chunk[q]
 * of group group[q] (metric-major index k*nscales + l) is sampled right after the distance stage produced it --
 * `abs.(pairwise(alg, m; dims=2)) .+ eps`, src/sequencer.jl:226, whichever kernel path (tiles or derivation from the
 * finer scale) computed it; chunk[q] = -1 samples the combined group matrix `Dkl` (src/sequencer.jl:247).  Read the
 * values with seqj_result_probes (NaN = entry not produced by that call).  n = 0 or group == NULL clears the probes. */
int seqj_set_probes(seqj_handle* h, const int32_t* group, const int32_t* chunk, const int32_t* i, const int32_t* j,
                    int64_t n);

/* ---- the hot call --------------------------------------------------------------------
 * Replaces `_sequence(A, scales, metrics, grid)`  (src/sequencer.jl:184-276) for the metric
 * instances of ALL_METRICS on the shared unit grid:  _splitnorm (:383-407), the metric x scale x
 * chunk loop with `abs.(pairwise(alg, m; dims=2)) .+ eps` (:194-236), `_measure_dm` (:284-290)
 * per chunk and per (metric, scale), the eta-weighted combination (:231-247), the proximity
 * fuse `_weighted_p_matrix` (:310-328), `inv.(Pw)` (:269), the final `_measure_dm` (:271) and
 * `unroll` (src/utils.jl:15-24).
 *
 * A          host pointer, M x N column-major, read-only.  The input conditioning of `sequence`
 *            (src/sequencer.jl:111,130) happens on the device copy: negative / NaN / infinite
 *            values return SEQJ_E_DOMAIN (the reference's AssertionError), and values below
 *            eps() = 2.22e-16 are clamped on load.  seqj_result_input_range reports min/max so the
 *            wrapper can apply the caller-visible `clamp!` to the host array only when needed.
 * metric_ids caller order is kept (it fixes the order of groups, src/sequencer.jl:253-254).
 * scales     number of chunks per column for each scale (src/sequencer.jl:389).
 * eps_floor  the reference's `const eps = 1e-6` (src/distancemetrics.jl:3).
 * l2_thresh  the reference's `L2THR = 1e-7`     (src/distancemetrics.jl:6).
 * group_mask NULL, or nmetrics*nscales flags (metric-major): only groups with a non-zero flag
 *            are computed and the final fuse is skipped (used to shard groups across processes;
 *            finish with seqj_fuse).
 */
int seqj_sequence(seqj_handle* h, const double* A, int64_t M, int64_t N,
                  const int32_t* metric_ids, int nmetrics, const int32_t* scales, int nscales,
                  double eps_floor, double l2_thresh, const uint8_t* group_mask,
                  seqj_result** out);

/* Float32 input (reference quirk Q2, SURVEY.md section 8: every image test of test/runtests.jl:56-186 passes a Float32
 * matrix).  The reference then clamps and normalises in Float32 (`clamp!`, src/sequencer.jl:130; `S ./ sum(S, dims=1)`,
 * :402-403), evaluates Euclidean / KL in Float32 and the ECDFs on Float32 weights, and only `.+ eps` (:226) promotes to
 * Float64.  Here the PDFs are the same Float32 values -- Float32(a) / Float32(chunk sum), clamp in Float32 -- and
 * everything after them runs in FP64, so results agree with the reference's to Float32 accuracy (~1e-6 relative)
 * instead of carrying its Float32 accumulation noise; half the bytes cross PCIe.  Coarse scales are not derived from
 * fine ones in this mode (the per-scale Float32 rounding breaks the exact nesting).  Single-device handles only. */
int seqj_sequence_f32(seqj_handle* h, const float* A, int64_t M, int64_t N, const int32_t* metric_ids, int nmetrics,
                      const int32_t* scales, int nscales, double eps_floor, double l2_thresh,
                      const uint8_t* group_mask, seqj_result** out);

/* Same, with A already resident on the handle's device (device pointer, same layout). */
int seqj_sequence_dev(seqj_handle* h, const double* A_dev, int64_t M, int64_t N,
                      const int32_t* metric_ids, int nmetrics, const int32_t* scales, int nscales,
                      double eps_floor, double l2_thresh, const uint8_t* group_mask,
                      seqj_result** out);

/* Chunk-major sharding (needed when one (metric, scale) group does not fit one GPU or
 * when there are fewer groups than GPUs): this rank computes only the chunks [chunk_lo[g], chunk_hi[g]) of
 * every group g = k*nscales + l (empty range = group not touched), i.e. the per-chunk part of the loop
 * src/sequencer.jl:210-236, and leaves the eta-weighted partial sums  sum_c clamp(eta_c * D_c)  (:231-235, no
 * division) in the caller's device buffer: the i-th touched group (metric-major) at S_out_dev + i*S_stride,
 * as an N x roundup(N,16) row-major matrix.  The caller all-gathers the chunk etas, reduces the partial sums
 * to one owner per group (NCCL over NVLink) and calls seqj_groups_finish there. */
int seqj_sequence_partial(seqj_handle* h, const double* A_dev, int64_t M, int64_t N,
                          const int32_t* metric_ids, int nmetrics, const int32_t* scales, int nscales,
                          double eps_floor, double l2_thresh, const int32_t* chunk_lo,
                          const int32_t* chunk_hi, double* S_out_dev, int64_t S_stride,
                          seqj_result** out);

/* Owner side of the chunk-major path: `Dkl = Dklms / sum(etas)` (src/sequencer.jl:247) and the group-level
 * `_measure_dm` (:251) for `ngroups` reduced accumulators S_dev + slots[g]*S_stride (slots == NULL: g).
 * eta_sum[g] = sum of ALL chunk etas of the group in chunk order (host array).  Outputs are host arrays:
 * eta[ngroups], root[ngroups], parents[ngroups*N], mst_*[ngroups*(N-1)]. */
int seqj_groups_finish(seqj_handle* h, double* S_dev, int64_t S_stride, int ngroups,
                       const int32_t* slots, int64_t N, const double* eta_sum, double eps_floor,
                       double* eta, int32_t* root, int32_t* parents, int32_t* mst_i,
                       int32_t* mst_j, double* mst_w);

/* Final fuse from per-group MSTs gathered by the caller: `_weighted_p_matrix` + `inv.` + final
 * `_measure_dm` + `unroll` (src/sequencer.jl:266-273).  mst_* hold ngroups*(N-1) entries. */
int seqj_fuse(seqj_handle* h, int64_t N, int ngroups, const double* eta_group,
              const int32_t* mst_i, const int32_t* mst_j, const double* mst_w,
              double eps_floor, seqj_result** out);

/* Group-major sharding without host bounces ("gather of per-group (eta_kl, N-1 MST edges)").
 * seqj_sequence_shard = seqj_sequence_dev with a group mask that ALSO leaves every computed group's (eta_kl, MST
 * edges) -- the inputs of `_weighted_p_matrix`, src/sequencer.jl:310-328 -- in the caller's device buffer: the q-th
 * computed group (metric-major order) at pack_dev + q * seqj_pack_stride(N) doubles, laid out as
 * { eta, w[N-1], (i, j) int32 pairs [N-1], pad }.  The ranks all-gather these buffers (NCCL over NVLink) and every rank
 * calls seqj_fuse_dev on the gathered buffer: group g (metric-major over ALL groups) sits in slot slot_of_group[g]
 * (host array).  Same results as seqj_fuse on the same data. */
int64_t seqj_pack_stride(int64_t N);
int seqj_sequence_shard(seqj_handle* h, const double* A_dev, int64_t M, int64_t N, const int32_t* metric_ids,
                        int nmetrics, const int32_t* scales, int nscales, double eps_floor, double l2_thresh,
                        const uint8_t* group_mask, double* pack_dev, seqj_result** out);
int seqj_fuse_dev(seqj_handle* h, int64_t N, int ngroups, const double* pack_dev, const int32_t* slot_of_group,
                  double eps_floor, seqj_result** out);

/* ---- result copy-out ------------------------------------------------------------------
 * Mirrors SequencerResult (src/sequencer.jl:13-20): EOSeg -> group chunks, EOAlgScale -> group,
 * D / mst / eta / order -> final.  NULL output pointers are skipped. */
int seqj_result_dims(const seqj_result* r, int64_t* N, int* ngroups);
int seqj_result_group_info(const seqj_result* r, int g, int* metric, int* scale, int* nchunks,
                           int* computed);
/* device time attributed to group g in milliseconds -- its own distance / derivation kernels plus its share of the
 * batched MST + combine launches it took part in: the `tt = @elapsed` the reference prints per group
 * (src/sequencer.jl:210,260) */
int seqj_result_group_elapsed(const seqj_result* r, int g, double* ms);
/* chunks of group g held by this result: [0, nchunks) except after seqj_sequence_partial */
int seqj_result_group_range(const seqj_result* r, int g, int* chunk_lo, int* chunk_hi);
/* eta_chunk[nchunks], root_chunk[nchunks], parents_chunk[nchunks*N] (BFS-tree parent of each
 * vertex, root -> itself): EOSeg[(alg, l)] = (etas, bfs trees)   (src/sequencer.jl:257) */
int seqj_result_group_chunks(const seqj_result* r, int g, double* eta_chunk, int32_t* root_chunk,
                             int32_t* parents_chunk);
/* EOAlgScale[(alg, l)] = (eta_kl, BFS_kl) (src/sequencer.jl:258) plus the group's MST edges
 * (i<j, weight), which feed `_weighted_p_matrix` */
int seqj_result_group(const seqj_result* r, int g, double* eta, int32_t* root, int32_t* parents,
                      int32_t* mst_i, int32_t* mst_j, double* mst_w);
/* final: eta (elong), root (stidx), order[N] (src/utils.jl:15-24), parents[N], mst edges[N-1] */
int seqj_result_final(const seqj_result* r, double* eta, int32_t* root, int32_t* order,
                      int32_t* parents, int32_t* mst_i, int32_t* mst_j, double* mst_w);
/* final distance matrix D (src/sequencer.jl:269) as triplets on the edge union (i<j); every
 * other off-diagonal entry is +Inf and the diagonal is eps_floor. */
int seqj_result_D_nnz(const seqj_result* r, int64_t* nnz);
int seqj_result_D_triplets(const seqj_result* r, int32_t* i, int32_t* j, double* val);
int seqj_result_timings(const seqj_result* r, double* ms /* SEQJ_NTIMERS */);
/* min / max of the input as seen by the device-side check (min < 2.22e-16 => values were clamped) */
int seqj_result_input_range(const seqj_result* r, double* minv, double* maxv);
/* values of the handle's probes (seqj_set_probes) as sampled by the call that produced this result; n must equal
 * the number of probes that were set */
int seqj_result_probes(const seqj_result* r, double* vals, int64_t n);
/* number of (metric, scale) groups whose chunk matrices were derived from the finest scale instead of being
 * computed by the tile kernels (they are excluded from the DMMA flop count of the roofline) */
int seqj_result_derived_groups(const seqj_result* r, int* n);
/* number of waves (HBM-fulls of chunk matrices) the call was streamed in: 1 when every chunk matrix was resident at
 * once, more under seqj_set_workspace_limit or when N is large (BASELINE config 4) */
int seqj_result_waves(const seqj_result* r, int* n);
/* number of kernels the library launched for this result */
int seqj_result_launches(const seqj_result* r, int64_t* launches);
void seqj_result_free(seqj_result* r);

/* ---- unit-level entry points (each layer parity-testable in isolation) ---------------- */

/* `_splitnorm(A, grid, scale)` (src/sequencer.jl:383-407) plus the ECDF on the shared unit grid
 * (`ecdf` inside EMD/Energy, src/distancemetrics.jl:86,147).  Outputs are M x N column-major like
 * A: P_out = normalised chunks stacked in row order, U_out = per-chunk CDFs.  nchunks_out = number
 * of chunks (can be < scale).  Any output may be NULL. */
int seqj_splitnorm(seqj_handle* h, const double* A, int64_t M, int64_t N, int32_t scale,
                   double* P_out, double* U_out, int32_t* nchunks_out);

/* `abs.(pairwise(alg, m; dims=2)) .+ eps` (src/sequencer.jl:226) for one chunk `m` (m x N
 * column-major).  normalize != 0 applies the _splitnorm column normalisation first.
 * kl_full != 0 computes both triangles of the (asymmetric) KL matrix as Distances.pairwise does;
 * otherwise the upper triangle is mirrored (what `Symmetric(dm)`, src/sequencer.jl:365, consumes).
 * D_out is N x N. */
int seqj_pairwise(seqj_handle* h, int32_t metric, const double* chunk, int64_t m, int64_t N,
                  int normalize, int kl_full, double eps_floor, double l2_thresh, double* D_out);

/* `_measure_dm(D)` (src/sequencer.jl:284-290): `_mst` (:358-374; clamp to [eps, Inf], upper
 * triangle wins), `_leastcentralpt` (:331), `elongation` (:348-354), `bfs_tree` (:288), and
 * optionally `unroll` (src/utils.jl:15-24).  D is N x N column-major; +Inf marks a non-edge. */
int seqj_measure_dm(seqj_handle* h, const double* D, int64_t N, double eps_floor, double* eta,
                    int32_t* root, int32_t* parents, int32_t* mst_i, int32_t* mst_j,
                    double* mst_w, int32_t* order);

/* `EMD(u, v, uw, vw)` / `Energy(u, v, uw, vw)` (src/distancemetrics.jl:82-87, 143-148): distance between two
 * weighted empirical CDFs on their OWN grids, `_cdf_distance(ecdf(u; weights = uw), ecdf(v; weights = vw), p)`
 * (:191-223) with p = 1 for SEQJ_EMD and sqrt(2) * (p = 2) for SEQJ_ENERGY.  u / v are grid values (any order,
 * any lengths nu / nv), uw / vw their weights.  Identical grids take the reference's shortcut branch (:206-209),
 * whose non-zero grid steps must number nu or one (SEQJ_E_DOMAIN otherwise, where the reference throws).  This is
 * the per-pair form of the reference's known-answer tests (test/algotests.jl:36-107); `EMD(grids)(u, v)` and
 * `emd(u, v)` with vectors of different lengths reduce to it. */
int seqj_cdf_distance(seqj_handle* h, int32_t metric, const double* u, const double* uw, int64_t nu,
                      const double* v, const double* vw, int64_t nv, double* out);

/* `Dklms .+= clamp(eta .* Dklm, eps(), Inf)` over chunks, then `/ sum(etas)`
 * (src/sequencer.jl:231-235,247).  Dchunks = n stacked N x N matrices. */
int seqj_accumulate(seqj_handle* h, const double* Dchunks, const double* etas, int n, int64_t N,
                    double* Dkl_out);

#ifdef __cplusplus
}
#endif
#endif /* SEQJ_H */
