/*
 * acav_b200.h -- C ABI of libacav_b200.so: the sm_100a CUDA implementation of ACAV100M's two
 * GPU-bound curation operators (mini-batch SGD k-means, exact greedy mutual-information selection).
 *
 * The reference (sangho-vision/acav100m) is pure Python: it has no FFI layer, its "plugin points"
 * are the Python classes `KMeans` (clustering/code/sgd_clustering.py:10-129) and the measures
 * returned by `get_measure` (subset_selection/code/measures/__init__.py:5-14).  Every entry point
 * below names the reference lines it replaces; INTEGRATION.md shows the ctypes binding a reference
 * maintainer would add.
 *
 * Conventions
 *   - plain pointers and sizes only; all data pointers are DEVICE pointers unless marked "host";
 *   - the caller owns every buffer; the library allocates only inside *_create handles
 *     (freed by *_destroy);
 *   - `stream` is a cudaStream_t passed as void*; all work is enqueued on it, nothing synchronises
 *     unless stated;
 *   - return value: 0 = success, > 0 = cudaError_t, < 0 = ACAV_E_* below; functions never throw;
 *   - one host thread per device at a time per handle.
 */
#ifndef ACAV_B200_H
#define ACAV_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ACAV_B200_ABI_VERSION 1

#define ACAV_OK              0
#define ACAV_E_INVALID      (-1)   /* bad argument (null pointer, negative size, shape mismatch)   */
#define ACAV_E_UNSUPPORTED  (-2)   /* shape/alignment outside what the kernel was built for         */
#define ACAV_E_STATE        (-3)   /* call order violated (e.g. run before load)                    */
#define ACAV_E_NO_DEVICE    (-4)   /* no sm_100 device / driver entry point missing                 */

int         acav_abi_version(void);
const char *acav_status_string(int status);                 /* static string, never NULL            */
int         acav_device_info(int *sm_count, int *cc_major, int *cc_minor);

/* ------------------------------------------------------------------------------------------------
 * k-means (clustering/code/sgd_clustering.py)
 * ---------------------------------------------------------------------------------------------- */

typedef struct acav_kmeans acav_kmeans_t;                    /* per-(k,d) workspace, no model state  */

/* Workspace for batches of up to `max_batch` rows against `k` centroids of dimension `d`.
 * Model state (centers[k,d], counts[k]) stays in caller-owned tensors, exactly like the attributes
 * of the reference object (sgd_clustering.py:24-26). */
int acav_kmeans_create(acav_kmeans_t **out, int32_t k, int32_t d, int64_t max_batch);
int acav_kmeans_destroy(acav_kmeans_t *h);
int64_t acav_kmeans_workspace_bytes(const acav_kmeans_t *h);

/* Tile shape of the tcgen05 distance GEMM behind ACAV_ASSIGN_TENSOR (the reference's `-2*matmul`,
 * sgd_clustering.py:72).  AUTO picks by K; the others pin one kernel (benchmarking / debugging).
 * Results are identical for every choice (the exact re-check decides near-ties). */
#define ACAV_TILE_AUTO      0
#define ACAV_TILE_SINGLE    1      /* one CTA per 128 x 256 tile (cta_group::1)                     */
#define ACAV_TILE_PAIR_256  2      /* CTA pair, 256 x 256 tile, two TMEM accumulator stages         */
#define ACAV_TILE_PAIR_512  3      /* CTA pair, 256 x 512 tile, X read once per 512 centroids       */
int acav_kmeans_set_tile_variant(acav_kmeans_t *h, int32_t variant);

/* Assignment modes */
#define ACAV_ASSIGN_EXACT   0      /* fp32 inputs, fp64-accumulated dot products on CUDA cores     */
#define ACAV_ASSIGN_TENSOR  1      /* tcgen05 bf16 distance GEMM + top-4 screening + exact refine  */

/* Replaces the distance branch of KMeans.calc_best (sgd_clustering.py:70-79):
 *   dist[i,j] = (-2<c_i,x_j> + |x_j|^2) + |c_i|^2 ; rows i with counts[i] < underused_threshold
 *   divided by reinit_r (:76-77) ; best[j] = first argmin_i ; *mean_dist = mean_j min_i dist.
 * x: [b, d] fp32 row-major with row stride ldx (elements).  best: int64[b] (torch.long, :78).
 * min_dist: fp32[b] or NULL.  mean_dist: fp32[1] on the device or NULL (no host sync here; the
 * reference's .item() at :79 is the caller's choice).  n_refined: int32[2] device or NULL -- in
 * ACAV_ASSIGN_TENSOR mode the number of rows re-checked on <= 16 candidates and the number sent
 * through the full exact kernel (both 0 in ACAV_ASSIGN_EXACT mode). */
int acav_kmeans_assign(acav_kmeans_t *h, const float *x, int64_t b, int64_t ldx,
                       const float *centers, const float *counts,
                       float underused_threshold, float reinit_r,
                       int64_t *best, float *min_dist, float *mean_dist, int32_t *n_refined,
                       int32_t mode, void *stream);

/* The three stages of acav_kmeans_assign(ACAV_ASSIGN_TENSOR), exposed so that a caller walking a
 * resident data set can (a) prepare the centroids once per pass and (b) run the memory-bound
 * preparation of chunk i+1 (fp32 -> bf16 copy + row norms) on a second stream and workspace while
 * the tensor-core kernel of chunk i is busy:
 *   prepare_centers : bf16 centroids, |c|^2, per-centroid epilogue parameters (re-init scaling)
 *   prepare_batch   : bf16 rows + |x|^2 into the workspace
 *   assign_prepared : tcgen05 distance GEMM + top-4 screen + exact re-check (+ exact distances if
 *                     min_dist / mean_dist are non-NULL); x must be the batch given to prepare_batch. */
int acav_kmeans_prepare_centers(acav_kmeans_t *h, const float *centers, const float *counts,
                                float underused_threshold, float reinit_r, void *stream);
int acav_kmeans_prepare_batch(acav_kmeans_t *h, const float *x, int64_t b, int64_t ldx, void *stream);
int acav_kmeans_assign_prepared(acav_kmeans_t *h, const float *x, int64_t b, int64_t ldx,
                                const float *centers, const float *counts,
                                float underused_threshold, float reinit_r,
                                int64_t *best, float *min_dist, float *mean_dist, int32_t *n_refined,
                                void *stream);

/* The same multi-GPU step WITHOUT NCCL (one process per GPU on one NVLink-connected node): every rank owns
 * ceil(k / world) consecutive centroids.  acav_kmeans_update_p2p, called by every rank after acav_kmeans_histogram
 * with its LOCAL counts_b, (1) exchanges the histograms through peer memory and takes the lr decision (:114-119),
 * (2) stores each rank's row-ordered deltas straight into the owning rank's receive buffer from inside the update
 * kernel (peer stores over NVLink), (3) has the owner add the world deltas IN RANK ORDER, apply decay and sum
 * (:121, :127) and store the new centroid rows into every rank's inbox, (4) copies the inbox into centers.
 * The result does not depend on timing and equals the single-process sum order of the rank-ordered reduction, so
 * all ranks hold bit-identical centers.  Nothing here touches the host: the call can be captured in a CUDA graph.
 * Setup as for acav_mi_comm_*: create, export (allocates the arena, returns a cudaIpcMemHandle_t in a HOST buffer of
 * acav_kmeans_comm_handle_bytes() bytes), all-gather the handles in the host language, connect.  d must be a
 * multiple of 4, world <= 16 (ACAV_E_UNSUPPORTED otherwise: use update_local + an all-reduce).  Waits on peers are
 * bounded (20 s, environment ACAV_KM_SPIN_TIMEOUT_MS); acav_kmeans_comm_status synchronises the stream and
 * returns 1 if a peer's data did not arrive in time since create. */
typedef struct acav_kmeans_comm acav_kmeans_comm_t;
int acav_kmeans_comm_create(acav_kmeans_comm_t **out, int32_t k, int32_t d, int32_t world, int32_t rank);
int acav_kmeans_comm_destroy(acav_kmeans_comm_t *h);
int acav_kmeans_comm_handle_bytes(void);
int acav_kmeans_comm_export(acav_kmeans_comm_t *h, void *handle_out);
int acav_kmeans_comm_connect(acav_kmeans_comm_t *h, const void *handles);
int acav_kmeans_comm_status(acav_kmeans_comm_t *h, int32_t *status_host, void *stream);
/* Alternative to connect for peers whose arenas the caller has mapped itself (several ranks driven from ONE process,
 * or memory shared by other means): arenas[r] = device pointer of rank r's arena as seen from this device
 * (acav_kmeans_comm_arena of that rank's handle); the entry of this rank is ignored. */
void *acav_kmeans_comm_arena(acav_kmeans_comm_t *h);
int acav_kmeans_comm_connect_ptrs(acav_kmeans_comm_t *h, void *const *arenas);
int acav_kmeans_update_p2p(acav_kmeans_t *h, acav_kmeans_comm_t *comm, const float *x, int64_t b, int64_t ldx,
                           const float *counts_b_local, double lr, float *centers, float *counts, int32_t *fallback,
                           void *stream);

/* For callers that replay the step from a captured CUDA graph (by-value arguments are frozen at capture):
 * flags[i] = counts[i] < *threshold_dev ? 0 : 1, with the per-step threshold (count/k)^p of sgd_clustering.py:77 in
 * DEVICE memory.  Passing (flags, 0.5f) as (counts, underused_threshold) to the assignment entry points above
 * selects exactly the centroids with counts[i] < threshold. */
int acav_kmeans_underused_flags(const float *counts, int32_t k, const float *threshold_dev, float *flags, void *stream);

/* Replaces the warm-up branch of calc_best (sgd_clustering.py:67-68,78-79): `noise` is the [k, b]
 * fp32 tensor the caller drew with torch.rand; best[j] = first argmin over rows. */
int acav_kmeans_assign_noise(const float *noise, int32_t k, int64_t b,
                             int64_t *best, float *min_dist, float *mean_dist, void *stream);

/* Step 1 of the "fast parallel update" (sgd_clustering.py:113): counts_b[i] = #{j : best[j] = i}
 * as fp32, and -- inside the workspace -- the rows of the batch partitioned by centroid in stable
 * (row) order, which acav_kmeans_update_* consume.  In a multi-GPU run the caller all-reduces
 * counts_b (sgd_clustering.py:114-115) before the next call. */
int acav_kmeans_histogram(acav_kmeans_t *h, const int64_t *best, int64_t b,
                          float *counts_b, void *stream);

/* Steps 2-4 (sgd_clustering.py:116-127), single process:
 *   lr_eff = lr, or 0.5/max(counts_b) when max(counts_b)*lr >= 1 (then ++*fallback)   [:116-119]
 *   counts += counts_b ; centers *= (1 - counts_b*lr_eff)                              [:120-121]
 *   centers += sum_{j: best[j]=i} fl32(x_j*lr_eff), summed in row order j             [:122-127]
 * The row-ordered fp32 sum is what torch-scatter's CPU kernel computes, so the result is
 * bit-identical to the reference's CPU path.  fallback: int32[1] device counter or NULL.
 * When counts_b is the very buffer the preceding acav_kmeans_histogram filled (up to 32768 rows), the
 * update kernels take the lr decision themselves from the maximum that call left in the workspace --
 * the buffer must then be unmodified; counts changed by the caller belong in another buffer (or in
 * acav_kmeans_update_local, which always reduces counts_b again). */
int acav_kmeans_update_fused(acav_kmeans_t *h, const float *x, int64_t b, int64_t ldx,
                             const float *counts_b, double lr,
                             float *centers, float *counts, int32_t *fallback, void *stream);

/* The reference's `sequential=True` branch (sgd_clustering.py:103-109): the rows of the batch are applied one at a
 * time, in batch order, each to its centroid: c = fl(fl(c * fl32(1 - lr)) + fl(fl32(lr) * x)); counts += 1.  No lr
 * fallback, no decay by the histogram.  Rows of different centroids commute, so the result equals the per-centroid
 * row-ordered recurrence the kernels run (same stable partition as the fast update). */
int acav_kmeans_update_sequential(acav_kmeans_t *h, const float *x, int64_t b, int64_t ldx, const float *counts_b,
                                  double lr, float *centers, float *counts, void *stream);

/* Multi-GPU split of the same step: `update_local` applies the decay with the GLOBAL counts_b and
 * writes this rank's deltas[k,d] (to be all-reduced, sgd_clustering.py:125-126);
 * `apply_deltas` is centers += deltas (:127). */
int acav_kmeans_update_local(acav_kmeans_t *h, const float *x, int64_t b, int64_t ldx,
                             const float *counts_b_global, double lr,
                             float *centers, float *counts, float *deltas, int32_t *fallback,
                             void *stream);
int acav_kmeans_apply_deltas(float *centers, const float *deltas, int64_t n, void *stream);

/* ------------------------------------------------------------------------------------------------
 * greedy mutual-information selection (subset_selection/code/measures/mi.py, EfficientMemMI)
 * ---------------------------------------------------------------------------------------------- */

typedef struct acav_mi acav_mi_t;

/* Engine for one clustering pair (P = 1): a K_a x K_v contingency table and `w` candidates.
 * `pos_base` is the global position of this engine's first candidate (multi-GPU: each rank holds a
 * contiguous range of the candidate list, so "first index wins ties" stays global). */
int acav_mi_create(acav_mi_t **out, int64_t w, int32_t k_a, int32_t k_v, int64_t max_picks,
                   int64_t pos_base);
int acav_mi_destroy(acav_mi_t *h);

/* Candidate cells as the reference holds them (EfficientMemMI.calc_N, mi.py:285-295): int64 [w, 2]
 * of (c1, c2), row-major, on the device.  Packed to 2 x uint16 per candidate inside the engine. */
int acav_mi_load_candidates(acav_mi_t *h, const int64_t *cells, void *stream);

/* Tables that make the fp32 arithmetic bit-identical to torch's CPU kernels:
 *   logs[k] = log(float(k)) for k in [0, n_logs) exactly as torch.log returns it (device pointer);
 *   consts (HOST pointer, 6 floats) = { x*log(x) of an empty table cell, of an empty marginal,
 *   n0, NlogN0, aloga0, blogb0 } from init_cache (mi.py:32-39, 297-308).
 * Resets the table to the empty state. */
int acav_mi_set_tables(acav_mi_t *h, const float *logs, int64_t n_logs, const float *consts,
                       void *stream);

/* Adds one sample to the table without selecting it (EfficientMemMI.add_samples, mi.py:408-412). */
int acav_mi_add_sample(acav_mi_t *h, int32_t c1, int32_t c2, void *stream);

/* One greedy iteration split for multi-GPU use (calc_measure, mi.py:108-114):
 *   local_best: scores every remaining local candidate (get_last :322-333, calc_MI :368-381) and
 *               writes key_cell[0] = (orderable(score) << 32) | (0xFFFFFFFF - global_position)
 *               (0 when no candidate remains) and key_cell[1] = (c1 << 16) | c2 of that candidate;
 *   apply:      given the `n` (key, cell) pairs of all ranks (all-gathered by the caller; n = 1 on
 *               one GPU) adopts the pair with the largest key -- highest score, earliest position --
 *               (update_cache :383-389, update_mats :401-406, remove_idx_all :104-106).  Every rank
 *               applies the same pairs; the rank owning the position tombstones the candidate.
 *               out_pos / out_gain: one element each (device) or NULL. */
int acav_mi_local_best(acav_mi_t *h, uint64_t *key_cell, void *stream);
int acav_mi_apply(acav_mi_t *h, const uint64_t *key_cells, int32_t n,
                  int64_t *out_pos, float *out_gain, void *stream);

/* Single-GPU greedy loop (EfficientMI.run_greedy, mi.py:150-192, body of the for loop x n_picks):
 * out_pos[i] = position in the candidate list of the i-th pick, out_gain[i] = its score.
 * mode 0: three kernels per iteration (reference-shaped); mode 1: persistent cooperative kernel
 * streaming every remaining candidate per iteration; mode 2: persistent kernel over a cell index --
 * candidates sorted once by table cell, every iteration scans the K_a x K_v cells (gain of the cell,
 * earliest remaining candidate of the cell) instead of the candidates; the score depends on a candidate
 * only through its cell (mi.py:322-381) and `max` keeps the first of equal scores (:79), so the picks and
 * gains are identical.  All modes continue from the engine's current state and can be mixed. */
#define ACAV_MI_LOOP_KERNELS     0
#define ACAV_MI_LOOP_PERSISTENT  1
#define ACAV_MI_LOOP_CELLS       2
/* mode 3: the persistent candidate stream with ONE byte per candidate -- the list stably partitioned by (c1, c2 / w)
 * sub-rows of w <= 255 columns, streamed through registers (no shared-memory ring), table counts of a CTA's sub-rows
 * cached in shared memory; half the HBM bytes per iteration of mode 1, same picks and gains.  Needs
 * k_a * ceil(k_v / 255) <= 24000 (acav_mi_loop_supported). */
#define ACAV_MI_LOOP_BYTES       3
int acav_mi_run(acav_mi_t *h, int64_t n_picks, int64_t *out_pos, float *out_gain,
                int32_t mode, void *stream);

/* Which loops can run a k_a x k_v table (1 / 0): ACAV_MI_LOOP_PERSISTENT needs at least one gain row of k_v + 1
 * floats next to the replicated marginals in shared memory and 4*k_v < 65536 (k_v <= 8140 at k_a = k_v);
 * ACAV_MI_LOOP_CELLS needs the marginals in shared memory (k <= 16384).  A host mirror picks the loop with this
 * instead of catching ACAV_E_UNSUPPORTED from acav_mi_run. */
int acav_mi_loop_supported(int32_t k_a, int32_t k_v, int32_t mode);

/* Builds the candidate layout of `mode` now (the stable partition / cell index that acav_mi_run would otherwise
 * build on first use) and returns its status, so that a multi-GPU caller can agree on success across ranks BEFORE
 * any rank enters the persistent kernel. */
int acav_mi_prepare(acav_mi_t *h, int32_t mode, void *stream);

/* Why the last persistent launch ended: synchronises `stream`, then *status_host = 0 (all iterations done or no
 * candidate left on any rank) or ACAV_MI_RUN_PEER_TIMEOUT (a peer GPU's per-iteration entry did not arrive within
 * the spin limit -- 20 s, environment ACAV_MI_SPIN_TIMEOUT_MS -- and the loop stopped; out_pos of the iterations
 * not run is -1).  The kernels never spin without bound on another GPU. */
#define ACAV_MI_RUN_OK            0
#define ACAV_MI_RUN_PEER_TIMEOUT  1
int acav_mi_status(acav_mi_t *h, int32_t *status_host, void *stream);

/* Multi-GPU persistent loop (one process per GPU on one node).  Every rank holds a contiguous range of
 * the candidate list (pos_base); inside ACAV_MI_LOOP_PERSISTENT each rank's per-iteration winner is
 * written straight into every peer's mailbox over NVLink (peer stores + release/acquire flags) -- no
 * host round trip, no NCCL launch per iteration.  Setup: every rank calls comm_export (allocates its
 * mailbox, returns a cudaIpcMemHandle_t, `acav_mi_comm_handle_bytes()` bytes, HOST pointer), the
 * host language all-gathers the handles, every rank calls comm_connect with the world x handle
 * array.  All ranks must then call acav_mi_run with the same n_picks. */
int acav_mi_comm_handle_bytes(void);
int acav_mi_comm_export(acav_mi_t *h, int32_t world, int32_t rank, void *handle_out);
int acav_mi_comm_connect(acav_mi_t *h, const void *handles);

/* Profiling aid: when `cycles` (device int64 [72 * #SMs]) is non-NULL the persistent loop records, per
 * CTA and for the last iteration it ran, SM cycles spent in {gain rows, candidate scan, block reduce +
 * publish, grid-barrier wait}, then {stream blocks, table rows} of the CTA's chunk and the cycles of
 * {per-iteration prologue, winner hand-over of the previous iteration} (the first 8 * #SMs words); the byte-stream
 * loop adds, from word 8 * #SMs on, 64 words per CTA: when each of its warps finished its span and when it had settled
 * its best candidate (cycles since the iteration began).  NULL switches it off (default). */
int acav_mi_debug_timers(acav_mi_t *h, int64_t *cycles);

/* Tuning of ACAV_MI_LOOP_BYTES (benchmarking): variant 0..5 = (threads per CTA, 16-byte loads in flight per thread)
 * (1024,3) (512,8) (768,4) (1024,2) (512,4) (1024,4); use_cache is ignored (the table counts of a CTA's sub-rows
 * always live in its shared memory).  Also settable through the environment (ACAV_MI_S8_VARIANT) before
 * acav_mi_create. */
int acav_mi_set_stream_variant(acav_mi_t *h, int32_t variant, int32_t use_cache);

/* Introspection for tests: copies table counts (uint32 [k_a*k_v], [k_v], [k_a]) and the four running
 * sums {NlogN, aloga, blogb, n} to device buffers (any may be NULL). */
int acav_mi_read_state(acav_mi_t *h, uint32_t *n_cells, uint32_t *a_cols, uint32_t *b_rows,
                       float *sums, void *stream);

/* ------------------------------------------------------------------------------------------------
 * greedy MI over P > 1 clustering pairs (EfficientMemMI with the `combination` / `bipartite` / `diagonal`
 * pairings of subset_selection/code/pairing.py:5-41; the reference default is P = 45 pairs of ten clusterings)
 * ---------------------------------------------------------------------------------------------- */

typedef struct acav_mi_pairs acav_mi_pairs_t;

/* Engine for `p` contingency tables of c x c cells over `w` candidates that carry `d` cluster ids each.
 * pairs: HOST int32 [p, 2], the id columns (0..d-1) holding (c1, c2) of every pair (pairing.py, gather_pairs
 * mi.py:310-320).  Limits: d <= 64, p <= 256, c <= 65535 (ACAV_E_UNSUPPORTED beyond).  `pos_base` as in
 * acav_mi_create. */
int acav_mi_pairs_create(acav_mi_pairs_t **out, int64_t w, int32_t d, int32_t c, int32_t p, const int32_t *pairs,
                         int64_t max_picks, int64_t pos_base);
int acav_mi_pairs_destroy(acav_mi_pairs_t *h);

/* Candidate rows as the reference holds them (get_assignments mi.py:41-45): int64 [w, d] row-major on the
 * device, row i = the d cluster ids of candidate i.  Stored as d uint16 columns inside the engine. */
int acav_mi_pairs_load_candidates(acav_mi_pairs_t *h, const int64_t *ids, void *stream);

/* logs as in acav_mi_set_tables; consts: HOST fp32 [p, 6] = per pair { x*log(x) of an empty cell, of an empty
 * marginal, n0, NlogN0, aloga0, blogb0 } as init_cache (mi.py:32-39, 297-308) produces them for THIS p and c
 * (the fp32 sums over the empty tables depend on the tensor shape).  Resets the tables. */
int acav_mi_pairs_set_tables(acav_mi_pairs_t *h, const float *logs, int64_t n_logs, const float *consts,
                             void *stream);

/* add_samples (mi.py:408-412): counts one sample (HOST int64 [d] ids) into every table. */
int acav_mi_pairs_add_sample(acav_mi_pairs_t *h, const int64_t *ids, void *stream);

/* One greedy iteration split for multi-GPU use, as acav_mi_local_best / acav_mi_apply.  A record is
 * acav_mi_pairs_record_words(h) = 1 + ceil(d / 4) uint64: the key ((orderable(mean score) << 32) |
 * (0xFFFFFFFF - global position), 0 = no candidate left) followed by the winner's d ids, 16 bits each.
 * The mean over pairs is added in the order torch's CPU `mean` uses (calc_score mi.py:76-80), so scores are
 * bit-identical to the reference's.  apply: `records` = n consecutive records (all-gathered by the caller;
 * n = 1 on one GPU); every rank applies the record with the largest key. */
int acav_mi_pairs_record_words(const acav_mi_pairs_t *h);
int acav_mi_pairs_local_best(acav_mi_pairs_t *h, uint64_t *record, void *stream);
int acav_mi_pairs_apply(acav_mi_pairs_t *h, const uint64_t *records, int32_t n,
                        int64_t *out_pos, float *out_gain, void *stream);

/* Single-GPU greedy loop (run_greedy mi.py:150-192): n_picks iterations of local_best + apply. */
int acav_mi_pairs_run(acav_mi_pairs_t *h, int64_t n_picks, int64_t *out_pos, float *out_gain, void *stream);

/* Introspection for tests: table counts (uint32 [p*c*c], [p*c], [p*c]) and running sums
 * {NlogN, aloga, blogb, n} per pair (fp32 [p, 4]) to device buffers (any may be NULL). */
int acav_mi_pairs_read_state(acav_mi_pairs_t *h, uint32_t *n_cells, uint32_t *a_cols, uint32_t *b_rows,
                             float *sums, void *stream);

/* ------------------------------------------------------------------------------------------------
 * batch_mi, the reference CLI's default measure (subset_selection/code/measures/batch.py)
 * ---------------------------------------------------------------------------------------------- */

typedef struct acav_mi_dense acav_mi_dense_t;

/* Running contingency tables of P clustering pairs with C centroids each (init_cache, mi.py:32-39). */
int acav_mi_dense_create(acav_mi_dense_t **out, int32_t p, int32_t c, void *stream);
int acav_mi_dense_destroy(acav_mi_dense_t *h);

/* Counts m samples into the tables (add_samples batch.py:190-193, update_cache :152-154).
 * cells: int64 [m, P, 2] of (c1, c2) per pair, device. */
int acav_mi_dense_add(acav_mi_dense_t *h, const int64_t *cells, int64_t m, void *stream);

/* scores[i] = mean over pairs of MI(table_p + one-hot(candidate i)) -- what sample_batch + get_last +
 * calc_MI + mean(dim=-1) compute (batch.py:34-54,143-144; mi.py:85-98) -- for nb candidates given as
 * int64 [nb, P, 2] device cells.  per_pair: fp32 [nb, P] device or NULL. */
int acav_mi_dense_score(acav_mi_dense_t *h, const int64_t *cells, int64_t nb, float *scores, float *per_pair,
                        void *stream);

/* The same scores BIT FOR BIT as the reference's CPU path computes them: every cell of
 *   (N / n * (N.log() + n.log() - (a.log() + b.log()))).sum([2, 3])          (mi.py:90)
 * with its five fp32 roundings, summed over the C*C cells in the order of torch's CPU reduction kernel (ATen
 * cascade_sum: 8 lanes x 4 interleaved accumulators, cascade levels), then averaged over the pairs in the same
 * kernel's order -- so a seeded run selects the very indices the reference selects (torch.topk / max see the same
 * bits).  logs / n_logs as in acav_mi_set_tables (logs[k] for every count the tables can reach + 1); consts: HOST
 * fp32[6] = { eps, a0, b0, log eps, log a0, log b0 }, the empty cell / column marginal / row marginal of init_cache
 * (mi.py:32-39) and torch's logs of them.  Cost O(nb * P * C * C), like the reference; P <= 256. */
int acav_mi_dense_score_exact(acav_mi_dense_t *h, const int64_t *cells, int64_t nb, const float *logs, int64_t n_logs,
                              const float *consts, float *scores, float *per_pair, void *stream);

/* The same for the adjusted-MI measure `ami` (EfficientAMI, mi.py:212-262): scores[i] = mean over pairs of
 * (MI - EMI) / max(generalized_mean(H_a, H_b) - EMI, eps) of (table_p + one-hot(candidate i)), with the reference's
 * single-term EMI (calc_EMI :217-231), evaluated in fp64.  average_method: 0 arithmetic, 1 max, 2 min
 * (generalized_mean :200-209).  Cost per call: O(P*C*C) for the row / column sums of the EMI terms plus O(1) per
 * (candidate, pair) -- the reference's dense evaluation is O(nb*P*C*C). */
#define ACAV_AMI_ARITHMETIC 0
#define ACAV_AMI_MAX        1
#define ACAV_AMI_MIN        2
int acav_mi_dense_score_ami(acav_mi_dense_t *h, const int64_t *cells, int64_t nb, int32_t average_method,
                            float *scores, float *per_pair, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* ACAV_B200_H */
