// Internal launch wrappers shared between the translation units of libacav_b200.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "common.cuh"

namespace acav {

// kmeans_exact.cu
int launch_row_norm2(const float *x, int64_t rows, int32_t d, int64_t ldx, const int32_t *rowlist,
                     float *out, cudaStream_t st);
int launch_assign_exact(const float *x, int64_t ldx, const int32_t *rowlist, int64_t nrows,
                        const int32_t *nrows_dev, const float *centers, int32_t k, int32_t d, const float *xn, const float *cn,
                        const float *counts, float thr, float r, int64_t *best, float *mind,
                        unsigned long long *packed, int32_t sm_count, cudaStream_t st, unsigned int *tickets = nullptr);
int launch_assign_noise(const float *noise, int32_t k, int64_t b, int64_t *best, float *mind,
                        cudaStream_t st);
int launch_mean(const float *v, int64_t n, float *out, cudaStream_t st);

// kmeans_umma.cu
constexpr int kMaxSplit = 8;
int make_bf16_tensor_map(void *out_map, const void *base, int64_t rows, int32_t dp, int32_t box_rows);
int umma_tile_n(int32_t k);
int64_t umma_partial_bytes(int64_t max_batch);
int64_t umma_param_bytes(int32_t k);
int launch_prep_rows(const float *x, int64_t rows, int32_t d, int64_t ldx, int32_t dp, void *xb, float *xn,
                     cudaStream_t st);
int launch_centroid_params(const float *cn, const float *counts, int32_t k, float thr, float r, void *params,
                           cudaStream_t st);
int launch_assign_umma(const void *tmap_x, const void *tmap_c, const float *xn, const void *cparams, int32_t b,
                       int32_t k, int32_t dp, int32_t sm_count, void *partial, int32_t *n_split_out, cudaStream_t st);
// kmeans_umma2.cu (CTA pairs): halves = 1 -> 256 x 256 pair tiles, 2 -> 256 x 512; *n_lists_out partial lists per row
int launch_assign_pair(const void *tmap_x, const void *tmap_c128, const float *xn, const void *cparams, int32_t b,
                       int32_t k, int32_t dp, int32_t sm_count, int32_t halves, void *partial, int32_t *n_lists_out,
                       cudaStream_t st);
int launch_merge_classify(const void *partial, int32_t b, int32_t n_split, const float *xn, const void *cparams,
                          const float *cn, int32_t k, int64_t *best, float *mind, int32_t *cand_rows, int32_t *cand_ids, int32_t *full_rows,
                          int32_t *counters, cudaStream_t st);
int launch_candidate_refine(const float *x, int64_t ldx, int32_t d, const float *centers, const float *xn,
                            const float *cn, const float *counts, float thr, float r, const int32_t *cand_rows,
                            const int32_t *cand_ids, const int32_t *counters, int32_t b, int64_t *best, float *mind,
                            cudaStream_t st);
int launch_exact_min_dist(const float *x, int64_t b, int32_t d, int64_t ldx, const float *centers,
                          const int64_t *best, const float *xn, const float *cn, const float *counts, float thr,
                          float r, float *mind, cudaStream_t st);

// kmeans_update.cu
int launch_partition(const int64_t *best, int64_t b, int32_t k, uint32_t *blockhist, uint32_t *lrank,
                     uint32_t *total, uint32_t *seg_start, uint32_t *sorted_rows, float *counts_b,
                     cudaStream_t st, float *hist_max = nullptr, bool *hist_max_written = nullptr);
int launch_effective_lr(const float *counts_b, int32_t k, double lr, float *lr_eff, int32_t *fallback,
                        cudaStream_t st);
constexpr int kKmMaxWorld = 16;
struct KmPush {                  // where the deltas of a multi-GPU step go: receive buffers of the centroid owners
    float *red[kKmMaxWorld];     // rank o's buffer [world][k_own][d] (peer-mapped), slot [src rank][c - o*k_own]
    int32_t k_own, rank;
    __host__ __device__ float *slot(int32_t c, int32_t d) const {
        const int32_t o = c / k_own;
        return red[o] + ((int64_t)rank * k_own + (c - o * k_own)) * d;
    }
};
// Side stream of a workspace: independent kernels of a step run next to each other (fork = the side stream waits for
// an event of the main stream, join = the other way round; both are capturable into a CUDA graph).
struct KmFork {
    cudaStream_t side;
    cudaEvent_t ev_fork, ev_join;
};
inline int km_fork(const KmFork *f, cudaStream_t st) {
    ACAV_CUDA_TRY(cudaEventRecord(f->ev_fork, st));
    ACAV_CUDA_TRY(cudaStreamWaitEvent(f->side, f->ev_fork, 0));
    return 0;
}
inline int km_join(const KmFork *f, cudaStream_t st) {
    ACAV_CUDA_TRY(cudaEventRecord(f->ev_join, f->side));
    ACAV_CUDA_TRY(cudaStreamWaitEvent(st, f->ev_join, 0));
    return 0;
}
int launch_update(const float *x, int64_t ldx, int32_t k, int32_t d, const uint32_t *seg_start,
                  const uint32_t *sorted_rows, const float *counts_b, const float *lr_eff,
                  float *centers, float *counts, float *deltas, const KmPush *push, bool sequential, cudaStream_t st, const KmFork *fork = nullptr,
                  double lr_in = -1.0, int32_t *fallback = nullptr);   // lr_in >= 0: the kernels decide the step's lr from lr_eff[2]
bool km_heavy_ring();
int launch_sequential_lr(double lr, float *lr_eff, cudaStream_t st);

// kmeans_comm.cu: the multi-GPU step over NVLink peer memory (no NCCL): histogram exchange + lr decision, and the
// owner-side rank-ordered reduction of the pushed deltas with the broadcast of the new centroid rows
struct KmComm {
    unsigned char *arena[kKmMaxWorld];   // the same layout on every rank (arena[rank] is local memory)
    int32_t world, rank, k, d, k_own;
    unsigned int *seq;                   // local step counter (device)
    int *status;                         // local status word (device): 0 ok, 1 a peer's flag did not arrive in time
    unsigned long long spin_limit_ns;
};
size_t km_comm_arena_bytes(int32_t world, int32_t k, int32_t d);
KmPush km_comm_push_target(const KmComm &c);
int launch_km_hist_exchange(const KmComm &c, const float *counts_b_local, double lr, float *counts_global,
                            float *lr_eff, int32_t *fallback, float *counts, cudaStream_t st);
int launch_km_reduce_broadcast(const KmComm &c, const float *counts_global, const float *lr_eff, float *centers,
                               cudaStream_t st);
int launch_apply_deltas(float *centers, const float *deltas, int64_t n, cudaStream_t st);
int launch_underused_flags(const float *counts, int32_t k, const float *thr_dev, float *flags, cudaStream_t st);

// mi_scan.cu
struct MiState {                 // device-resident table + running sums of one clustering pair
    uint32_t *cells;             // [w] packed (c1 << 16 | c2); 0xFFFFFFFF = removed
    uint32_t *n_cells;           // [k_a * k_v] contingency counts  N   (mi.py cache['N'])
    uint32_t *a_cols;            // [k_v] column marginals           a   (cache['a'], index c2)
    uint32_t *b_rows;            // [k_a] row marginals              b   (cache['b'], index c1)
    float *gain;                 // [k_a * k_v] per-iteration score of adding one sample to a cell
    float *col_term;             // [k_v]
    float *row_term;             // [k_a]
    float *sums;                 // {NlogN, aloga, blogb, n, fN0, fa0}
    const float *logs;           // torch-CPU log table
    int64_t n_logs;
    unsigned long long *key;     // [2] running argmax key + packed cell of the local winner
    int64_t w;
    int64_t pos_base;
    int32_t k_a, k_v;
};
int launch_mi_pack(const int64_t *cells, int64_t w, uint32_t *packed, cudaStream_t st);
int launch_mi_reset(const MiState &s, const float *consts_dev, cudaStream_t st);
int launch_mi_add_sample(const MiState &s, int32_t c1, int32_t c2, cudaStream_t st);
int launch_mi_refresh_terms(const MiState &s, cudaStream_t st, void *pub = nullptr, size_t pub_bytes = 0,
                            unsigned int *bar = nullptr);   // pub/bar given: zeroed for the next run of a persistent loop
int launch_mi_gain_table(const MiState &s, cudaStream_t st);
int launch_mi_scan(const MiState &s, int sm_count, cudaStream_t st);
int launch_mi_emit(const MiState &s, unsigned long long *out, cudaStream_t st);
int launch_mi_apply(const MiState &s, const unsigned long long *key_cells, int32_t n,
                    int64_t *out_pos, float *out_gain, cudaStream_t st);

// mi_persistent.cu
int mi_partition_scratch_tiles(int64_t w);
int launch_mi_partition(const uint32_t *cells, int64_t w, int32_t k_a, uint32_t *tilehist, uint32_t *row_total,
                        uint32_t *row_start, uint16_t *c2s, uint32_t *pos_s, int64_t stream_capacity,
                        uint16_t pad_marker, cudaStream_t st);
int launch_mi_block_sort(uint16_t *c2s, uint32_t *pos_s, int64_t w_padded, cudaStream_t st);
int64_t mi_stream_capacity(int64_t w, int32_t k_a);
int mi_stream_block();
int mi_persistent_rows_that_fit(int32_t k_a, int32_t k_v);
size_t mi_pub_bytes(int32_t grid);
size_t mi_mail_bytes(int32_t world);
constexpr int kMiMaxWorld = 16;
int launch_mi_persistent(const MiState &s, uint32_t *n_alt, uint16_t *c2s, const uint32_t *pos_s,
                         const uint32_t *row_start, const uint32_t *chunk_start, int32_t grid, void *pub,
                         unsigned int *bar, int64_t n_picks, int64_t *out_pos, float *out_gain, int32_t rows_smem,
                         int32_t world, int32_t rank, unsigned int seq_base, void *mail_local, void *const *mail_peer,
                         long long *dbg, int *status, unsigned long long spin_limit_ns, cudaStream_t st, bool sync_clean = false);

// mi_stream8.cu (one-byte candidate stream, register-staged)
int mi_s8_k_rows(int32_t k_a, int32_t k_v);
int mi_s8_block();
int mi_s8_tiles(int64_t w);
int64_t mi_s8_stream_capacity(int64_t w, int32_t k_a, int32_t k_v);
int mi_s8_rows_that_fit(int32_t k_a, int32_t k_v);
bool mi_s8_supported(int32_t k_a, int32_t k_v);
int mi_s8_slots_for_rows(int32_t rows);
int launch_mi_s8_count(const uint32_t *cells, int64_t w, int32_t k_a, int32_t k_v, uint32_t *tilehist,
                       uint32_t *row_total, cudaStream_t st);
int launch_mi_s8_scatter(const uint32_t *cells, int64_t w, int32_t k_a, int32_t k_v, const uint32_t *tilehist,
                         const uint32_t *row_start, uint8_t *stage_stream, uint32_t *stage_pos, int64_t stream_capacity,
                         cudaStream_t st);
int launch_mi_s8_block_arrange(const uint8_t *stage_stream, const uint32_t *stage_pos, const uint32_t *blk_src,
                               uint8_t *stream, uint32_t *pos_s, unsigned long long *vrank, int64_t w_padded,
                               cudaStream_t st);
// chunks: [grid + 1] records of four uint32 {first slot, slots, distinct sub-rows, stream offset} (S8Chunk)
int launch_mi_stream8(const MiState &s, uint32_t *n_alt, uint8_t *stream, const uint32_t *pos_s,
                      const unsigned long long *vrank, const uint32_t *slot_start, const uint32_t *slot_row, const uint32_t *slot_u, const void *chunks,
                      int32_t grid, void *pub, unsigned int *bar, int64_t n_picks, int64_t *out_pos, float *out_gain,
                      int32_t rows_smem, int32_t variant, int32_t world, int32_t rank, unsigned int seq_base,
                      void *mail_local, void *const *mail_peer, long long *dbg, int *status,
                      unsigned long long spin_limit_ns, cudaStream_t st, bool sync_clean = false);

// mi_cells.cu (cell-index loop)
int mi_cells_tiles(int64_t w);
bool mi_cells_smem_fits(int32_t k_a, int32_t k_v);
int launch_mi_cells_build(const MiState &s, uint32_t *tilehist, uint32_t *total, uint32_t *start, uint32_t *tmp_cells,
                          uint32_t *tmp_pos, uint32_t *sorted_cells, uint32_t *sorted_pos, uint32_t *cell_start,
                          uint32_t *head, uint32_t *first_pos, int64_t *n_live_host, cudaStream_t st);
int launch_mi_cells(const MiState &s, const uint32_t *cell_start, const uint32_t *sorted_pos, uint32_t *head,
                    uint32_t *first_pos, int32_t grid, void *pub, unsigned int *bar, int64_t n_picks,
                    int64_t *out_pos, float *out_gain, int32_t world, int32_t rank, unsigned int seq_base,
                    void *mail_local, void *const *mail_peer, int *status, unsigned long long spin_limit_ns,
                    cudaStream_t st, bool sync_clean = false);

// mi_dense.cu
struct MiDense {
    uint32_t *n_cells;     // [P, C, C]
    uint32_t *a_cols;      // [P, C]  (index c2)
    uint32_t *b_rows;      // [P, C]  (index c1)
    uint32_t *n;           // [P]
    double *sums;          // [P, 3]  NlogN, aloga, blogb
    int32_t p, c;
};
int launch_mi_dense_reset(const MiDense &s, cudaStream_t st);
int launch_mi_dense_add(const MiDense &s, const int64_t *cells, int64_t m, cudaStream_t st);
int launch_mi_dense_score(const MiDense &s, const int64_t *cells, int64_t nb, float *per_pair, float *scores,
                          cudaStream_t st);
int launch_mi_dense_score_exact(const MiDense &s, const int64_t *cells, int64_t nb, const float *logs,
                                const float *consts_host, float *per_pair, float *scores, cudaStream_t st);
// scratch: 4*P*C + P doubles (row / column sums of the EMI terms, plain and with the marginal bumped, and their total)
int launch_mi_dense_score_ami(const MiDense &s, const int64_t *cells, int64_t nb, double *scratch, int32_t average_method,
                              float *per_pair, float *scores, cudaStream_t st);

}  // namespace acav
