// Dense-table MI scoring for the reference's default CLI measure `batch_mi`
// (subset_selection/code/measures/batch.py:10-260 on top of measures/mi.py:85-98).
//
// Per iteration the reference builds, for each of B = 20 sampled candidates and each clustering pair, the
// dense C x C table "current table + one-hot(candidate)" and evaluates
//     MI = sum_ij N_ij/n * (log N_ij + log n - log a_j - log b_i)                       (mi.py:85-91)
// i.e. 4*B*P*C*C logs (126 ms at C = 1024 on the reference's CPU path).  Because
// sum_ij N_ij log a_j = sum_j a_j log a_j, the same number is (NlogN - aloga - blogb)/n + log n, and adding
// one sample changes one cell, one column marginal and one row marginal: O(1) per (candidate, pair) from
// three running sums.  The sums are kept in fp64, so the score is the exact value the reference's fp32
// dense sum approximates (agreement ~1e-6 relative; the reference's own CPU and CUDA paths differ from
// each other by as much).  Which of several mathematically tied candidates torch.topk returns in the
// reference depends on fp32 summation noise of its dense sum, so index-for-index parity is not defined
// for this measure; scores are checked to 1e-5 (tests/test_batch_mi_gpu.py).
#include "common.cuh"
#include "kernels.cuh"

namespace acav {

__device__ __forceinline__ double dense_eps() { return 2.220446049250313e-16; }     // np.finfo('float64').eps, mi.py:25
// x log x of a table entry holding `k` samples on top of its "empty" value e0 (eps, C*eps or C*C*eps):
// k >= 1 absorbs e0 in fp32, exactly as in the reference's fp32 tables
__device__ __forceinline__ double xlogx_d(uint32_t k, double e0) {
    const double v = k == 0 ? e0 : (double)k;
    return v * log(v);
}

__global__ void mi_dense_reset_kernel(MiDense s) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= s.p) return;
    const double e = dense_eps(), c = (double)s.c;
    s.sums[3 * p + 0] = c * c * (e * log(e));
    s.sums[3 * p + 1] = c * ((c * e) * log(c * e));
    s.sums[3 * p + 2] = c * ((c * e) * log(c * e));
    s.n[p] = 0;
}

// count m samples into the tables (add_samples batch.py:190-193 / update_cache :152-154); one thread per pair
__global__ void mi_dense_add_kernel(MiDense s, const int64_t *__restrict__ cells, int64_t m) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= s.p) return;
    const double e = dense_eps(), c = (double)s.c;
    double nlogn = s.sums[3 * p], aloga = s.sums[3 * p + 1], blogb = s.sums[3 * p + 2];
    for (int64_t i = 0; i < m; ++i) {
        const int32_t c1 = (int32_t)cells[(i * s.p + p) * 2], c2 = (int32_t)cells[(i * s.p + p) * 2 + 1];
        uint32_t *x = s.n_cells + ((int64_t)p * s.c + c1) * s.c + c2;
        uint32_t *y = s.a_cols + (int64_t)p * s.c + c2;
        uint32_t *z = s.b_rows + (int64_t)p * s.c + c1;
        nlogn += xlogx_d(*x + 1, e) - xlogx_d(*x, e);
        aloga += xlogx_d(*y + 1, c * e) - xlogx_d(*y, c * e);
        blogb += xlogx_d(*z + 1, c * e) - xlogx_d(*z, c * e);
        *x += 1; *y += 1; *z += 1;
    }
    s.sums[3 * p] = nlogn; s.sums[3 * p + 1] = aloga; s.sums[3 * p + 2] = blogb;
    s.n[p] += (uint32_t)m;
}

// MI of (table + one sample) for every (candidate, pair)
__global__ void mi_dense_pair_score_kernel(MiDense s, const int64_t *__restrict__ cells, int64_t nb,
                                           float *__restrict__ per_pair) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nb * s.p) return;
    const int p = (int)(i % s.p);
    const double e = dense_eps(), c = (double)s.c;
    const int32_t c1 = (int32_t)cells[i * 2], c2 = (int32_t)cells[i * 2 + 1];
    const uint32_t x = s.n_cells[((int64_t)p * s.c + c1) * s.c + c2];
    const uint32_t y = s.a_cols[(int64_t)p * s.c + c2], z = s.b_rows[(int64_t)p * s.c + c1];
    const double nlogn = s.sums[3 * p] + xlogx_d(x + 1, e) - xlogx_d(x, e);
    const double aloga = s.sums[3 * p + 1] + xlogx_d(y + 1, c * e) - xlogx_d(y, c * e);
    const double blogb = s.sums[3 * p + 2] + xlogx_d(z + 1, c * e) - xlogx_d(z, c * e);
    const double n1 = (double)s.n[p] + 1.0;
    per_pair[i] = (float)((nlogn - aloga - blogb) / n1 + log(n1));
}

// scores.mean(dim=-1) (batch.py:144): fp32, pairs added in order, one division
__global__ void mi_dense_mean_kernel(const float *__restrict__ per_pair, int64_t nb, int32_t p,
                                     float *__restrict__ scores) {
    const int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= nb) return;
    float acc = 0.f;
    for (int32_t j = 0; j < p; ++j) acc = __fadd_rn(acc, per_pair[b * p + j]);
    scores[b] = __fdiv_rn(acc, (float)p);
}

int launch_mi_dense_reset(const MiDense &s, cudaStream_t st) {
    const size_t cells = (size_t)s.p * s.c * s.c;
    ACAV_CUDA_TRY(cudaMemsetAsync(s.n_cells, 0, sizeof(uint32_t) * cells, st));
    ACAV_CUDA_TRY(cudaMemsetAsync(s.a_cols, 0, sizeof(uint32_t) * (size_t)s.p * s.c, st));
    ACAV_CUDA_TRY(cudaMemsetAsync(s.b_rows, 0, sizeof(uint32_t) * (size_t)s.p * s.c, st));
    mi_dense_reset_kernel<<<(unsigned)ceil_div(s.p, 64), 64, 0, st>>>(s);
    ACAV_LAUNCH_CHECK();
    return 0;
}

int launch_mi_dense_add(const MiDense &s, const int64_t *cells, int64_t m, cudaStream_t st) {
    if (m == 0) return 0;
    mi_dense_add_kernel<<<(unsigned)ceil_div(s.p, 64), 64, 0, st>>>(s, cells, m);
    ACAV_LAUNCH_CHECK();
    return 0;
}

int launch_mi_dense_score(const MiDense &s, const int64_t *cells, int64_t nb, float *per_pair, float *scores,
                          cudaStream_t st) {
    if (nb == 0) return 0;
    mi_dense_pair_score_kernel<<<(unsigned)ceil_div(nb * s.p, 128), 128, 0, st>>>(s, cells, nb, per_pair);
    ACAV_LAUNCH_CHECK();
    mi_dense_mean_kernel<<<(unsigned)ceil_div(nb, 128), 128, 0, st>>>(per_pair, nb, s.p, scores);
    ACAV_LAUNCH_CHECK();
    return 0;
}

}  // namespace acav
