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
#include "mi_ami_math.h"
#include "mi_dense_exact_math.h"
#include "mi_pairs_math.h"

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

// ---- adjusted MI (`ami`, reference measures/mi.py:212-262) -------------------------------------------------------------
// The reference evaluates nine lgamma per cell of every candidate's dense table: O(W*P*C*C) per iteration.  Adding one
// sample to cell (i, j) changes n (every cell), a_j (column j) and b_i (row i), so with the per-iteration sums of
// mi_ami_math.h -- the per-cell shares of MI - EMI over the whole table, per row and per column, each with and without
// the marginal bumped, all at n + 1 samples -- a candidate's MI - EMI is four corrections: O(P*C*C) per iteration for
// the sums, O(1) per candidate.

__device__ __forceinline__ double block_sum_f64(double v, double *scratch) {
    v = warp_sum_f64(v);
    if (threadIdx.x % kWarp == 0) scratch[threadIdx.x / kWarp] = v;
    __syncthreads();
    double t = 0.0;
    if (threadIdx.x == 0)
        for (int wi = 0; wi < (int)(blockDim.x / kWarp); ++wi) t += scratch[wi];
    __syncthreads();
    return t;                                   // valid in thread 0
}

// grid (C, P, 2): z = 0 sums row blockIdx.x over its columns, z = 1 sums column blockIdx.x over its rows
__global__ void __launch_bounds__(128) mi_dense_ami_lines_kernel(MiDense s, double *__restrict__ line,
                                                                 double *__restrict__ line_up) {
    __shared__ double scratch[4];
    const int32_t q = blockIdx.x, p = blockIdx.y;
    const bool rows = blockIdx.z == 0;
    const int64_t o = (int64_t)p * s.c;
    const uint32_t m = s.n[p] + 1u;
    const double c = (double)s.c;
    double acc = 0.0, acc_up = 0.0;
    for (int32_t t = threadIdx.x; t < s.c; t += blockDim.x) {
        const int32_t i = rows ? q : t, j = rows ? t : q;
        const uint32_t x = s.n_cells[(o + i) * s.c + j], y = s.a_cols[o + j], z = s.b_rows[o + i];
        acc += ami_gap_term(x, y, z, m, c);
        acc_up += rows ? ami_gap_term(x, y, z + 1u, m, c) : ami_gap_term(x, y + 1u, z, m, c);
    }
    const int64_t out = ((int64_t)blockIdx.z * s.p + p) * s.c + q;
    acc = block_sum_f64(acc, scratch);
    acc_up = block_sum_f64(acc_up, scratch);
    if (threadIdx.x == 0) { line[out] = acc; line_up[out] = acc_up; }
}

// base[p] = sum over the rows of pair p
__global__ void __launch_bounds__(128) mi_dense_ami_base_kernel(MiDense s, const double *__restrict__ line,
                                                                double *__restrict__ base) {
    __shared__ double scratch[4];
    const int32_t p = blockIdx.x;
    double acc = 0.0;
    for (int32_t i = threadIdx.x; i < s.c; i += blockDim.x) acc += line[(int64_t)p * s.c + i];
    acc = block_sum_f64(acc, scratch);
    if (threadIdx.x == 0) base[p] = acc;
}

// AMI of (table + one sample) for every (candidate, pair)
__global__ void mi_dense_ami_score_kernel(MiDense s, const int64_t *__restrict__ cells, int64_t nb,
                                          const double *__restrict__ line, const double *__restrict__ line_up,
                                          const double *__restrict__ base, int32_t average_method,
                                          float *__restrict__ per_pair) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nb * s.p) return;
    const int p = (int)(i % s.p);
    const double e = dense_eps(), c = (double)s.c;
    const int32_t c1 = (int32_t)cells[i * 2], c2 = (int32_t)cells[i * 2 + 1];
    const int64_t o = (int64_t)p * s.c;
    const uint32_t x = s.n_cells[(o + c1) * s.c + c2], y = s.a_cols[o + c2], z = s.b_rows[o + c1];
    const double nlogn = s.sums[3 * p] + xlogx_d(x + 1, e) - xlogx_d(x, e);
    const double aloga = s.sums[3 * p + 1] + xlogx_d(y + 1, c * e) - xlogx_d(y, c * e);
    const double blogb = s.sums[3 * p + 2] + xlogx_d(z + 1, c * e) - xlogx_d(z, c * e);
    const uint32_t m = s.n[p] + 1u;
    const double n1 = (double)m, logn = log(n1);
    const double mi = (nlogn - aloga - blogb) / n1 + logn;
    const double ha = logn - aloga / n1, hb = logn - blogb / n1;         // -sum a/n log(a/n), sum a = n
    const int64_t cols = (int64_t)s.p * s.c;                               // lines of z = 1 follow those of z = 0
    const double gap = ami_gap_with_sample(base[p], line[o + c1], line_up[o + c1], line[cols + o + c2],
                                           line_up[cols + o + c2], x, y, z, m, c);
    per_pair[i] = (float)ami_from_parts(mi, gap, ha, hb, average_method);
}

int launch_mi_dense_score_ami(const MiDense &s, const int64_t *cells, int64_t nb, double *scratch, int32_t average_method,
                              float *per_pair, float *scores, cudaStream_t st) {
    if (nb == 0) return 0;
    const int64_t lines = 2 * (int64_t)s.p * s.c;
    double *line = scratch, *line_up = scratch + lines, *base = scratch + 2 * lines;
    mi_dense_ami_lines_kernel<<<dim3((unsigned)s.c, (unsigned)s.p, 2), 128, 0, st>>>(s, line, line_up);
    ACAV_LAUNCH_CHECK();
    mi_dense_ami_base_kernel<<<(unsigned)s.p, 128, 0, st>>>(s, line, base);
    ACAV_LAUNCH_CHECK();
    mi_dense_ami_score_kernel<<<(unsigned)ceil_div(nb * s.p, 128), 128, 0, st>>>(s, cells, nb, line, line_up, base,
                                                                                 average_method, per_pair);
    ACAV_LAUNCH_CHECK();
    mi_dense_mean_kernel<<<(unsigned)ceil_div(nb, 128), 128, 0, st>>>(per_pair, nb, s.p, scores);
    ACAV_LAUNCH_CHECK();
    return 0;
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

// ---- bit-exact dense MI (mi_dense_exact_math.h): one warp per (candidate, pair) ---------------------------------------
// Lane k*L + l owns accumulator stream (k, l) of ATen's cascade sum over the C*C cells of the candidate's table and walks
// 32 consecutive cells per step with the rest of the warp; the 4 x L partials are folded with shuffles in ATen's order.
__global__ void __launch_bounds__(32)
mi_dense_exact_kernel(MiDense s, const int64_t *__restrict__ cells, const float *__restrict__ logs,
                      DenseExactConsts k, float *__restrict__ per_pair) {
    const int64_t bi = blockIdx.x;
    const int p = blockIdx.y;
    const int lane = threadIdx.x;
    const int64_t cc = (int64_t)s.c * s.c;
    DenseExactView v;
    v.N = s.n_cells + (int64_t)p * cc; v.a = s.a_cols + (int64_t)p * s.c; v.b = s.b_rows + (int64_t)p * s.c; v.C = s.c;
    v.c1 = (int32_t)cells[(bi * s.p + p) * 2]; v.c2 = (int32_t)cells[(bi * s.p + p) * 2 + 1];
    const uint32_t n1 = s.n[p] + 1u;
    v.nf = (float)n1; v.ln = logs[n1];
    auto elem = [&](int64_t e) { return dense_exact_elem(v, k, logs, e); };
    const DenseExactShape sh = dense_exact_shape(cc);
    const int kk = lane / sh.L, l = lane % sh.L;
    float part = lane < 4 * sh.L ? dense_exact_stream(sh, kk, l, elem) : 0.f;
    if (lane < sh.L) part = dense_exact_leftover(sh, l, part, elem);
#pragma unroll
    for (int q = 1; q < 4; ++q) {                                // partial[0] += partial[q], lanes 0 .. L-1
        const float other = __shfl_sync(0xffffffffu, part, (q * sh.L + l) & 31);
        if (lane < sh.L) part = part + other;
    }
    float out = part;                                            // L == 1: the row sum is lane 0's partial
    if (sh.L == 8) {
        float acc = 0.f;
        if (lane == 0)
            for (int64_t e = sh.V * sh.L; e < cc; ++e) acc = acc + elem(e);
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            const float pq = __shfl_sync(0xffffffffu, part, q);
            acc = acc + pq;
        }
        out = acc;
    }
    if (lane == 0) per_pair[bi * s.p + p] = out;
}

// scores.mean(dim=-1) in torch's CPU summation order (pairs_mean, mi_pairs_math.h)
__global__ void mi_dense_exact_mean_kernel(const float *__restrict__ per_pair, int64_t nb, int32_t p,
                                           float *__restrict__ scores) {
    const int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= nb) return;
    const float *row = per_pair + b * p;
    scores[b] = pairs_mean(p, [&](int j) { return row[j]; });
}

int launch_mi_dense_score_exact(const MiDense &s, const int64_t *cells, int64_t nb, const float *logs,
                                const float *consts_host, float *per_pair, float *scores, cudaStream_t st) {
    if (nb == 0) return 0;
    if (s.p > 65535 || s.p > kPairsMax) return ACAV_E_UNSUPPORTED;
    DenseExactConsts k{consts_host[0], consts_host[1], consts_host[2], consts_host[3], consts_host[4], consts_host[5]};
    mi_dense_exact_kernel<<<dim3((unsigned)nb, (unsigned)s.p), 32, 0, st>>>(s, cells, logs, k, per_pair);
    ACAV_LAUNCH_CHECK();
    mi_dense_exact_mean_kernel<<<(unsigned)ceil_div(nb, 128), 128, 0, st>>>(per_pair, nb, s.p, scores);
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
