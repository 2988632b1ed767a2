// Arithmetic of the adjusted-MI measure `ami` (reference subset_selection/code/measures/mi.py:212-262, EfficientAMI),
// shared between the device kernels (mi_dense.cu) and a host build used by the CPU tests
// (tests/native/mi_ami_math_host.cpp).  fp64 throughout: the reference evaluates these expressions on fp32 tensors,
// where the nine lgamma terms of every cell cancel to a few units out of ~n*log(n); this is the value they approximate.
//
// For one clustering pair with table N (column marginals a_j, row marginals b_i, n samples) the reference computes,
// for the table "current + one candidate sample":
//   MI  = sum_ij T_ij,                T_ij = N_ij/n * (log N_ij + log n - log a_j - log b_i)          calc_MI  :85-91
//   EMI = sum_ij T_ij * exp(L_ij),    L_ij = lgamma(a_j+1) + lgamma(b_i+1) + lgamma(n-a_j+1) + lgamma(n-b_i+1)
//                                            - lgamma(n+1) - lgamma(N_ij+1) - lgamma(a_j-N_ij+1) - lgamma(b_i-N_ij+1)
//                                            - lgamma(n-a_j-b_i+N_ij+1)                                 calc_EMI :217-231
//   H_a = -sum_j a_j/n log(a_j/n),  H_b likewise                                                     calc_entropy :233-236
//   AMI = (MI - EMI) / max(mean(H_a, H_b) - EMI, eps)                                                calc_AMI :247-262
// Entries that hold no sample carry the reference's "empty" values (eps, C*eps; init_cache :32-39); a count >= 1
// absorbs them in the reference's fp32 tables, so an entry is its integer count or its empty value.
//
// The numerator is accumulated as the GAP  MI - EMI = sum_ij T_ij * (1 - exp(L_ij)) = -sum_ij T_ij * expm1(L_ij), not as
// the difference of the two sums (better conditioned where EMI is close to MI).  One regime is singular: while ALL samples
// sit in one cell (a candidate joining the cell of the first picks) both entropies vanish and the denominator is a
// difference of ~1e-14 "empty value" terms; there every L_ij is exactly 0 -- in the reference's fp32 too: its lgamma terms
// cancel pairwise -- so the reference gets MI == EMI bit for bit and AMI = 0.  That case returns a gap of exactly 0.
#pragma once
#include <math.h>
#include <stdint.h>

#if defined(__CUDACC__)
#define ACAV_AMI_HD __host__ __device__ __forceinline__
#else
#define ACAV_AMI_HD inline
#endif

namespace acav {

constexpr double kAmiEps = 2.220446049250313e-16;         // np.finfo('float64').eps, mi.py:25

ACAV_AMI_HD double ami_value(uint32_t k, double empty) { return k == 0 ? empty : (double)k; }

// one cell's share of MI - EMI for counts x = N_ij, y = a_j, z = b_i out of m samples (x <= y, x <= z, y + z - x <= m)
ACAV_AMI_HD double ami_gap_term(uint32_t x, uint32_t y, uint32_t z, uint32_t m, double c) {
    const double vn = ami_value(x, kAmiEps), va = ami_value(y, c * kAmiEps), vb = ami_value(z, c * kAmiEps);
    const double n = (double)m;
    const double t = vn / n * (log(vn) + log(n) - (log(va) + log(vb)));
    const double l = lgamma((double)y + 1.0) + lgamma((double)z + 1.0) + lgamma((double)(m - y) + 1.0) +
                     lgamma((double)(m - z) + 1.0) -
                     (lgamma(n + 1.0) + lgamma((double)x + 1.0) + lgamma((double)(y - x) + 1.0) +
                      lgamma((double)(z - x) + 1.0) + lgamma((double)(m - y - z + x) + 1.0));
    return -t * expm1(l);
}

// Running pieces of the gap of one pair for the table "current counts, n + 1 samples" (everything a candidate does not
// touch), G = ami_gap_term:
//   base = sum_ij G(N_ij, a_j, b_i), row[i] = sum_j G(N_ij, a_j, b_i), row_up[i] = sum_j G(N_ij, a_j, b_i + 1),
//   col[j] / col_up[j] likewise over i with a_j + 1.
// Gap of the table with one more sample in cell (i, j), N = N_ij, a = a_j, b = b_i, m = n + 1:
ACAV_AMI_HD double ami_gap_with_sample(double base, double row_i, double row_up_i, double col_j, double col_up_j,
                                       uint32_t N, uint32_t a, uint32_t b, uint32_t m, double c) {
    if (a + 1 == m && b + 1 == m) return 0.0;                      // every sample in cell (i, j): L == 0 everywhere
    const double untouched = base - row_i - col_j + ami_gap_term(N, a, b, m, c);
    const double row_part = row_up_i - ami_gap_term(N, a, b + 1, m, c);
    const double col_part = col_up_j - ami_gap_term(N, a + 1, b, m, c);
    return untouched + row_part + col_part + ami_gap_term(N + 1, a + 1, b + 1, m, c);
}

// average_method of generalized_mean (mi.py:200-209): 0 arithmetic (default), 1 max, 2 min
// gap = MI - EMI
ACAV_AMI_HD double ami_from_parts(double mi, double gap, double ha, double hb, int average_method) {
    double normalizer;
    if (average_method == 1) normalizer = ha > hb ? ha : hb;
    else if (average_method == 2) normalizer = ha < hb ? ha : hb;
    else normalizer = (ha + hb) / 2.0;
    double denominator = normalizer - (mi - gap);                 // normalizer - EMI
    if (!(denominator > kAmiEps)) denominator = kAmiEps;          // ensure_nonzero :193-198 (torch.max with eps)
    return gap / denominator;
}

}  // namespace acav
