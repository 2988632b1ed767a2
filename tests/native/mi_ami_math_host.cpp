// Host build of acav100m_b200/csrc/mi_ami_math.h for the CPU test-suite (tests/test_ami_cpu.py): the same functions
// the `ami` kernels of csrc/mi_dense.cu are built from, walked through the kernels' data flow (row / column sums of
// the EMI terms at n + 1 samples, then four corrections per candidate).  TEST CODE, not part of libacav_b200.so.
#include <cstdint>
#include <vector>

#include "../../acav100m_b200/csrc/mi_ami_math.h"

using namespace acav;

static double xlogx_host(uint32_t k, double empty) {
    const double v = k == 0 ? empty : (double)k;
    return v * log(v);
}

extern "C" {

// counts: uint32 N [C, C] (row c1, column c2) of ONE pair; cells: int32 [nb, 2] = (c1, c2); out: double [nb] AMI of
// (table + candidate), fp64
void host_ami_scores(const uint32_t *N, int32_t C, const int32_t *cells, int64_t nb, int32_t average_method,
                     double *out) {
    const double c = (double)C, e = kAmiEps;
    std::vector<uint32_t> a(C, 0), b(C, 0);
    uint32_t n = 0;
    for (int i = 0; i < C; ++i)
        for (int j = 0; j < C; ++j) { a[j] += N[i * C + j]; b[i] += N[i * C + j]; n += N[i * C + j]; }
    double nlogn = 0, aloga = 0, blogb = 0;
    for (int i = 0; i < C * C; ++i) nlogn += xlogx_host(N[i], e);
    for (int i = 0; i < C; ++i) { aloga += xlogx_host(a[i], c * e); blogb += xlogx_host(b[i], c * e); }
    const uint32_t m = n + 1;
    std::vector<double> row(C, 0), row_up(C, 0), col(C, 0), col_up(C, 0);      // mi_dense_ami_lines_kernel
    for (int i = 0; i < C; ++i)
        for (int j = 0; j < C; ++j) {
            const uint32_t x = N[i * C + j];
            const double f = ami_gap_term(x, a[j], b[i], m, c);
            row[i] += f; col[j] += f;
            row_up[i] += ami_gap_term(x, a[j], b[i] + 1, m, c);
            col_up[j] += ami_gap_term(x, a[j] + 1, b[i], m, c);
        }
    double base = 0;                                                              // mi_dense_ami_base_kernel
    for (int i = 0; i < C; ++i) base += row[i];
    for (int64_t w = 0; w < nb; ++w) {                                            // mi_dense_ami_score_kernel
        const int c1 = cells[2 * w], c2 = cells[2 * w + 1];
        const uint32_t x = N[c1 * C + c2], y = a[c2], z = b[c1];
        const double nl = nlogn + xlogx_host(x + 1, e) - xlogx_host(x, e);
        const double al = aloga + xlogx_host(y + 1, c * e) - xlogx_host(y, c * e);
        const double bl = blogb + xlogx_host(z + 1, c * e) - xlogx_host(z, c * e);
        const double n1 = (double)m, logn = log(n1);
        const double mi = (nl - al - bl) / n1 + logn;
        const double ha = logn - al / n1, hb = logn - bl / n1;
        const double gap = ami_gap_with_sample(base, row[c1], row_up[c1], col[c2], col_up[c2], x, y, z, m, c);
        out[w] = ami_from_parts(mi, gap, ha, hb, average_method);
    }
}

}  // extern "C"
