// Host build of acav100m_b200/csrc/mi_dense_exact_math.h for the CPU test-suite (tests/test_dense_exact_cpu.py): the
// reference's dense MI of (table + one-hot) per (candidate, pair), evaluated with the header the CUDA kernel is built
// from -- serially, and lane by lane the way the warp kernel splits the sum -- so that both the arithmetic and the
// decomposition are compared with torch bit for bit without a GPU.  TEST CODE: nothing here is linked into
// libacav_b200.so.
#include <cstdint>
#include <vector>

#include "../../acav100m_b200/csrc/mi_dense_exact_math.h"

using namespace acav;

extern "C" {

float host_dense_exact_row_sum(const float *x, int64_t n) {
    return dense_exact_row_sum_serial(n, [&](int64_t e) { return x[e]; });
}

// N [P, C, C], a [P, C], b [P, C] counts, n [P]; cells int64 [nb, P, 2]; consts = {eps, a0, b0, log_eps, log_a0, log_b0}
// per_pair [nb, P].  mode 0: serial order; mode 1: emulation of the warp kernel (32 lanes, shuffles as array reads)
void host_dense_exact_score(const uint32_t *N, const uint32_t *a, const uint32_t *b, const uint32_t *n, int32_t P,
                            int32_t C, const int64_t *cells, int64_t nb, const float *logs, const float *consts,
                            int32_t mode, float *per_pair) {
    DenseExactConsts k{consts[0], consts[1], consts[2], consts[3], consts[4], consts[5]};
    const int64_t cc = (int64_t)C * C;
    for (int64_t bi = 0; bi < nb; ++bi) {
        for (int p = 0; p < P; ++p) {
            DenseExactView v;
            v.N = N + p * cc; v.a = a + (int64_t)p * C; v.b = b + (int64_t)p * C; v.C = C;
            v.c1 = (int32_t)cells[(bi * P + p) * 2]; v.c2 = (int32_t)cells[(bi * P + p) * 2 + 1];
            v.nf = (float)(n[p] + 1u); v.ln = logs[n[p] + 1u];
            auto elem = [&](int64_t e) { return dense_exact_elem(v, k, logs, e); };
            float out;
            if (mode == 0) {
                out = dense_exact_row_sum_serial(cc, elem);
            } else {
                const DenseExactShape s = dense_exact_shape(cc);
                float part[32];
                for (int lane = 0; lane < 32; ++lane) {
                    const int kk = lane / s.L, l = lane % s.L;
                    part[lane] = lane < 4 * s.L ? dense_exact_stream(s, kk, l, elem) : 0.f;
                    if (lane < s.L) part[lane] = dense_exact_leftover(s, l, part[lane], elem);
                }
                for (int kk = 1; kk < 4; ++kk)
                    for (int l = 0; l < s.L; ++l) part[l] = part[l] + part[kk * s.L + l];
                if (s.L == 1) out = part[0];
                else {
                    float acc = 0.f;
                    for (int64_t e = s.V * s.L; e < cc; ++e) acc = acc + elem(e);
                    for (int l = 0; l < s.L; ++l) acc = acc + part[l];
                    out = acc;
                }
            }
            per_pair[bi * P + p] = out;
        }
    }
}

}  // extern "C"
