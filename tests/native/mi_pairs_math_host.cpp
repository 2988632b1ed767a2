// Host build of acav100m_b200/csrc/mi_pairs_math.h for the CPU test-suite (tests/test_mi_pairs_math_cpu.py).
//
// The multi-pair greedy-MI kernels (csrc/mi_pairs.cu) take all their arithmetic from that header; this file
// compiles the same functions with g++ -ffp-contract=off and walks them through the kernels' data flow --
// marginal terms per (pair, id), one score per table cell, candidates gathered and averaged with pairs_mean,
// winner applied with pairs_bump -- so that the decomposition itself (not only the formulas) is compared with
// the oracle bit for bit without a GPU.  TEST CODE: nothing here is linked into libacav_b200.so.
#include <cstdint>
#include <cstring>
#include <vector>

#include "../../acav100m_b200/csrc/mi_pairs_math.h"

using namespace acav;

extern "C" {

float host_pairs_mean(const float *g, int32_t P) {
    return pairs_mean(P, [&](int p) { return g[p]; });
}

// ids: int32 [W, D]; pairs: int32 [P, 2]; consts: [P, 6] {fN0, fa0, n0, NlogN0, aloga0, blogb0}
int64_t host_pairs_greedy(const int32_t *ids, int64_t W, int32_t D, int32_t C, const int32_t *pairs, int32_t P,
                          const float *logs, const float *consts, int64_t n_picks, int64_t *out_pos, float *out_gain,
                          float *out_sums) {
    const int64_t cc = (int64_t)C * C;
    std::vector<uint32_t> N((size_t)P * cc, 0), a((size_t)P * C, 0), b((size_t)P * C, 0);
    std::vector<float> gain((size_t)P * cc), col((size_t)P * C), row((size_t)P * C), sums((size_t)P * 4);
    std::vector<uint8_t> gone((size_t)(W > 0 ? W : 1), 0);
    for (int p = 0; p < P; ++p) {
        sums[4 * p] = consts[6 * p + 3]; sums[4 * p + 1] = consts[6 * p + 4]; sums[4 * p + 2] = consts[6 * p + 5];
        sums[4 * p + 3] = consts[6 * p + 2];
    }
    int64_t it = 0;
    for (; it < n_picks; ++it) {
        for (int p = 0; p < P; ++p) {                                   // mip_terms + mip_gain_kernel
            const float n1 = sums[4 * p + 3] + 1.0f, fn0 = consts[6 * p], fa0 = consts[6 * p + 1];
            for (int i = 0; i < C; ++i) {
                col[(size_t)p * C + i] = pairs_marginal_term(sums[4 * p + 1], a[(size_t)p * C + i], fa0, n1, logs);
                row[(size_t)p * C + i] = pairs_marginal_term(sums[4 * p + 2], b[(size_t)p * C + i], fa0, n1, logs);
            }
            for (int c1 = 0; c1 < C; ++c1)
                for (int c2 = 0; c2 < C; ++c2)
                    gain[p * cc + (int64_t)c1 * C + c2] =
                        pairs_cell_score(sums[4 * p], N[p * cc + (int64_t)c1 * C + c2], fn0, n1, col[(size_t)p * C + c2],
                                         row[(size_t)p * C + c1], logs);
        }
        int64_t best = -1;                                              // mip_scan_kernel
        float bs = 0.f;
        for (int64_t w = 0; w < W; ++w) {
            if (gone[w]) continue;
            const int32_t *r = ids + w * D;
            const float sc = pairs_mean(P, [&](int p) {
                return gain[p * cc + (int64_t)r[pairs[2 * p]] * C + r[pairs[2 * p + 1]]];
            });
            if (best < 0 || sc > bs) { best = w; bs = sc; }
        }
        if (best < 0) break;
        out_pos[it] = best; out_gain[it] = bs;
        for (int p = 0; p < P; ++p) {                                   // mip_apply_kernel
            const int32_t c1 = ids[best * D + pairs[2 * p]], c2 = ids[best * D + pairs[2 * p + 1]];
            const float fn0 = consts[6 * p], fa0 = consts[6 * p + 1];
            uint32_t &x = N[p * cc + (int64_t)c1 * C + c2], &y = a[(size_t)p * C + c2], &z = b[(size_t)p * C + c1];
            sums[4 * p] = pairs_bump(sums[4 * p], x, fn0, logs);
            sums[4 * p + 1] = pairs_bump(sums[4 * p + 1], y, fa0, logs);
            sums[4 * p + 2] = pairs_bump(sums[4 * p + 2], z, fa0, logs);
            sums[4 * p + 3] = sums[4 * p + 3] + 1.0f;
            x += 1; y += 1; z += 1;
        }
        gone[best] = 1;
    }
    if (out_sums) std::memcpy(out_sums, sums.data(), sizeof(float) * sums.size());
    return it;
}

}  // extern "C"
