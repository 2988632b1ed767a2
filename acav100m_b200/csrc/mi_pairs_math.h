// Arithmetic of the multi-pair (P > 1) exact greedy-MI engine, shared between the device kernels
// (mi_pairs.cu, nvcc --fmad=false) and a host build of the same functions that the CPU test-suite
// compares with the oracle (tests/native/mi_pairs_math_host.cpp, g++ -ffp-contract=off).  Plain IEEE
// fp32 operators only: no contraction, no intrinsics, so both builds produce the same bits.
//
// Reference: subset_selection/code/measures/mi.py
//   per-pair score of adding one sample to cell (c1, c2)      get_last :322-333, calc_MI :368-381
//   mean over the P clustering pairs                           calc_score :76-78 (`scores.mean(dim=-1)`)
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define ACAV_HD __host__ __device__ __forceinline__
#define ACAV_UNROLL _Pragma("unroll")
#else
#define ACAV_HD inline
#define ACAV_UNROLL
#endif

namespace acav {

constexpr int kPairsMax = 256;          // P above this would enter ATen's cascade levels (see pairs_mean)
constexpr int kPairColsMax = 64;        // distinct clustering columns an engine can hold per candidate

// x*log(x) of a table entry: `k` samples on top of an "empty" value whose x*log(x) is f0; counts >= 1 absorb
// the empty value in fp32 (eps + 1 == 1), so the entry is the exact integer k and log comes from the table
// torch's CPU kernel produced (DESIGN.md "MI exactness").
ACAV_HD float pairs_xlogx(uint32_t k, float f0, const float *logs) {
    return k == 0 ? f0 : (float)k * logs[k];
}
// prev - f(k) + f(k+1), left to right (update_nlogn, mi.py:339-340)
ACAV_HD float pairs_bump(float prev, uint32_t k, float f0, const float *logs) {
    return (prev - pairs_xlogx(k, f0, logs)) + pairs_xlogx(k + 1, 0.f, logs);
}
// (-aloga')/n'  resp.  (-blogb')/n'   (calc_MI :376-377)
ACAV_HD float pairs_marginal_term(float sum, uint32_t k, float fa0, float n1, const float *logs) {
    return (-pairs_bump(sum, k, fa0, logs)) / n1;
}
// ((NlogN'/n' + ta) + tb) + log n'   (calc_MI :375-380)
ACAV_HD float pairs_cell_score(float nlogn, uint32_t x, float fn0, float n1, float ta, float tb,
                               const float *logs) {
    const float tn = pairs_bump(nlogn, x, fn0, logs) / n1;
    return ((tn + ta) + tb) + logs[(int64_t)n1];
}

// `scores.mean(dim=-1)` of a contiguous fp32 [W, P] tensor as torch's CPU kernel evaluates it, per row:
// ATen/native/cpu/SumKernel.cpp (cascade_sum; the kernel is built for 8-lane vectors on AVX2 and AVX-512 hosts
// alike -- checked against torch itself in tests/test_oracle_golden.py) followed by one division by P.
//   P < 8 : four interleaved partial sums over the first 4*floor(P/4) values, the rest added to partial 0,
//           then partial 0 += partial 1, 2, 3;
//   P >= 8: V = floor(P/8) vectors of 8 lanes; lane-wise the same four-partial scheme over the vectors; then
//           a scalar accumulator takes the P - 8V trailing values in order and finally lanes 0..7.
// The cascade levels of the ATen kernel only engage from 64 vectors per partial, i.e. P >= 512 (kPairsMax).
// `g(p)` returns the p-th value; every index is a compile-time constant after unrolling, so the 32
// accumulators live in registers.
template <typename G>
ACAV_HD float pairs_mean(int P, G g) {
    float s;
    if (P < 8) {
        float q0 = 0.f, q1 = 0.f, q2 = 0.f, q3 = 0.f;
        int i = 0;
        if (P >= 4) {
            q0 = q0 + g(0); q1 = q1 + g(1); q2 = q2 + g(2); q3 = q3 + g(3);
            i = 4;
        }
        for (; i < P; ++i) q0 = q0 + g(i);
        q0 = q0 + q1; q0 = q0 + q2; q0 = q0 + q3;
        s = q0;
    } else {
        const int V = P >> 3, Vf = V & ~3;
        float acc[4][8];
ACAV_UNROLL
        for (int k = 0; k < 4; ++k)
ACAV_UNROLL
            for (int l = 0; l < 8; ++l) acc[k][l] = 0.f;
        for (int v = 0; v < Vf; v += 4) {
ACAV_UNROLL
            for (int k = 0; k < 4; ++k)
ACAV_UNROLL
                for (int l = 0; l < 8; ++l) acc[k][l] = acc[k][l] + g((v + k) * 8 + l);
        }
        for (int v = Vf; v < V; ++v) {
ACAV_UNROLL
            for (int l = 0; l < 8; ++l) acc[0][l] = acc[0][l] + g(v * 8 + l);
        }
ACAV_UNROLL
        for (int k = 1; k < 4; ++k)
ACAV_UNROLL
            for (int l = 0; l < 8; ++l) acc[0][l] = acc[0][l] + acc[k][l];
        s = 0.f;
        for (int p = V * 8; p < P; ++p) s = s + g(p);
ACAV_UNROLL
        for (int l = 0; l < 8; ++l) s = s + acc[0][l];
    }
    return s / (float)P;
}

}  // namespace acav
