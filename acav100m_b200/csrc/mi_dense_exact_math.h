// Bit-exact arithmetic of the reference's DENSE mutual information (measures/mi.py:85-91), shared between the device
// kernel (mi_dense.cu, nvcc --fmad=false) and a host build for the CPU test-suite (tests/native/mi_dense_exact_host.cpp,
// g++ -ffp-contract=off).  Plain IEEE fp32 operators only.
//
//   scores = (N / n * (N.log() + n.log() - (a.log() + b.log()))).sum([2, 3])            # mi.py:90, [W, P, C, C] -> [W, P]
//
// of the table "cache + one-hot(candidate)".  Two things decide which of several near-equal candidates torch.topk / max
// returns, so both are reproduced exactly:
//   * every elementwise operator rounds to fp32 on its own (five roundings per cell, no contraction);
//   * `.sum([2, 3])` over a contiguous tensor is ATen's cascade_sum over C*C contiguous floats per output
//     (ATen/native/cpu/SumKernel.cpp: 8-lane vectors, 4 interleaved accumulators, cascade levels of 2^level_power
//     steps) -- restated independently in oracle/mi_oracle.c (mi_oracle_aten_row_sum) and checked against torch.
// Table entries are exact integers (counts) or the three "empty" values eps / a0 / b0 of init_cache (mi.py:32-39), so
// log() comes from the table torch's CPU kernel produced (DESIGN.md "MI exactness") and from three host constants.
//
// The sum maps onto one warp: lane = k * L + l owns accumulator stream (k, l) of the kernel's 4 x L partial sums
// (L = 8 lanes, or 1 when C*C < 8) and walks elements e = (4 i + k) L + l, i = 0 .. S-1 -- 32 consecutive floats per
// step across the warp; the partials are then folded in ATen's order (dense_exact_fold_*).
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define ACAV_DX __host__ __device__ __forceinline__
#else
#define ACAV_DX inline
#endif

namespace acav {

struct DenseExactConsts {
    float eps, a0, b0;                 // value of an empty cell / column marginal / row marginal (fp32, as torch sums them)
    float log_eps, log_a0, log_b0;     // torch.log of those
};

struct DenseExactView {                // one clustering pair's running table (counts) and the candidate's cell
    const uint32_t *N, *a, *b;         // [C*C], [C] (index c2), [C] (index c1)
    int32_t C, c1, c2;
    float nf, ln;                      // n + 1 as float and its log
};

// one cell of N / n * (N.log() + n.log() - (a.log() + b.log()))
ACAV_DX float dense_exact_elem(const DenseExactView &v, const DenseExactConsts &k, const float *logs, int64_t e) {
    const int32_t i = (int32_t)(e / v.C), j = (int32_t)(e - (int64_t)i * v.C);
    const uint32_t cn = v.N[e] + ((i == v.c1 && j == v.c2) ? 1u : 0u);
    const uint32_t ca = v.a[j] + (j == v.c2 ? 1u : 0u);
    const uint32_t cb = v.b[i] + (i == v.c1 ? 1u : 0u);
    const float Nf = cn ? (float)cn : k.eps, lN = cn ? logs[cn] : k.log_eps;
    const float la = ca ? logs[ca] : k.log_a0, lb = cb ? logs[cb] : k.log_b0;
    const float t1 = Nf / v.nf;
    const float t4 = lN + v.ln;
    const float t5 = la + lb;
    const float t6 = t4 - t5;
    return t1 * t6;
}

ACAV_DX int dense_exact_ceil_log2(int64_t x) {      // c10::utils::CeilLog2
    if (x <= 2) return 1;
    --x;
    int r = 0;
    while (x > 0) { x >>= 1; ++r; }
    return r;
}

struct DenseExactShape {               // how ATen cuts a row of n_el contiguous floats
    int64_t n_el, V, S;                // elements, vectors of L lanes, steps per accumulator stream (V / 4)
    int32_t L, level_power;
};
ACAV_DX DenseExactShape dense_exact_shape(int64_t n_el) {
    DenseExactShape s;
    s.n_el = n_el;
    s.L = n_el >= 8 ? 8 : 1;
    s.V = n_el / s.L;
    s.S = s.V / 4;
    const int lp = dense_exact_ceil_log2(s.S) / 4;
    s.level_power = lp < 4 ? 4 : lp;
    return s;
}

// multi_row_sum for ONE accumulator stream (k, l): elements e = (4 i + k) L + l, i < S, with the cascade levels
template <typename F>
ACAV_DX float dense_exact_stream(const DenseExactShape &s, int k, int l, F elem) {
    const int64_t T = (int64_t)1 << s.level_power, mask = T - 1;
    float acc0 = 0.f, acc1 = 0.f, acc2 = 0.f, acc3 = 0.f;
    int64_t i = 0;
    while (i + T <= s.S) {
        for (int64_t j = 0; j < T; ++j, ++i) acc0 = acc0 + elem(((4 * i + k) * s.L) + l);
        acc1 = acc1 + acc0; acc0 = 0.f;
        if ((i & (mask << s.level_power)) == 0) {
            acc2 = acc2 + acc1; acc1 = 0.f;
            if ((i & (mask << (2 * s.level_power))) == 0) { acc3 = acc3 + acc2; acc2 = 0.f; }
        }
    }
    for (; i < s.S; ++i) acc0 = acc0 + elem(((4 * i + k) * s.L) + l);
    acc0 = acc0 + acc1;
    acc0 = acc0 + acc2;
    acc0 = acc0 + acc3;
    return acc0;
}

// row_sum's leftover vectors (v = 4 S .. V-1), added to stream k = 0 of lane l in order
template <typename F>
ACAV_DX float dense_exact_leftover(const DenseExactShape &s, int l, float part0, F elem) {
    for (int64_t v = 4 * s.S; v < s.V; ++v) part0 = part0 + elem(v * s.L + l);
    return part0;
}

// Host-side (and reference) evaluation of the whole row: the order the warp version reproduces with shuffles.
template <typename F>
inline float dense_exact_row_sum_serial(int64_t n_el, F elem) {
    const DenseExactShape s = dense_exact_shape(n_el);
    float part[4][8];
    for (int k = 0; k < 4; ++k)
        for (int l = 0; l < s.L; ++l) part[k][l] = dense_exact_stream(s, k, l, elem);
    for (int l = 0; l < s.L; ++l) part[0][l] = dense_exact_leftover(s, l, part[0][l], elem);
    for (int k = 1; k < 4; ++k)
        for (int l = 0; l < s.L; ++l) part[0][l] = part[0][l] + part[k][l];
    if (s.L == 1) return part[0][0];
    float acc = 0.f;
    for (int64_t e = s.V * s.L; e < n_el; ++e) acc = acc + elem(e);
    for (int l = 0; l < s.L; ++l) acc = acc + part[0][l];
    return acc;
}

}  // namespace acav
