// Types and tile constants shared by the tcgen05 k-means assignment kernels (kmeans_umma.cu: one CTA per
// tile; kmeans_umma2.cu: CTA pairs, cta_group::2).
#pragma once
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

namespace acav {

constexpr int kBM = 128;                 // rows of X per tile (UMMA M)
constexpr int kBK = 64;                  // bf16 elements per k-block = one 128-byte swizzle row
constexpr int kBNMax = 256;              // centroids per accumulator stage (UMMA N <= 256)
constexpr int kStages = 4;
constexpr int kABytes = kBM * kBK * 2;           // 16 KiB
constexpr int kBBytesMax = kBNMax * kBK * 2;     // 32 KiB
constexpr int kStageBytes = kABytes + kBBytesMax;
constexpr int kTmemCols = 512;
constexpr int kUmmaThreads = 192;
constexpr int kEpiThreads = 128;

struct UmmaSmem {
    // dynamic smem, 1024-byte aligned base:
    //   [kStages][A 16K | B 32K] | cparams[2][256] float4 | barriers
    static constexpr int kParamsOff = kStages * kStageBytes;
    static constexpr int kBarOff = kParamsOff + 2 * kBNMax * 16;
    static constexpr int kBytes = kBarOff + 256;
};

// Epilogue parameters of one centroid.  The kernels rank centroids by a LOWER BOUND of the exact distance
//     L = s*|x|^2 + (a*dot + b) - e*|x|
// where a = -2 s_c, s = s_c (1 - eps32), b = s_c |c|^2 (1 - eps32), s_c in {1, 1/r} (re-init scaling), and
// e*|x| + eps32 s_c (|x|^2 + |c|^2) bounds the error of the bf16 screen against the exact evaluation:
//     |<x,c>_bf16 - <x,c>| <= 1.03 * 2^-7 |x| |c|   (bf16 has 8 significant bits: unit roundoff 2^-8 per
//     operand, 2^-7 per product; Cauchy-Schwarz; 3 % slack for the second-order term and the fp32
//     accumulation over D <= 64 K terms), times 2 s_c for the distance  ->  e = 2.06 * 2^-7 s_c |c| ;
//     eps32 = 2^-19 covers the fp32 roundings of both evaluations of the three-term formula (< 8e-7 relative
//     to |x|^2 + |c|^2 in total).
constexpr float kScreenKappa = 2.06f * 0.0078125f;
constexpr float kScreenEps32 = 1.9073486328125e-6f;
struct __align__(16) CentroidParam {
    float a, b, s, e;
};

// Screening state per row: the four smallest distance LOWER BOUNDS (sorted, earliest index first on ties)
// and the fifth smallest value.  With U = upper bound of centroid i[0] (= d[0] + 2 * its error), the exact
// arg-min is one of the centroids whose lower bound is <= U; if d5 > U those are all in this list.
struct Top4 {
    float d[4];
    int32_t i[4];
    float d5;
};

constexpr int kMaxCand = 16;             // candidates per row the re-check evaluates exactly (over all partial lists)

__device__ __forceinline__ void top4_init(Top4 &t) {
#pragma unroll
    for (int s = 0; s < 4; ++s) { t.d[s] = INFINITY; t.i[s] = 0x7fffffff; }
    t.d5 = INFINITY;
}

// stable insertion (strict <: an equal value stays behind the earlier index)
__device__ __forceinline__ void top4_insert(Top4 &t, float v, int32_t vi) {
#pragma unroll
    for (int s = 0; s < 4; ++s) {
        const bool lt = v < t.d[s];
        const float dv = lt ? t.d[s] : v;
        const int32_t di = lt ? t.i[s] : vi;
        t.d[s] = lt ? v : t.d[s];
        t.i[s] = lt ? vi : t.i[s];
        v = dv; vi = di;
    }
    t.d5 = fminf(t.d5, v);
}

}  // namespace acav
