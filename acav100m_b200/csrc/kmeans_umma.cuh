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

struct __align__(16) CentroidParam {     // dist = s*|x|^2 + (a*dot + b)
    float a, b, s, pad;
};

// Screening state per row: the four smallest approximate distances (sorted, earliest index first on
// ties) and the fifth smallest value.  The exact arg-min is guaranteed to be among the entries within
// the error bound of d[0]; if d5 is outside the bound those are all in this list.
struct Top4 {
    float d[4];
    int32_t i[4];
    float d5;
};

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
