// k-means assignment distance-GEMM on CTA PAIRS (tcgen05 cta_group::2), the large-shape variant of
// kmeans_umma.cu.  Same contract: replaces -2*matmul + broadcasts + re-init scaling + min of
// KMeans.calc_best (clustering/code/sgd_clustering.py:72-78) by a screen that keeps the four smallest
// approximate distances per row; near-ties are re-checked exactly afterwards.
//
// Why pairs: a single CTA computing a 128 x 256 tile must pull (128 + 256) x 64 bf16 = 48 KiB through
// L2 -> shared memory for every 512 tensor-pipe cycles (96 B/clk/SM), more than the L2 can deliver to
// 148 SMs at once (~43 B/clk/SM), so the one-CTA kernel tops out near 45 % tensor-pipe activity.  A
// pair shares operands: every CTA keeps 128 rows of X and HALF of the centroid tile, the MMA (issued
// by the leader CTA only, M = 256) reads both halves across the pair.
//   kHalves = 1: pair tile 256 x 256, two TMEM accumulator stages (epilogue of tile t overlaps the MMAs of
//                tile t+1), 32 KiB per CTA per k-block  -> 64 B/clk/SM
//   kHalves = 2: pair tile 256 x 512, both accumulators form ONE tile (X read once for 512 centroids),
//                48 KiB per CTA per 1024 cycles         -> 48 B/clk/SM, epilogue not overlapped
//
// Roles per CTA (320 threads): warp 0 = TMA producer (own 128 rows of X + own half of every centroid
// sub-tile; completion is signalled on the LEADER's full barrier), warp 1 = TMEM allocator (both CTAs)
// and MMA issuer (leader only; tcgen05.commit multicast frees the smem slot / publishes the accumulator
// in both CTAs), warps 2-9 = epilogue: two warps per TMEM lane quarter, each taking half of the tile's
// columns, so a row's top-4 list comes out as two partial lists (merged by km_merge_classify_kernel).
#include <cuda.h>
#include <cuda_bf16.h>

#include "common.cuh"
#include "kernels.cuh"
#include "kmeans_umma.cuh"
#include "sm100_ptx.cuh"

namespace acav {

constexpr int kPairThreads = 320;
constexpr int kPairMaxRegs = 128;                       // 320 x 128 = 40 K registers: leaves room for the blocks of the
                                                        // background preparation kernel on the same SM
constexpr int kPairEpiWarps = 8;
constexpr int kPairEpiThreads = kPairEpiWarps * kWarp;
constexpr int kPairBHalfBytes = 128 * kBK * 2;          // one CTA's half of a 256-centroid sub-tile: 16 KiB

template <int kHalves>
struct PairCfg {
    static constexpr int kTileN = 256 * kHalves;                       // centroids per pair tile
    static constexpr int kAccStages = 2 / kHalves;
    static constexpr int kStageBytes = kABytes + kHalves * kPairBHalfBytes;
    static constexpr int kStages = kHalves == 2 ? 4 : 6;               // 192 KiB of operand ring either way
    static constexpr int kParamsOff = kStages * kStageBytes;
    static constexpr int kBarOff = kParamsOff + 2 * kTileN * 16;       // double-buffered CentroidParam[kTileN]
    static constexpr int kBytes = kBarOff + 256;
    static constexpr int kColsPerWarpSet = kTileN / 2;                 // columns one epilogue warp set scans
};

__device__ __forceinline__ void tmem_ld_wait_dep(uint32_t (&v)[32]) {
    // wait::ld with the destination registers as in/out operands: uses of v cannot be hoisted above it
    asm volatile(
        "tcgen05.wait::ld.sync.aligned;"
        : "+r"(v[0]), "+r"(v[1]), "+r"(v[2]), "+r"(v[3]), "+r"(v[4]), "+r"(v[5]), "+r"(v[6]), "+r"(v[7]),
          "+r"(v[8]), "+r"(v[9]), "+r"(v[10]), "+r"(v[11]), "+r"(v[12]), "+r"(v[13]), "+r"(v[14]), "+r"(v[15]),
          "+r"(v[16]), "+r"(v[17]), "+r"(v[18]), "+r"(v[19]), "+r"(v[20]), "+r"(v[21]), "+r"(v[22]), "+r"(v[23]),
          "+r"(v[24]), "+r"(v[25]), "+r"(v[26]), "+r"(v[27]), "+r"(v[28]), "+r"(v[29]), "+r"(v[30]), "+r"(v[31])
        :
        : "memory");
}

__device__ __forceinline__ void screen_32(Top4 &t4, const uint32_t (&v)[32], const CentroidParam *__restrict__ sp,
                                          float xnr, float nxs, int32_t c_first) {
#pragma unroll
    for (int j = 0; j < 32; ++j) {
        const CentroidParam p = sp[j];
        const float dist = fmaf(p.e, nxs, fmaf(p.s, xnr, fmaf(p.a, __uint_as_float(v[j]), p.b)));   // lower bound
        if (dist < t4.d5) top4_insert(t4, dist, c_first + j);           // rare after the first columns
    }
}

template <int kHalves>
__global__ void __cluster_dims__(2, 1, 1) __maxnreg__(kPairMaxRegs)
km_assign_pair_kernel(const __grid_constant__ CUtensorMap tmap_x, const __grid_constant__ CUtensorMap tmap_c,
                      const float *__restrict__ xn, const CentroidParam *__restrict__ cparams,
                      int32_t b, int32_t k, int32_t num_kb, int32_t n_tiles, int32_t n_split,
                      Top4 *__restrict__ partial) {
    using Cfg = PairCfg<kHalves>;
    constexpr int kSt = Cfg::kStages;
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    CentroidParam *sparams = reinterpret_cast<CentroidParam *>(smem + Cfg::kParamsOff);
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem + Cfg::kBarOff);
    uint64_t *full_bar = bars;                       // [kSt]  used in the leader: bytes of BOTH CTAs land here
    uint64_t *empty_bar = bars + kSt;                // [kSt]  per CTA, signalled by the leader's multicast commit
    uint64_t *tfull_bar = bars + 2 * kSt;            // [2]    per CTA, multicast commit
    uint64_t *tempty_bar = bars + 2 * kSt + 2;       // [2]    used in the leader: epilogue warps of both CTAs arrive
    uint32_t *tmem_ptr = reinterpret_cast<uint32_t *>(bars + 2 * kSt + 4);

    const int warp = threadIdx.x / kWarp, lane = threadIdx.x % kWarp;
    const uint32_t cta_rank = ptx::cluster_ctarank();
    const int32_t cluster_id = blockIdx.x >> 1, num_clusters = gridDim.x >> 1;
    const int32_t num_m = (b + 2 * kBM - 1) / (2 * kBM);
    const int32_t tpg = (n_tiles + n_split - 1) / n_split;        // pair tiles per centroid group
    const int32_t num_units = num_m * n_split;

    if (warp == 0 && lane == 0) {
        ptx::prefetch_tensormap(&tmap_x);
        ptx::prefetch_tensormap(&tmap_c);
        for (int s = 0; s < kSt; ++s) { ptx::mbar_init(&full_bar[s], 1); ptx::mbar_init(&empty_bar[s], 1); }
        for (int s = 0; s < 2; ++s) { ptx::mbar_init(&tfull_bar[s], 1); ptx::mbar_init(&tempty_bar[s], 2 * kPairEpiWarps); }
        ptx::fence_barrier_init();
    }
    if (warp == 1) ptx::tmem_alloc_pair(tmem_ptr, kTmemCols);
    ptx::tc_fence_before();
    ptx::cluster_sync_all();                                     // peer barriers initialised, TMEM allocated
    ptx::tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr;

    if (warp == 0) {
        // ===== TMA producer (both CTAs) =====
        if (lane == 0) {
            uint32_t stage = 0, phase = 0;
            for (int32_t u = cluster_id; u < num_units; u += num_clusters) {
                const int32_t mb = u / n_split, g = u % n_split;
                const int32_t nt_end = min(n_tiles, (g + 1) * tpg);
                const int32_t row0 = mb * 2 * kBM + (int32_t)cta_rank * kBM;
                for (int32_t nt = g * tpg; nt < nt_end; ++nt) {
                    const int32_t crow0 = nt * Cfg::kTileN + (int32_t)cta_rank * 128;
                    for (int32_t kb = 0; kb < num_kb; ++kb) {
                        ptx::mbar_wait(&empty_bar[stage], phase ^ 1);
                        if (cta_rank == 0) ptx::mbar_arrive_expect_tx(&full_bar[stage], 2u * (uint32_t)Cfg::kStageBytes);
                        const uint32_t lead_full = ptx::mapa_u32(ptx::smem_u32(&full_bar[stage]), 0);
                        uint8_t *sa = smem + stage * Cfg::kStageBytes;
                        ptx::tma_load_2d_pair(sa, &tmap_x, kb * kBK, row0, lead_full);
#pragma unroll
                        for (int h = 0; h < kHalves; ++h)
                            ptx::tma_load_2d_pair(sa + kABytes + h * kPairBHalfBytes, &tmap_c, kb * kBK,
                                                  crow0 + h * 256, lead_full);
                        if (++stage == kSt) { stage = 0; phase ^= 1; }
                    }
                }
            }
            // tail: do not leave while the leader's multicast commits may still arrive on my barriers
            for (int s = 0; s < kSt; ++s) {
                ptx::mbar_wait(&empty_bar[stage], phase ^ 1);
                if (++stage == kSt) { stage = 0; phase ^= 1; }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer (leader CTA only) =====
        if (lane == 0 && cta_rank == 0) {
            const uint32_t idesc = ptx::umma_idesc_bf16(2 * kBM, 256);
            uint32_t stage = 0, phase = 0, acc = 0, acc_phase = 0;
            for (int32_t u = cluster_id; u < num_units; u += num_clusters) {
                const int32_t g = u % n_split;
                const int32_t nt_end = min(n_tiles, (g + 1) * tpg);
                for (int32_t nt = g * tpg; nt < nt_end; ++nt) {
                    ptx::mbar_wait(&tempty_bar[acc], acc_phase ^ 1);
                    ptx::tc_fence_after();
                    const uint32_t tmem_d = tmem_base + acc * 256;
                    for (int32_t kb = 0; kb < num_kb; ++kb) {
                        ptx::mbar_wait(&full_bar[stage], phase);
                        ptx::tc_fence_after();
                        const uint32_t sa = ptx::smem_u32(smem + stage * Cfg::kStageBytes);
#pragma unroll
                        for (int k4 = 0; k4 < kBK / 16; ++k4) {
                            const uint64_t adesc = ptx::umma_smem_desc_sw128(sa + k4 * 32);
#pragma unroll
                            for (int h = 0; h < kHalves; ++h) {
                                const uint64_t bdesc =
                                    ptx::umma_smem_desc_sw128(sa + kABytes + h * kPairBHalfBytes + k4 * 32);
                                ptx::umma_bf16_pair(tmem_d + h * 256, adesc, bdesc, idesc, (kb | k4) != 0 ? 1u : 0u);
                            }
                        }
                        ptx::umma_commit_pair(&empty_bar[stage], 3);     // both CTAs may refill this slot
                        if (++stage == kSt) { stage = 0; phase ^= 1; }
                    }
                    ptx::umma_commit_pair(&tfull_bar[acc], 3);           // accumulator ready in both CTAs
                    if (++acc == Cfg::kAccStages) { acc = 0; acc_phase ^= 1; }
                }
            }
        }
    } else {
        // ===== epilogue (warps 2..9) =====
        const int q = warp % 4;                                       // TMEM lane quarter of this warp
        const int wset = (warp - 2) / 4;                              // which half of the tile's columns
        const int et = threadIdx.x - 2 * kWarp;                       // 0..255
        const int32_t row_in_cta = q * 32 + lane;
        const int32_t col0 = wset * Cfg::kColsPerWarpSet;
        uint32_t acc = 0, acc_phase = 0, pbuf = 0;
        for (int32_t u = cluster_id; u < num_units; u += num_clusters) {
            const int32_t mb = u / n_split, g = u % n_split;
            const int32_t nt_end = min(n_tiles, (g + 1) * tpg);
            const int32_t row = mb * 2 * kBM + (int32_t)cta_rank * kBM + row_in_cta;
            const float xnr = row < b ? xn[row] : 0.f;
            const float nxs = -sqrtf(xnr) * 1.000001f;                 // -|x| (rounded away from zero)
            Top4 t4;
            top4_init(t4);
            for (int32_t nt = g * tpg; nt < nt_end; ++nt) {
                CentroidParam *sp = sparams + pbuf * Cfg::kTileN;
                pbuf ^= 1;
                // stage this tile's centroid parameters; the buffer's previous user (two tiles ago) is done:
                // every epilogue thread has passed the barrier of the tile in between after reading it
                for (int32_t i = et; i < Cfg::kTileN; i += kPairEpiThreads) {
                    const int32_t c = nt * Cfg::kTileN + i;
                    CentroidParam p;
                    if (c < k) p = cparams[c];
                    else { p.a = 0.f; p.b = INFINITY; p.s = 0.f; p.e = 0.f; }
                    sp[i] = p;
                }
                asm volatile("bar.sync 1, %0;" ::"n"(kPairEpiThreads) : "memory");
                ptx::mbar_wait(&tfull_bar[acc], acc_phase);
                ptx::tc_fence_after();
                const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + acc * 256 + col0;
                const int32_t cbase = nt * Cfg::kTileN + col0;
                const CentroidParam *spw = sp + col0;
                uint32_t va[32], vb[32];
                ptx::tmem_ld_32x32(taddr, va);
#pragma unroll 1
                for (int32_t c0 = 0; c0 < Cfg::kColsPerWarpSet; c0 += 64) {
                    tmem_ld_wait_dep(va);
                    ptx::tmem_ld_32x32(taddr + c0 + 32, vb);              // in flight while va is screened
                    screen_32(t4, va, spw + c0, xnr, nxs, cbase + c0);
                    tmem_ld_wait_dep(vb);
                    if (c0 + 64 < Cfg::kColsPerWarpSet) ptx::tmem_ld_32x32(taddr + c0 + 64, va);
                    screen_32(t4, vb, spw + c0 + 32, xnr, nxs, cbase + c0 + 32);
                }
                ptx::tc_fence_before();
                __syncwarp();
                if (lane == 0) ptx::mbar_arrive_cluster(ptx::mapa_u32(ptx::smem_u32(&tempty_bar[acc]), 0));
                if (++acc == Cfg::kAccStages) { acc = 0; acc_phase ^= 1; }
            }
            if (row < b) partial[(int64_t)(g * 2 + wset) * b + row] = t4;
        }
    }
    ptx::tc_fence_before();
    ptx::cluster_sync_all();
    if (warp == 1) ptx::tmem_dealloc_pair(tmem_base, kTmemCols);
}

// ---- host side ----------------------------------------------------------------------------------

template <int kHalves>
static int launch_pair_t(const void *tmap_x, const void *tmap_c128, const float *xn, const void *cparams, int32_t b,
                         int32_t k, int32_t dp, int32_t sm_count, void *partial, int32_t *n_lists_out,
                         cudaStream_t st) {
    using Cfg = PairCfg<kHalves>;
    static size_t attr_done[kMaxDevices];
    const int smem_bytes = Cfg::kBytes + 1024;
    { int rc = ensure_dynamic_smem(km_assign_pair_kernel<kHalves>, (size_t)smem_bytes, attr_done); if (rc) return rc; }
    const int32_t n_tiles = (int32_t)ceil_div(k, Cfg::kTileN);
    const int32_t num_m = (int32_t)ceil_div(b, 2 * kBM);
    const int32_t clusters = sm_count / 2;
    int32_t n_split = 1;                                   // every group yields two partial lists per row
    while (n_split < n_tiles && 4 * n_split <= kMaxSplit && num_m * n_split < clusters) n_split *= 2;
    if (n_split > n_tiles) n_split = n_tiles;
    *n_lists_out = 2 * n_split;
    const int32_t units = num_m * n_split;
    const int32_t grid = 2 * (units < clusters ? units : clusters);
    if (grid == 0) return 0;
    km_assign_pair_kernel<kHalves><<<grid, kPairThreads, smem_bytes, st>>>(
        *reinterpret_cast<const CUtensorMap *>(tmap_x), *reinterpret_cast<const CUtensorMap *>(tmap_c128), xn,
        reinterpret_cast<const CentroidParam *>(cparams), b, k, dp / kBK, n_tiles, n_split,
        reinterpret_cast<Top4 *>(partial));
    ACAV_LAUNCH_CHECK();
    return 0;
}

int launch_assign_pair(const void *tmap_x, const void *tmap_c128, const float *xn, const void *cparams, int32_t b,
                       int32_t k, int32_t dp, int32_t sm_count, int32_t halves, void *partial, int32_t *n_lists_out,
                       cudaStream_t st) {
    if (halves == 2)
        return launch_pair_t<2>(tmap_x, tmap_c128, xn, cparams, b, k, dp, sm_count, partial, n_lists_out, st);
    return launch_pair_t<1>(tmap_x, tmap_c128, xn, cparams, b, k, dp, sm_count, partial, n_lists_out, st);
}

}  // namespace acav
