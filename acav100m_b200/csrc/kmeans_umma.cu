// k-means assignment as a tcgen05 distance-GEMM fused with a per-row top-2 (arg)min.
//
// Replaces the hot part of KMeans.calc_best (clustering/code/sgd_clustering.py:72-78): the reference
// materialises dist[k, b] = -2 * centers @ batch.T + |x|^2 + |c|^2 (cuBLAS SGEMM + three elementwise
// passes + a min reduction).  Here one persistent, warp-specialised kernel per SM does
//     TMA (bf16 tiles of X and C, 128B-swizzled)  ->  tcgen05.mma (fp32 accumulators in TMEM)
//     ->  epilogue straight out of TMEM: dist = s_c*|x|^2 + (a_c*acc + b_c), running (d1, i1, d2)
// and never writes the distance matrix.  bf16 operands make the distances approximate, so the kernel
// is a SCREEN: rows whose best/second-best margin is inside the worst-case bf16 error bound are
// re-evaluated by the exact fp32/fp64 kernel (kmeans_exact.cu); all other rows provably have the same
// arg-min as an exact evaluation (DESIGN.md "assignment exactness").
//
// Roles (192 threads): warp 0 = TMA producer, warp 1 = TMEM allocator + MMA issuer (one elected
// lane), warps 2-5 = epilogue (TMEM lane quarter = warp_idx % 4).  Pipelines: smem full/empty ring
// (kStages), TMEM full/empty (2 accumulator stages of 256 columns).
#include <cuda.h>
#include <cuda_bf16.h>

#include "common.cuh"
#include "kernels.cuh"
#include "kmeans_umma.cuh"
#include "sm100_ptx.cuh"

namespace acav {

// fp32 [rows, d] -> bf16 [rows, dp] (zero padded) + |row|^2 as the reference computes it (norm ** 2).
__global__ void km_prep_rows_kernel(const float *__restrict__ x, int64_t rows, int32_t d, int64_t ldx,
                                    int32_t dp, __nv_bfloat16 *__restrict__ xb, float *__restrict__ xn) {
    pdl_begin();
    const int64_t r = (int64_t)blockIdx.x * (blockDim.x / kWarp) + threadIdx.x / kWarp;
    const int lane = threadIdx.x % kWarp;
    if (r >= rows) return;
    const float *p = x + r * ldx;
    __nv_bfloat16 *q = xb + r * dp;
    double s = 0.0;
    const bool vec = (d % 4 == 0) && (ldx % 4 == 0) && ((reinterpret_cast<uintptr_t>(x) & 15) == 0);
    if (vec) {
        for (int32_t i = lane * 4; i < dp; i += kWarp * 4) {
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (i < d) v = __ldg(reinterpret_cast<const float4 *>(p + i));
            s += (double)v.x * v.x + (double)v.y * v.y + (double)v.z * v.z + (double)v.w * v.w;
            __nv_bfloat162 lo = __floats2bfloat162_rn(v.x, v.y), hi = __floats2bfloat162_rn(v.z, v.w);
            uint2 pk;
            pk.x = *reinterpret_cast<uint32_t *>(&lo);
            pk.y = *reinterpret_cast<uint32_t *>(&hi);
            *reinterpret_cast<uint2 *>(q + i) = pk;
        }
    } else {
        for (int32_t i = lane; i < dp; i += kWarp) {
            float v = i < d ? __ldg(p + i) : 0.f;
            s += (double)v * v;
            q[i] = __float2bfloat16_rn(v);
        }
    }
    s = warp_sum_f64(s);
    if (lane == 0) {
        float nrm = sqrtf((float)s);
        xn[r] = __fmul_rn(nrm, nrm);
    }
}

// Per-centroid epilogue parameters (see CentroidParam): scale s_c = 1/r for under-used centroids
// (sgd_clustering.py:76-77), else 1, and the per-centroid error term of the bf16 screen.
__global__ void km_centroid_params_kernel(const float *__restrict__ cn, const float *__restrict__ counts, int32_t k,
                                          float thr, float r, CentroidParam *__restrict__ params) {
    const int32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= k) return;
    const float s = counts[i] < thr ? 1.0f / r : 1.0f;
    CentroidParam p;
    p.a = -2.0f * s;
    p.b = cn[i] * s * (1.0f - kScreenEps32);
    p.s = s * (1.0f - kScreenEps32);
    p.e = kScreenKappa * s * sqrtf(cn[i]) * 1.000001f;
    params[i] = p;
}

__global__ void __launch_bounds__(kUmmaThreads, 1)
km_assign_umma_kernel(const __grid_constant__ CUtensorMap tmap_x, const __grid_constant__ CUtensorMap tmap_c,
                      const float *__restrict__ xn, const CentroidParam *__restrict__ cparams,
                      int32_t b, int32_t k, int32_t num_kb, int32_t bn, int32_t n_tiles, int32_t n_split,
                      Top4 *__restrict__ partial) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    CentroidParam *sparams = reinterpret_cast<CentroidParam *>(smem + UmmaSmem::kParamsOff);
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem + UmmaSmem::kBarOff);
    uint64_t *full_bar = bars;                       // [kStages]
    uint64_t *empty_bar = bars + kStages;            // [kStages]
    uint64_t *tfull_bar = bars + 2 * kStages;        // [2]
    uint64_t *tempty_bar = bars + 2 * kStages + 2;   // [2]
    uint32_t *tmem_ptr = reinterpret_cast<uint32_t *>(bars + 2 * kStages + 4);

    const int warp = threadIdx.x / kWarp, lane = threadIdx.x % kWarp;
    const int32_t num_m = (b + kBM - 1) / kBM;
    const int32_t tpg = (n_tiles + n_split - 1) / n_split;        // n-tiles per group
    const int32_t num_units = num_m * n_split;

    if (warp == 0 && lane == 0) {
        ptx::prefetch_tensormap(&tmap_x);
        ptx::prefetch_tensormap(&tmap_c);
        for (int s = 0; s < kStages; ++s) { ptx::mbar_init(&full_bar[s], 1); ptx::mbar_init(&empty_bar[s], 1); }
        for (int s = 0; s < 2; ++s) { ptx::mbar_init(&tfull_bar[s], 1); ptx::mbar_init(&tempty_bar[s], kEpiThreads); }
        ptx::fence_barrier_init();
    }
    if (warp == 1) ptx::tmem_alloc(tmem_ptr, kTmemCols);
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr;

    if (warp == 0) {
        // ===== TMA producer =====
        if (lane == 0) {
            uint32_t stage = 0, phase = 0;
            const uint32_t tx_bytes = (uint32_t)kABytes + (uint32_t)bn * kBK * 2;
            for (int32_t u = blockIdx.x; u < num_units; u += gridDim.x) {
                const int32_t mb = u / n_split, g = u % n_split;
                const int32_t nt_end = min(n_tiles, (g + 1) * tpg);
                for (int32_t nt = g * tpg; nt < nt_end; ++nt) {
                    for (int32_t kb = 0; kb < num_kb; ++kb) {
                        ptx::mbar_wait(&empty_bar[stage], phase ^ 1);
                        ptx::mbar_arrive_expect_tx(&full_bar[stage], tx_bytes);
                        uint8_t *sa = smem + stage * kStageBytes;
                        ptx::tma_load_2d(sa, &tmap_x, kb * kBK, mb * kBM, &full_bar[stage]);
                        ptx::tma_load_2d(sa + kABytes, &tmap_c, kb * kBK, nt * kBNMax, &full_bar[stage]);
                        if (++stage == kStages) { stage = 0; phase ^= 1; }
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer =====
        if (lane == 0) {
            const uint32_t idesc = ptx::umma_idesc_bf16(kBM, (uint32_t)bn);
            uint32_t stage = 0, phase = 0, acc = 0, acc_phase = 0;
            for (int32_t u = blockIdx.x; u < num_units; u += gridDim.x) {
                const int32_t g = u % n_split;
                const int32_t nt_end = min(n_tiles, (g + 1) * tpg);
                for (int32_t nt = g * tpg; nt < nt_end; ++nt) {
                    ptx::mbar_wait(&tempty_bar[acc], acc_phase ^ 1);
                    ptx::tc_fence_after();
                    const uint32_t tmem_d = tmem_base + acc * kBNMax;
                    for (int32_t kb = 0; kb < num_kb; ++kb) {
                        ptx::mbar_wait(&full_bar[stage], phase);
                        ptx::tc_fence_after();
                        const uint32_t sa = ptx::smem_u32(smem + stage * kStageBytes);
                        const uint32_t sb = sa + kABytes;
#pragma unroll
                        for (int k4 = 0; k4 < kBK / 16; ++k4) {
                            const uint64_t adesc = ptx::umma_smem_desc_sw128(sa + k4 * 32);
                            const uint64_t bdesc = ptx::umma_smem_desc_sw128(sb + k4 * 32);
                            ptx::umma_bf16(tmem_d, adesc, bdesc, idesc, (kb | k4) != 0 ? 1u : 0u);
                        }
                        ptx::umma_commit(&empty_bar[stage]);          // smem slot free once these MMAs retire
                        if (++stage == kStages) { stage = 0; phase ^= 1; }
                    }
                    ptx::umma_commit(&tfull_bar[acc]);                // accumulator ready for the epilogue
                    if (++acc == 2) { acc = 0; acc_phase ^= 1; }
                }
            }
        }
    } else {
        // ===== epilogue (warps 2..5) =====
        const int q = warp % 4;                                       // TMEM lane quarter of this warp
        const int et = threadIdx.x - 2 * kWarp;                       // 0..127
        const int32_t row_in_tile = q * 32 + lane;
        uint32_t acc = 0, acc_phase = 0;
        for (int32_t u = blockIdx.x; u < num_units; u += gridDim.x) {
            const int32_t mb = u / n_split, g = u % n_split;
            const int32_t nt_end = min(n_tiles, (g + 1) * tpg);
            const int32_t row = mb * kBM + row_in_tile;
            const float xnr = row < b ? xn[row] : 0.f;
            const float nxs = -sqrtf(xnr) * 1.000001f;                 // -|x| (rounded away from zero)
            Top4 t4;
            top4_init(t4);
            for (int32_t nt = g * tpg; nt < nt_end; ++nt) {
                CentroidParam *sp = sparams + acc * kBNMax;
                // stage this tile's centroid parameters (the buffer's previous user, two tiles ago, is done:
                // every epilogue thread passed the barrier below after reading it)
                for (int32_t i = et; i < kBNMax; i += kEpiThreads) {
                    const int32_t c = nt * kBNMax + i;
                    CentroidParam p;
                    if (c < k) p = cparams[c];
                    else { p.a = 0.f; p.b = INFINITY; p.s = 0.f; p.e = 0.f; }
                    sp[i] = p;
                }
                asm volatile("bar.sync 1, %0;" ::"n"(kEpiThreads) : "memory");
                ptx::mbar_wait(&tfull_bar[acc], acc_phase);
                ptx::tc_fence_after();
                const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + acc * kBNMax;
                for (int32_t c0 = 0; c0 < bn; c0 += 32) {
                    uint32_t v[32];
                    ptx::tmem_ld_32x32(taddr + c0, v);
                    ptx::tmem_ld_wait();
#pragma unroll
                    for (int j = 0; j < 32; ++j) {
                        const CentroidParam p = sp[c0 + j];
                        const float dist = fmaf(p.e, nxs, fmaf(p.s, xnr, fmaf(p.a, __uint_as_float(v[j]), p.b)));
                        if (dist < t4.d5) top4_insert(t4, dist, nt * kBNMax + c0 + j);   // rare after the first columns
                    }
                }
                ptx::tc_fence_before();
                ptx::mbar_arrive(&tempty_bar[acc]);
                if (++acc == 2) { acc = 0; acc_phase ^= 1; }
            }
            if (row < b) partial[(int64_t)g * b + row] = t4;
        }
    }
    ptx::tc_fence_before();
    __syncthreads();
    if (warp == 1) ptx::tmem_dealloc(tmem_base, kTmemCols);
}

// Classify every row from its partial top-4 lists (one list per centroid group / epilogue warp set) and route
// near-ties.  The lists rank centroids by the lower bound L_c of their exact distance (see CentroidParam).
// With c0 = the centroid of the smallest L over all lists:
//   U = L_c0 + 2 * err_c0(row)  is an upper bound of the exact distance of c0, hence of the exact minimum;
//   every centroid with L <= U is a possible winner.  A partial list holds ALL of its group's possible winners
//   iff its fifth-smallest value d5 is > U;
//   only c0 has L <= U                         -> c0 is the arg-min for certain;
//   all lists complete, <= kMaxCand candidates -> queue the row for the candidate re-check (exact dot products
//                                                 of exactly those centroids);
//   otherwise                                  -> queue for the full exact kernel.
__global__ void km_merge_classify_kernel(const Top4 *__restrict__ partial, int32_t b, int32_t n_split,
                                         const float *__restrict__ xn, const CentroidParam *__restrict__ cparams,
                                         const float *__restrict__ cn, int32_t k,
                                         int64_t *__restrict__ best, float *__restrict__ mind,
                                         int32_t *__restrict__ cand_rows, int32_t *__restrict__ cand_ids,
                                         int32_t *__restrict__ full_rows, int32_t *__restrict__ counters) {
    pdl_begin();
    const int32_t row = blockIdx.x * blockDim.x + threadIdx.x;
    if (row >= b) return;
    float d0 = INFINITY, d5min = INFINITY;
    int32_t i0 = 0x7fffffff;
    for (int32_t g = 0; g < n_split; ++g) {
        const Top4 &o = partial[(int64_t)g * b + row];
        const float od = o.d[0];
        const int32_t oi = o.i[0];
        if (od < d0 || (od == d0 && oi < i0)) { d0 = od; i0 = oi; }
        d5min = fminf(d5min, o.d5);
    }
    best[row] = i0;
    if (mind) mind[row] = d0;
    if (i0 < 0 || i0 >= k) {                                   // no finite distance at all (NaN / inf input)
        full_rows[atomicAdd(&counters[1], 1)] = row;
        return;
    }
    const CentroidParam p0 = cparams[i0];
    const float xnr = xn[row];
    const float s0 = p0.a * -0.5f;
    const float err0 = p0.e * sqrtf(xnr) * 1.000001f + kScreenEps32 * s0 * (xnr + cn[i0]);
    const float upper = d0 + 2.0f * err0 * 1.000001f + 1e-30f;
    int32_t n_cand = 0;
    for (int32_t g = 0; g < n_split; ++g) {
        const Top4 &o = partial[(int64_t)g * b + row];
#pragma unroll
        for (int s = 0; s < 4; ++s) n_cand += (o.d[s] <= upper) ? 1 : 0;
    }
    if (n_cand <= 1) return;
    if (d5min > upper && n_cand <= kMaxCand) {
        const int32_t slot = atomicAdd(&counters[0], 1);
        cand_rows[slot] = row;
        int32_t w = 0;
        for (int32_t g = 0; g < n_split; ++g) {
            const Top4 &o = partial[(int64_t)g * b + row];
#pragma unroll
            for (int s = 0; s < 4; ++s)
                if (o.d[s] <= upper) cand_ids[kMaxCand * slot + w++] = o.i[s];
        }
        for (; w < kMaxCand; ++w) cand_ids[kMaxCand * slot + w] = -1;
    } else {
        full_rows[atomicAdd(&counters[1], 1)] = row;
    }
}

// Candidate re-check: kMaxCand threads per row, one per candidate, each evaluates the exact distance with the
// SAME operation order as assign_exact_kernel (sequential fp64 FMA over k, one rounding to fp32, then the
// reference's three fp32 operations), so both kernels agree bit for bit; lowest index wins ties.
__global__ void __launch_bounds__(128)
km_candidate_refine_kernel(const float *__restrict__ x, int64_t ldx, int32_t d,
                           const float *__restrict__ centers, const float *__restrict__ xn,
                           const float *__restrict__ cn, const float *__restrict__ counts, float thr, float r,
                           const int32_t *__restrict__ cand_rows, const int32_t *__restrict__ cand_ids,
                           const int32_t *__restrict__ counters, int32_t capacity,
                           int64_t *__restrict__ best, float *__restrict__ mind) {
    const int32_t n = min(counters[0], capacity);
    const int32_t slot = (blockIdx.x * blockDim.x + threadIdx.x) / kMaxCand;
    const int sub = threadIdx.x % kMaxCand;
    const bool live = slot < n;
    float dist = INFINITY;
    int32_t c = 0x7fffffff;
    int32_t row = 0;
    if (live) {
        row = cand_rows[slot];
        const int32_t cid = cand_ids[kMaxCand * slot + sub];
        if (cid >= 0) {
            c = cid;
            const float *p = x + (int64_t)row * ldx, *q = centers + (int64_t)c * d;
            double acc = 0.0;
            if ((d % 4 == 0) && (ldx % 4 == 0) && ((reinterpret_cast<uintptr_t>(x) & 15) == 0) &&
                ((reinterpret_cast<uintptr_t>(centers) & 15) == 0)) {
                // the chain of d fp64 FMAs in k order is the definition of "exact" (same order as assign_exact_kernel)
                // and stays serial; the loads are double-buffered in registers, 32 elements (16 x 16 bytes) per
                // buffer, so the L2 latency of batch i+1 hides behind the FMA chain of batch i
                float4 a0[8], w0[8], a1[8], w1[8];
                auto ld = [&](float4 (&av)[8], float4 (&wv)[8], int32_t at) {
#pragma unroll
                    for (int u = 0; u < 8; ++u) {
                        av[u] = __ldg(reinterpret_cast<const float4 *>(p + at) + u);
                        wv[u] = __ldg(reinterpret_cast<const float4 *>(q + at) + u);
                    }
                };
                auto fm = [&](const float4 (&av)[8], const float4 (&wv)[8]) {
#pragma unroll
                    for (int u = 0; u < 8; ++u) {
                        acc = fma((double)av[u].x, (double)wv[u].x, acc);
                        acc = fma((double)av[u].y, (double)wv[u].y, acc);
                        acc = fma((double)av[u].z, (double)wv[u].z, acc);
                        acc = fma((double)av[u].w, (double)wv[u].w, acc);
                    }
                };
                const int32_t nb = d / 32;
                int32_t i = nb * 32;
                if (nb > 0) {
                    ld(a0, w0, 0);
                    int32_t bi = 0;
                    while (true) {
                        if (bi + 1 < nb) ld(a1, w1, (bi + 1) * 32);
                        fm(a0, w0);
                        if (++bi >= nb) break;
                        if (bi + 1 < nb) ld(a0, w0, (bi + 1) * 32);
                        fm(a1, w1);
                        if (++bi >= nb) break;
                    }
                }
                for (; i < d; i += 4) {
                    const float4 a = __ldg(reinterpret_cast<const float4 *>(p + i));
                    const float4 w = __ldg(reinterpret_cast<const float4 *>(q + i));
                    acc = fma((double)a.x, (double)w.x, acc);
                    acc = fma((double)a.y, (double)w.y, acc);
                    acc = fma((double)a.z, (double)w.z, acc);
                    acc = fma((double)a.w, (double)w.w, acc);
                }
            } else {
                for (int32_t i = 0; i < d; ++i) acc = fma((double)__ldg(p + i), (double)__ldg(q + i), acc);
            }
            dist = __fadd_rn(__fadd_rn(__fmul_rn(-2.f, (float)acc), xn[row]), cn[c]);
            if (counts[c] < thr) dist = __fdiv_rn(dist, r);
        }
    }
#pragma unroll
    for (int o = kMaxCand / 2; o > 0; o >>= 1) {
        const float od = __shfl_xor_sync(0xffffffffu, dist, o);
        const int32_t oc = __shfl_xor_sync(0xffffffffu, c, o);
        if (od < dist || (od == dist && oc < c)) { dist = od; c = oc; }
    }
    if (live && sub == 0) {
        best[row] = c;
        if (mind) mind[row] = dist;
    }
}

// exact distance of every row to its assigned centroid (for the returned mean distance)
__global__ void km_exact_min_dist_kernel(const float *__restrict__ x, int64_t b, int32_t d, int64_t ldx,
                                         const float *__restrict__ centers, const int64_t *__restrict__ best,
                                         const float *__restrict__ xn, const float *__restrict__ cn,
                                         const float *__restrict__ counts, float thr, float r,
                                         float *__restrict__ mind) {
    const int64_t row = (int64_t)blockIdx.x * (blockDim.x / kWarp) + threadIdx.x / kWarp;
    const int lane = threadIdx.x % kWarp;
    if (row >= b) return;
    const int64_t c = best[row];
    const float *p = x + row * ldx, *q = centers + c * d;
    double s = 0.0;
    for (int32_t i = lane; i < d; i += kWarp) s = fma((double)__ldg(p + i), (double)__ldg(q + i), s);
    s = warp_sum_f64(s);
    if (lane == 0) {
        float dist = __fadd_rn(__fadd_rn(__fmul_rn(-2.f, (float)s), xn[row]), cn[c]);
        if (counts[c] < thr) dist = __fdiv_rn(dist, r);
        mind[row] = dist;
    }
}

// ---- host side ----------------------------------------------------------------------------------

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_tiled_fn() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    }
    return fn;
}

// bf16 [rows, dp] row-major, box = 64 elements x box_rows, 128-byte swizzle, OOB reads as zero.
int make_bf16_tensor_map(void *out_map, const void *base, int64_t rows, int32_t dp, int32_t box_rows) {
    EncodeTiledFn fn = encode_tiled_fn();
    if (!fn) return ACAV_E_NO_DEVICE;
    cuuint64_t dims[2] = {(cuuint64_t)dp, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)dp * 2};
    cuuint32_t box[2] = {(cuuint32_t)kBK, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = fn(reinterpret_cast<CUtensorMap *>(out_map), CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2,
                    const_cast<void *>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? 0 : ACAV_E_INVALID;
}

int umma_tile_n(int32_t k) { return k >= kBNMax ? kBNMax : (int)ceil_div(k, 16) * 16; }

int launch_prep_rows(const float *x, int64_t rows, int32_t d, int64_t ldx, int32_t dp, void *xb, float *xn,
                     cudaStream_t st) {
    if (rows == 0) return 0;
    const int wpb = 8;
    ACAV_CUDA_TRY(launch_pdl(km_prep_rows_kernel, dim3((unsigned)ceil_div(rows, wpb)), dim3(wpb * kWarp), 0, st,
                             x, rows, d, ldx, dp, reinterpret_cast<__nv_bfloat16 *>(xb), xn));
    return 0;
}

int launch_centroid_params(const float *cn, const float *counts, int32_t k, float thr, float r, void *params,
                           cudaStream_t st) {
    km_centroid_params_kernel<<<(unsigned)ceil_div(k, 256), 256, 0, st>>>(cn, counts, k, thr, r,
                                                                         reinterpret_cast<CentroidParam *>(params));
    ACAV_LAUNCH_CHECK();
    return 0;
}

int launch_assign_umma(const void *tmap_x, const void *tmap_c, const float *xn, const void *cparams, int32_t b,
                       int32_t k, int32_t dp, int32_t sm_count, void *partial, int32_t *n_split_out, cudaStream_t st) {
    static size_t attr_done[kMaxDevices];
    const int smem_bytes = UmmaSmem::kBytes + 1024;
    { int rc = ensure_dynamic_smem(km_assign_umma_kernel, (size_t)smem_bytes, attr_done); if (rc) return rc; }
    const int32_t bn = umma_tile_n(k);
    const int32_t n_tiles = (int32_t)ceil_div(k, kBNMax);
    const int32_t num_m = (int32_t)ceil_div(b, kBM);
    int32_t n_split = 1;
    while (n_split < n_tiles && n_split < kMaxSplit && num_m * n_split < sm_count) n_split *= 2;
    if (n_split > n_tiles) n_split = n_tiles;
    *n_split_out = n_split;
    const int32_t units = num_m * n_split;
    const int32_t grid = units < sm_count ? units : sm_count;
    if (grid == 0) return 0;
    km_assign_umma_kernel<<<grid, kUmmaThreads, smem_bytes, st>>>(
        *reinterpret_cast<const CUtensorMap *>(tmap_x), *reinterpret_cast<const CUtensorMap *>(tmap_c), xn,
        reinterpret_cast<const CentroidParam *>(cparams), b, k, dp / kBK, bn, n_tiles, n_split,
        reinterpret_cast<Top4 *>(partial));
    ACAV_LAUNCH_CHECK();
    return 0;
}

int launch_merge_classify(const void *partial, int32_t b, int32_t n_split, const float *xn, const void *cparams,
                          const float *cn, int32_t k, int64_t *best, float *mind, int32_t *cand_rows, int32_t *cand_ids, int32_t *full_rows,
                          int32_t *counters, cudaStream_t st) {
    // counters[0..1] were zeroed by the caller BEFORE the distance kernel (a memset between the two kernels would
    // make this launch an ordinary one)
    if (b == 0) return 0;
    ACAV_CUDA_TRY(launch_pdl(km_merge_classify_kernel, dim3((unsigned)ceil_div(b, 256)), dim3(256), 0, st,
                             reinterpret_cast<const Top4 *>(partial), b, n_split, xn,
                             reinterpret_cast<const CentroidParam *>(cparams), cn, k, best, mind, cand_rows, cand_ids,
                             full_rows, counters));
    return 0;
}

// grid covers the worst case (every row queued); blocks past the device-side count exit at once
int launch_candidate_refine(const float *x, int64_t ldx, int32_t d, const float *centers, const float *xn,
                            const float *cn, const float *counts, float thr, float r, const int32_t *cand_rows,
                            const int32_t *cand_ids, const int32_t *counters, int32_t b, int64_t *best, float *mind,
                            cudaStream_t st) {
    if (b == 0) return 0;
    km_candidate_refine_kernel<<<(unsigned)ceil_div((int64_t)b * kMaxCand, 128), 128, 0, st>>>(
        x, ldx, d, centers, xn, cn, counts, thr, r, cand_rows, cand_ids, counters, b, best, mind);
    ACAV_LAUNCH_CHECK();
    return 0;
}

int launch_exact_min_dist(const float *x, int64_t b, int32_t d, int64_t ldx, const float *centers,
                          const int64_t *best, const float *xn, const float *cn, const float *counts, float thr,
                          float r, float *mind, cudaStream_t st) {
    if (b == 0) return 0;
    const int wpb = 8;
    km_exact_min_dist_kernel<<<(unsigned)ceil_div(b, wpb), wpb * kWarp, 0, st>>>(x, b, d, ldx, centers, best, xn,
                                                                                cn, counts, thr, r, mind);
    ACAV_LAUNCH_CHECK();
    return 0;
}

int64_t umma_partial_bytes(int64_t max_batch) { return (int64_t)sizeof(Top4) * kMaxSplit * max_batch; }
int64_t umma_param_bytes(int32_t k) { return (int64_t)sizeof(CentroidParam) * k; }

}  // namespace acav
