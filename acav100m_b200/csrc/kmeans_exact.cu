// Exact k-means assignment on CUDA cores (fp32 inputs, fp64-accumulated dot products).
//
// Replaces KMeans.calc_best of the reference (clustering/code/sgd_clustering.py:63-79).  This is the
// canonical evaluation of the reference's distance formula
//     dist[i,j] = (-2*<c_i,x_j> + |x_j|^2) + |c_i|^2            (:72-74, three fp32 roundings)
//     dist[i,:] /= r   where counts[i] < threshold               (:76-77)
//     best[j] = first argmin_i dist[i,j]                         (:78)
// with the inner product rounded ONCE from an fp64 accumulation (what an ideal sgemm returns).  It is
// (a) the whole assignment in ACAV_ASSIGN_EXACT mode, (b) the re-check of rows the tensor-core kernel
// marks ambiguous in ACAV_ASSIGN_TENSOR mode, (c) the only path for tiny batches.
#include "common.cuh"
#include "kernels.cuh"

namespace acav {

// |row|^2 the way the reference gets it: torch.norm(.., dim=1) ** 2, i.e. sqrt then square (:73-74).
__global__ void row_norm2_kernel(const float *__restrict__ x, int64_t rows, int32_t d, int64_t ldx,
                                 const int32_t *__restrict__ rowlist, float *__restrict__ out) {
    int64_t r = (int64_t)blockIdx.x * (blockDim.x / kWarp) + threadIdx.x / kWarp;
    int lane = threadIdx.x % kWarp;
    if (r >= rows) return;
    int64_t src = rowlist ? (int64_t)rowlist[r] : r;
    const float *p = x + src * ldx;
    double s = 0.0;
    for (int32_t i = lane; i < d; i += kWarp) {
        double v = (double)__ldg(p + i);
        s += v * v;
    }
    s = warp_sum_f64(s);
    if (lane == 0) {
        float nrm = sqrtf((float)s);
        out[r] = __fmul_rn(nrm, nrm);
    }
}

constexpr int kTR = 64;      // rows per block tile
constexpr int kTC = 64;      // centroids per inner tile
constexpr int kKC = 16;      // reduction chunk
constexpr int kPad = 66;     // smem row stride (doubles), keeps double2 reads aligned

struct BestPair {
    float d;
    int32_t i;
};

__device__ __forceinline__ bool better(float d, int32_t i, float bd, int32_t bi) {
    return d < bd || (d == bd && i < bi);
}

// One block = 64 rows (optionally gathered through rowlist) against all k centroids.
__global__ void __launch_bounds__(256)
assign_exact_kernel(const float *__restrict__ x, int64_t ldx, const int32_t *__restrict__ rowlist,
                    int64_t nrows, const int32_t *__restrict__ nrows_dev,
                    const float *__restrict__ centers, int32_t k, int32_t d,
                    const float *__restrict__ xn, const float *__restrict__ cn,
                    const float *__restrict__ counts, float thr, float r,
                    unsigned long long *__restrict__ packed, unsigned int *__restrict__ tickets,
                    int64_t *__restrict__ best, float *__restrict__ mind) {
    pdl_begin();
    // tiles are widened to fp64 once, on the way into shared memory (F2F.F64 is a slow pipe: doing it
    // per FMA operand made the kernel conversion-bound)
    __shared__ __align__(16) double Xs[kKC][kPad];
    __shared__ __align__(16) double Cs[kKC][kPad];
    const int tid = threadIdx.x;
    const int tx = tid % 16, ty = tid / 16;
    const int64_t row0 = (int64_t)blockIdx.x * kTR;
    if (nrows_dev) nrows = min(nrows, (int64_t)*nrows_dev);      // list length lives on the device
    if (row0 >= nrows) return;

    // loader mapping: 64 rows x 16 k per tile, 4 scalars per thread
    const int lrow = tid / 4, lk = (tid % 4) * 4;
    int64_t xsrc = -1;
    if (row0 + lrow < nrows) xsrc = rowlist ? (int64_t)rowlist[row0 + lrow] : row0 + lrow;

    float bd[4];
    int32_t bi[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) { bd[i] = INFINITY; bi[i] = 0x7fffffff; }

    float xnr[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        int64_t rr = row0 + ty * 4 + i;
        xnr[i] = rr < nrows ? xn[rowlist ? (int64_t)rowlist[rr] : rr] : 0.f;     // xn is per source row
    }

    // blockIdx.y selects a contiguous group of centroid tiles (more CTAs in flight than row blocks alone)
    const int32_t tiles = (k + kTC - 1) / kTC;
    const int32_t per = (tiles + (int32_t)gridDim.y - 1) / (int32_t)gridDim.y;
    const int32_t c_begin = (int32_t)blockIdx.y * per * kTC, c_end = min(k, c_begin + per * kTC);
    for (int32_t c0 = c_begin; c0 < c_end; c0 += kTC) {
        double acc[4][4];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) acc[i][j] = 0.0;
        const int32_t csrc = c0 + lrow;
        for (int32_t k0 = 0; k0 < d; k0 += kKC) {
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                int32_t kk = k0 + lk + q;
                float xv = 0.f, cv = 0.f;
                if (kk < d) {
                    if (xsrc >= 0) xv = __ldg(x + xsrc * ldx + kk);
                    if (csrc < k) cv = __ldg(centers + (int64_t)csrc * d + kk);
                }
                Xs[lk + q][lrow] = (double)xv;
                Cs[lk + q][lrow] = (double)cv;
            }
            __syncthreads();
#pragma unroll
            for (int kk = 0; kk < kKC; ++kk) {
                const double2 xa = *reinterpret_cast<const double2 *>(&Xs[kk][ty * 4]);
                const double2 xb = *reinterpret_cast<const double2 *>(&Xs[kk][ty * 4 + 2]);
                const double2 ca = *reinterpret_cast<const double2 *>(&Cs[kk][tx * 4]);
                const double2 cb = *reinterpret_cast<const double2 *>(&Cs[kk][tx * 4 + 2]);
                const double xd[4] = {xa.x, xa.y, xb.x, xb.y};
                const double cd[4] = {ca.x, ca.y, cb.x, cb.y};
#pragma unroll
                for (int i = 0; i < 4; ++i)
#pragma unroll
                    for (int j = 0; j < 4; ++j) acc[i][j] = fma(xd[i], cd[j], acc[i][j]);
            }
            __syncthreads();
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            int32_t c = c0 + tx * 4 + j;
            if (c >= k) continue;
            float cnc = __ldg(cn + c);
            bool under = __ldg(counts + c) < thr;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                float dotf = (float)acc[i][j];
                float dist = __fadd_rn(__fadd_rn(__fmul_rn(-2.f, dotf), xnr[i]), cnc);
                if (under) dist = __fdiv_rn(dist, r);
                if (better(dist, c, bd[i], bi[i])) { bd[i] = dist; bi[i] = c; }
            }
        }
    }
    // combine the 16 tx-threads of each row (16 consecutive lanes)
#pragma unroll
    for (int i = 0; i < 4; ++i) {
#pragma unroll
        for (int o = 8; o > 0; o >>= 1) {
            float od = __shfl_xor_sync(0xffffffffu, bd[i], o);
            int32_t oi = __shfl_xor_sync(0xffffffffu, bi[i], o);
            if (better(od, oi, bd[i], bi[i])) { bd[i] = od; bi[i] = oi; }
        }
        int64_t rr = row0 + ty * 4 + i;
        if (tx == 0 && rr < nrows && bi[i] != 0x7fffffff) {
            // smallest distance, then smallest index: unsigned order of (orderable(dist) << 32 | index)
            const unsigned long long key = ((unsigned long long)orderable(bd[i]) << 32) | (uint32_t)bi[i];
            atomicMin(packed + rr, key);
        }
    }
    if (tickets) {
        // one launch instead of init + main + finish: the LAST centroid group to finish a row block writes its rows out
        // and leaves packed[] / the ticket in their idle state (~0 / 0) for the next launch
        __shared__ bool last;
        __threadfence();
        __syncthreads();
        if (tid == 0) last = atomicAdd(&tickets[blockIdx.x], 1u) == gridDim.y - 1;
        __syncthreads();
        if (last) {
            __threadfence();
            for (int r2 = tid; r2 < kTR; r2 += 256) {
                const int64_t rr = row0 + r2;
                if (rr >= nrows) break;
                const unsigned long long key = __ldcg(packed + rr);
                packed[rr] = ~0ull;
                const int64_t dst = rowlist ? (int64_t)rowlist[rr] : rr;
                best[dst] = (int64_t)(key & 0xFFFFFFFFull);
                if (mind) mind[dst] = from_orderable((uint32_t)(key >> 32));
            }
            if (tid == 0) tickets[blockIdx.x] = 0u;
        }
    }
}

__global__ void assign_exact_init_kernel(unsigned long long *__restrict__ packed, int64_t n) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) packed[i] = ~0ull;
}

__global__ void assign_exact_finish_kernel(unsigned long long *__restrict__ packed,
                                           const int32_t *__restrict__ rowlist, int64_t nrows,
                                           const int32_t *__restrict__ nrows_dev,
                                           int64_t *__restrict__ best, float *__restrict__ mind) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (nrows_dev) nrows = min(nrows, (int64_t)*nrows_dev);
    if (i >= nrows) return;
    const unsigned long long key = packed[i];
    packed[i] = ~0ull;                                   // idle state for the single-launch variant (tickets != null)
    const int64_t dst = rowlist ? (int64_t)rowlist[i] : i;
    best[dst] = (int64_t)(key & 0xFFFFFFFFull);
    if (mind) mind[dst] = from_orderable((uint32_t)(key >> 32));
}

// Warm-up branch (:67-68,78): column-wise first argmin of noise[k, b].
__global__ void assign_noise_kernel(const float *__restrict__ noise, int32_t k, int64_t b,
                                    int64_t *__restrict__ best, float *__restrict__ mind) {
    int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= b) return;
    float bd = INFINITY;
    int32_t bi = 0;
    for (int32_t i = 0; i < k; ++i) {
        float v = __ldg(noise + (int64_t)i * b + j);
        if (v < bd) { bd = v; bi = i; }
    }
    best[j] = bi;
    if (mind) mind[j] = bd;
}

// Deterministic mean of n floats: fixed strided partials in fp64, fixed-order block reduction.
__global__ void __launch_bounds__(1024) mean_kernel(const float *__restrict__ v, int64_t n,
                                                   float *__restrict__ out) {
    __shared__ double part[32];
    double s = 0.0;
    for (int64_t i = threadIdx.x; i < n; i += blockDim.x) s += (double)v[i];
    s = warp_sum_f64(s);
    if (threadIdx.x % kWarp == 0) part[threadIdx.x / kWarp] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int w = 0; w < (int)(blockDim.x / kWarp); ++w) t += part[w];
        out[0] = n > 0 ? (float)(t / (double)n) : nanf("");
    }
}

int launch_row_norm2(const float *x, int64_t rows, int32_t d, int64_t ldx, const int32_t *rowlist,
                     float *out, cudaStream_t st) {
    if (rows == 0) return 0;
    const int wpb = 8;
    row_norm2_kernel<<<(unsigned)ceil_div(rows, wpb), wpb * kWarp, 0, st>>>(x, rows, d, ldx, rowlist, out);
    ACAV_LAUNCH_CHECK();
    return 0;
}

int launch_assign_exact(const float *x, int64_t ldx, const int32_t *rowlist, int64_t nrows,
                        const int32_t *nrows_dev,
                        const float *centers, int32_t k, int32_t d, const float *xn, const float *cn,
                        const float *counts, float thr, float r, int64_t *best, float *mind,
                        unsigned long long *packed, int32_t sm_count, cudaStream_t st, unsigned int *tickets) {
    if (nrows == 0) return 0;
    const unsigned row_blocks = (unsigned)ceil_div(nrows, kTR);
    const int32_t tiles = (int32_t)ceil_div(k, kTC);
    int32_t split = 1;                                   // aim for >= 4 CTAs per SM
    while (split < tiles && (int64_t)row_blocks * split < 4ll * sm_count) split *= 2;
    if (split > tiles) split = tiles;
    if (tickets) {                                       // packed[] idles at ~0, tickets at 0 (acav_kmeans_create)
        ACAV_CUDA_TRY(launch_pdl(assign_exact_kernel, dim3(row_blocks, (unsigned)split), dim3(256), 0, st,
                                 x, ldx, rowlist, nrows, nrows_dev, centers, k, d, xn, cn, counts, thr, r, packed, tickets, best, mind));
        return 0;
    }
    assign_exact_init_kernel<<<(unsigned)ceil_div(nrows, 256), 256, 0, st>>>(packed, nrows);
    ACAV_LAUNCH_CHECK();
    assign_exact_kernel<<<dim3(row_blocks, (unsigned)split), 256, 0, st>>>(
        x, ldx, rowlist, nrows, nrows_dev, centers, k, d, xn, cn, counts, thr, r, packed, nullptr, nullptr, nullptr);
    ACAV_LAUNCH_CHECK();
    assign_exact_finish_kernel<<<(unsigned)ceil_div(nrows, 256), 256, 0, st>>>(packed, rowlist, nrows, nrows_dev,
                                                                              best, mind);
    ACAV_LAUNCH_CHECK();
    return 0;
}

int launch_assign_noise(const float *noise, int32_t k, int64_t b, int64_t *best, float *mind,
                        cudaStream_t st) {
    if (b == 0) return 0;
    assign_noise_kernel<<<(unsigned)ceil_div(b, 256), 256, 0, st>>>(noise, k, b, best, mind);
    ACAV_LAUNCH_CHECK();
    return 0;
}

int launch_mean(const float *v, int64_t n, float *out, cudaStream_t st) {
    mean_kernel<<<1, 1024, 0, st>>>(v, n, out);
    ACAV_LAUNCH_CHECK();
    return 0;
}

}  // namespace acav
