// Centroid update of the mini-batch SGD k-means (the "fast parallel update" branch of KMeans.add,
// clustering/code/sgd_clustering.py:113-127).
//
// The reference accumulates  deltas[best[j]] += fl32(x_j * lr)  with torch-scatter's scatter_add;
// its CPU kernel walks the rows in order, so every centroid's sum is a strict row-order fp32 chain.
// To be bit-identical we (1) partition the batch rows by centroid with a STABLE counting sort and
// (2) let one thread own one (centroid, 4 columns) slot and add its member rows sequentially.  Reads
// are coalesced across columns; the only serial dependence is the fp32 add chain itself.
#include <cstdlib>

#include "common.cuh"
#include "kernels.cuh"
#include "sm100_ptx.cuh"

namespace acav {

constexpr int kRankRows = 512;   // rows per block in the rank / scatter kernels

// Stable local rank of every row among the rows of the same centroid inside its 512-row block, and
// the block's histogram.
// <false>: warps take turns in row order over one shared histogram (k counters of shared memory): sixteen
// __match_any_sync in a row -- each keeps its scheduler's MIO queue busy for ~600 cycles when the 32 ids differ.
// <true> (k up to kRankParMaxK): every warp matches its own 32 rows at once and leaves its counts in its own uint16
// histogram; a row's rank is its rank inside the warp plus the earlier warps' counts of its centroid.
constexpr int kRankParMaxK = 6144;                 // 16 warps x k x 2 bytes <= 192 KiB
template <bool kPar>
__global__ void __launch_bounds__(kRankRows)
km_block_rank_kernel(const int64_t *__restrict__ best, int64_t b, int32_t k,
                     uint32_t *__restrict__ blockhist, uint32_t *__restrict__ lrank) {
    pdl_begin();
    extern __shared__ __align__(16) uint32_t hist[];
    const int64_t row = (int64_t)blockIdx.x * kRankRows + threadIdx.x;
    const int warp = threadIdx.x / kWarp, lane = threadIdx.x % kWarp;
    uint32_t *dst = blockhist + (int64_t)blockIdx.x * k;
    if constexpr (kPar) {
        constexpr int kW = kRankRows / kWarp;
        uint16_t *whist = reinterpret_cast<uint16_t *>(hist);          // [kW][k]
        for (int32_t i = threadIdx.x; i < kW * k / 2; i += blockDim.x) hist[i] = 0u;
        int64_t key = -1;
        if (row < b) {
            key = best[row];
            if (key < 0 || key >= k) key = -1;      // out-of-range ids are ignored (never produced)
        }
        __syncthreads();
        const unsigned m = __match_any_sync(0xffffffffu, key);
        uint32_t rank = __popc(m & ((1u << lane) - 1u));
        if (key >= 0 && rank == 0) whist[warp * k + key] = (uint16_t)__popc(m);
        __syncthreads();
        if (key >= 0) {
            for (int w = 0; w < warp; ++w) rank += whist[w * k + key];
            lrank[row] = rank;
        }
        for (int32_t i = threadIdx.x; i < k; i += blockDim.x) {
            uint32_t t = 0;
#pragma unroll
            for (int w = 0; w < kW; ++w) t += whist[w * k + i];
            dst[i] = t;
        }
        return;
    }
    for (int32_t i = threadIdx.x; i < k; i += blockDim.x) hist[i] = 0;
    __syncthreads();
    int64_t key = -1;
    if (row < b) {
        key = best[row];
        if (key < 0 || key >= k) key = -1;          // out-of-range ids are ignored (never produced)
    }
    for (int w = 0; w < kRankRows / kWarp; ++w) {
        if (warp == w) {
            unsigned m = __match_any_sync(0xffffffffu, key);
            int leader = __ffs(m) - 1;
            uint32_t rank = __popc(m & ((1u << lane) - 1u));
            uint32_t base = 0;
            if (key >= 0 && lane == leader) {
                base = hist[key];
                hist[key] = base + __popc(m);
            }
            base = __shfl_sync(0xffffffffu, base, leader);
            if (key >= 0) lrank[row] = base + rank;
        }
        __syncthreads();
    }
    for (int32_t i = threadIdx.x; i < k; i += blockDim.x) dst[i] = hist[i];
}

// Per centroid: exclusive prefix over blocks (in place) and the batch histogram (fp32, :113).
__global__ void km_block_prefix_kernel(uint32_t *__restrict__ blockhist, int32_t nblk, int32_t k,
                                       uint32_t *__restrict__ total, float *__restrict__ counts_b) {
    pdl_begin();
    int32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= k) return;
    uint32_t run = 0;
    for (int32_t blk = 0; blk < nblk; ++blk) {
        uint32_t t = blockhist[(int64_t)blk * k + i];
        blockhist[(int64_t)blk * k + i] = run;
        run += t;
    }
    total[i] = run;
    counts_b[i] = (float)run;
}

constexpr uint32_t kUpdHeavyRows = 128;              // centroids with at least this many batch rows take the ring kernel

// seg_start[0..k] = exclusive scan of total[0..k) (single block, chunks of 1024 with carry); also the list of
// "heavy" centroids (>= kUpdHeavyRows rows of this batch): seg_start[k+1] = their number, seg_start[k+2..] = ids.
// With `blockhist` given (small batches: nblk <= kFusedPrefixBlocks) the kernel also does km_block_prefix_kernel's work for
// its centroid first -- one launch less on the training step's chain.
constexpr int32_t kFusedPrefixBlocks = 64;
__global__ void __launch_bounds__(1024)
km_segment_start_kernel(uint32_t *__restrict__ total, int32_t k, uint32_t *__restrict__ seg_start,
                        uint32_t *__restrict__ blockhist, int32_t nblk, float *__restrict__ counts_b,
                        float *__restrict__ hist_max) {
    __shared__ uint32_t warp_sums[32];
    __shared__ uint32_t warp_max[32];
    __shared__ uint32_t carry;
    pdl_begin();
    __shared__ uint32_t n_heavy;
    if (threadIdx.x == 0) { carry = 0; n_heavy = 0; }
    __syncthreads();
    const int lane = threadIdx.x % kWarp, warp = threadIdx.x / kWarp;
    uint32_t vmax = 0u;                            // this thread's largest count (for the learning-rate decision, :116)
    for (int32_t base = 0; base < k; base += 1024) {
        int32_t i = base + threadIdx.x;
        uint32_t v = 0u;
        if (i < k) {
            if (blockhist) {                       // exclusive prefix over blocks (in place) + batch histogram (:113)
                uint32_t run = 0;
                for (int32_t blk = 0; blk < nblk; ++blk) {
                    const uint32_t t = blockhist[(int64_t)blk * k + i];
                    blockhist[(int64_t)blk * k + i] = run;
                    run += t;
                }
                total[i] = run;
                counts_b[i] = (float)run;
                v = run;
            } else {
                v = total[i];
            }
        }
        vmax = max(vmax, v);
        uint32_t inc = v;
#pragma unroll
        for (int o = 1; o < kWarp; o <<= 1) {
            uint32_t t = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= o) inc += t;
        }
        if (lane == kWarp - 1) warp_sums[warp] = inc;
        __syncthreads();
        if (warp == 0) {
            uint32_t ws = warp_sums[lane];
            uint32_t winc = ws;
#pragma unroll
            for (int o = 1; o < kWarp; o <<= 1) {
                uint32_t t = __shfl_up_sync(0xffffffffu, winc, o);
                if (lane >= o) winc += t;
            }
            warp_sums[lane] = winc - ws;           // exclusive warp offsets
        }
        __syncthreads();
        uint32_t excl = carry + warp_sums[warp] + inc - v;
        if (i < k) seg_start[i] = excl;
        if (i < k && v >= kUpdHeavyRows) seg_start[k + 2 + atomicAdd(&n_heavy, 1u)] = (uint32_t)i;
        __syncthreads();
        if (threadIdx.x == 1023) carry = excl + v;
        __syncthreads();
    }
    if (threadIdx.x == 0) { seg_start[k] = carry; seg_start[k + 1] = n_heavy; }
    if (hist_max) {                                // max_c counts_b[c] as fp32: what km_effective_lr_kernel would reduce
        vmax = __reduce_max_sync(0xffffffffu, vmax);
        if (lane == 0) warp_max[warp] = vmax;
        __syncthreads();
        if (warp == 0) {
            vmax = __reduce_max_sync(0xffffffffu, warp_max[lane]);
            if (lane == 0) *hist_max = (float)vmax;
        }
    }
}

__global__ void __launch_bounds__(kRankRows)
km_scatter_rows_kernel(const int64_t *__restrict__ best, int64_t b, int32_t k,
                       const uint32_t *__restrict__ blockhist, const uint32_t *__restrict__ lrank,
                       const uint32_t *__restrict__ seg_start, uint32_t *__restrict__ sorted_rows) {
    pdl_begin();
    const int64_t row = (int64_t)blockIdx.x * kRankRows + threadIdx.x;
    if (row >= b) return;
    int64_t key = best[row];
    if (key < 0 || key >= k) return;
    uint32_t dst = seg_start[key] + blockhist[(int64_t)blockIdx.x * k + key] + lrank[row];
    sorted_rows[dst] = (uint32_t)row;
}

// lr fallback of sgd_clustering.py:116-119 evaluated on the device in the same python-float (double)
// arithmetic:  if max(counts_b) * lr >= 1.0: lr = 0.5 / max(counts_b); fallback += 1.
// `lr * counts` in the reference is an fp32 tensor times a python scalar: torch casts the scalar to
// fp32 first, so lr_eff is stored as float.
__global__ void __launch_bounds__(1024)
km_effective_lr_kernel(const float *__restrict__ counts_b, int32_t k, double lr,
                       float *__restrict__ lr_eff, int32_t *__restrict__ fallback) {
    pdl_begin();
    __shared__ float wmax[32];
    float m = 0.f;
    for (int32_t i = threadIdx.x; i < k; i += blockDim.x) m = fmaxf(m, counts_b[i]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    if (threadIdx.x % kWarp == 0) wmax[threadIdx.x / kWarp] = m;
    __syncthreads();
    if (threadIdx.x == 0) {
        float mm = 0.f;
        for (int w = 0; w < (int)(blockDim.x / kWarp); ++w) mm = fmaxf(mm, wmax[w]);
        double eff = lr;
        if ((double)mm * lr >= 1.0) {
            eff = 0.5 / (double)mm;
            if (fallback) *fallback += 1;
        }
        *lr_eff = (float)eff;
    }
}

// The step's learning rate: decided by an earlier kernel (lr_in < 0: *lr_eff_p), or -- one launch less on the step's
// chain -- here, from the batch histogram's maximum that km_segment_start_kernel left in lr_eff_p[2], with the arithmetic
// of km_effective_lr_kernel (:116-119).
__device__ __forceinline__ float step_lr(const float *__restrict__ lr_eff_p, double lr_in, bool *fell_back) {
    *fell_back = false;
    if (lr_in < 0.0) return *lr_eff_p;
    const float mm = lr_eff_p[2];
    double eff = lr_in;
    if ((double)mm * lr_in >= 1.0) { eff = 0.5 / (double)mm; *fell_back = true; }
    return (float)eff;
}

// Where a (centroid, columns) delta goes:
//   kUpdFused : centers = centers*decay + delta (:121,:127), one process;
//   kUpdSplit : centers *= decay and the local delta is written out for an all-reduce (:125-126);
//   kUpdPush  : the delta is stored straight into the receive buffer of the rank that OWNS the centroid (peer memory
//               over NVLink); centers and counts are not touched here -- the owner adds the ranks' deltas in rank
//               order, applies the decay and hands the new rows to everybody (kmeans_comm.cu).
//   kUpdSeq   : the reference's `sequential=True` branch (:103-109): the rows of a centroid are applied one after the
//               other, c = fl(fl(c * w) + fl(lr * x)) with w = fl32(1 - lr); lr_eff_p[1] holds w.  Per centroid this is
//               the same strict row-order chain, started from the centroid instead of from zero.
enum { kUpdFused = 0, kUpdSplit = 1, kUpdPush = 2, kUpdSeq = 3 };

// One thread = one centroid x VEC columns.
// The fp32 add chain per (centroid, column) is inherently serial (that IS the reference's sum order);
// everything around it is parallel: row indices are staged through shared memory 256 at a time and 16
// row loads are kept in flight per thread, so a heavily skewed batch is bound by the 4-cycle add chain
// rather than by memory latency.
template <int VEC, int MODE>
__global__ void __launch_bounds__(128, VEC == 4 ? 8 : 4)
km_update_kernel(const float *__restrict__ x, int64_t ldx, int32_t d,
                 const uint32_t *__restrict__ seg_start, const uint32_t *__restrict__ sorted_rows,
                 const float *__restrict__ counts_b, const float *__restrict__ lr_eff_p,
                 float *__restrict__ centers, float *__restrict__ counts, float *__restrict__ deltas, KmPush push,
                 double lr_in, int32_t *__restrict__ fallback) {
    pdl_begin();
    // rows per register buffer: 2 x 4 x 16 bytes in flight per thread and <= 64 registers, so that eight blocks share
    // an SM -- the batch's average centroid owns b / K = 8 rows, and what bounds the kernel then is how many blocks'
    // dependent chains (segment -> row ids -> rows -> centroid) run side by side, not the depth of one chain
    constexpr int kChunk = 256, kGroup = VEC == 4 ? 4 : 32;
    __shared__ uint32_t sidx[kChunk];
    const int32_t c = blockIdx.x;
    bool fell_back;
    const float lr = step_lr(lr_eff_p, lr_in, &fell_back);
    if (fell_back && fallback && blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 0) *fallback += 1;
    if (VEC == 4 && seg_start[c + 1] - seg_start[c] >= kUpdHeavyRows) return;   // heavy centroid: km_update_stream_kernel's
    const int32_t col = (blockIdx.y * blockDim.x + threadIdx.x) * VEC;
    const bool active = col < d;
    const float cb = counts_b[c];
    if (MODE != kUpdPush && blockIdx.y == 0 && threadIdx.x == 0) counts[c] = __fadd_rn(counts[c], cb);          // :120
    const uint32_t lo = seg_start[c], hi = seg_start[c + 1];
    float acc[VEC];
#pragma unroll
    for (int v = 0; v < VEC; ++v) acc[v] = 0.f;
    float wseq = 0.f;
    if constexpr (MODE == kUpdSeq) {
        wseq = lr_eff_p[1];
        if (active) {
#pragma unroll
            for (int v = 0; v < VEC; ++v) acc[v] = centers[(int64_t)c * d + col + v];
        }
    }
    float cold[VEC];                                        // the centroid's old values: asked for before the rows
    if constexpr (MODE == kUpdFused || MODE == kUpdSplit) {
        if (active) {
            if constexpr (VEC == 4) {
                const float4 t = *reinterpret_cast<const float4 *>(centers + (int64_t)c * d + col);
                cold[0] = t.x; cold[1] = t.y; cold[2] = t.z; cold[3] = t.w;
            } else {
                cold[0] = centers[(int64_t)c * d + col];
            }
        }
    }
    const float *xcol = x + col;
    float buf[2][kGroup][VEC];                              // double buffer: loads of group g+1 fly during adds of g
    auto load_group = [&](float (&dst)[kGroup][VEC], uint32_t s, uint32_t n) {
#pragma unroll
        for (int u = 0; u < kGroup; ++u) {
            if (s + u < n) {
                const float *p = xcol + (int64_t)sidx[s + u] * ldx;
                if constexpr (VEC == 4) {
                    const float4 t = __ldg(reinterpret_cast<const float4 *>(p));
                    dst[u][0] = t.x; dst[u][1] = t.y; dst[u][2] = t.z; dst[u][3] = t.w;
                } else {
                    dst[u][0] = __ldg(p);
                }
            }
        }
    };
    auto add_group = [&](const float (&src)[kGroup][VEC], uint32_t s, uint32_t n) {
#pragma unroll
        for (int u = 0; u < kGroup; ++u) {
            if (s + u < n) {
#pragma unroll
                for (int v = 0; v < VEC; ++v)
                    acc[v] = MODE == kUpdSeq ? __fadd_rn(__fmul_rn(acc[v], wseq), __fmul_rn(src[u][v], lr))   // :107-108
                                             : __fadd_rn(acc[v], __fmul_rn(src[u][v], lr));                   // :123
            }
        }
    };
    for (uint32_t chunk = lo; chunk < hi; chunk += kChunk) {
        const uint32_t n = min((uint32_t)kChunk, hi - chunk);
        __syncthreads();
        for (uint32_t t = threadIdx.x; t < n; t += blockDim.x) sidx[t] = sorted_rows[chunk + t];
        __syncthreads();
        if (!active) continue;
        load_group(buf[0], 0, n);
        for (uint32_t s = 0; s < n; s += 2 * kGroup) {
            if (s + kGroup < n) load_group(buf[1], s + kGroup, n);
            add_group(buf[0], s, n);
            if (s + 2 * kGroup < n) load_group(buf[0], s + 2 * kGroup, n);
            if (s + kGroup < n) add_group(buf[1], s + kGroup, n);
        }
    }
    if (!active) return;
    if constexpr (MODE == kUpdPush) {
        static_assert(VEC == 4, "the push path needs 16-byte columns");
        *reinterpret_cast<float4 *>(push.slot(c, d) + col) = make_float4(acc[0], acc[1], acc[2], acc[3]);
        return;
    }
    if constexpr (MODE == kUpdSeq) {
#pragma unroll
        for (int v = 0; v < VEC; ++v) centers[(int64_t)c * d + col + v] = acc[v];
        return;
    }
    const float decay = __fsub_rn(1.f, __fmul_rn(cb, lr));                                    // :121
    float *cp = centers + (int64_t)c * d + col;
    float out[VEC];
#pragma unroll
    for (int v = 0; v < VEC; ++v) {
        const float scaled = __fmul_rn(cold[v], decay);
        out[v] = MODE == kUpdFused ? __fadd_rn(scaled, acc[v]) : scaled;                       // :127
    }
    if constexpr (VEC == 4) {
        *reinterpret_cast<float4 *>(cp) = make_float4(out[0], out[1], out[2], out[3]);
        if (MODE != kUpdFused)
            *reinterpret_cast<float4 *>(deltas + (int64_t)c * d + col) = make_float4(acc[0], acc[1], acc[2], acc[3]);
    } else {
        cp[0] = out[0];
        if (MODE != kUpdFused) deltas[(int64_t)c * d + col] = acc[0];
    }
}

// Same result as km_update_kernel<4, FUSED>, for centroids that own MANY rows of the batch: the per-(centroid,
// column) add chain is serial, so what bounds a skewed batch is how many row loads a thread keeps in flight.
// Here every thread streams its 16 bytes of each row through its own shared-memory ring with cp.async --
// kUpdGroups x kUpdGroupRows = 128 rows in flight per thread instead of the 16 that fit in registers -- and
// consumes the groups in order (wait_group), so the fp32 sum order is unchanged.  One warp x 4 columns per
// block (16 blocks per centroid at D = 2048, so a handful of heavy centroids still spreads over dozens of SMs),
// 64 KiB of ring.  Centroids below kUpdHeavyRows stay with km_update_kernel (blocks of the other class exit).
constexpr int kUpdThreads = 32;                      // one warp x 4 columns = 128 columns per block
constexpr int kUpdGroupRows = 8;
constexpr int kUpdGroups = 16;                       // 128 rows x 512 B = 64 KiB in flight per block
constexpr int kUpdChunk = 1024;                      // row indices staged per pass

template <int MODE>
__global__ void __launch_bounds__(kUpdThreads)
km_update_stream_kernel(const float *__restrict__ x, int64_t ldx, int32_t k, int32_t d,
                        const uint32_t *__restrict__ seg_start, const uint32_t *__restrict__ sorted_rows,
                        const float *__restrict__ counts_b, const float *__restrict__ lr_eff_p,
                        float *__restrict__ centers, float *__restrict__ counts, float *__restrict__ deltas, KmPush push) {
    extern __shared__ __align__(16) unsigned char usmem[];
    float4 *ring = reinterpret_cast<float4 *>(usmem);                                   // [groups][rows][threads]
    uint32_t *sidx = reinterpret_cast<uint32_t *>(usmem + (size_t)kUpdGroups * kUpdGroupRows * kUpdThreads * 16);
    const int32_t col = (blockIdx.y * kUpdThreads + threadIdx.x) * 4;
    const bool active = col < d;
    const float lr = *lr_eff_p;
    const uint32_t n_heavy = seg_start[k + 1];
    for (uint32_t hidx = blockIdx.x; hidx < n_heavy; hidx += gridDim.x) {
    const int32_t c = (int32_t)seg_start[k + 2 + hidx];
    const float cb = counts_b[c];
    if (MODE != kUpdPush && blockIdx.y == 0 && threadIdx.x == 0) counts[c] = __fadd_rn(counts[c], cb);          // :120
    const uint32_t lo = seg_start[c], hi = seg_start[c + 1];
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    const float *xcol = x + (active ? col : 0);
    const uint32_t my = (uint32_t)__cvta_generic_to_shared(ring + threadIdx.x);
    // cp.async of one group of rows into ring slot `slot`: the row indices are read first (independent shared-memory
    // loads), then the copies are issued back to back -- a load-then-issue pair per row serialised the single warp
    auto issue_group = [&](int slot, uint32_t s, uint32_t n) {
        uint32_t idx[kUpdGroupRows];
#pragma unroll
        for (int u = 0; u < kUpdGroupRows; ++u) idx[u] = s + u < n ? sidx[s + u] : 0u;
#pragma unroll
        for (int u = 0; u < kUpdGroupRows; ++u) {
            if (s + u < n) {
                const float *p = xcol + (int64_t)idx[u] * ldx;
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(my + (uint32_t)((slot * kUpdGroupRows + u) * kUpdThreads * 16)),
                             "l"(p));
            }
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
    auto load_group = [&](float4 (&dst)[kUpdGroupRows], int slot) {
#pragma unroll
        for (int u = 0; u < kUpdGroupRows; ++u) dst[u] = ring[(slot * kUpdGroupRows + u) * kUpdThreads + threadIdx.x];
    };
    for (uint32_t chunk = lo; chunk < hi; chunk += kUpdChunk) {
        const uint32_t n = min((uint32_t)kUpdChunk, hi - chunk);
        __syncwarp();
        for (uint32_t t = threadIdx.x; t < n; t += kUpdThreads) sidx[t] = sorted_rows[chunk + t];
        __syncwarp();
#pragma unroll
        for (int g = 0; g < kUpdGroups - 1; ++g) issue_group(g, (uint32_t)g * kUpdGroupRows, n);
        asm volatile("cp.async.wait_group %0;" ::"n"(kUpdGroups - 2) : "memory");          // group 0 has landed
        float4 cur[kUpdGroupRows], nxt[kUpdGroupRows];
        load_group(cur, 0);
        int slot = 0;
        // software pipeline: while the fp32 add chain of group g runs (the serial part: the reference's sum order),
        // the rows of group g+1 are already on their way from shared memory to registers and group g+15 is being
        // fetched into the slot group g-1 left
        for (uint32_t s = 0; s < n; s += kUpdGroupRows) {
            int pre = slot + kUpdGroups - 1;
            if (pre >= kUpdGroups) pre -= kUpdGroups;
            issue_group(pre, s + (kUpdGroups - 1) * kUpdGroupRows, n);
            asm volatile("cp.async.wait_group %0;" ::"n"(kUpdGroups - 2) : "memory");      // the next group has landed
            int nslot = slot + 1;
            if (nslot == kUpdGroups) nslot = 0;
            load_group(nxt, nslot);
#pragma unroll
            for (int u = 0; u < kUpdGroupRows; ++u) {
                if (s + u < n) {
                    acc[0] = __fadd_rn(acc[0], __fmul_rn(cur[u].x, lr));                       // :123
                    acc[1] = __fadd_rn(acc[1], __fmul_rn(cur[u].y, lr));
                    acc[2] = __fadd_rn(acc[2], __fmul_rn(cur[u].z, lr));
                    acc[3] = __fadd_rn(acc[3], __fmul_rn(cur[u].w, lr));
                }
            }
#pragma unroll
            for (int u = 0; u < kUpdGroupRows; ++u) cur[u] = nxt[u];
            slot = nslot;
        }
        asm volatile("cp.async.wait_group 0;" ::: "memory");
    }
    if (active && MODE == kUpdPush) {
        *reinterpret_cast<float4 *>(push.slot(c, d) + col) = make_float4(acc[0], acc[1], acc[2], acc[3]);
    } else if (active) {
        const float decay = __fsub_rn(1.f, __fmul_rn(cb, lr));                                // :121
        float *cp = centers + (int64_t)c * d + col;
#pragma unroll
        for (int v = 0; v < 4; ++v) {
            float scaled = __fmul_rn(cp[v], decay);
            if (MODE == kUpdFused) {
                cp[v] = __fadd_rn(scaled, acc[v]);                                             // :127
            } else {
                cp[v] = scaled;
                deltas[(int64_t)c * d + col + v] = acc[v];
            }
        }
    }
    }   // heavy centroids of this block
}

// Heavy centroids, TMA version (default): one CTA = (heavy centroid, 64 columns), five warps.  Warps 0-3 are
// producers: per stage of 32 rows each of them issues eight bulk async copies (cp.async.bulk, 256 bytes of one row each,
// one per lane 0..7; the instruction takes uniform operands, so a warp issues its lanes one after the other -- four
// warps issue four at a time) completing on the stage's mbarrier; the row indices of the NEXT stage are loaded before
// the copies of this one are issued.  The copy engine keeps kBulkStages x 32 rows in flight per CTA without occupying
// load/store slots (the cp.async ring above is limited by the outstanding-miss capacity of the SM: ~24 rows in flight
// measured).  Warp 4 is the consumer: lane l owns columns 2l, 2l+1 and adds its 8 bytes of every row in stage order --
// the same strict row-order fp32 chain.  32 CTAs per heavy centroid at D = 2048, so one centroid owning a third of
// the batch is still pulled by 32 SMs.
constexpr int kBulkCols = 64;                        // columns per CTA (256 bytes per row)
constexpr int kBulkRows = 32;                        // rows per stage
constexpr int kBulkStages = 8;                       // 256 rows x 256 B = 64 KiB in flight per CTA
constexpr int kBulkProducers = 4;                    // producer warps, kBulkRows / kBulkProducers rows each per stage
constexpr int kBulkThreads = (kBulkProducers + 1) * kWarp;

template <int MODE>
__global__ void __launch_bounds__(kBulkThreads)
km_update_bulk_kernel(const float *__restrict__ x, int64_t ldx, int32_t k, int32_t d,
                      const uint32_t *__restrict__ seg_start, const uint32_t *__restrict__ sorted_rows,
                      const float *__restrict__ counts_b, const float *__restrict__ lr_eff_p,
                      float *__restrict__ centers, float *__restrict__ counts, float *__restrict__ deltas, KmPush push,
                      double lr_in) {
    extern __shared__ __align__(128) unsigned char bsmem[];
    float *ring = reinterpret_cast<float *>(bsmem);                                     // [stages][rows][64]
    uint64_t *full = reinterpret_cast<uint64_t *>(bsmem + (size_t)kBulkStages * kBulkRows * kBulkCols * 4);
    uint64_t *empty = full + kBulkStages;
    const int warp = threadIdx.x / kWarp, lane = threadIdx.x % kWarp;
    const int32_t col0 = blockIdx.y * kBulkCols;
    const uint32_t row_bytes = (uint32_t)min(kBulkCols, d - col0) * 4u;                 // multiple of 16 (d % 4 == 0)
    bool fell_back;
    const float lr = step_lr(lr_eff_p, lr_in, &fell_back);
    const uint32_t n_heavy = seg_start[k + 1];
    if (threadIdx.x == 0) {
        for (int s = 0; s < kBulkStages; ++s) { ptx::mbar_init(&full[s], kBulkProducers); ptx::mbar_init(&empty[s], 1); }
        ptx::fence_barrier_init();
    }
    __syncthreads();
    constexpr int kPer = kBulkRows / kBulkProducers;  // rows of a stage issued by one producer warp
    uint32_t stage = 0, phase = 0;                    // every warp walks the stages in the same order, across centroids
    for (uint32_t hidx = blockIdx.x; hidx < n_heavy; hidx += gridDim.x) {
        const int32_t c = (int32_t)seg_start[k + 2 + hidx];
        const uint32_t lo = seg_start[c], hi = seg_start[c + 1];
        if (warp < kBulkProducers) {
            // ===== producers =====
            const uint32_t mine = (uint32_t)(warp * kPer + lane);                          // my row of a stage (lanes 0..7)
            uint32_t row = (lane < kPer && lo + mine < hi) ? __ldg(sorted_rows + lo + mine) : 0u;
            for (uint32_t base = lo; base < hi; base += kBulkRows) {
                const uint32_t n = min((uint32_t)kBulkRows, hi - base);
                const uint32_t nxt = base + kBulkRows + mine;
                const uint32_t row_next = (lane < kPer && nxt < hi) ? __ldg(sorted_rows + nxt) : 0u;
                const uint32_t first = (uint32_t)(warp * kPer);
                const uint32_t my_n = n > first ? min((uint32_t)kPer, n - first) : 0u;      // rows this warp issues
                ptx::mbar_wait(&empty[stage], phase ^ 1u);                              // slot free (first lap passes)
                if (lane == 0) ptx::mbar_arrive_expect_tx(&full[stage], my_n * row_bytes);
                __syncwarp();
                if ((uint32_t)lane < my_n) {
                    const float *src = x + (int64_t)row * ldx + col0;
                    float *dst = ring + ((size_t)stage * kBulkRows + mine) * kBulkCols;
                    asm volatile(
                        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                        ::"r"(ptx::smem_u32(dst)), "l"(src), "r"(row_bytes), "r"(ptx::smem_u32(&full[stage])) : "memory");
                }
                row = row_next;
                if (++stage == kBulkStages) { stage = 0; phase ^= 1u; }
            }
        } else {
            // ===== consumer =====
            const int32_t col = col0 + 2 * lane;
            const bool active = col < d;
            float acc0 = 0.f, acc1 = 0.f;
            float wseq = 0.f;
            if constexpr (MODE == kUpdSeq) {
                wseq = lr_eff_p[1];
                if (active) { acc0 = centers[(int64_t)c * d + col]; acc1 = centers[(int64_t)c * d + col + 1]; }
            }
            for (uint32_t base = lo; base < hi; base += kBulkRows) {
                const uint32_t n = min((uint32_t)kBulkRows, hi - base);
                ptx::mbar_wait(&full[stage], phase);
                const float2 *rows = reinterpret_cast<const float2 *>(ring + (size_t)stage * kBulkRows * kBulkCols) + lane;
                if (n == kBulkRows) {
                    float2 v[kBulkRows];
#pragma unroll
                    for (int r = 0; r < kBulkRows; ++r) v[r] = rows[r * (kBulkCols / 2)];
#pragma unroll
                    for (int r = 0; r < kBulkRows; ++r) {
                        if constexpr (MODE == kUpdSeq) {
                            acc0 = __fadd_rn(__fmul_rn(acc0, wseq), __fmul_rn(v[r].x, lr));   // :107-108
                            acc1 = __fadd_rn(__fmul_rn(acc1, wseq), __fmul_rn(v[r].y, lr));
                        } else {
                            acc0 = __fadd_rn(acc0, __fmul_rn(v[r].x, lr));                 // :123
                            acc1 = __fadd_rn(acc1, __fmul_rn(v[r].y, lr));
                        }
                    }
                } else {
                    for (uint32_t r = 0; r < n; ++r) {
                        const float2 v = rows[r * (kBulkCols / 2)];
                        if constexpr (MODE == kUpdSeq) {
                            acc0 = __fadd_rn(__fmul_rn(acc0, wseq), __fmul_rn(v.x, lr));
                            acc1 = __fadd_rn(__fmul_rn(acc1, wseq), __fmul_rn(v.y, lr));
                        } else {
                            acc0 = __fadd_rn(acc0, __fmul_rn(v.x, lr));
                            acc1 = __fadd_rn(acc1, __fmul_rn(v.y, lr));
                        }
                    }
                }
                __syncwarp();
                if (lane == 0) ptx::mbar_arrive(&empty[stage]);
                if (++stage == kBulkStages) { stage = 0; phase ^= 1u; }
            }
            const float cb = counts_b[c];
            if (MODE != kUpdPush && blockIdx.y == 0 && lane == 0) counts[c] = __fadd_rn(counts[c], cb);      // :120
            if (active && MODE == kUpdPush) {
                *reinterpret_cast<float2 *>(push.slot(c, d) + col) = make_float2(acc0, acc1);
            } else if (active && MODE == kUpdSeq) {
                centers[(int64_t)c * d + col] = acc0; centers[(int64_t)c * d + col + 1] = acc1;
            } else if (active) {
                const float decay = __fsub_rn(1.f, __fmul_rn(cb, lr));                    // :121
                float *cp = centers + (int64_t)c * d + col;
                const float s0 = __fmul_rn(cp[0], decay), s1 = __fmul_rn(cp[1], decay);
                if (MODE == kUpdFused) {
                    cp[0] = __fadd_rn(s0, acc0); cp[1] = __fadd_rn(s1, acc1);               // :127
                } else {
                    cp[0] = s0; cp[1] = s1;
                    deltas[(int64_t)c * d + col] = acc0; deltas[(int64_t)c * d + col + 1] = acc1;
                }
            }
        }
    }
}

__global__ void km_apply_deltas_kernel(float *__restrict__ centers, const float *__restrict__ deltas,
                                       int64_t n) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) centers[i] = __fadd_rn(centers[i], deltas[i]);
}

int launch_partition(const int64_t *best, int64_t b, int32_t k, uint32_t *blockhist, uint32_t *lrank,
                     uint32_t *total, uint32_t *seg_start, uint32_t *sorted_rows, float *counts_b,
                     cudaStream_t st, float *hist_max, bool *hist_max_written) {
    const int32_t nblk = (int32_t)ceil_div(b, kRankRows);
    const bool par = k <= kRankParMaxK;
    const size_t smem = par ? (size_t)(kRankRows / kWarp) * k * sizeof(uint16_t) : (size_t)k * sizeof(uint32_t);
    if (smem > 200 * 1024) return ACAV_E_UNSUPPORTED;
    static size_t rdone[2][kMaxDevices];
    if (nblk > 0 && par) {
        { int rc = ensure_dynamic_smem(km_block_rank_kernel<true>, smem, rdone[1]); if (rc) return rc; }
        ACAV_CUDA_TRY(launch_pdl(km_block_rank_kernel<true>, dim3(nblk), dim3(kRankRows), smem, st, best, b, k, blockhist, lrank));
    } else if (nblk > 0) {
        { int rc = ensure_dynamic_smem(km_block_rank_kernel<false>, smem, rdone[0]); if (rc) return rc; }
        ACAV_CUDA_TRY(launch_pdl(km_block_rank_kernel<false>, dim3(nblk), dim3(kRankRows), smem, st, best, b, k, blockhist, lrank));
    }
    if (nblk <= kFusedPrefixBlocks) {
        ACAV_CUDA_TRY(launch_pdl(km_segment_start_kernel, dim3(1), dim3(1024), 0, st, total, k, seg_start, blockhist, nblk, counts_b, hist_max));
        if (hist_max_written) *hist_max_written = hist_max != nullptr;
    } else {
        ACAV_CUDA_TRY(launch_pdl(km_block_prefix_kernel, dim3((unsigned)ceil_div(k, 256)), dim3(256), 0, st, blockhist, nblk, k, total, counts_b));
        ACAV_CUDA_TRY(launch_pdl(km_segment_start_kernel, dim3(1), dim3(1024), 0, st, total, k, seg_start, nullptr, 0, nullptr, nullptr));
        if (hist_max_written) *hist_max_written = false;
    }
    if (nblk > 0) {
        ACAV_CUDA_TRY(launch_pdl(km_scatter_rows_kernel, dim3(nblk), dim3(kRankRows), 0, st, best, b, k, blockhist, lrank, seg_start, sorted_rows));
    }
    return 0;
}

// sequential branch: lr_eff[0] = fl32(lr), lr_eff[1] = fl32(1 - lr) (python-float subtraction, then the cast torch
// applies to a python scalar multiplying an fp32 tensor)
__global__ void km_sequential_lr_kernel(double lr, float *__restrict__ lr_eff) {
    lr_eff[0] = (float)lr;
    lr_eff[1] = (float)(1.0 - lr);
}

int launch_sequential_lr(double lr, float *lr_eff, cudaStream_t st) {
    km_sequential_lr_kernel<<<1, 1, 0, st>>>(lr, lr_eff);
    ACAV_LAUNCH_CHECK();
    return 0;
}

int launch_effective_lr(const float *counts_b, int32_t k, double lr, float *lr_eff, int32_t *fallback,
                        cudaStream_t st) {
    ACAV_CUDA_TRY(launch_pdl(km_effective_lr_kernel, dim3(1), dim3(1024), 0, st, counts_b, k, lr, lr_eff, fallback));
    return 0;
}

bool km_heavy_ring() {
    static int use_ring = -1;
    if (use_ring < 0) {
        const char *e = std::getenv("ACAV_KM_HEAVY");
        use_ring = (e && e[0] == 'r') ? 1 : 0;
    }
    return use_ring == 1;
}

int launch_update(const float *x, int64_t ldx, int32_t k, int32_t d, const uint32_t *seg_start,
                  const uint32_t *sorted_rows, const float *counts_b, const float *lr_eff,
                  float *centers, float *counts, float *deltas, const KmPush *push, bool sequential, cudaStream_t st,
                  const KmFork *fork, double lr_in, int32_t *fallback) {
    const bool vec4 = (d % 4 == 0) && (ldx % 4 == 0) &&
                      (((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(centers) |
                         reinterpret_cast<uintptr_t>(deltas)) & 15) == 0);
    const int vec = vec4 ? 4 : 1;
    if (push && !vec4) return ACAV_E_UNSUPPORTED;
    if (sequential) {                                  // :103-109: no histogram decay, no deltas
        if (vec4) {
            const size_t bsmem = (size_t)kBulkStages * kBulkRows * kBulkCols * 4 + 2 * kBulkStages * sizeof(uint64_t);
            static size_t sdone[kMaxDevices];
            { int rc = ensure_dynamic_smem(km_update_bulk_kernel<kUpdSeq>, bsmem, sdone); if (rc) return rc; }
            dim3 bgrid((unsigned)(k < 96 ? k : 96), (unsigned)ceil_div(d, kBulkCols));
            km_update_bulk_kernel<kUpdSeq><<<bgrid, kBulkThreads, bsmem, st>>>(x, ldx, k, d, seg_start, sorted_rows, counts_b, lr_eff,
                                                                                centers, counts, nullptr, KmPush(), -1.0);
            ACAV_LAUNCH_CHECK();
            dim3 grid((unsigned)k, (unsigned)ceil_div(d, 128 * 4));
            ACAV_CUDA_TRY(launch_pdl(km_update_kernel<4, kUpdSeq>, grid, dim3(128), 0, st, x, ldx, d, seg_start, sorted_rows, counts_b, lr_eff, centers, counts, nullptr, KmPush(), -1.0, nullptr));
        } else {
            dim3 grid((unsigned)k, (unsigned)ceil_div(d, 128));
            ACAV_CUDA_TRY(launch_pdl(km_update_kernel<1, kUpdSeq>, grid, dim3(128), 0, st, x, ldx, d, seg_start, sorted_rows, counts_b, lr_eff, centers, counts, nullptr, KmPush(), -1.0, nullptr));
        }
        ACAV_LAUNCH_CHECK();
        return 0;
    }
    KmPush pz = push ? *push : KmPush();
    bool forked = false;
    if (vec4) {
        // heavy centroids (>= kUpdHeavyRows rows of this batch): cp.async ring kernel; its blocks for light
        // centroids exit at once, and km_update_kernel below skips the heavy ones
        const size_t smem = (size_t)kUpdGroups * kUpdGroupRows * kUpdThreads * 16 + (size_t)kUpdChunk * 4;
        static size_t done_fused[kMaxDevices], done_split[kMaxDevices], done_push[kMaxDevices];
        { int rc = ensure_dynamic_smem(km_update_stream_kernel<kUpdFused>, smem, done_fused); if (rc) return rc; }
        { int rc = ensure_dynamic_smem(km_update_stream_kernel<kUpdSplit>, smem, done_split); if (rc) return rc; }
        { int rc = ensure_dynamic_smem(km_update_stream_kernel<kUpdPush>, smem, done_push); if (rc) return rc; }
        const int use_ring = km_heavy_ring() ? 1 : 0;                                    // "ring": the cp.async kernel
        if (use_ring && lr_in >= 0.0) return ACAV_E_STATE;                               // that kernel reads *lr_eff
        if (!use_ring) {
            const size_t bsmem = (size_t)kBulkStages * kBulkRows * kBulkCols * 4 + 2 * kBulkStages * sizeof(uint64_t);
            static size_t bdone[3][kMaxDevices];
            { int rc = ensure_dynamic_smem(km_update_bulk_kernel<kUpdFused>, bsmem, bdone[0]); if (rc) return rc; }
            { int rc = ensure_dynamic_smem(km_update_bulk_kernel<kUpdSplit>, bsmem, bdone[1]); if (rc) return rc; }
            { int rc = ensure_dynamic_smem(km_update_bulk_kernel<kUpdPush>, bsmem, bdone[2]); if (rc) return rc; }
            dim3 bgrid((unsigned)(k < 96 ? k : 96), (unsigned)ceil_div(d, kBulkCols));     // loops over the heavy list
            // heavy and light centroids are disjoint rows of centers / counts / deltas: the two kernels run side by side
            cudaStream_t bst = st;
            if (fork) { int rc = km_fork(fork, st); if (rc) return rc; bst = fork->side; }
            if (push)
                km_update_bulk_kernel<kUpdPush><<<bgrid, kBulkThreads, bsmem, bst>>>(x, ldx, k, d, seg_start, sorted_rows, counts_b, lr_eff,
                                                                         centers, counts, nullptr, pz, -1.0);
            else if (deltas)
                km_update_bulk_kernel<kUpdSplit><<<bgrid, kBulkThreads, bsmem, bst>>>(x, ldx, k, d, seg_start, sorted_rows, counts_b, lr_eff,
                                                                          centers, counts, deltas, pz, lr_in);
            else
                km_update_bulk_kernel<kUpdFused><<<bgrid, kBulkThreads, bsmem, bst>>>(x, ldx, k, d, seg_start, sorted_rows, counts_b, lr_eff,
                                                                          centers, counts, nullptr, pz, lr_in);
            ACAV_LAUNCH_CHECK();
            forked = fork != nullptr;
        }
        dim3 sgrid((unsigned)(k < 96 ? k : 96), (unsigned)ceil_div(d, kUpdThreads * 4));   // loops over the heavy list
        if (!use_ring) {
        } else if (push)
            km_update_stream_kernel<kUpdPush><<<sgrid, kUpdThreads, smem, st>>>(x, ldx, k, d, seg_start, sorted_rows, counts_b, lr_eff,
                                                                                  centers, counts, nullptr, pz);
        else if (deltas)
            km_update_stream_kernel<kUpdSplit><<<sgrid, kUpdThreads, smem, st>>>(x, ldx, k, d, seg_start, sorted_rows, counts_b, lr_eff,
                                                                                   centers, counts, deltas, pz);
        else
            km_update_stream_kernel<kUpdFused><<<sgrid, kUpdThreads, smem, st>>>(x, ldx, k, d, seg_start, sorted_rows, counts_b, lr_eff,
                                                                                   centers, counts, nullptr, pz);
        ACAV_LAUNCH_CHECK();
    }
    dim3 grid((unsigned)k, (unsigned)ceil_div(d, 128 * vec));
    if (push) {
        ACAV_CUDA_TRY(launch_pdl(km_update_kernel<4, kUpdPush>, grid, dim3(128), 0, st, x, ldx, d, seg_start, sorted_rows, counts_b, lr_eff, centers, counts, nullptr, pz, -1.0, nullptr));
    } else if (deltas) {
        if (vec4)
            ACAV_CUDA_TRY(launch_pdl(km_update_kernel<4, kUpdSplit>, grid, dim3(128), 0, st, x, ldx, d, seg_start, sorted_rows, counts_b, lr_eff, centers, counts, deltas, pz, lr_in, fallback));
        else
            ACAV_CUDA_TRY(launch_pdl(km_update_kernel<1, kUpdSplit>, grid, dim3(128), 0, st, x, ldx, d, seg_start, sorted_rows, counts_b, lr_eff, centers, counts, deltas, pz, lr_in, fallback));
    } else {
        if (vec4)
            ACAV_CUDA_TRY(launch_pdl(km_update_kernel<4, kUpdFused>, grid, dim3(128), 0, st, x, ldx, d, seg_start, sorted_rows, counts_b, lr_eff, centers, counts, nullptr, pz, lr_in, fallback));
        else
            ACAV_CUDA_TRY(launch_pdl(km_update_kernel<1, kUpdFused>, grid, dim3(128), 0, st, x, ldx, d, seg_start, sorted_rows, counts_b, lr_eff, centers, counts, nullptr, pz, lr_in, fallback));
    }
    ACAV_LAUNCH_CHECK();
    if (forked) return km_join(fork, st);
    return 0;
}

// flags[i] = counts[i] < *thr ? 0 : 1.  A caller replaying a captured CUDA graph cannot change by-value kernel
// arguments, so it keeps the per-step threshold (count/k)^p of sgd_clustering.py:77 in device memory and hands
// (flags, 0.5f) to the assignment entry points in place of (counts, threshold): flags[i] < 0.5 <=> counts[i] < thr.
__global__ void km_underused_flags_kernel(const float *__restrict__ counts, int32_t k, const float *__restrict__ thr,
                                          float *__restrict__ flags) {
    pdl_begin();
    const int32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < k) flags[i] = counts[i] < *thr ? 0.f : 1.f;
}

int launch_underused_flags(const float *counts, int32_t k, const float *thr_dev, float *flags, cudaStream_t st) {
    ACAV_CUDA_TRY(launch_pdl(km_underused_flags_kernel, dim3((unsigned)ceil_div(k, 256)), dim3(256), 0, st, counts, k, thr_dev, flags));
    return 0;
}

int launch_apply_deltas(float *centers, const float *deltas, int64_t n, cudaStream_t st) {
    if (n == 0) return 0;
    km_apply_deltas_kernel<<<(unsigned)ceil_div(n, 256), 256, 0, st>>>(centers, deltas, n);
    ACAV_LAUNCH_CHECK();
    return 0;
}

}  // namespace acav
