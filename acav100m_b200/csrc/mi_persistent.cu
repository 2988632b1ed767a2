// Persistent greedy-MI kernel: the whole selection loop of EfficientMI.run_greedy
// (subset_selection/code/measures/mi.py:150-192 with EfficientMemMI :284-412) in ONE cooperative launch.
//
// Data layout (built once per engine by the partition kernels below, DESIGN.md "MI layout"):
//   * candidates are STABLY partitioned by table row c1; the stream holds only c2 as uint16
//     (2 bytes per candidate per iteration instead of the reference's 16-byte int64 pair), 0xFFFF =
//     removed.  Inside a row the stream keeps list order, so "first maximum wins" (mi.py:79) is
//     "first in stream" within a row and "smallest original position" (pos[] side array, read only on
//     ties and for the per-thread winner) across rows.
//   * the stream is cut into one contiguous chunk per CTA, balanced by (candidates + 3 * K_v per row
//     touched).  A CTA stages the gain rows of the table rows its chunk touches in SHARED memory
//     (up to kMaxRowsSmem at a time) and streams its chunk with 128-bit loads, gathering gains from
//     shared memory.
// Per iteration: [all CTAs] gain rows + scan + one 64-bit atomicMax  ->  grid barrier  ->  [owner CTA:
// the one whose local best equals the global winner] table/sums update, tombstone, next iteration's
// row/column terms  ->  grid barrier.  Scores use the same fp32 operation sequence and the same
// torch-CPU log table as mi_scan.cu, so picks and gains are bit-identical to the reference.
#include <cooperative_groups.h>

#include "common.cuh"
#include "kernels.cuh"

namespace acav {

constexpr uint16_t kGone = 0xFFFFu;
constexpr int kPersistThreads = 1024;
constexpr int kTileElems = 32768;            // elements per partition tile
constexpr int kPartThreads = 512;
constexpr int kSmallCounts = 256;            // per-iteration table of tN(x) for counts below this

__device__ __forceinline__ float xlogx_cnt(uint32_t k, float f0, const float *__restrict__ logs) {
    return k == 0 ? f0 : __fmul_rn((float)k, __ldg(logs + k));
}
__device__ __forceinline__ float bump_sum(float prev, uint32_t k, float f0, const float *__restrict__ logs) {
    return __fadd_rn(__fsub_rn(prev, xlogx_cnt(k, f0, logs)), xlogx_cnt(k + 1, 0.f, logs));
}

// ---- stable partition of the candidate list by table row ------------------------------------------

__global__ void __launch_bounds__(kPartThreads)
mi_part_count_kernel(const uint32_t *__restrict__ cells, int64_t w, int32_t k_a,
                     uint32_t *__restrict__ tilehist) {
    extern __shared__ uint32_t hist[];
    for (int32_t i = threadIdx.x; i < k_a; i += blockDim.x) hist[i] = 0;
    __syncthreads();
    const int64_t lo = (int64_t)blockIdx.x * kTileElems;
    const int64_t hi = min(w, lo + kTileElems);
    for (int64_t e = lo + threadIdx.x; e < hi; e += blockDim.x) {
        const uint32_t r = cells[e] >> 16;
        if (r < (uint32_t)k_a) atomicAdd(&hist[r], 1u);         // removed entries (0xFFFF) are dropped
    }
    __syncthreads();
    uint32_t *dst = tilehist + (int64_t)blockIdx.x * k_a;
    for (int32_t i = threadIdx.x; i < k_a; i += blockDim.x) dst[i] = hist[i];
}

// per row: exclusive prefix over tiles (in place) and the row total
__global__ void mi_part_prefix_kernel(uint32_t *__restrict__ tilehist, int32_t ntiles, int32_t k_a,
                                      uint32_t *__restrict__ row_total) {
    const int32_t r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= k_a) return;
    uint32_t run = 0;
    for (int32_t t = 0; t < ntiles; ++t) {
        const uint32_t v = tilehist[(int64_t)t * k_a + r];
        tilehist[(int64_t)t * k_a + r] = run;
        run += v;
    }
    row_total[r] = run;
}

// row_start[0..k_a] = exclusive scan of row_total (single block)
__global__ void __launch_bounds__(1024)
mi_part_rowstart_kernel(const uint32_t *__restrict__ total, int32_t k, uint32_t *__restrict__ row_start) {
    __shared__ uint32_t warp_sums[32];
    __shared__ uint32_t carry;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    const int lane = threadIdx.x % kWarp, warp = threadIdx.x / kWarp;
    for (int32_t base = 0; base < k; base += 1024) {
        const int32_t i = base + threadIdx.x;
        const uint32_t v = i < k ? total[i] : 0u;
        uint32_t inc = v;
#pragma unroll
        for (int o = 1; o < kWarp; o <<= 1) {
            const uint32_t t = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= o) inc += t;
        }
        if (lane == kWarp - 1) warp_sums[warp] = inc;
        __syncthreads();
        if (warp == 0) {
            const uint32_t ws = warp_sums[lane];
            uint32_t winc = ws;
#pragma unroll
            for (int o = 1; o < kWarp; o <<= 1) {
                const uint32_t t = __shfl_up_sync(0xffffffffu, winc, o);
                if (lane >= o) winc += t;
            }
            warp_sums[lane] = winc - ws;
        }
        __syncthreads();
        const uint32_t excl = carry + warp_sums[warp] + inc - v;
        if (i < k) row_start[i] = excl;
        __syncthreads();
        if (threadIdx.x == 1023) carry = excl + v;
        __syncthreads();
    }
    if (threadIdx.x == 0) row_start[k] = carry;
}

// stable scatter: tiles in list order, chunks of 512 in list order, warps in order, lanes in order
__global__ void __launch_bounds__(kPartThreads)
mi_part_scatter_kernel(const uint32_t *__restrict__ cells, int64_t w, int32_t k_a,
                       const uint32_t *__restrict__ tilehist, const uint32_t *__restrict__ row_start,
                       uint16_t *__restrict__ c2s, uint32_t *__restrict__ pos_s) {
    extern __shared__ uint32_t cursor[];
    const uint32_t *tp = tilehist + (int64_t)blockIdx.x * k_a;
    for (int32_t i = threadIdx.x; i < k_a; i += blockDim.x) cursor[i] = row_start[i] + tp[i];
    __syncthreads();
    const int64_t lo = (int64_t)blockIdx.x * kTileElems;
    const int64_t hi = min(w, lo + kTileElems);
    const int warp = threadIdx.x / kWarp, lane = threadIdx.x % kWarp;
    for (int64_t base = lo; base < hi; base += kPartThreads) {
        const int64_t e = base + threadIdx.x;
        uint32_t cell = 0xFFFFFFFFu;
        if (e < hi) cell = cells[e];
        const uint32_t r = cell >> 16;
        const bool live = r < (uint32_t)k_a;
        const int32_t key = live ? (int32_t)r : -1;
        for (int ww = 0; ww < kPartThreads / kWarp; ++ww) {
            if (warp == ww) {
                const unsigned m = __match_any_sync(0xffffffffu, key);
                const int leader = __ffs(m) - 1;
                const uint32_t rank = __popc(m & ((1u << lane) - 1u));
                uint32_t basev = 0;
                if (live && lane == leader) {
                    basev = cursor[key];
                    cursor[key] = basev + __popc(m);
                }
                basev = __shfl_sync(0xffffffffu, basev, leader);
                if (live) {
                    c2s[basev + rank] = (uint16_t)(cell & 0xFFFFu);
                    pos_s[basev + rank] = (uint32_t)e;
                }
            }
            __syncthreads();
        }
    }
}

__global__ void mi_fill_u16_kernel(uint16_t *p, int64_t lo, int64_t hi, uint16_t v) {
    const int64_t i = lo + (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < hi) p[i] = v;
}

// ---- the persistent kernel -----------------------------------------------------------------------

struct MiPersist {
    MiState s;
    const uint16_t *c2s_ro;          // same memory as c2s (loads bypass L1)
    uint16_t *c2s;
    const uint32_t *pos_s;
    const uint32_t *row_start;       // [k_a + 1]
    const uint32_t *chunk_start;     // [grid + 1] element offsets, multiples of 8
    unsigned long long *slots;       // [2] winner key per iteration parity
    unsigned int *bar;               // [2] {count, generation}
    int64_t w_sorted;                // live candidates in the stream
    int64_t n_picks;
    int64_t *out_pos;
    float *out_gain;
    int32_t rows_smem;               // gain rows that fit in shared memory
};

__device__ __forceinline__ void grid_barrier(unsigned int *bar, unsigned int nblocks) {
    __syncthreads();
    if (threadIdx.x == 0) {
        volatile unsigned int *gen = bar + 1;
        const unsigned int g = *gen;
        __threadfence();
        if (atomicAdd(bar, 1u) == nblocks - 1) {
            bar[0] = 0;
            __threadfence();
            atomicAdd(bar + 1, 1u);
        } else {
            while (*gen == g) { }
        }
        __threadfence();
    }
    __syncthreads();
}

__device__ __forceinline__ uint4 ldcg_u4(const uint4 *p) { return __ldcg(p); }

__global__ void __launch_bounds__(kPersistThreads, 1) mi_persistent_kernel(MiPersist P) {
    extern __shared__ __align__(16) unsigned char psmem[];
    const MiState &s = P.s;
    const int32_t k_v = s.k_v;
    float *col_term = reinterpret_cast<float *>(psmem);                 // [k_v]
    float *tn_small = col_term + k_v;                                   // [kSmallCounts]
    uint32_t *rs_local = reinterpret_cast<uint32_t *>(tn_small + kSmallCounts);   // [rows_smem + 1]
    float *gain = reinterpret_cast<float *>(rs_local + P.rows_smem + 1 + ((P.rows_smem + 1) & 1));  // [rows_smem][k_v]
    __shared__ unsigned long long wkey[32];
    __shared__ uint32_t widx[32];
    __shared__ unsigned long long sh_best_key;
    __shared__ uint32_t sh_best_idx;

    const uint32_t e_lo = P.chunk_start[blockIdx.x], e_hi = P.chunk_start[blockIdx.x + 1];
    // rows touched by this chunk: first row with row_start[r+1] > e_lo ... last row with row_start[r] < e_hi
    int32_t r_lo = 0, r_hi = -1;
    if (e_hi > e_lo) {
        int32_t a = 0, b = s.k_a;                                        // upper_bound(row_start, e_lo) - 1
        while (a < b) { const int32_t m = (a + b) >> 1; if (P.row_start[m + 1] > e_lo) b = m; else a = m + 1; }
        r_lo = a;
        a = r_lo; b = s.k_a;
        while (a < b) { const int32_t m = (a + b) >> 1; if (P.row_start[m] < e_hi) a = m + 1; else b = m; }
        r_hi = a - 1;
    }
    const uint32_t base_pos = (uint32_t)s.pos_base;

    for (int64_t it = 0; it < P.n_picks; ++it) {
        // ---------------- phase A: score my chunk ----------------
        const float NlogN = __ldcg(s.sums + 0), nn = __ldcg(s.sums + 3), fN0 = __ldcg(s.sums + 4);
        const float np = __fadd_rn(nn, 1.0f);
        const float lognp = __ldg(s.logs + (int64_t)np);
        for (int32_t i = threadIdx.x; i < k_v; i += blockDim.x) col_term[i] = __ldcg(s.col_term + i);
        for (int32_t i = threadIdx.x; i < kSmallCounts; i += blockDim.x)
            tn_small[i] = __fdiv_rn(bump_sum(NlogN, (uint32_t)i, fN0, s.logs), np);
        float bs = 0.f;                 // best gain of this thread
        uint32_t bi = 0xFFFFFFFFu;      // its stream index
        uint32_t bend = 0;              // end of the row segment that holds it (ties inside are never earlier)
        uint32_t bp = 0;                // its original position (valid when bpv)
        bool bpv = false;
        for (int32_t rb = r_lo; rb <= r_hi; rb += P.rows_smem) {
            const int32_t nr = min(P.rows_smem, r_hi - rb + 1);
            __syncthreads();            // previous sub-batch finished reading gain rows / first use of col_term
            for (int32_t i = threadIdx.x; i <= nr; i += blockDim.x) rs_local[i] = P.row_start[rb + i];
            for (int32_t i = threadIdx.x; i < nr * k_v; i += blockDim.x) {
                const int32_t rr = i / k_v, c2 = i - rr * k_v;
                const uint32_t x = __ldcg(s.n_cells + (int64_t)(rb + rr) * k_v + c2);
                const float tN = x < (uint32_t)kSmallCounts ? tn_small[x]
                                                           : __fdiv_rn(bump_sum(NlogN, x, fN0, s.logs), np);
                const float rt = __ldcg(s.row_term + rb + rr);
                gain[i] = __fadd_rn(__fadd_rn(__fadd_rn(tN, col_term[c2]), rt), lognp);
            }
            __syncthreads();
            const uint32_t s_lo = max(e_lo, rs_local[0]), s_hi = min(e_hi, rs_local[nr]);
            if (s_hi <= s_lo) continue;
            const uint32_t v_lo = s_lo >> 3, v_hi = (s_hi + 7) >> 3;
            int32_t crow = 0;                                  // cached local row of this thread
            for (uint32_t v = v_lo + threadIdx.x; v < v_hi; v += blockDim.x) {
                const uint4 q = ldcg_u4(reinterpret_cast<const uint4 *>(P.c2s_ro) + v);
                const uint32_t words[4] = {q.x, q.y, q.z, q.w};
                const uint32_t e0 = v << 3;
                uint32_t ef = max(e0, s_lo);
                if (!(ef >= rs_local[crow] && ef < rs_local[crow + 1])) {
                    int32_t a = 0, b = nr;                     // row with rs_local[row] <= ef < rs_local[row+1]
                    while (a < b) { const int32_t m = (a + b) >> 1; if (rs_local[m + 1] > ef) b = m; else a = m + 1; }
                    crow = a;
                }
                uint32_t rend = rs_local[crow + 1];
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const uint32_t e = e0 + j;
                    const uint32_t c2 = (words[j >> 1] >> ((j & 1) * 16)) & 0xFFFFu;
                    if (e < s_lo || e >= s_hi) continue;
                    while (e >= rend) { ++crow; rend = rs_local[crow + 1]; }
                    if (c2 == kGone) continue;
                    const float g = gain[crow * k_v + c2];
                    if (bi == 0xFFFFFFFFu || g > bs) {
                        bs = g; bi = e; bend = rend; bpv = false;
                    } else if (g == bs && e >= bend) {           // tie with a candidate of a later row segment
                        if (!bpv) { bp = __ldg(P.pos_s + bi); bpv = true; }
                        const uint32_t pe = __ldg(P.pos_s + e);
                        if (pe < bp) { bi = e; bp = pe; bend = rend; }
                    }
                }
            }
        }
        unsigned long long key = 0ull;
        if (bi != 0xFFFFFFFFu) {
            if (!bpv) bp = __ldg(P.pos_s + bi);
            key = make_key(bs, base_pos + bp);
        }
        // block arg-max of (key, stream index)
        {
            unsigned long long k2 = key;
            uint32_t i2 = bi;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                const unsigned long long ok = __shfl_xor_sync(0xffffffffu, k2, o);
                const uint32_t oi = __shfl_xor_sync(0xffffffffu, i2, o);
                if (ok > k2) { k2 = ok; i2 = oi; }
            }
            if (threadIdx.x % kWarp == 0) { wkey[threadIdx.x / kWarp] = k2; widx[threadIdx.x / kWarp] = i2; }
            __syncthreads();
            if (threadIdx.x < kWarp) {
                k2 = wkey[threadIdx.x]; i2 = widx[threadIdx.x];
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) {
                    const unsigned long long ok = __shfl_xor_sync(0xffffffffu, k2, o);
                    const uint32_t oi = __shfl_xor_sync(0xffffffffu, i2, o);
                    if (ok > k2) { k2 = ok; i2 = oi; }
                }
                if (threadIdx.x == 0) {
                    sh_best_key = k2; sh_best_idx = i2;
                    if (k2) atomicMax(P.slots + (it & 1), k2);
                }
            }
        }
        grid_barrier(P.bar, gridDim.x);
        // ---------------- phase B: the owner applies the winner ----------------
        const unsigned long long win = __ldcg(P.slots + (it & 1));
        if (win == 0ull) {                                     // nothing left anywhere: record and stop
            if (blockIdx.x == 0 && threadIdx.x == 0)
                for (int64_t j = it; j < P.n_picks; ++j) { P.out_pos[j] = -1; P.out_gain[j] = nanf(""); }
            break;
        }
        if (sh_best_key == win) {                              // keys are unique: exactly one owner
            const uint32_t widx_s = sh_best_idx;
            // locate the winner's row (c1) from the stream index
            if (threadIdx.x == 0) {
                int32_t a = 0, b = s.k_a;
                while (a < b) { const int32_t m = (a + b) >> 1; if (P.row_start[m + 1] > widx_s) b = m; else a = m + 1; }
                const int32_t c1 = a;
                const int32_t c2 = (int32_t)__ldcg(reinterpret_cast<const unsigned short *>(P.c2s) + widx_s);
                const uint32_t x = __ldcg(s.n_cells + (int64_t)c1 * k_v + c2), y = __ldcg(s.a_cols + c2),
                               z = __ldcg(s.b_rows + c1);
                const float fN0w = __ldcg(s.sums + 4), fa0w = __ldcg(s.sums + 5);
                s.sums[0] = bump_sum(__ldcg(s.sums + 0), x, fN0w, s.logs);        // update_cache mi.py:383-389
                s.sums[1] = bump_sum(__ldcg(s.sums + 1), y, fa0w, s.logs);
                s.sums[2] = bump_sum(__ldcg(s.sums + 2), z, fa0w, s.logs);
                s.sums[3] = __fadd_rn(__ldcg(s.sums + 3), 1.0f);                  // update_mats :401-406
                s.n_cells[(int64_t)c1 * k_v + c2] = x + 1; s.a_cols[c2] = y + 1; s.b_rows[c1] = z + 1;
                P.c2s[widx_s] = kGone;                                            // remove_idx_all :104-106
                const int64_t pos = (int64_t)key_pos(win);
                s.cells[pos - s.pos_base] = 0xFFFFFFFFu;                          // keep the list-order view in sync
                P.out_pos[it] = pos;
                P.out_gain[it] = key_score(win);
                P.slots[(it + 1) & 1] = 0ull;
                __threadfence();
            }
            __syncthreads();
            // next iteration's per-row / per-column terms (same code path as mi_scan.cu: mi_terms)
            {
                const float npn = __fadd_rn(__ldcg(s.sums + 3), 1.0f);
                const float aloga = __ldcg(s.sums + 1), blogb = __ldcg(s.sums + 2), fa0 = __ldcg(s.sums + 5);
                for (int32_t i = threadIdx.x; i < s.k_v; i += blockDim.x)
                    s.col_term[i] = __fdiv_rn(-bump_sum(aloga, __ldcg(s.a_cols + i), fa0, s.logs), npn);
                for (int32_t i = threadIdx.x; i < s.k_a; i += blockDim.x)
                    s.row_term[i] = __fdiv_rn(-bump_sum(blogb, __ldcg(s.b_rows + i), fa0, s.logs), npn);
            }
            __threadfence();
        }
        grid_barrier(P.bar, gridDim.x);
    }
}

// ---- host side -----------------------------------------------------------------------------------

int mi_partition_scratch_tiles(int64_t w) { return (int)ceil_div(w > 0 ? w : 1, kTileElems); }

int launch_mi_partition(const uint32_t *cells, int64_t w, int32_t k_a, uint32_t *tilehist, uint32_t *row_total,
                        uint32_t *row_start, uint16_t *c2s, uint32_t *pos_s, int64_t stream_capacity,
                        cudaStream_t st) {
    const int ntiles = mi_partition_scratch_tiles(w);
    const size_t smem = (size_t)k_a * sizeof(uint32_t);
    if (smem > 96 * 1024) return ACAV_E_UNSUPPORTED;
    static bool attr = false;
    if (!attr && smem > 48 * 1024) {
        ACAV_CUDA_TRY(cudaFuncSetAttribute(mi_part_count_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
        ACAV_CUDA_TRY(cudaFuncSetAttribute(mi_part_scatter_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
        attr = true;
    }
    mi_part_count_kernel<<<ntiles, kPartThreads, smem, st>>>(cells, w, k_a, tilehist);
    ACAV_LAUNCH_CHECK();
    mi_part_prefix_kernel<<<(unsigned)ceil_div(k_a, 256), 256, 0, st>>>(tilehist, ntiles, k_a, row_total);
    ACAV_LAUNCH_CHECK();
    mi_part_rowstart_kernel<<<1, 1024, 0, st>>>(row_total, k_a, row_start);
    ACAV_LAUNCH_CHECK();
    mi_part_scatter_kernel<<<ntiles, kPartThreads, smem, st>>>(cells, w, k_a, tilehist, row_start, c2s, pos_s);
    ACAV_LAUNCH_CHECK();
    (void)stream_capacity;
    return 0;
}

int launch_mi_fill_gone(uint16_t *c2s, int64_t lo, int64_t hi, cudaStream_t st) {
    if (hi <= lo) return 0;
    mi_fill_u16_kernel<<<(unsigned)ceil_div(hi - lo, 256), 256, 0, st>>>(c2s, lo, hi, kGone);
    ACAV_LAUNCH_CHECK();
    return 0;
}

// shared memory needed for `rows` gain rows
static size_t persist_smem_bytes(int32_t k_v, int32_t rows) {
    size_t words = (size_t)k_v + kSmallCounts + (size_t)rows + 1 + ((rows + 1) & 1) + (size_t)rows * k_v;
    return words * 4 + 16;
}

int mi_persistent_rows_that_fit(int32_t k_v) {
    const size_t budget = 200 * 1024;
    int32_t rows = 0;
    while (persist_smem_bytes(k_v, rows + 1) <= budget && rows < 4096) ++rows;
    return rows;
}

int launch_mi_persistent(const MiState &s, uint16_t *c2s, const uint32_t *pos_s, const uint32_t *row_start,
                         const uint32_t *chunk_start, int32_t grid, unsigned long long *slots, unsigned int *bar,
                         int64_t w_sorted, int64_t n_picks, int64_t *out_pos, float *out_gain, int32_t rows_smem,
                         cudaStream_t st) {
    MiPersist P;
    P.s = s; P.c2s_ro = c2s; P.c2s = c2s; P.pos_s = pos_s; P.row_start = row_start; P.chunk_start = chunk_start;
    P.slots = slots; P.bar = bar; P.w_sorted = w_sorted; P.n_picks = n_picks; P.out_pos = out_pos;
    P.out_gain = out_gain; P.rows_smem = rows_smem;
    const size_t smem = persist_smem_bytes(s.k_v, rows_smem);
    static size_t attr_set = 0;
    if (smem > attr_set) {
        ACAV_CUDA_TRY(cudaFuncSetAttribute(mi_persistent_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr_set = smem;
    }
    ACAV_CUDA_TRY(cudaMemsetAsync(slots, 0, 2 * sizeof(unsigned long long), st));
    ACAV_CUDA_TRY(cudaMemsetAsync(bar, 0, 2 * sizeof(unsigned int), st));
    void *args[] = {&P};
    ACAV_CUDA_TRY(cudaLaunchCooperativeKernel((void *)mi_persistent_kernel, dim3(grid), dim3(kPersistThreads), args,
                                              smem, st));
    return 0;
}

}  // namespace acav
