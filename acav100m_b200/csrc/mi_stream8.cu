// Persistent greedy-MI kernel, ONE BYTE per candidate (ACAV_MI_LOOP_BYTES): the selection loop of
// EfficientMI.run_greedy (subset_selection/code/measures/mi.py:150-192 with EfficientMemMI :284-412) in one
// cooperative launch, like mi_persistent.cu, with half the HBM bytes per iteration and no shared-memory ring.
//
// Layout (built once per engine):
//   * the K_a x K_v table is cut into SUB-ROWS of sub_w <= 255 columns (n_sub = ceil(K_v / 255) per table row,
//     sub_w = ceil(K_v / n_sub): 5 x 205 at K_v = 1024).  Candidates are STABLY partitioned by sub-row; the stream holds
//     the column inside the sub-row as ONE byte (255 = removed / padding; that slot of a gain row holds -inf, so the
//     hot loop has no branch).  100 MB per iteration at W = 1e8 instead of the reference's 1.6 GB of int64 pairs.
//   * sub-rows are padded to whole 512-candidate blocks, every block is sorted by the shared-memory bank of its gain
//     slot (lane l owns entries [16l, 16l+16): gather step j reads entries 16 apart, i.e. ~32 different banks); ties are
//     settled by original position (pos[], permuted along, read only on that rare path).
//   * the stream is cut into one contiguous chunk per CTA; a CTA stages the gain rows (256 floats) of the sub-rows its
//     chunk touches in shared memory, 1 KiB aligned so that a gather address is (byte << 2) | row_base: shift, LOP3, LDS.
//   * stream loads go through REGISTERS: DEPTH 16-byte loads per thread in flight (ld.global.cg, L2 evict-first), 512
//     threads x 128 registers per CTA -- the shared-memory pipe only serves the gathers.
//   * a CTA whose sub-rows all fit keeps their table COUNTS in shared memory for the whole launch (every CTA learns
//     every winner and bumps its copy), so building the gain rows of an iteration reads no global memory; the others
//     read the double-buffered global table like mi_persistent.cu.
//   * the per-thread best remembers position, sub-row and byte of its candidate, so publishing the CTA's winner needs
//     no dependent global loads in the common case.
// Scores use the same fp32 operation sequence and the same torch-CPU log table as mi_scan.cu: picks and gains are
// bit-identical to the reference (tests/test_mi_gpu.py).
#include "common.cuh"
#include "kernels.cuh"
#include "mi_loop.cuh"

namespace acav {

namespace {

constexpr int kS8Removed = 255;             // byte of a removed / padding entry
constexpr int kS8GainStride = 256;          // floats per staged gain row (1 KiB), slot 255 = -inf
constexpr int kS8Blk = 512;                 // candidates per stream block: 32 lanes x 16 bytes
constexpr int kS8Tile = 32768;              // candidates per partition tile
constexpr int kS8PartThreads = 512;
constexpr size_t kS8SmemBudget = 222 * 1024;

struct S8Geom {
    int32_t n_sub, sub_w, k_rows;
};
__host__ __device__ inline S8Geom s8_geom(int32_t k_a, int32_t k_v) {
    S8Geom g;
    g.n_sub = (k_v + kS8Removed - 1) / kS8Removed;
    g.sub_w = (k_v + g.n_sub - 1) / g.n_sub;
    g.k_rows = k_a * g.n_sub;
    return g;
}

// ---- stable partition of the candidate list by sub-row ---------------------------------------------------------

__global__ void __launch_bounds__(kS8PartThreads)
s8_count_kernel(const uint32_t *__restrict__ cells, int64_t w, S8Geom g, uint32_t *__restrict__ tilehist) {
    extern __shared__ uint32_t s8_hist[];
    for (int32_t i = threadIdx.x; i < g.k_rows; i += blockDim.x) s8_hist[i] = 0;
    __syncthreads();
    const int64_t lo = (int64_t)blockIdx.x * kS8Tile, hi = min(w, lo + kS8Tile);
    for (int64_t e = lo + threadIdx.x; e < hi; e += blockDim.x) {
        const uint32_t cell = cells[e];
        if (cell != 0xFFFFFFFFu)                                  // removed entries are dropped
            atomicAdd(&s8_hist[(cell >> 16) * (uint32_t)g.n_sub + (cell & 0xFFFFu) / (uint32_t)g.sub_w], 1u);
    }
    __syncthreads();
    uint32_t *dst = tilehist + (int64_t)blockIdx.x * g.k_rows;
    for (int32_t i = threadIdx.x; i < g.k_rows; i += blockDim.x) dst[i] = s8_hist[i];
}

// per sub-row: exclusive prefix over tiles (in place) and the total
__global__ void s8_prefix_kernel(uint32_t *__restrict__ tilehist, int32_t ntiles, int32_t k_rows,
                                 uint32_t *__restrict__ row_total) {
    const int32_t r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= k_rows) return;
    uint32_t run = 0;
    for (int32_t t = 0; t < ntiles; ++t) {
        const uint32_t v = tilehist[(int64_t)t * k_rows + r];
        tilehist[(int64_t)t * k_rows + r] = run;
        run += v;
    }
    row_total[r] = run;
}

// row_start[0..k] = exclusive scan of the totals, each padded to whole blocks (single CTA)
__global__ void __launch_bounds__(1024)
s8_rowstart_kernel(const uint32_t *__restrict__ total, int32_t k, uint32_t *__restrict__ row_start) {
    __shared__ uint32_t warp_sums[32];
    __shared__ uint32_t carry;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    const int lane = threadIdx.x % kWarp, warp = threadIdx.x / kWarp;
    for (int32_t base = 0; base < k; base += 1024) {
        const int32_t i = base + threadIdx.x;
        const uint32_t v = i < k ? ((total[i] + (kS8Blk - 1)) & ~(uint32_t)(kS8Blk - 1)) : 0u;
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

// stable scatter: tiles in list order, 512-chunks in list order, warps in order, lanes in order
__global__ void __launch_bounds__(kS8PartThreads)
s8_scatter_kernel(const uint32_t *__restrict__ cells, int64_t w, S8Geom g, const uint32_t *__restrict__ tilehist,
                  const uint32_t *__restrict__ row_start, uint8_t *__restrict__ stream, uint32_t *__restrict__ pos_s) {
    extern __shared__ uint32_t s8_cursor[];
    const uint32_t *tp = tilehist + (int64_t)blockIdx.x * g.k_rows;
    for (int32_t i = threadIdx.x; i < g.k_rows; i += blockDim.x) s8_cursor[i] = row_start[i] + tp[i];
    __syncthreads();
    const int64_t lo = (int64_t)blockIdx.x * kS8Tile, hi = min(w, lo + kS8Tile);
    const int warp = threadIdx.x / kWarp, lane = threadIdx.x % kWarp;
    for (int64_t base = lo; base < hi; base += kS8PartThreads) {
        const int64_t e = base + threadIdx.x;
        uint32_t cell = 0xFFFFFFFFu;
        if (e < hi) cell = cells[e];
        const bool live = cell != 0xFFFFFFFFu;
        const uint32_t c2 = cell & 0xFFFFu;
        const uint32_t sub = c2 / (uint32_t)g.sub_w;
        const int32_t key = live ? (int32_t)((cell >> 16) * (uint32_t)g.n_sub + sub) : -1;
        for (int ww = 0; ww < kS8PartThreads / kWarp; ++ww) {
            if (warp == ww) {
                const unsigned m = __match_any_sync(0xffffffffu, key);
                const int leader = __ffs(m) - 1;
                const uint32_t rank = __popc(m & ((1u << lane) - 1u));
                uint32_t basev = 0;
                if (live && lane == leader) {
                    basev = s8_cursor[key];
                    s8_cursor[key] = basev + __popc(m);
                }
                basev = __shfl_sync(0xffffffffu, basev, leader);
                if (live) {
                    stream[basev + rank] = (uint8_t)(c2 - sub * (uint32_t)g.sub_w);       // column inside the sub-row
                    pos_s[basev + rank] = (uint32_t)e;
                }
            }
            __syncthreads();
        }
    }
}

__global__ void s8_fill_kernel(uint4 *p, int64_t n16, uint32_t v) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n16) p[i] = make_uint4(v, v, v, v);
}

// Sort every 512-candidate block by (shared-memory bank of its gain slot, column, original order): in the scan lane l
// owns entries [16l, 16l+16) and gather step j reads entry 16l+j in all lanes -- 16 apart in bank order.
__global__ void __launch_bounds__(kS8Blk)
s8_block_sort_kernel(uint8_t *__restrict__ stream, uint32_t *__restrict__ pos_s, int64_t n_blocks) {
    __shared__ uint32_t key[kS8Blk];
    __shared__ uint8_t colv[kS8Blk];
    __shared__ uint32_t posv[kS8Blk];
    for (int64_t blk = blockIdx.x; blk < n_blocks; blk += gridDim.x) {
        const int64_t base = blk * kS8Blk;
        const int t = threadIdx.x;
        const uint32_t c = stream[base + t];
        colv[t] = (uint8_t)c;
        posv[t] = pos_s[base + t];
        key[t] = ((c & 31u) << 20) | (c << 9) | (uint32_t)t;        // c < 2^8, t < 2^9
        __syncthreads();
        for (int size = 2; size <= kS8Blk; size <<= 1) {
            for (int stride = size >> 1; stride > 0; stride >>= 1) {
                const int partner = t ^ stride;
                if (partner > t) {
                    const uint32_t a = key[t], b = key[partner];
                    const bool up = (t & size) == 0;
                    if ((a > b) == up) { key[t] = b; key[partner] = a; }
                }
                __syncthreads();
            }
        }
        const int src = (int)(key[t] & 511u);
        stream[base + t] = colv[src];
        pos_s[base + t] = posv[src];
        __syncthreads();
    }
}

// ---- the persistent kernel ----------------------------------------------------------------------------------------

struct MiS8 {
    MiState s;
    uint32_t *n_alt;                 // second copy of the table counts (double buffering as in mi_persistent.cu)
    uint8_t *stream;
    const uint32_t *pos_s;
    const uint32_t *row_start;       // [k_rows + 1] stream offsets of the sub-rows
    const uint32_t *chunk_start;     // [grid + 1]
    MiPub *pub;
    unsigned int *bar;
    int64_t n_picks;
    int64_t *out_pos;
    float *out_gain;
    S8Geom g;
    int32_t rows_smem;               // sub-rows whose gain row + counts fit in shared memory
    int32_t fixed_bytes;             // bytes of the arrays in front of the gain rows
    int32_t use_cache;               // 0: never keep counts in shared memory (debug / comparison)
    int32_t world, rank;
    unsigned int seq_base;
    MiMail *mail_local;
    MiMail *mail_peer[kMaxWorld];
    long long *dbg;
    int *status;
    unsigned long long spin_limit_ns;
};

__device__ __forceinline__ uint64_t s8_policy_evict_first() {
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
// 16-byte stream load into registers: L2-coherent (.cg: removals written by another SM before the grid barrier are
// seen), evict-first in L2 so the 100 MB stream does not push the table, positions and log table out
__device__ __forceinline__ uint4 s8_ld_stream(const uint4 *p, uint64_t pol) {
    uint4 v;
    asm volatile("ld.global.cg.L2::cache_hint.v4.u32 {%0, %1, %2, %3}, [%4], %5;"
                 : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p), "l"(pol));
    return v;
}
__device__ __forceinline__ float s8_gather(uint32_t word, int byte, uint32_t row_base) {
    // address = ((word >> 8*byte) & 0xFF) << 2 | row_base (row_base is 1 KiB aligned): one shift, one LOP3, one LDS
    const uint32_t sh = byte == 0 ? (word << 2) : (word >> (8 * byte - 2));
    uint32_t addr;
    asm("lop3.b32 %0, %1, 0x3FC, %2, 0xEA;" : "=r"(addr) : "r"(sh), "r"(row_base));      // (a & b) | c
    float g;
    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(g) : "r"(addr));
    return g;
}

// Per-thread running arg-max.  `bi` = stream index of the first candidate (of its sub-row segment) holding the best gain
// `bs`; bpos / brow / bbyte describe that candidate while `meta` is true.  Equal-gain candidates of LATER segments are
// parked in tie[] and only compared by original position if this thread ends up holding the block maximum.
struct S8Best {
    float bs;
    uint32_t bi, bend, bpos;
    int32_t brow;
    uint32_t bbyte;
    uint32_t tie[3];
    int ntie;
    bool meta;
};

__device__ __forceinline__ void s8_tie(S8Best &b, uint32_t e, uint32_t rend, const uint32_t *__restrict__ pos_s) {
    if (b.ntie < 3) {
        if (b.ntie == 0) b.tie[0] = e;
        else if (b.ntie == 1) b.tie[1] = e;
        else b.tie[2] = e;
        ++b.ntie;
    } else {                                   // list full: settle by original position now
        uint32_t bp = b.meta ? b.bpos : __ldg(pos_s + b.bi);
        uint32_t nb = b.bi;
#pragma unroll
        for (int t = 0; t < 3; ++t) {
            const uint32_t pt = __ldg(pos_s + b.tie[t]);
            if (pt < bp) { bp = pt; nb = b.tie[t]; }
        }
        const uint32_t pe = __ldg(pos_s + e);
        if (pe < bp) { bp = pe; nb = e; }
        if (nb != b.bi) { b.bi = nb; b.meta = false; }
        b.bpos = bp;
        b.ntie = 0;
    }
    b.bend = rend;
}

template <int THREADS, int DEPTH>
__global__ void __launch_bounds__(THREADS, 1) mi_stream8_kernel(MiS8 P) {
    extern __shared__ __align__(16) unsigned char s8_smem[];
    const MiState &s = P.s;
    const int32_t k_v = s.k_v, k_a = s.k_a;
    const int32_t n_sub = P.g.n_sub, sub_w = P.g.sub_w, k_rows = P.g.k_rows;
    float *col_term = reinterpret_cast<float *>(s8_smem);                          // [k_v]
    float *tn_small = col_term + k_v;                                              // [kSmallCounts]
    uint32_t *a_cnt = reinterpret_cast<uint32_t *>(tn_small + kSmallCounts);       // [k_v]
    uint32_t *b_cnt = a_cnt + k_v;                                                 // [k_a]
    uint32_t *rs_loc = b_cnt + k_a;                                                // [rows_smem + 1] stream offsets
    float *rt_local = reinterpret_cast<float *>(rs_loc + P.rows_smem + 1);         // [rows_smem]
    int32_t *row_c1 = reinterpret_cast<int32_t *>(rt_local + P.rows_smem);         // [rows_smem] table row of the sub-row
    int32_t *row_c2 = row_c1 + P.rows_smem;                                        // [rows_smem] first column of the sub-row
    // gain rows: 1 KiB aligned in the shared window (the gather ORs the row base into the byte offset)
    const uint32_t smem_b = (uint32_t)__cvta_generic_to_shared(s8_smem);
    const uint32_t gain_b = (smem_b + (uint32_t)P.fixed_bytes + 1023u) & ~1023u;
    float *gain = reinterpret_cast<float *>(s8_smem + (gain_b - smem_b));          // [rows_smem][256]
    uint32_t *cnt = reinterpret_cast<uint32_t *>(gain + (size_t)P.rows_smem * kS8GainStride);   // [rows_smem][256]
    __shared__ unsigned long long wkey[32];
    __shared__ unsigned long long wpay[32];
    __shared__ uint32_t widx[32];
    __shared__ unsigned long long sh_best_key, sh_win_key, sh_win_pay;
    __shared__ uint32_t sh_best_idx;
    __shared__ int32_t sh_rows[2];
    __shared__ float ps[6];                                   // {NlogN, aloga, blogb, n, fN0, fa0}
    constexpr int kWarps = THREADS / kWarp;

    for (int32_t i = threadIdx.x; i < k_v; i += THREADS) a_cnt[i] = __ldcg(s.a_cols + i);
    for (int32_t i = threadIdx.x; i < k_a; i += THREADS) b_cnt[i] = __ldcg(s.b_rows + i);
    if (threadIdx.x < 6) ps[threadIdx.x] = __ldcg(s.sums + threadIdx.x);
    const uint32_t e_lo = P.chunk_start[blockIdx.x], e_hi = P.chunk_start[blockIdx.x + 1];
    if (threadIdx.x == 0) {                                    // sub-rows touched by this chunk
        int32_t r_lo = 0, r_hi = -1;
        if (e_hi > e_lo) {
            int32_t a = 0, b = k_rows;
            while (a < b) { const int32_t m = (a + b) >> 1; if (__ldg(P.row_start + m + 1) > e_lo) b = m; else a = m + 1; }
            r_lo = a;
            a = r_lo; b = k_rows;
            while (a < b) { const int32_t m = (a + b) >> 1; if (__ldg(P.row_start + m) < e_hi) a = m + 1; else b = m; }
            r_hi = a - 1;
        }
        sh_rows[0] = r_lo; sh_rows[1] = r_hi;
    }
    __syncthreads();
    const int32_t r_lo = sh_rows[0], r_hi = sh_rows[1];
    const int32_t n_mine = r_hi - r_lo + 1;
    const bool cached = P.use_cache && n_mine > 0 && n_mine <= P.rows_smem;
    if (cached) {                                             // table counts of my sub-rows stay here for the whole launch
        for (int32_t idx = threadIdx.x; idx < n_mine * kS8GainStride; idx += THREADS) {
            const int32_t i = idx >> 8, j = idx & 255;
            const int32_t grow = r_lo + i, c1 = grow / n_sub, c2 = (grow - c1 * n_sub) * sub_w + j;
            cnt[idx] = (j < sub_w && c2 < k_v) ? __ldcg(s.n_cells + (int64_t)c1 * k_v + c2) : 0u;
        }
    }
    const uint32_t base_pos = (uint32_t)s.pos_base;
    const uint32_t grid = gridDim.x;
    const uint64_t stream_pol = s8_policy_evict_first();
    const uint32_t lane = threadIdx.x % kWarp;
    int32_t prev1 = -1, prev2 = -1;                            // table cell of picks it-1, it-2
    int64_t done = 0;
    bool broke = false;
    long long t_learn = 0;
    __syncthreads();

    for (int64_t it = 0; it < P.n_picks; ++it) {
        const int cur = (int)(it & 1);
        const long long t0 = P.dbg ? clock64() : 0;
        long long t_gain = 0, t_pre = 0;
        const uint32_t *Tcur = cur ? P.n_alt : s.n_cells;
        uint32_t *Toth = cur ? s.n_cells : P.n_alt;
        if (blockIdx.x == 0 && threadIdx.x == 0) {             // lagged writer of the other table copy
            if (prev2 >= 0) Toth[prev2] += 1;
            if (prev1 >= 0) Toth[prev1] += 1;
        }
        // ---------------- score my chunk ----------------
        const float NlogN = ps[0], aloga = ps[1], blogb = ps[2], fN0 = ps[4], fa0 = ps[5];
        const float np = __fadd_rn(ps[3], 1.0f);
        const float lognp = __ldg(s.logs + (int64_t)np);
        for (int32_t i = threadIdx.x; i < k_v; i += THREADS)
            col_term[i] = __fdiv_rn(-bump_sum(aloga, a_cnt[i], fa0, s.logs), np);
        for (int32_t i = threadIdx.x; i < kSmallCounts; i += THREADS)
            tn_small[i] = __fdiv_rn(bump_sum(NlogN, (uint32_t)i, fN0, s.logs), np);
        S8Best B;
        B.bs = -INFINITY; B.bi = 0xFFFFFFFFu; B.bend = 0; B.bpos = 0; B.brow = 0; B.bbyte = 0; B.ntie = 0; B.meta = false;
        B.tie[0] = B.tie[1] = B.tie[2] = 0;
        if (P.dbg) t_pre = clock64() - t0;
        for (int32_t rb = r_lo; rb <= r_hi; rb += P.rows_smem) {
            const int32_t nr = min(P.rows_smem, r_hi - rb + 1);
            __syncthreads();            // previous batch finished reading gain rows; col_term / tn_small ready
            for (int32_t i = threadIdx.x; i <= nr; i += THREADS) {
                if (!cached || it == 0) rs_loc[i] = __ldg(P.row_start + rb + i);
                if (i < nr) {
                    const int32_t grow = rb + i, c1 = grow / n_sub;
                    row_c1[i] = c1;
                    row_c2[i] = (grow - c1 * n_sub) * sub_w;
                    rt_local[i] = __fdiv_rn(-bump_sum(blogb, b_cnt[c1], fa0, s.logs), np);
                }
            }
            __syncthreads();
            const long long tg0 = P.dbg ? clock64() : 0;
            if (cached) {               // counts from shared memory: no global loads on this path
                for (int32_t idx = threadIdx.x; idx < nr * kS8GainStride; idx += THREADS) {
                    const int32_t i = idx >> 8, j = idx & 255;
                    const int32_t c2 = row_c2[i] + j;
                    float gv = -INFINITY;                       // slot 255 and columns beyond the sub-row
                    if (j < sub_w && c2 < k_v) {
                        const uint32_t x = cnt[idx];
                        const float tN = x < (uint32_t)kSmallCounts ? tn_small[x]
                                                                    : __fdiv_rn(bump_sum(NlogN, x, fN0, s.logs), np);
                        gv = __fadd_rn(__fadd_rn(__fadd_rn(tN, col_term[c2]), rt_local[i]), lognp);
                    }
                    gain[idx] = gv;
                }
            } else {                    // counts from the global table copy of this iteration, 4 loads in flight
                for (int32_t base = threadIdx.x; base < nr * kS8GainStride; base += 4 * THREADS) {
                    uint32_t xs[4];
                    int32_t cl[4];
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        const int32_t idx = base + u * THREADS;
                        cl[u] = -1;
                        xs[u] = 0u;
                        if (idx < nr * kS8GainStride) {
                            const int32_t i = idx >> 8, j = idx & 255;
                            const int32_t c2 = row_c2[i] + j;
                            if (j < sub_w && c2 < k_v) { cl[u] = row_c1[i] * k_v + c2; xs[u] = __ldcg(Tcur + cl[u]); }
                        }
                    }
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        const int32_t idx = base + u * THREADS;
                        if (idx < nr * kS8GainStride) {
                            float gv = -INFINITY;
                            if (cl[u] >= 0) {
                                const int32_t i = idx >> 8, j = idx & 255;
                                const uint32_t x = xs[u] + (cl[u] == prev1 ? 1u : 0u);            // pick it-1
                                const float tN = x < (uint32_t)kSmallCounts ? tn_small[x]
                                                                            : __fdiv_rn(bump_sum(NlogN, x, fN0, s.logs), np);
                                gv = __fadd_rn(__fadd_rn(__fadd_rn(tN, col_term[row_c2[i] + j]), rt_local[i]), lognp);
                            }
                            gain[idx] = gv;
                        }
                    }
                }
            }
            __syncthreads();
            if (P.dbg) t_gain += clock64() - tg0;
            // sub-rows are padded to whole blocks and chunk edges are block aligned: a block never straddles a sub-row or
            // a chunk, so the gain row is uniform per warp-iteration and there is no slow path
            const uint32_t s_lo = max(e_lo, rs_loc[0]), s_hi = min(e_hi, rs_loc[nr]);
            if (s_hi <= s_lo) continue;
            const uint32_t b_lo = s_lo / kS8Blk, b_hi = s_hi / kS8Blk;
            // each WARP walks its own contiguous span of blocks (coalesced 512-byte loads); lane l owns vector l
            const uint32_t span = (b_hi - b_lo + kWarps - 1) / kWarps;
            const uint32_t wb_lo = min(b_hi, b_lo + (threadIdx.x / kWarp) * span);
            const uint32_t wb_hi = min(b_hi, wb_lo + span);
            int32_t crow = 0;
            if (wb_lo < wb_hi) {                                 // sub-row of my first block (uniform per warp)
                const uint32_t ef = wb_lo * kS8Blk;
                int32_t a = 0, b = nr;
                while (a < b) { const int32_t m = (a + b) >> 1; if (rs_loc[m + 1] > ef) b = m; else a = m + 1; }
                crow = a;
            }
            uint32_t rend_blk = rs_loc[crow + 1] / kS8Blk;       // first block of the next sub-row
            uint32_t grow_b = gain_b + (uint32_t)crow * (kS8GainStride * 4u);
            const uint4 *src4 = reinterpret_cast<const uint4 *>(P.stream) + (size_t)wb_lo * kWarp + lane;
            uint4 stage[DEPTH];
#pragma unroll
            for (int r = 0; r < DEPTH; ++r)
                stage[r] = wb_lo + (uint32_t)r < wb_hi ? s8_ld_stream(src4 + (size_t)r * kWarp, stream_pol)
                                                       : make_uint4(0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu);
            for (uint32_t blk0 = wb_lo; blk0 < wb_hi; blk0 += DEPTH) {
#pragma unroll
                for (int r = 0; r < DEPTH; ++r) {
                    const uint32_t blk = blk0 + (uint32_t)r;
                    if (blk < wb_hi) {
                        const uint4 q = stage[r];
                        if (blk + DEPTH < wb_hi)
                            stage[r] = s8_ld_stream(src4 + (size_t)(blk - wb_lo + DEPTH) * kWarp, stream_pol);
                        while (blk >= rend_blk) {                // next non-empty sub-row (uniform per warp)
                            ++crow;
                            rend_blk = rs_loc[crow + 1] / kS8Blk;
                            grow_b = gain_b + (uint32_t)crow * (kS8GainStride * 4u);
                        }
                        float g[16];
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            g[j] = s8_gather(q.x, j, grow_b);
                            g[4 + j] = s8_gather(q.y, j, grow_b);
                            g[8 + j] = s8_gather(q.z, j, grow_b);
                            g[12 + j] = s8_gather(q.w, j, grow_b);
                        }
                        float m = fmaxf(fmaxf(fmaxf(g[0], g[1]), fmaxf(g[2], g[3])), fmaxf(fmaxf(g[4], g[5]), fmaxf(g[6], g[7])));
                        m = fmaxf(m, fmaxf(fmaxf(fmaxf(g[8], g[9]), fmaxf(g[10], g[11])),
                                           fmaxf(fmaxf(g[12], g[13]), fmaxf(g[14], g[15]))));
                        if (m >= B.bs) {                         // rare once the thread has seen a good candidate
                            const uint32_t e0 = blk * kS8Blk + lane * 16u;
                            if (m > -INFINITY && (m > B.bs || (B.bi != 0xFFFFFFFFu && e0 >= B.bend))) {
                                // of the maxima in this vector take the one that came first in the candidate list
                                uint32_t bj = 0, bp = 0xFFFFFFFFu;
#pragma unroll
                                for (int j = 0; j < 16; ++j) {
                                    if (g[j] == m) {
                                        const uint32_t pj = __ldg(P.pos_s + e0 + j);
                                        if (pj < bp) { bp = pj; bj = (uint32_t)j; }
                                    }
                                }
                                if (m > B.bs) {
                                    const uint32_t wsel = bj < 4 ? q.x : bj < 8 ? q.y : bj < 12 ? q.z : q.w;
                                    B.bs = m; B.bi = e0 + bj; B.bend = rend_blk * kS8Blk; B.ntie = 0;
                                    B.bpos = bp; B.brow = rb + crow; B.bbyte = (wsel >> (8 * (bj & 3u))) & 0xFFu; B.meta = true;
                                } else {
                                    s8_tie(B, e0 + bj, rend_blk * kS8Blk, P.pos_s);
                                }
                            }
                        }
                    }
                }
            }
        }
        const long long t1 = P.dbg ? clock64() : 0;
        // ---------------- block arg-max ----------------
        const uint32_t my32 = B.bi == 0xFFFFFFFFu ? 0u : orderable(B.bs);
        uint32_t m32 = my32;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) m32 = max(m32, __shfl_xor_sync(0xffffffffu, m32, o));
        if (lane == 0) widx[threadIdx.x / kWarp] = m32;
        __syncthreads();
        m32 = lane < kWarps ? widx[lane] : 0u;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) m32 = max(m32, __shfl_xor_sync(0xffffffffu, m32, o));
        __syncthreads();                                       // widx is reused below
        unsigned long long key = 0ull, pay = 0ull;
        uint32_t bi = B.bi;
        if (my32 != 0u && my32 == m32) {
            // this thread holds the block-maximum gain.  Common case (no parked ties): everything about its candidate
            // is in registers and -- with the counts cached -- nothing is loaded from global memory here.
            uint32_t bp = B.meta ? B.bpos : __ldg(P.pos_s + bi);
            bool meta = B.meta;
#pragma unroll
            for (int t = 0; t < 3; ++t) {
                if (t < B.ntie) {
                    const uint32_t pt = __ldg(P.pos_s + B.tie[t]);
                    if (pt < bp) { bp = pt; bi = B.tie[t]; meta = false; }
                }
            }
            int32_t row = B.brow;
            uint32_t byte = B.bbyte;
            if (!meta) {                                       // a parked tie won: look its sub-row and byte up
                byte = (uint32_t)__ldcg(reinterpret_cast<const unsigned char *>(P.stream) + bi);
                int32_t a = 0, b = k_rows;
                while (a < b) { const int32_t m = (a + b) >> 1; if (__ldg(P.row_start + m + 1) > bi) b = m; else a = m + 1; }
                row = a;
            }
            const int32_t c1 = row / n_sub;
            const uint32_t c2 = (uint32_t)(row - c1 * n_sub) * (uint32_t)sub_w + byte;
            const int32_t cell = c1 * k_v + (int32_t)c2;
            const uint32_t x = cached ? cnt[(row - r_lo) * kS8GainStride + (int32_t)byte]
                                      : __ldcg(Tcur + cell) + (cell == prev1 ? 1u : 0u);
            key = ((unsigned long long)m32 << 32) | (unsigned long long)(0xFFFFFFFFu - (base_pos + bp));
            pay = ((unsigned long long)c1 << 48) | ((unsigned long long)c2 << 32) | x;
        }
        {   // block arg-max of (key, payload, stream index); thread 0 publishes the CTA's candidate
            unsigned long long k2 = key, p2 = pay;
            uint32_t i2 = bi;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                const unsigned long long ok = __shfl_xor_sync(0xffffffffu, k2, o);
                const unsigned long long op = __shfl_xor_sync(0xffffffffu, p2, o);
                const uint32_t oi = __shfl_xor_sync(0xffffffffu, i2, o);
                if (ok > k2) { k2 = ok; p2 = op; i2 = oi; }
            }
            if (lane == 0) { wkey[threadIdx.x / kWarp] = k2; wpay[threadIdx.x / kWarp] = p2; widx[threadIdx.x / kWarp] = i2; }
            __syncthreads();
            if (threadIdx.x < kWarp) {
                const bool live = (int)threadIdx.x < kWarps;
                k2 = live ? wkey[threadIdx.x] : 0ull; p2 = live ? wpay[threadIdx.x] : 0ull; i2 = live ? widx[threadIdx.x] : 0u;
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) {
                    const unsigned long long ok = __shfl_xor_sync(0xffffffffu, k2, o);
                    const unsigned long long op = __shfl_xor_sync(0xffffffffu, p2, o);
                    const uint32_t oi = __shfl_xor_sync(0xffffffffu, i2, o);
                    if (ok > k2) { k2 = ok; p2 = op; i2 = oi; }
                }
                if (threadIdx.x == 0) {
                    sh_best_key = k2; sh_best_idx = i2;
                    MiPub *pb = P.pub + (size_t)cur * grid + blockIdx.x;
                    pb->key = k2; pb->payload = k2 ? p2 : 0ull;
                }
            }
        }
        const long long t2 = P.dbg ? clock64() : 0;
        if (grid_barrier(P.bar, grid, P.world > 1 ? P.status : nullptr)) { broke = true; break; }   // a CTA gave up on a peer GPU
        const long long t3 = P.dbg ? clock64() : 0;
        if (P.dbg && threadIdx.x == 0) {
            long long *d = P.dbg + 8 * blockIdx.x;
            d[0] = t_gain; d[1] = t1 - t0 - t_gain; d[2] = t2 - t1; d[3] = clock64() - t2;
            d[4] = (long long)(e_hi - e_lo) / kS8Blk; d[5] = n_mine; d[6] = t_pre; d[7] = t_learn;
        }
        // ---------------- everyone learns the winner ----------------
        {
            unsigned long long k2 = 0ull, p2 = 0ull;
            for (uint32_t t = threadIdx.x; t < grid; t += THREADS) {
                const MiPub *pb = P.pub + (size_t)cur * grid + t;
                const unsigned long long kk = __ldcg(&pb->key);
                const unsigned long long pp = __ldcg(&pb->payload);      // unconditional: one memory latency, not two
                if (kk > k2) { k2 = kk; p2 = pp; }
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                const unsigned long long ok = __shfl_xor_sync(0xffffffffu, k2, o);
                const unsigned long long op = __shfl_xor_sync(0xffffffffu, p2, o);
                if (ok > k2) { k2 = ok; p2 = op; }
            }
            if (lane == 0) { wkey[threadIdx.x / kWarp] = k2; wpay[threadIdx.x / kWarp] = p2; }
            __syncthreads();
            if (threadIdx.x < kWarp) {
                const bool live = (int)threadIdx.x < kWarps;
                k2 = live ? wkey[threadIdx.x] : 0ull; p2 = live ? wpay[threadIdx.x] : 0ull;
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) {
                    const unsigned long long ok = __shfl_xor_sync(0xffffffffu, k2, o);
                    const unsigned long long op = __shfl_xor_sync(0xffffffffu, p2, o);
                    if (ok > k2) { k2 = ok; p2 = op; }
                }
                if (P.world > 1) {
                    // push this GPU's winner into every rank's mailbox (NVLink stores), then lanes 0..world-1 each wait
                    // (bounded) for one rank's entry of this iteration in the local mailbox
                    const unsigned int tag = P.seq_base + (unsigned int)it + 1u;
                    k2 = __shfl_sync(0xffffffffu, k2, 0);
                    p2 = __shfl_sync(0xffffffffu, p2, 0);
                    if (blockIdx.x == 0 && (int)threadIdx.x < P.world) {
                        MiMail *m = P.mail_peer[threadIdx.x] + (size_t)cur * P.world + P.rank;
                        m->key = k2; m->payload = p2;
                        st_release_sys(&m->seq, tag);
                    }
                    unsigned long long gk = 0ull, gp = 0ull;
                    bool timed_out = false;
                    if ((int)threadIdx.x < P.world) {
                        const MiMail *m = P.mail_local + (size_t)cur * P.world + threadIdx.x;
                        timed_out = !wait_peer_tag(&m->seq, tag, P.spin_limit_ns);
                        gk = *reinterpret_cast<const volatile unsigned long long *>(&m->key);
                        gp = *reinterpret_cast<const volatile unsigned long long *>(&m->payload);
                    }
                    if (__any_sync(0xffffffffu, timed_out)) {      // a peer never delivered: stop here, say why
                        gk = 0ull; gp = 0ull;
                        if (threadIdx.x == 0) *reinterpret_cast<volatile int *>(P.status) = kMiRunPeerTimeout;
                    }
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) {
                        const unsigned long long ok = __shfl_xor_sync(0xffffffffu, gk, o);
                        const unsigned long long op = __shfl_xor_sync(0xffffffffu, gp, o);
                        if (ok > gk) { gk = ok; gp = op; }
                    }
                    k2 = gk; p2 = gp;
                }
                if (threadIdx.x == 0) { sh_win_key = k2; sh_win_pay = p2; }
            }
            __syncthreads();
        }
        const unsigned long long win = sh_win_key, wpayload = sh_win_pay;
        if (win == 0ull) { broke = true; break; }              // nothing left on any rank
        const int32_t c1 = (int32_t)(wpayload >> 48), c2w = (int32_t)((wpayload >> 32) & 0xFFFFu);
        if (threadIdx.x == 0) {
            if (sh_best_key == win) {                          // keys are unique: exactly one owner CTA
                P.stream[sh_best_idx] = (uint8_t)kS8Removed;   // remove_idx_all mi.py:104-106 (the -inf slot)
                s.cells[(int64_t)key_pos(win) - s.pos_base] = 0xFFFFFFFFu;     // list-order view stays in sync
            }
            const uint32_t x = (uint32_t)(wpayload & 0xFFFFFFFFull), y = a_cnt[c2w], z = b_cnt[c1];
            ps[0] = bump_sum(ps[0], x, ps[4], s.logs);         // update_cache mi.py:383-389
            ps[1] = bump_sum(ps[1], y, ps[5], s.logs);
            ps[2] = bump_sum(ps[2], z, ps[5], s.logs);
            ps[3] = __fadd_rn(ps[3], 1.0f);                    // update_mats :401-406
            a_cnt[c2w] = y + 1; b_cnt[c1] = z + 1;
            if (cached) {                                      // my copy of the cell's count, if the cell is mine
                const int32_t sub = c2w / sub_w, wrow = c1 * n_sub + sub;
                if (wrow >= r_lo && wrow <= r_hi) cnt[(wrow - r_lo) * kS8GainStride + (c2w - sub * sub_w)] = x + 1;
            }
            if (blockIdx.x == 0) {
                P.out_pos[it] = (int64_t)key_pos(win);
                P.out_gain[it] = key_score(win);
            }
        }
        prev2 = prev1;
        prev1 = c1 * k_v + c2w;
        done = it + 1;
        __syncthreads();
        if (P.dbg) t_learn = clock64() - t3;
    }
    // ---------------- write the replicated state back (CTA 0) ----------------
    if (blockIdx.x == 0) {
        if (threadIdx.x == 0) {
            for (int64_t j = done; j < P.n_picks; ++j) { P.out_pos[j] = -1; P.out_gain[j] = nanf(""); }
            // canonical copy 0 holds picks <= done-2 if `done` is even (it was the read copy of iteration `done`),
            // <= done-3 if odd (an early exit at an odd `done` happened after CTA 0 had already completed copy 0)
            if (!(broke && (done & 1))) {
                if (prev1 >= 0) s.n_cells[prev1] += 1;
                if ((done & 1) && prev2 >= 0) s.n_cells[prev2] += 1;
            }
            for (int i = 0; i < 4; ++i) s.sums[i] = ps[i];
        }
        for (int32_t i = threadIdx.x; i < k_v; i += THREADS) s.a_cols[i] = a_cnt[i];
        for (int32_t i = threadIdx.x; i < k_a; i += THREADS) s.b_rows[i] = b_cnt[i];
    }
}

size_t s8_fixed_bytes(int32_t k_a, int32_t k_v, int32_t rows) {
    const size_t words = 2 * (size_t)k_v + kSmallCounts + (size_t)k_a + ((size_t)rows + 1) + 3 * (size_t)rows;
    return words * 4;
}

}  // namespace

int mi_s8_k_rows(int32_t k_a, int32_t k_v) { return s8_geom(k_a, k_v).k_rows; }
int mi_s8_block() { return kS8Blk; }
int mi_s8_tiles(int64_t w) { return (int)ceil_div(w > 0 ? w : 1, kS8Tile); }
int64_t mi_s8_stream_capacity(int64_t w, int32_t k_a, int32_t k_v) {
    return w + (int64_t)kS8Blk * s8_geom(k_a, k_v).k_rows + kS8Blk;
}

int mi_s8_rows_that_fit(int32_t k_a, int32_t k_v) {
    int32_t rows = 0;
    while (rows < 4096 &&
           s8_fixed_bytes(k_a, k_v, rows + 1) + 1024 + (size_t)(rows + 1) * kS8GainStride * 8 <= kS8SmemBudget)
        ++rows;
    return rows;
}

bool mi_s8_supported(int32_t k_a, int32_t k_v) {
    const S8Geom g = s8_geom(k_a, k_v);
    return g.k_rows <= 24000 && mi_s8_rows_that_fit(k_a, k_v) >= 1 && (int64_t)k_a * k_v < (1ll << 31);   // partition histogram: 96 KB of shared memory
}

int launch_mi_s8_partition(const uint32_t *cells, int64_t w, int32_t k_a, int32_t k_v, uint32_t *tilehist,
                           uint32_t *row_total, uint32_t *row_start, uint8_t *stream, uint32_t *pos_s,
                           int64_t stream_capacity, cudaStream_t st) {
    const S8Geom g = s8_geom(k_a, k_v);
    const int ntiles = mi_s8_tiles(w);
    const size_t smem = (size_t)g.k_rows * sizeof(uint32_t);
    if (smem > 96 * 1024) return ACAV_E_UNSUPPORTED;
    static size_t done_count[kMaxDevices], done_scatter[kMaxDevices];
    if (smem > 48 * 1024) {
        int rc = ensure_dynamic_smem(s8_count_kernel, (size_t)96 * 1024, done_count);
        if (!rc) rc = ensure_dynamic_smem(s8_scatter_kernel, (size_t)96 * 1024, done_scatter);
        if (rc) return rc;
    }
    s8_count_kernel<<<ntiles, kS8PartThreads, smem, st>>>(cells, w, g, tilehist);
    ACAV_LAUNCH_CHECK();
    s8_prefix_kernel<<<(unsigned)ceil_div(g.k_rows, 256), 256, 0, st>>>(tilehist, ntiles, g.k_rows, row_total);
    ACAV_LAUNCH_CHECK();
    s8_rowstart_kernel<<<1, 1024, 0, st>>>(row_total, g.k_rows, row_start);
    ACAV_LAUNCH_CHECK();
    // padding entries of every sub-row: the "removed" byte (stream_capacity is a multiple of 16)
    const int64_t n16 = stream_capacity / 16;
    s8_fill_kernel<<<(unsigned)ceil_div(n16, 256), 256, 0, st>>>(reinterpret_cast<uint4 *>(stream), n16, 0xFFFFFFFFu);
    ACAV_LAUNCH_CHECK();
    s8_scatter_kernel<<<ntiles, kS8PartThreads, smem, st>>>(cells, w, g, tilehist, row_start, stream, pos_s);
    ACAV_LAUNCH_CHECK();
    return 0;
}

int launch_mi_s8_block_sort(uint8_t *stream, uint32_t *pos_s, int64_t w_padded, cudaStream_t st) {
    const int64_t n_blocks = w_padded / kS8Blk;
    if (n_blocks == 0) return 0;
    const unsigned grid = (unsigned)(n_blocks < 65535 * 16 ? n_blocks : 65535 * 16);
    s8_block_sort_kernel<<<grid, kS8Blk, 0, st>>>(stream, pos_s, n_blocks);
    ACAV_LAUNCH_CHECK();
    return 0;
}

template <int THREADS, int DEPTH>
static int launch_s8_variant(const MiS8 &P, size_t smem, int32_t grid, cudaStream_t st) {
    static size_t attr_done[kMaxDevices];
    { int rc = ensure_dynamic_smem(mi_stream8_kernel<THREADS, DEPTH>, smem, attr_done); if (rc) return rc; }
    MiS8 Pc = P;
    void *args[] = {&Pc};
    ACAV_CUDA_TRY(cudaLaunchCooperativeKernel((void *)mi_stream8_kernel<THREADS, DEPTH>, dim3(grid), dim3(THREADS), args,
                                              smem, st));
    return 0;
}

int launch_mi_stream8(const MiState &s, uint32_t *n_alt, uint8_t *stream, const uint32_t *pos_s,
                      const uint32_t *row_start, const uint32_t *chunk_start, int32_t grid, void *pub,
                      unsigned int *bar, int64_t n_picks, int64_t *out_pos, float *out_gain, int32_t rows_smem,
                      int32_t variant, int32_t use_cache, int32_t world, int32_t rank, unsigned int seq_base,
                      void *mail_local, void *const *mail_peer, long long *dbg, int *status,
                      unsigned long long spin_limit_ns, cudaStream_t st) {
    MiS8 P;
    P.s = s; P.n_alt = n_alt; P.stream = stream; P.pos_s = pos_s; P.row_start = row_start; P.chunk_start = chunk_start;
    P.pub = reinterpret_cast<MiPub *>(pub); P.bar = bar; P.n_picks = n_picks; P.out_pos = out_pos; P.out_gain = out_gain;
    P.g = s8_geom(s.k_a, s.k_v); P.rows_smem = rows_smem; P.use_cache = use_cache;
    P.fixed_bytes = (int32_t)s8_fixed_bytes(s.k_a, s.k_v, rows_smem);
    P.world = world; P.rank = rank; P.seq_base = seq_base;
    P.mail_local = reinterpret_cast<MiMail *>(mail_local);
    for (int r = 0; r < kMaxWorld; ++r)
        P.mail_peer[r] = (world > 1 && r < world) ? reinterpret_cast<MiMail *>(mail_peer[r]) : nullptr;
    P.dbg = dbg; P.status = status; P.spin_limit_ns = spin_limit_ns;
    const size_t smem = (size_t)P.fixed_bytes + 1024 + (size_t)rows_smem * kS8GainStride * 8;
    // both table copies start equal; the barrier words start at zero
    ACAV_CUDA_TRY(cudaMemcpyAsync(n_alt, s.n_cells, sizeof(uint32_t) * (size_t)s.k_a * s.k_v, cudaMemcpyDeviceToDevice, st));
    ACAV_CUDA_TRY(cudaMemsetAsync(pub, 0, mi_pub_bytes(grid), st));
    ACAV_CUDA_TRY(cudaMemsetAsync(bar, 0, 2 * sizeof(unsigned int), st));
    ACAV_CUDA_TRY(cudaMemsetAsync(status, 0, sizeof(int), st));
    switch (variant) {
        case 1: return launch_s8_variant<512, 8>(P, smem, grid, st);
        case 2: return launch_s8_variant<768, 4>(P, smem, grid, st);
        case 3: return launch_s8_variant<1024, 2>(P, smem, grid, st);
        case 4: return launch_s8_variant<512, 4>(P, smem, grid, st);
        default: return launch_s8_variant<512, 6>(P, smem, grid, st);
    }
}

}  // namespace acav
