// Greedy-MI selection over a CELL INDEX: the same picks and fp32 gains as the candidate-stream loops
// (mi_scan.cu, mi_persistent.cu) and the reference (subset_selection/code/measures/mi.py:150-192, 284-412),
// at O(K_a * K_v) work per iteration instead of O(remaining candidates).
//
// The reference scores every remaining candidate, but the score depends on the candidate only through its
// table cell (c1, c2) (mi.py:322-381): all candidates of a cell tie, and `max` returns the first of them
// (:79).  So the arg-max over candidates is the arg-max over NON-EMPTY cells of (gain(cell), earliest
// remaining candidate of the cell).  Built once: a stable LSD counting sort of the candidate list by (c1, c2)
// (two 16-bit passes), giving every cell its candidates in list order; per cell a head counter and the
// position of its first remaining candidate.  Per iteration every CTA scans the cells of the table rows it
// owns (count + first position: 8 bytes per cell, L2 resident), publishes its best, one grid barrier, every
// CTA learns the winner and applies it to its replicated marginals / running sums; the CTA that owns the
// winner's row bumps the cell count, the CTA that published the winner pops the cell's head.
// Multi-GPU: candidates sharded by position range exactly as in mi_persistent.cu (same NVLink mailbox).
#include <cooperative_groups.h>

#include "common.cuh"
#include "kernels.cuh"
#include "mi_loop.cuh"

namespace acav {

constexpr int kCellTile = 32768;             // elements per sort tile
constexpr int kCellSortThreads = 512;
constexpr int kCellThreads = 1024;
constexpr uint32_t kNoPos = 0xFFFFFFFFu;

// ---- stable counting sort by a 16-bit digit of the packed cell --------------------------------------

__device__ __forceinline__ uint32_t digit_of(uint32_t cell, int shift) { return (cell >> shift) & 0xFFFFu; }

__global__ void __launch_bounds__(kCellSortThreads)
cells_count_kernel(const uint32_t *__restrict__ cells, int64_t w, int shift, int32_t k,
                   uint32_t *__restrict__ tilehist) {
    extern __shared__ uint32_t hist[];
    for (int32_t i = threadIdx.x; i < k; i += blockDim.x) hist[i] = 0;
    __syncthreads();
    const int64_t lo = (int64_t)blockIdx.x * kCellTile;
    const int64_t hi = min(w, lo + kCellTile);
    for (int64_t e = lo + threadIdx.x; e < hi; e += blockDim.x) {
        const uint32_t c = cells[e];
        if (c != 0xFFFFFFFFu) atomicAdd(&hist[digit_of(c, shift)], 1u);      // removed entries are dropped
    }
    __syncthreads();
    uint32_t *dst = tilehist + (int64_t)blockIdx.x * k;
    for (int32_t i = threadIdx.x; i < k; i += blockDim.x) dst[i] = hist[i];
}

// per digit value: exclusive prefix over tiles (in place) and the total
__global__ void cells_prefix_kernel(uint32_t *__restrict__ tilehist, int32_t ntiles, int32_t k,
                                    uint32_t *__restrict__ total) {
    const int32_t r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= k) return;
    uint32_t run = 0;
    for (int32_t t = 0; t < ntiles; ++t) {
        const uint32_t v = tilehist[(int64_t)t * k + r];
        tilehist[(int64_t)t * k + r] = run;
        run += v;
    }
    total[r] = run;
}

// start[0..k] = exclusive scan of total (single block)
__global__ void __launch_bounds__(1024)
cells_scan_kernel(const uint32_t *__restrict__ total, int32_t k, uint32_t *__restrict__ start) {
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
        if (i < k) start[i] = excl;
        __syncthreads();
        if (threadIdx.x == 1023) carry = excl + v;
        __syncthreads();
    }
    if (threadIdx.x == 0) start[k] = carry;
}

// stable scatter: tiles in list order, chunks of 512 in list order, warps in order, lanes in order.
// pos_in == nullptr: the payload is the element's own index (first pass).
__global__ void __launch_bounds__(kCellSortThreads)
cells_scatter_kernel(const uint32_t *__restrict__ cells, const uint32_t *__restrict__ pos_in, int64_t w, int shift,
                     int32_t k, const uint32_t *__restrict__ tilehist, const uint32_t *__restrict__ start,
                     uint32_t *__restrict__ cells_out, uint32_t *__restrict__ pos_out) {
    extern __shared__ uint32_t cursor[];
    const uint32_t *tp = tilehist + (int64_t)blockIdx.x * k;
    for (int32_t i = threadIdx.x; i < k; i += blockDim.x) cursor[i] = start[i] + tp[i];
    __syncthreads();
    const int64_t lo = (int64_t)blockIdx.x * kCellTile;
    const int64_t hi = min(w, lo + kCellTile);
    const int warp = threadIdx.x / kWarp, lane = threadIdx.x % kWarp;
    for (int64_t base = lo; base < hi; base += kCellSortThreads) {
        const int64_t e = base + threadIdx.x;
        uint32_t cell = 0xFFFFFFFFu;
        if (e < hi) cell = cells[e];
        const bool live = cell != 0xFFFFFFFFu;
        const int32_t key = live ? (int32_t)digit_of(cell, shift) : -1;
        const uint32_t payload = live ? (pos_in ? pos_in[e] : (uint32_t)e) : 0u;
        for (int ww = 0; ww < kCellSortThreads / kWarp; ++ww) {
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
                    cells_out[basev + rank] = cell;
                    pos_out[basev + rank] = payload;
                }
            }
            __syncthreads();
        }
    }
}

// cell_start[c] = index of the first sorted entry with cell id >= c  (c = c1 * k_v + c2), cell_start[K] = n
__global__ void cells_start_kernel(const uint32_t *__restrict__ sorted, int64_t n, int32_t k_v, int64_t n_cells,
                                   uint32_t *__restrict__ cell_start) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i > n) return;
    const int64_t cur = i < n ? (int64_t)(sorted[i] >> 16) * k_v + (sorted[i] & 0xFFFFu) : n_cells;
    const int64_t prev = i > 0 ? (int64_t)(sorted[i - 1] >> 16) * k_v + (sorted[i - 1] & 0xFFFFu) : -1;
    for (int64_t c = prev + 1; c <= cur; ++c) cell_start[c] = (uint32_t)i;
}

__global__ void cells_init_state_kernel(const uint32_t *__restrict__ cell_start, const uint32_t *__restrict__ sorted_pos,
                                        int64_t n_cells, uint32_t *__restrict__ head, uint32_t *__restrict__ first_pos) {
    const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= n_cells) return;
    head[c] = 0;
    first_pos[c] = cell_start[c + 1] > cell_start[c] ? sorted_pos[cell_start[c]] : kNoPos;
}

// ---- the loop -----------------------------------------------------------------------------------------

struct MiCells {
    MiState s;
    const uint32_t *cell_start;      // [k_a * k_v + 1]
    const uint32_t *sorted_pos;      // local positions of the candidates, sorted by cell, list order inside a cell
    uint32_t *head;                  // [k_a * k_v] candidates of the cell already selected
    uint32_t *first_pos;             // [k_a * k_v] local position of the first remaining candidate, kNoPos = none
    MiPub *pub;                      // [2][grid]
    unsigned int *bar;
    int64_t n_picks;
    int64_t *out_pos;
    float *out_gain;
    int32_t rows_per_cta;
    int32_t cache_rows;              // 1: the CTA keeps counts + first positions of its rows in shared memory
    int32_t world, rank;
    unsigned int seq_base;
    MiMail *mail_local;
    MiMail *mail_peer[kMaxWorld];
    int *status;                     // kMiRun* word of this launch (device)
    unsigned long long spin_limit_ns;
};

__global__ void __launch_bounds__(kCellThreads, 1) mi_cells_kernel(MiCells P) {
    extern __shared__ __align__(16) unsigned char csmem[];
    const MiState &s = P.s;
    const int32_t k_v = s.k_v, k_a = s.k_a;
    float *col_term = reinterpret_cast<float *>(csmem);                             // [k_v]
    float *tn_small = col_term + k_v;                                               // [kSmallCounts]
    uint32_t *a_cnt = reinterpret_cast<uint32_t *>(tn_small + kSmallCounts);        // [k_v] column marginals
    uint32_t *b_cnt = a_cnt + k_v;                                                  // [k_a] row marginals
    float *rt_local = reinterpret_cast<float *>(b_cnt + k_a);                       // [rows_per_cta]
    uint32_t *n_loc = reinterpret_cast<uint32_t *>(rt_local + P.rows_per_cta);      // [rows_per_cta * k_v] if cached
    uint32_t *fp_loc = n_loc + (P.cache_rows ? P.rows_per_cta * k_v : 0);           // [rows_per_cta * k_v] if cached
    __shared__ unsigned long long wkey[32];
    __shared__ unsigned long long wpay[32];
    __shared__ unsigned long long sh_best_key, sh_win_key, sh_win_pay;
    __shared__ float ps[6];                                   // {NlogN, aloga, blogb, n, fN0, fa0}

    for (int32_t i = threadIdx.x; i < k_v; i += blockDim.x) a_cnt[i] = __ldcg(s.a_cols + i);
    for (int32_t i = threadIdx.x; i < k_a; i += blockDim.x) b_cnt[i] = __ldcg(s.b_rows + i);
    if (threadIdx.x < 6) ps[threadIdx.x] = __ldcg(s.sums + threadIdx.x);
    __syncthreads();

    const int32_t r_lo = min(k_a, (int32_t)blockIdx.x * P.rows_per_cta);
    const int32_t r_hi = min(k_a, r_lo + P.rows_per_cta);                 // my table rows [r_lo, r_hi)
    const int32_t n_my = (r_hi - r_lo) * k_v;
    if (P.cache_rows) {                                                   // only this CTA ever touches these cells
        for (int32_t j = threadIdx.x; j < n_my; j += kCellThreads) {
            n_loc[j] = __ldcg(s.n_cells + (int64_t)r_lo * k_v + j);
            fp_loc[j] = __ldcg(P.first_pos + (int64_t)r_lo * k_v + j);
        }
        __syncthreads();
    }
    const uint32_t base_pos = (uint32_t)s.pos_base;
    const uint32_t grid = gridDim.x;
    int64_t done = 0;

    for (int64_t it = 0; it < P.n_picks; ++it) {
        const int cur = (int)(it & 1);
        const float NlogN = ps[0], aloga = ps[1], blogb = ps[2], fN0 = ps[4], fa0 = ps[5];
        const float np = __fadd_rn(ps[3], 1.0f);
        const float lognp = __ldg(s.logs + (int64_t)np);
        for (int32_t i = threadIdx.x; i < k_v; i += blockDim.x)
            col_term[i] = __fdiv_rn(-bump_sum(aloga, a_cnt[i], fa0, s.logs), np);
        for (int32_t i = threadIdx.x; i < kSmallCounts; i += blockDim.x)
            tn_small[i] = __fdiv_rn(bump_sum(NlogN, (uint32_t)i, fN0, s.logs), np);
        for (int32_t i = threadIdx.x; i < r_hi - r_lo; i += blockDim.x)
            rt_local[i] = __fdiv_rn(-bump_sum(blogb, b_cnt[r_lo + i], fa0, s.logs), np);
        __syncthreads();
        // ---------------- score my cells ----------------
        unsigned long long key = 0ull, pay = 0ull;
        for (int32_t j = threadIdx.x; j < n_my; j += kCellThreads) {
            const int32_t rr = j / k_v, c2 = j - rr * k_v;
            const int32_t c1 = r_lo + rr;
            const int64_t cell = (int64_t)c1 * k_v + c2;
            const uint32_t fp = P.cache_rows ? fp_loc[j] : __ldcg(P.first_pos + cell);
            if (fp == kNoPos) continue;
            const uint32_t x = P.cache_rows ? n_loc[j] : __ldcg(s.n_cells + cell);
            const float tN = x < (uint32_t)kSmallCounts ? tn_small[x] : __fdiv_rn(bump_sum(NlogN, x, fN0, s.logs), np);
            const float g = __fadd_rn(__fadd_rn(__fadd_rn(tN, col_term[c2]), rt_local[rr]), lognp);
            const unsigned long long kk = make_key(g, base_pos + fp);
            if (kk > key) { key = kk; pay = ((unsigned long long)c1 << 48) | ((unsigned long long)c2 << 32) | x; }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const unsigned long long ok = __shfl_xor_sync(0xffffffffu, key, o);
            const unsigned long long op = __shfl_xor_sync(0xffffffffu, pay, o);
            if (ok > key) { key = ok; pay = op; }
        }
        if (threadIdx.x % kWarp == 0) { wkey[threadIdx.x / kWarp] = key; wpay[threadIdx.x / kWarp] = pay; }
        __syncthreads();
        if (threadIdx.x < kWarp) {
            key = wkey[threadIdx.x]; pay = wpay[threadIdx.x];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                const unsigned long long ok = __shfl_xor_sync(0xffffffffu, key, o);
                const unsigned long long op = __shfl_xor_sync(0xffffffffu, pay, o);
                if (ok > key) { key = ok; pay = op; }
            }
            if (threadIdx.x == 0) {
                sh_best_key = key;
                MiPub *pb = P.pub + (size_t)cur * grid + blockIdx.x;
                pb->key = key; pb->payload = pay;
            }
        }
        if (grid_barrier(P.bar, grid, P.world > 1 ? P.status : nullptr)) break;      // a CTA gave up on a peer GPU
        // ---------------- everyone learns the winner ----------------
        {
            unsigned long long k2 = 0ull, p2 = 0ull;
            for (uint32_t t = threadIdx.x; t < grid; t += blockDim.x) {
                const MiPub *pb = P.pub + (size_t)cur * grid + t;
                const unsigned long long kk = __ldcg(&pb->key);
                const unsigned long long pp = __ldcg(&pb->payload);
                if (kk > k2) { k2 = kk; p2 = pp; }
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                const unsigned long long ok = __shfl_xor_sync(0xffffffffu, k2, o);
                const unsigned long long op = __shfl_xor_sync(0xffffffffu, p2, o);
                if (ok > k2) { k2 = ok; p2 = op; }
            }
            __syncthreads();                                   // wkey / wpay of the block reduce are free again
            if (threadIdx.x % kWarp == 0) { wkey[threadIdx.x / kWarp] = k2; wpay[threadIdx.x / kWarp] = p2; }
            __syncthreads();
            if (threadIdx.x < kWarp) {
                k2 = wkey[threadIdx.x]; p2 = wpay[threadIdx.x];
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) {
                    const unsigned long long ok = __shfl_xor_sync(0xffffffffu, k2, o);
                    const unsigned long long op = __shfl_xor_sync(0xffffffffu, p2, o);
                    if (ok > k2) { k2 = ok; p2 = op; }
                }
                if (P.world > 1) {
                    // push this GPU's winner into every rank's mailbox (NVLink stores), then lanes 0..world-1
                    // each wait for one rank's entry of this iteration in the local mailbox
                    const unsigned int tag = P.seq_base + (unsigned int)it + 1u;
                    k2 = __shfl_sync(0xffffffffu, k2, 0);
                    p2 = __shfl_sync(0xffffffffu, p2, 0);
                    if (blockIdx.x == 0 && (int)threadIdx.x < P.world) {
                        mail_store(P.mail_peer[threadIdx.x] + (size_t)cur * P.world + P.rank, k2, p2, tag);
                    }
                    unsigned long long gk = 0ull, gp = 0ull;
                    bool timed_out = false;
                    if ((int)threadIdx.x < P.world) {
                        timed_out = !mail_wait(P.mail_local + (size_t)cur * P.world + threadIdx.x, tag, P.spin_limit_ns, gk, gp);
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
        if (win == 0ull) break;                                // nothing left on any rank
        const int32_t c1 = (int32_t)(wpayload >> 48), c2w = (int32_t)((wpayload >> 32) & 0xFFFFu);
        if (threadIdx.x == 0) {
            const int64_t cell = (int64_t)c1 * k_v + c2w;
            if (sh_best_key == win) {                          // keys are unique: exactly one publishing CTA (on one rank)
                const uint32_t h = __ldcg(P.head + cell) + 1u;             // remove_idx_all mi.py:104-106
                const uint32_t lo = __ldg(P.cell_start + cell), hi = __ldg(P.cell_start + cell + 1);
                const uint32_t nfp = lo + h < hi ? __ldg(P.sorted_pos + lo + h) : kNoPos;
                __stcg(P.head + cell, h);
                __stcg(P.first_pos + cell, nfp);
                if (P.cache_rows) fp_loc[cell - (int64_t)r_lo * k_v] = nfp;     // the publisher owns the cell's row
                s.cells[(int64_t)key_pos(win) - s.pos_base] = 0xFFFFFFFFu;  // list-order view stays in sync
            }
            const uint32_t x = (uint32_t)(wpayload & 0xFFFFFFFFull), y = a_cnt[c2w], z = b_cnt[c1];
            if (c1 >= r_lo && c1 < r_hi) {                     // update_mats :401-406 (row owner, on every rank)
                __stcg(s.n_cells + cell, x + 1);
                if (P.cache_rows) n_loc[cell - (int64_t)r_lo * k_v] = x + 1;
            }
            ps[0] = bump_sum(ps[0], x, ps[4], s.logs);         // update_cache mi.py:383-389
            ps[1] = bump_sum(ps[1], y, ps[5], s.logs);
            ps[2] = bump_sum(ps[2], z, ps[5], s.logs);
            ps[3] = __fadd_rn(ps[3], 1.0f);
            a_cnt[c2w] = y + 1; b_cnt[c1] = z + 1;
            if (blockIdx.x == 0) {
                P.out_pos[it] = (int64_t)key_pos(win);
                P.out_gain[it] = key_score(win);
            }
        }
        done = it + 1;
        __syncthreads();
    }
    // ---------------- write the replicated state back (CTA 0) ----------------
    if (blockIdx.x == 0) {
        if (threadIdx.x == 0) {
            for (int64_t j = done; j < P.n_picks; ++j) { P.out_pos[j] = -1; P.out_gain[j] = nanf(""); }
            for (int i = 0; i < 4; ++i) s.sums[i] = ps[i];
        }
        for (int32_t i = threadIdx.x; i < k_v; i += blockDim.x) s.a_cols[i] = a_cnt[i];
        for (int32_t i = threadIdx.x; i < k_a; i += blockDim.x) s.b_rows[i] = b_cnt[i];
    }
}

// ---- host side ------------------------------------------------------------------------------------------

int mi_cells_tiles(int64_t w) { return (int)ceil_div(w > 0 ? w : 1, kCellTile); }

// replicated marginals + per-CTA row terms must fit in shared memory (the per-row caches are optional)
bool mi_cells_smem_fits(int32_t k_a, int32_t k_v) {
    const size_t smem = ((size_t)2 * k_v + kSmallCounts + (size_t)k_a + (size_t)k_a / 64 + 1) * 4 + 16;   // rows per CTA <= k_a / 64 + 1
    return smem <= 200 * 1024;
}

static int cells_sort_pass(const uint32_t *cells_in, const uint32_t *pos_in, int64_t w, int shift, int32_t k,
                           uint32_t *tilehist, uint32_t *total, uint32_t *start, uint32_t *cells_out,
                           uint32_t *pos_out, cudaStream_t st) {
    const int ntiles = mi_cells_tiles(w);
    const size_t smem = (size_t)k * sizeof(uint32_t);
    if (smem > 200 * 1024) return ACAV_E_UNSUPPORTED;
    static size_t done_count[kMaxDevices], done_scatter[kMaxDevices];
    if (smem > 48 * 1024) {
        int rc = ensure_dynamic_smem(cells_count_kernel, smem, done_count);
        if (!rc) rc = ensure_dynamic_smem(cells_scatter_kernel, smem, done_scatter);
        if (rc) return rc;
    }
    cells_count_kernel<<<ntiles, kCellSortThreads, smem, st>>>(cells_in, w, shift, k, tilehist);
    ACAV_LAUNCH_CHECK();
    cells_prefix_kernel<<<(unsigned)ceil_div(k, 256), 256, 0, st>>>(tilehist, ntiles, k, total);
    ACAV_LAUNCH_CHECK();
    cells_scan_kernel<<<1, 1024, 0, st>>>(total, k, start);
    ACAV_LAUNCH_CHECK();
    cells_scatter_kernel<<<ntiles, kCellSortThreads, smem, st>>>(cells_in, pos_in, w, shift, k, tilehist, start,
                                                                   cells_out, pos_out);
    ACAV_LAUNCH_CHECK();
    return 0;
}

// Build the cell index from the list-order view (removed entries are skipped).  tmp_* / sorted_* hold w entries.
int launch_mi_cells_build(const MiState &s, uint32_t *tilehist, uint32_t *total, uint32_t *start, uint32_t *tmp_cells,
                          uint32_t *tmp_pos, uint32_t *sorted_cells, uint32_t *sorted_pos, uint32_t *cell_start,
                          uint32_t *head, uint32_t *first_pos, int64_t *n_live_host, cudaStream_t st) {
    const int32_t kmax = s.k_a > s.k_v ? s.k_a : s.k_v;
    int rc = cells_sort_pass(s.cells, nullptr, s.w, 0, s.k_v, tilehist, total, start, tmp_cells, tmp_pos, st);
    if (rc) return rc;
    uint32_t n_live = 0;
    ACAV_CUDA_TRY(cudaMemcpyAsync(&n_live, start + s.k_v, sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
    ACAV_CUDA_TRY(cudaStreamSynchronize(st));
    rc = cells_sort_pass(tmp_cells, tmp_pos, n_live, 16, s.k_a, tilehist, total, start, sorted_cells, sorted_pos, st);
    if (rc) return rc;
    const int64_t n_cells = (int64_t)s.k_a * s.k_v;
    cells_start_kernel<<<(unsigned)ceil_div((int64_t)n_live + 1, 256), 256, 0, st>>>(sorted_cells, n_live, s.k_v, n_cells,
                                                                                       cell_start);
    ACAV_LAUNCH_CHECK();
    cells_init_state_kernel<<<(unsigned)ceil_div(n_cells, 256), 256, 0, st>>>(cell_start, sorted_pos, n_cells, head,
                                                                                first_pos);
    ACAV_LAUNCH_CHECK();
    *n_live_host = n_live;
    (void)kmax;
    return 0;
}

int launch_mi_cells(const MiState &s, const uint32_t *cell_start, const uint32_t *sorted_pos, uint32_t *head,
                    uint32_t *first_pos, int32_t grid, void *pub, unsigned int *bar, int64_t n_picks,
                    int64_t *out_pos, float *out_gain, int32_t world, int32_t rank, unsigned int seq_base,
                    void *mail_local, void *const *mail_peer, int *status, unsigned long long spin_limit_ns,
                    cudaStream_t st, bool sync_clean) {
    MiCells P;
    P.status = status; P.spin_limit_ns = spin_limit_ns;
    P.s = s; P.cell_start = cell_start; P.sorted_pos = sorted_pos; P.head = head; P.first_pos = first_pos;
    P.pub = reinterpret_cast<MiPub *>(pub); P.bar = bar; P.n_picks = n_picks; P.out_pos = out_pos; P.out_gain = out_gain;
    P.rows_per_cta = (int32_t)ceil_div(s.k_a, grid);
    P.world = world; P.rank = rank; P.seq_base = seq_base;
    P.mail_local = reinterpret_cast<MiMail *>(mail_local);
    for (int r = 0; r < kMaxWorld; ++r)
        P.mail_peer[r] = (world > 1 && r < world) ? reinterpret_cast<MiMail *>(mail_peer[r]) : nullptr;
    size_t smem = ((size_t)2 * s.k_v + kSmallCounts + (size_t)s.k_a + P.rows_per_cta) * 4 + 16;
    if (smem > 200 * 1024) return ACAV_E_UNSUPPORTED;
    const size_t cache = (size_t)P.rows_per_cta * s.k_v * 8;
    P.cache_rows = smem + cache <= 200 * 1024 ? 1 : 0;
    if (P.cache_rows) smem += cache;
    static size_t attr_done[kMaxDevices];
    if (smem > 48 * 1024) { int rc = ensure_dynamic_smem(mi_cells_kernel, smem, attr_done); if (rc) return rc; }
    // the barrier words and the records start at zero: mi_refresh_kernel leaves them so after every run (sync_clean);
    // the status word can only be set by a peer timeout
    if (!sync_clean) {
        ACAV_CUDA_TRY(cudaMemsetAsync(pub, 0, mi_pub_bytes(grid), st));
        ACAV_CUDA_TRY(cudaMemsetAsync(bar, 0, 2 * sizeof(unsigned int), st));
    }
    if (world > 1 || !sync_clean) ACAV_CUDA_TRY(cudaMemsetAsync(status, 0, sizeof(int), st));
    void *args[] = {&P};
    ACAV_CUDA_TRY(cudaLaunchCooperativeKernel((void *)mi_cells_kernel, dim3(grid), dim3(kCellThreads), args, smem, st));
    return 0;
}

}  // namespace acav
