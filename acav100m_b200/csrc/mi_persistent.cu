// Persistent greedy-MI kernel: the whole selection loop of EfficientMI.run_greedy
// (subset_selection/code/measures/mi.py:150-192 with EfficientMemMI :284-412) in ONE cooperative launch.
//
// Data layout (built once per engine by the partition kernels below, DESIGN.md "MI layout"):
//   * candidates are STABLY partitioned by table row c1; the stream holds only c2 as uint16
//     (2 bytes per candidate per iteration instead of the reference's 16-byte int64 pair), stored as the
//     byte offset 4*c2 into a gain row; 4*K_v = removed (that slot holds -inf: no branch in the hot loop).  Inside a row the stream keeps list order, so "first maximum wins" (mi.py:79) is
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
#include "mi_loop.cuh"

namespace acav {

constexpr int kPersistThreads = 1024;
constexpr int kTileElems = 32768;            // elements per partition tile
constexpr int kPartThreads = 512;
constexpr int kRing = 4;                     // 16-byte stream loads in flight per thread (cp.async ring)
constexpr int kBlk = 256;                    // stream block: 32 lanes x one 16-byte vector of 8 candidates; rows are
                                             // padded to whole blocks and each block is sorted by shared-memory bank

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
        const uint32_t v = i < k ? ((total[i] + (kBlk - 1)) & ~(uint32_t)(kBlk - 1)) : 0u;    // rows padded to blocks
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
                    c2s[basev + rank] = (uint16_t)((cell & 0xFFFFu) << 2);      // byte offset into a gain row
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

// Sort every 256-candidate block of the (row-partitioned, row-padded) stream by (shared-memory bank of its gain
// slot, c2, original order).  In the scan, lane l of a warp owns the 8 consecutive entries [8l, 8l+8) of a block
// and gather step j reads entry 8l+j in all lanes: in a bank-sorted block those 32 entries lie 8 apart, i.e. in
// (about) 32 different banks, or they are the same c2 (one address, broadcast) -- the random gather of the
// unsorted stream needed ~3.5 shared-memory wavefronts per load, this needs ~1.  Order inside a block no longer
// is list order, so the scan breaks gain ties inside a vector by original position (pos_s is permuted along).
__global__ void __launch_bounds__(kBlk)
mi_block_sort_kernel(uint16_t *__restrict__ c2s, uint32_t *__restrict__ pos_s, int64_t n_blocks) {
    __shared__ uint32_t key[kBlk];
    __shared__ uint16_t c2v[kBlk];
    __shared__ uint32_t posv[kBlk];
    for (int64_t blk = blockIdx.x; blk < n_blocks; blk += gridDim.x) {
        const int64_t base = blk * kBlk;
        const int t = threadIdx.x;
        const uint16_t c = c2s[base + t];
        c2v[t] = c;
        posv[t] = pos_s[base + t];
        const uint32_t idx = (uint32_t)c >> 2;                   // column index (k_v for padding / removed)
        key[t] = ((idx & 31u) << 24) | (idx << 8) | (uint32_t)t;  // idx < 2^14 (checked on the host)
        __syncthreads();
        for (int size = 2; size <= kBlk; size <<= 1) {
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
        const int src = (int)(key[t] & 255u);
        c2s[base + t] = c2v[src];
        pos_s[base + t] = posv[src];
        __syncthreads();
    }
}

// ---- the persistent kernel -----------------------------------------------------------------------

struct MiPersist {
    MiState s;                       // canonical state in global memory (read at entry, written back at exit)
    uint32_t *n_alt;                 // second copy of the table counts (double buffering, see kernel)
    const uint16_t *c2s_ro;          // same memory as c2s (loads bypass L1)
    uint16_t *c2s;
    const uint32_t *pos_s;
    const uint32_t *row_start;       // [k_a + 1]
    const uint32_t *chunk_start;     // [grid + 1] element offsets
    MiPub *pub;                      // [2][grid]
    unsigned int *bar;               // [2] {count, generation}
    int64_t n_picks;
    int64_t *out_pos;
    float *out_gain;
    int32_t rows_smem;               // gain rows that fit in shared memory
    int32_t ring_offset;             // byte offset of the cp.async ring in dynamic shared memory
    // multi-GPU (world > 1): every rank's winner is pushed into every peer's mailbox
    int32_t world, rank;
    unsigned int seq_base;
    MiMail *mail_local;              // [2][world] in this GPU's memory
    MiMail *mail_peer[kMaxWorld];    // the same array on every rank (peer-mapped pointers)
    long long *dbg;                  // optional [grid][8] counters of the LAST iteration (nullptr = off)
    int *status;                     // kMiRun* word of this launch (device)
    unsigned long long spin_limit_ns;    // how long a CTA waits for a peer GPU's entry before giving up
};

__device__ __forceinline__ uint4 ldcg_u4(const uint4 *p) { return __ldcg(p); }

// 16-byte asynchronous global->shared copy that bypasses L1 (coherent at L2: removals written by
// another SM before the grid barrier are seen), used to keep kRing loads per thread in flight while
// the previous vectors are being scored.
__device__ __forceinline__ void cp_async16(void *smem_dst, const void *gmem_src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)),
                 "l"(gmem_src)
                 : "memory");
}
__device__ __forceinline__ void cp_async16_u32(uint32_t smem_dst, const void *gmem_src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_dst), "l"(gmem_src) : "memory");
}
// same, with an L2 eviction-priority hint: the candidate stream (hundreds of MB, read once per iteration) is marked
// evict-first so that it does not push the table counts, row offsets and positions out of L2
__device__ __forceinline__ uint64_t l2_policy_evict_first() {
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
__device__ __forceinline__ void cp_async16_u32_hint(uint32_t smem_dst, const void *gmem_src, uint64_t pol) {
    asm volatile("cp.async.cg.shared.global.L2::cache_hint [%0], [%1], 16, %2;" ::"r"(smem_dst), "l"(gmem_src), "l"(pol)
                 : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// Per-thread running arg-max.  `bi` is the stream index of the first candidate (of its row segment)
// holding the best gain `bs`; equal-gain candidates of LATER row segments are parked in tie[] and only
// compared by original position if this thread ends up holding the block maximum.
struct ScanBest {
    float bs;
    uint32_t bi, bend;
    uint32_t tie[3];
    int ntie;
};

__device__ __forceinline__ void scan_tie(ScanBest &b, uint32_t e, uint32_t rend, const uint32_t *__restrict__ pos_s) {
    if (b.ntie < 3) {                          // predicated stores: no dynamically indexed local array
        if (b.ntie == 0) b.tie[0] = e;
        else if (b.ntie == 1) b.tie[1] = e;
        else b.tie[2] = e;
        ++b.ntie;
    } else {                                   // list full: settle by original position now
        uint32_t bp = __ldg(pos_s + b.bi);
#pragma unroll
        for (int t = 0; t < 3; ++t) {
            const uint32_t pt = __ldg(pos_s + b.tie[t]);
            if (pt < bp) { bp = pt; b.bi = b.tie[t]; }
        }
        if (__ldg(pos_s + e) < bp) b.bi = e;
        b.ntie = 0;
    }
    b.bend = rend;
}

__device__ __forceinline__ void scan_consider(ScanBest &b, float g, uint32_t e, uint32_t rend,
                                              const uint32_t *__restrict__ pos_s) {
    if (g > b.bs) { b.bs = g; b.bi = e; b.bend = rend; b.ntie = 0; }
    else if (g == b.bs && b.bi != 0xFFFFFFFFu && e >= b.bend) scan_tie(b, e, rend, pos_s);
}

// One cooperative launch = n_picks greedy iterations with ONE grid barrier each.
//
// State placement: the marginals a[], b[] and the running sums are replicated in every CTA's shared
// memory (every CTA sees every winner and applies it itself, same fp32 ops => identical copies).  The
// K_a x K_v table counts stay in global memory in TWO copies: iteration `it` reads copy it&1, which
// holds all picks up to it-2, and patches the one cell of pick it-1 locally; CTA 0 meanwhile brings the
// other copy up to date (picks it-2, it-1), which nobody reads until the barrier has passed.  So no
// second barrier is needed to order the table update against the next iteration's reads.
__global__ void __launch_bounds__(kPersistThreads, 1) mi_persistent_kernel(MiPersist P) {
    extern __shared__ __align__(16) unsigned char psmem[];
    const MiState &s = P.s;
    const int32_t k_v = s.k_v, k_a = s.k_a;
    float *col_term = reinterpret_cast<float *>(psmem);                             // [k_v]
    float *tn_small = col_term + k_v;                                               // [kSmallCounts]
    uint32_t *a_cnt = reinterpret_cast<uint32_t *>(tn_small + kSmallCounts);        // [k_v] column marginals
    uint32_t *b_cnt = a_cnt + k_v;                                                  // [k_a] row marginals
    uint32_t *rs_all = b_cnt + k_a;                                                 // [k_a + 1] row offsets
    float *rt_local = reinterpret_cast<float *>(rs_all + k_a + 1);                  // [rows_smem]
    float *gain = rt_local + P.rows_smem;                                           // [rows_smem][k_v + 1]
    uint4 *ring = reinterpret_cast<uint4 *>(psmem + P.ring_offset) + threadIdx.x;   // [kRing][threads] uint4
    const int32_t gstride = k_v + 1;                                                // slot k_v holds -inf
    __shared__ unsigned long long wkey[32];
    __shared__ unsigned long long wpay[32];
    __shared__ uint32_t widx[32];
    __shared__ unsigned long long sh_best_key, sh_win_key, sh_win_pay;
    __shared__ uint32_t sh_best_idx;
    __shared__ float ps[6];                                   // {NlogN, aloga, blogb, n, fN0, fa0}

    for (int32_t i = threadIdx.x; i < k_v; i += blockDim.x) a_cnt[i] = __ldcg(s.a_cols + i);
    for (int32_t i = threadIdx.x; i < k_a; i += blockDim.x) b_cnt[i] = __ldcg(s.b_rows + i);
    for (int32_t i = threadIdx.x; i <= k_a; i += blockDim.x) rs_all[i] = P.row_start[i];
    if (threadIdx.x < 6) ps[threadIdx.x] = __ldcg(s.sums + threadIdx.x);
    __syncthreads();

    const uint32_t e_lo = P.chunk_start[blockIdx.x], e_hi = P.chunk_start[blockIdx.x + 1];
    int32_t r_lo = 0, r_hi = -1;                               // rows touched by this chunk
    if (e_hi > e_lo) {
        int32_t a = 0, b = k_a;
        while (a < b) { const int32_t m = (a + b) >> 1; if (rs_all[m + 1] > e_lo) b = m; else a = m + 1; }
        r_lo = a;
        a = r_lo; b = k_a;
        while (a < b) { const int32_t m = (a + b) >> 1; if (rs_all[m] < e_hi) a = m + 1; else b = m; }
        r_hi = a - 1;
    }
    const uint32_t base_pos = (uint32_t)s.pos_base;
    const uint32_t grid = gridDim.x;
    const uint64_t stream_pol = l2_policy_evict_first();
    int32_t prev1 = -1, prev2 = -1;                            // table cell of picks it-1, it-2
    int64_t done = 0;
    bool broke = false;
    long long t_learn = 0;

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
        for (int32_t i = threadIdx.x; i < k_v; i += blockDim.x)
            col_term[i] = __fdiv_rn(-bump_sum(aloga, a_cnt[i], fa0, s.logs), np);
        for (int32_t i = threadIdx.x; i < kSmallCounts; i += blockDim.x)
            tn_small[i] = __fdiv_rn(bump_sum(NlogN, (uint32_t)i, fN0, s.logs), np);
        ScanBest B;
        B.bs = -INFINITY; B.bi = 0xFFFFFFFFu; B.bend = 0; B.ntie = 0;
        B.tie[0] = B.tie[1] = B.tie[2] = 0;
        if (P.dbg) t_pre = clock64() - t0;
        for (int32_t rb = r_lo; rb <= r_hi; rb += P.rows_smem) {
            const int32_t nr = min(P.rows_smem, r_hi - rb + 1);
            const uint32_t *rs_local = rs_all + rb;
            __syncthreads();            // previous sub-batch finished reading gain rows; col_term/tn_small ready
            for (int32_t i = threadIdx.x; i < nr; i += blockDim.x) {
                rt_local[i] = __fdiv_rn(-bump_sum(blogb, b_cnt[rb + i], fa0, s.logs), np);
                gain[i * gstride + k_v] = -INFINITY;           // slot read by removed entries (c2 == k_v)
            }
            __syncthreads();
            const long long tg0 = P.dbg ? clock64() : 0;
            {   // gain rows: thread owns columns c2 = tid, tid + T, ...; 8 rows' counts in flight per thread
                const int32_t cell0 = rb * k_v;                            // < 2^31 (checked on the host)
                const uint32_t *nbase = Tcur + cell0;
                for (int32_t c2 = threadIdx.x; c2 < k_v; c2 += kPersistThreads) {
                    const float ct = col_term[c2];
                    for (int32_t r0 = 0; r0 < nr; r0 += 8) {
                        uint32_t xs[8];
#pragma unroll
                        for (int u = 0; u < 8; ++u)
                            xs[u] = r0 + u < nr ? __ldcg(nbase + (r0 + u) * k_v + c2) : 0u;
#pragma unroll
                        for (int u = 0; u < 8; ++u) {
                            if (r0 + u < nr) {
                                const int32_t rr = r0 + u;
                                const uint32_t x = xs[u] + (cell0 + rr * k_v + c2 == prev1 ? 1u : 0u);   // pick it-1
                                const float tN = x < (uint32_t)kSmallCounts
                                                     ? tn_small[x]
                                                     : __fdiv_rn(bump_sum(NlogN, x, fN0, s.logs), np);
                                gain[rr * gstride + c2] = __fadd_rn(__fadd_rn(__fadd_rn(tN, ct), rt_local[rr]), lognp);
                            }
                        }
                    }
                }
            }
            __syncthreads();
            if (P.dbg) t_gain += clock64() - tg0;
            // rows are padded to whole 256-candidate blocks and chunk edges are block aligned: a block never
            // straddles a row or a chunk, so the row is uniform per warp-iteration and there is no slow path
            const uint32_t s_lo = max(e_lo, rs_local[0]), s_hi = min(e_hi, rs_local[nr]);
            if (s_hi <= s_lo) continue;
            const uint32_t b_lo = s_lo / kBlk, b_hi = s_hi / kBlk;
            const uint4 *vec = reinterpret_cast<const uint4 *>(P.c2s_ro);
            // each WARP walks its own contiguous span of blocks (coalesced 512-byte loads); lane l owns vector l.
            // Measured alternatives on B200 (W = 1e8): warps interleaved over one contiguous CTA window 2.3x
            // slower; ring depth 6 / 8 / 12 instead of 4 (more bytes in flight) 10-30 % slower.
            const uint32_t nwarps = kPersistThreads / kWarp;
            const uint32_t span = (b_hi - b_lo + nwarps - 1) / nwarps;
            const uint32_t wb_lo = min(b_hi, b_lo + (threadIdx.x / kWarp) * span);
            const uint32_t wb_hi = min(b_hi, wb_lo + span);
            const uint32_t lane = threadIdx.x % kWarp;
            const uint32_t gain_b = (uint32_t)__cvta_generic_to_shared(gain);
            int32_t crow = 0;
            if (wb_lo < wb_hi) {                                 // row of my first block (uniform per warp)
                const uint32_t ef = wb_lo * kBlk;
                int32_t a = 0, b = nr;
                while (a < b) { const int32_t m = (a + b) >> 1; if (rs_local[m + 1] > ef) b = m; else a = m + 1; }
                crow = a;
            }
            uint32_t rend_blk = rs_local[crow + 1] / kBlk;       // first block of the next row
            uint32_t grow_b = gain_b + (uint32_t)(crow * gstride) * 4u;
            // pointer form of the ring: shared addresses as 32-bit offsets (slot stride 16 KiB, 4 slots = 64 KiB,
            // so "next slot" is an add and a mask), the global source as a running pointer
            static_assert(kRing == 4 && kPersistThreads * 16 == 0x4000, "ring arithmetic below assumes 4 x 16 KiB");
            const uint32_t ring_u32 = (uint32_t)__cvta_generic_to_shared(ring);
            const char *src = reinterpret_cast<const char *>(vec + (size_t)wb_lo * kWarp + lane);
#pragma unroll
            for (int r = 0; r < kRing - 1; ++r) {
                if (wb_lo + (uint32_t)r < wb_hi) cp_async16_u32_hint(ring_u32 + r * 0x4000u, src + r * 512, stream_pol);
                cp_async_commit();
            }
            src += (kRing - 1) * 512;
            uint32_t slot_off = 0, pre_off = (kRing - 1) * 0x4000u;
            for (uint32_t blk = wb_lo; blk < wb_hi; ++blk) {
                if (blk + (kRing - 1) < wb_hi) cp_async16_u32_hint(ring_u32 + pre_off, src, stream_pol);
                cp_async_commit();
                src += 512;
                pre_off = (pre_off + 0x4000u) & 0xFFFFu;
                cp_async_wait<kRing - 1>();
                uint4 q;
                asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];"
                             : "=r"(q.x), "=r"(q.y), "=r"(q.z), "=r"(q.w) : "r"(ring_u32 + slot_off));
                slot_off = (slot_off + 0x4000u) & 0xFFFFu;
                while (blk >= rend_blk) {                        // next non-empty row (uniform per warp)
                    ++crow;
                    rend_blk = rs_local[crow + 1] / kBlk;
                    grow_b = gain_b + (uint32_t)(crow * gstride) * 4u;
                }
                // the stream holds byte offsets into a gain row; removed / padding entries point at its -inf slot
                float g[8];
                const uint32_t words[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const uint32_t off = (j & 1) ? (words[j >> 1] >> 16) : (words[j >> 1] & 0xFFFFu);
                    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(g[j]) : "r"(grow_b + off));
                }
                const float m = fmaxf(fmaxf(fmaxf(g[0], g[1]), fmaxf(g[2], g[3])),
                                      fmaxf(fmaxf(g[4], g[5]), fmaxf(g[6], g[7])));
                if (m >= B.bs) {                                 // rare once the thread has seen a good candidate
                    const uint32_t e0 = blk * kBlk + lane * 8u;
                    if (m > B.bs || (B.bi != 0xFFFFFFFFu && e0 >= B.bend)) {
                        // of the maxima in this vector take the one that came first in the candidate list
                        uint32_t bj = 0, bp = 0xFFFFFFFFu;
#pragma unroll
                        for (int j = 0; j < 8; ++j) {
                            if (g[j] == m) {
                                const uint32_t pj = __ldg(P.pos_s + e0 + j);
                                if (pj < bp) { bp = pj; bj = (uint32_t)j; }
                            }
                        }
                        if (m > -INFINITY) scan_consider(B, m, e0 + bj, rend_blk * kBlk, P.pos_s);
                    }
                }
            }
            cp_async_wait<0>();
        }
        const long long t1 = P.dbg ? clock64() : 0;
        float bs = B.bs;
        uint32_t bi = B.bi;
        const int ntie = B.ntie;
        uint32_t tie[3] = {B.tie[0], B.tie[1], B.tie[2]};
        // block maximum of the gain, then only the threads holding it settle their ties by position
        const uint32_t my32 = bi == 0xFFFFFFFFu ? 0u : orderable(bs);
        uint32_t m32 = my32;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) m32 = max(m32, __shfl_xor_sync(0xffffffffu, m32, o));
        if (threadIdx.x % kWarp == 0) widx[threadIdx.x / kWarp] = m32;
        __syncthreads();
        m32 = widx[threadIdx.x % kWarp];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) m32 = max(m32, __shfl_xor_sync(0xffffffffu, m32, o));
        __syncthreads();                                       // widx is reused below
        unsigned long long key = 0ull, pay = 0ull;
        if (my32 != 0u && my32 == m32) {
            // this thread holds the block-maximum gain: its candidate's original position and column are loaded
            // together, the table count of its cell right after (two memory latencies; thread 0 only publishes)
            uint32_t bp = __ldg(P.pos_s + bi);
#pragma unroll
            for (int t = 0; t < 3; ++t) {
                if (t < ntie) {
                    const uint32_t pt = __ldg(P.pos_s + tie[t]);
                    if (pt < bp) { bp = pt; bi = tie[t]; }
                }
            }
            const uint32_t c2 = (uint32_t)__ldcg(reinterpret_cast<const unsigned short *>(P.c2s) + bi) >> 2;
            int32_t a = 0, b = k_a;                            // table row of the stream index
            while (a < b) { const int32_t m = (a + b) >> 1; if (rs_all[m + 1] > bi) b = m; else a = m + 1; }
            const int32_t cell = a * k_v + (int32_t)c2;
            const uint32_t x = __ldcg(Tcur + cell) + (cell == prev1 ? 1u : 0u);
            key = ((unsigned long long)m32 << 32) | (unsigned long long)(0xFFFFFFFFu - (base_pos + bp));
            pay = ((unsigned long long)a << 48) | ((unsigned long long)c2 << 32) | x;
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
            if (threadIdx.x % kWarp == 0) {
                wkey[threadIdx.x / kWarp] = k2; wpay[threadIdx.x / kWarp] = p2; widx[threadIdx.x / kWarp] = i2;
            }
            __syncthreads();
            if (threadIdx.x < kWarp) {
                k2 = wkey[threadIdx.x]; p2 = wpay[threadIdx.x]; i2 = widx[threadIdx.x];
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
            d[4] = (long long)(e_hi - e_lo) / kBlk; d[5] = r_hi - r_lo + 1; d[6] = t_pre; d[7] = t_learn;
        }
        // ---------------- everyone learns the winner ----------------
        {
            unsigned long long k2 = 0ull, p2 = 0ull;
            for (uint32_t t = threadIdx.x; t < grid; t += blockDim.x) {
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
        if (win == 0ull) { broke = true; break; }              // nothing left on any rank
        const int32_t c1 = (int32_t)(wpayload >> 48), c2w = (int32_t)((wpayload >> 32) & 0xFFFFu);
        if (threadIdx.x == 0) {
            if (sh_best_key == win) {                          // keys are unique: exactly one owner CTA
                P.c2s[sh_best_idx] = (uint16_t)(k_v << 2);     // remove_idx_all mi.py:104-106 (offset of the -inf slot)
                s.cells[(int64_t)key_pos(win) - s.pos_base] = 0xFFFFFFFFu;     // list-order view stays in sync
            }
            const uint32_t x = (uint32_t)(wpayload & 0xFFFFFFFFull), y = a_cnt[c2w], z = b_cnt[c1];
            ps[0] = bump_sum(ps[0], x, ps[4], s.logs);         // update_cache mi.py:383-389
            ps[1] = bump_sum(ps[1], y, ps[5], s.logs);
            ps[2] = bump_sum(ps[2], z, ps[5], s.logs);
            ps[3] = __fadd_rn(ps[3], 1.0f);                    // update_mats :401-406
            a_cnt[c2w] = y + 1; b_cnt[c1] = z + 1;
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
            // canonical copy 0 holds picks <= done-2 if `done` is even (it was the read copy of
            // iteration `done`), <= done-3 if odd
            // (an early exit at an odd `done` happened after CTA 0 had already completed copy 0)
            if (!(broke && (done & 1))) {
                if (prev1 >= 0) s.n_cells[prev1] += 1;
                if ((done & 1) && prev2 >= 0) s.n_cells[prev2] += 1;
            }
            for (int i = 0; i < 4; ++i) s.sums[i] = ps[i];
        }
        for (int32_t i = threadIdx.x; i < k_v; i += blockDim.x) s.a_cols[i] = a_cnt[i];
        for (int32_t i = threadIdx.x; i < k_a; i += blockDim.x) s.b_rows[i] = b_cnt[i];
    }
}

// ---- host side -----------------------------------------------------------------------------------

int mi_partition_scratch_tiles(int64_t w) { return (int)ceil_div(w > 0 ? w : 1, kTileElems); }

int launch_mi_partition(const uint32_t *cells, int64_t w, int32_t k_a, uint32_t *tilehist, uint32_t *row_total,
                        uint32_t *row_start, uint16_t *c2s, uint32_t *pos_s, int64_t stream_capacity,
                        uint16_t pad_marker, cudaStream_t st) {
    const int ntiles = mi_partition_scratch_tiles(w);
    const size_t smem = (size_t)k_a * sizeof(uint32_t);
    if (smem > 96 * 1024) return ACAV_E_UNSUPPORTED;
    static size_t done_count[kMaxDevices], done_scatter[kMaxDevices];
    if (smem > 48 * 1024) {
        int rc = ensure_dynamic_smem(mi_part_count_kernel, (size_t)96 * 1024, done_count);
        if (!rc) rc = ensure_dynamic_smem(mi_part_scatter_kernel, (size_t)96 * 1024, done_scatter);
        if (rc) return rc;
    }
    mi_part_count_kernel<<<ntiles, kPartThreads, smem, st>>>(cells, w, k_a, tilehist);
    ACAV_LAUNCH_CHECK();
    mi_part_prefix_kernel<<<(unsigned)ceil_div(k_a, 256), 256, 0, st>>>(tilehist, ntiles, k_a, row_total);
    ACAV_LAUNCH_CHECK();
    mi_part_rowstart_kernel<<<1, 1024, 0, st>>>(row_total, k_a, row_start);
    ACAV_LAUNCH_CHECK();
    // padding entries of every row: "removed" marker (the -inf slot of a gain row)
    mi_fill_u16_kernel<<<(unsigned)ceil_div(stream_capacity, 256), 256, 0, st>>>(c2s, 0, stream_capacity, pad_marker);
    ACAV_LAUNCH_CHECK();
    mi_part_scatter_kernel<<<ntiles, kPartThreads, smem, st>>>(cells, w, k_a, tilehist, row_start, c2s, pos_s);
    ACAV_LAUNCH_CHECK();
    return 0;
}

// second step, once the padded stream length is known on the host (row_start[k_a]): bank-sort every block
int launch_mi_block_sort(uint16_t *c2s, uint32_t *pos_s, int64_t w_padded, cudaStream_t st) {
    const int64_t n_blocks = w_padded / kBlk;
    if (n_blocks == 0) return 0;
    const unsigned grid = (unsigned)(n_blocks < 65535 * 16 ? n_blocks : 65535 * 16);
    mi_block_sort_kernel<<<grid, kBlk, 0, st>>>(c2s, pos_s, n_blocks);
    ACAV_LAUNCH_CHECK();
    return 0;
}

int64_t mi_stream_capacity(int64_t w, int32_t k_a) { return w + (int64_t)kBlk * k_a + kBlk; }
int mi_stream_block() { return kBlk; }


// dynamic shared memory: [col_term k_v | tn_small | a_cnt k_v | b_cnt k_a | rs_all k_a+1 | rt_local rows |
// gain rows*(k_v+1)] | cp.async ring
static size_t persist_table_bytes(int32_t k_a, int32_t k_v, int32_t rows) {
    size_t words = 2 * (size_t)k_v + kSmallCounts + 2 * (size_t)k_a + 1 + (size_t)rows + (size_t)rows * (k_v + 1);
    return (words * 4 + 15) & ~(size_t)15;
}
static size_t persist_ring_bytes() { return (size_t)kRing * kPersistThreads * 16; }

int mi_persistent_rows_that_fit(int32_t k_a, int32_t k_v) {
    const size_t budget = 224 * 1024 - persist_ring_bytes();
    int32_t rows = 0;
    while (persist_table_bytes(k_a, k_v, rows + 1) <= budget && rows < 4096) ++rows;
    return rows;
}

size_t mi_pub_bytes(int32_t grid) { return sizeof(MiPub) * 2 * (size_t)grid; }
size_t mi_mail_bytes(int32_t world) { return sizeof(MiMail) * 2 * (size_t)world; }

int launch_mi_persistent(const MiState &s, uint32_t *n_alt, uint16_t *c2s, const uint32_t *pos_s,
                         const uint32_t *row_start, const uint32_t *chunk_start, int32_t grid, void *pub,
                         unsigned int *bar, int64_t n_picks, int64_t *out_pos, float *out_gain, int32_t rows_smem,
                         int32_t world, int32_t rank, unsigned int seq_base, void *mail_local, void *const *mail_peer,
                         long long *dbg, int *status, unsigned long long spin_limit_ns, cudaStream_t st, bool sync_clean) {
    MiPersist P;
    P.dbg = dbg; P.status = status; P.spin_limit_ns = spin_limit_ns;
    P.s = s; P.n_alt = n_alt; P.c2s_ro = c2s; P.c2s = c2s; P.pos_s = pos_s; P.row_start = row_start;
    P.chunk_start = chunk_start; P.pub = reinterpret_cast<MiPub *>(pub); P.bar = bar; P.n_picks = n_picks;
    P.out_pos = out_pos; P.out_gain = out_gain; P.rows_smem = rows_smem;
    P.world = world; P.rank = rank; P.seq_base = seq_base;
    P.mail_local = reinterpret_cast<MiMail *>(mail_local);
    for (int r = 0; r < kMaxWorld; ++r)
        P.mail_peer[r] = (world > 1 && r < world) ? reinterpret_cast<MiMail *>(mail_peer[r]) : nullptr;
    P.ring_offset = (int32_t)persist_table_bytes(s.k_a, s.k_v, rows_smem);
    const size_t smem = persist_table_bytes(s.k_a, s.k_v, rows_smem) + persist_ring_bytes();
    static size_t attr_done[kMaxDevices];
    { int rc = ensure_dynamic_smem(mi_persistent_kernel, smem, attr_done); if (rc) return rc; }
    // both table copies start equal; the barrier words start at zero
    ACAV_CUDA_TRY(cudaMemcpyAsync(n_alt, s.n_cells, sizeof(uint32_t) * (size_t)s.k_a * s.k_v, cudaMemcpyDeviceToDevice, st));
    // the barrier words and the records start at zero: mi_refresh_kernel leaves them so after every run (sync_clean);
    // the status word can only be set by a peer timeout
    if (!sync_clean) {
        ACAV_CUDA_TRY(cudaMemsetAsync(pub, 0, mi_pub_bytes(grid), st));
        ACAV_CUDA_TRY(cudaMemsetAsync(bar, 0, 2 * sizeof(unsigned int), st));
    }
    if (world > 1 || !sync_clean) ACAV_CUDA_TRY(cudaMemsetAsync(status, 0, sizeof(int), st));
    void *args[] = {&P};
    ACAV_CUDA_TRY(cudaLaunchCooperativeKernel((void *)mi_persistent_kernel, dim3(grid), dim3(kPersistThreads), args,
                                              smem, st));
    return 0;
}

}  // namespace acav
