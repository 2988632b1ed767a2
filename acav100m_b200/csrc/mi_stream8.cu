// Persistent greedy-MI kernel, ONE BYTE per candidate (ACAV_MI_LOOP_BYTES): the selection loop of
// EfficientMI.run_greedy (subset_selection/code/measures/mi.py:150-192 with EfficientMemMI :284-412) in one
// cooperative launch, like mi_persistent.cu, with half the HBM bytes per iteration and no shared-memory ring.
//
// Layout (built once per engine):
//   * the K_a x K_v table is cut into SUB-ROWS of sub_w <= 255 columns (n_sub = ceil(K_v / 255) per table row,
//     sub_w = ceil(K_v / n_sub): 5 x 205 at K_v = 1024).  Candidates are STABLY partitioned by sub-row; the stream holds
//     the column inside the sub-row as ONE byte (255 = removed / padding; that slot of a gain row holds -inf, so the
//     hot loop has no branch).  100 MB per iteration at W = 1e8 instead of the reference's 1.6 GB of int64 pairs.
//   * the sub-rows lie in the stream in an order chosen on the host (`perm`: stream slot -> sub-row): big and small
//     sub-rows interleaved so that every stretch of the stream holds about the same number of sub-rows per block.  Every
//     CTA's contiguous chunk then has the same number of blocks AND about the same number of gain rows to build, and
//     all of its gain rows and table counts fit in its shared memory (the cut enforces it).
//   * sub-rows are padded to whole 512-candidate blocks, every block is sorted by the shared-memory bank of its gain
//     slot (lane l owns entries [16l, 16l+16): gather step j reads entries 16 apart, i.e. ~32 different banks); ties are
//     settled by original position (pos[] and the per-vector rank words, permuted along, read only after the scan).
//   * a CTA stages the gain rows (256 floats) of its sub-rows in shared memory, 1 KiB aligned so that a gather address
//     is (byte << 2) | row_base: shift, LOP3, LDS; their table COUNTS stay in shared memory for the whole launch (every
//     CTA learns every winner and bumps its copy), so building the gain rows of an iteration reads no global memory.
//   * stream loads go through REGISTERS: DEPTH 16-byte loads per thread in flight (ld.global.cg, L2 evict-first) -- the
//     shared-memory pipe only serves the gathers.
//   * the scan loop keeps the arg-max in four registers (best gain, the block it was first seen in, the end of that
//     block's sub-row segment, the number of parked ties); blocks of LATER segments that hold the same gain are parked
//     in a small per-thread list.  Nothing is loaded on behalf of the arg-max while streaming; after the scan the
//     recorded blocks are read again and the entry with the smallest original position wins (mi.py:79: first maximum).
// Scores use the same fp32 operation sequence and the same torch-CPU log table as mi_scan.cu: picks and gains are
// bit-identical to the reference (tests/test_mi_gpu.py).
#include <cstdlib>

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

// stable scatter into the STAGING layout (sub-rows in (c1, sub) order, each padded to whole blocks): tiles in list
// order, 512-chunks in list order, warps in order, lanes in order.  The block sort then moves every block to its place
// in the stream the host laid out (s8_build_layout).
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

// Arrange every 512-candidate block for the scan's gathers.  In the scan lane l owns the 16 bytes [16l, 16l+16) of a
// block and gather step j reads byte 16l+j in all lanes; a step is one shared-memory wavefront iff, bank by bank, the
// lanes that hit the bank read the SAME gain slot (one address: broadcast).  So the block is sorted by (bank, slot) and
// its RUNS (entries of one slot) are dealt to the 16 steps: a run goes, whole, to the step with the most room among
// those where its bank is still free (or already serves the same slot); what does not fit goes on to the next such
// step.  A 205-column sub-row has 6-7 slots per bank, so a conflict-free schedule nearly always exists -- where the
// plain "sort by bank, 16 consecutive entries per lane" layout puts two slots of a bank in one step whenever a bank
// holds more than 16 entries of the block (every other bank of a flat sub-row: 1.5-2 wavefronts per step).
// Order inside a block is not list order any more; vrank holds, per vector, every entry's rank in list order (16 x 4
// bits) so that settling a vector after the scan takes one position load.
constexpr int kS8Steps = 16;
constexpr int kS8ArrangeWarps = 8;

// step 1: sort the block by (bank, slot, original order) -- one CTA per block; block b of the stream is read from
// block blk_src[b] of the staging layout
__global__ void __launch_bounds__(kS8Blk)
s8_block_sort_kernel(const uint8_t *__restrict__ stage_stream, const uint32_t *__restrict__ stage_pos,
                     const uint32_t *__restrict__ blk_src, uint8_t *__restrict__ stream, uint32_t *__restrict__ pos_s,
                     int64_t n_blocks) {
    __shared__ uint32_t key[kS8Blk];
    __shared__ uint8_t colv[kS8Blk];
    __shared__ uint32_t posv[kS8Blk];
    for (int64_t blk = blockIdx.x; blk < n_blocks; blk += gridDim.x) {
        const int64_t base = blk * kS8Blk, sbase = (int64_t)blk_src[blk] * kS8Blk;
        const int t = threadIdx.x;
        const uint32_t c = stage_stream[sbase + t];
        colv[t] = (uint8_t)c;
        posv[t] = stage_pos[sbase + t];
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

// step 2: deal the runs of the sorted block to the 16 gather steps -- one WARP per block (the dealing is a sequential
// greedy over ~200 runs; 40 warps per SM keep it off the critical path of the build)
__global__ void __launch_bounds__(kS8ArrangeWarps *kWarp)
s8_block_arrange_kernel(uint8_t *__restrict__ stream, uint32_t *__restrict__ pos_s, unsigned long long *__restrict__ vrank,
                        int64_t n_blocks) {
    __shared__ uint32_t posv_s[kS8ArrangeWarps][kS8Blk];
    __shared__ uint8_t colv_s[kS8ArrangeWarps][kS8Blk];
    __shared__ uint16_t dest_s[kS8ArrangeWarps][kS8Blk];        // sorted index -> byte of the block (lane * 16 + step)
    // run starts [514] + per (step, bank) the slot served (0xFE = none yet) [512 bytes]; after the dealing the same
    // memory holds the inverse map, byte of the block -> sorted index [512]
    __shared__ uint16_t scratch_s[kS8ArrangeWarps][kS8Blk + 2 + kS8Steps * 16];
    const int lane = threadIdx.x % kWarp, warp = threadIdx.x / kWarp;
    uint32_t *posv = posv_s[warp];
    uint8_t *colv = colv_s[warp];
    uint16_t *dest = dest_s[warp], *inv = scratch_s[warp], *run_start = scratch_s[warp];
    uint8_t *bank_slot = reinterpret_cast<uint8_t *>(scratch_s[warp] + kS8Blk + 2);
    const int64_t warps_total = (int64_t)gridDim.x * kS8ArrangeWarps;
    for (int64_t blk = (int64_t)blockIdx.x * kS8ArrangeWarps + warp; blk < n_blocks; blk += warps_total) {
        const int64_t base = blk * kS8Blk;
        // load the sorted block (16 bytes of slots and 16 positions per lane), find the runs of equal slot
        {
            const uint4 q = *reinterpret_cast<const uint4 *>(stream + base + lane * 16);
            *reinterpret_cast<uint4 *>(colv + lane * 16) = q;
#pragma unroll
            for (int j = 0; j < 4; ++j)
                *reinterpret_cast<uint4 *>(posv + lane * 16 + j * 4) =
                    *reinterpret_cast<const uint4 *>(pos_s + base + lane * 16 + j * 4);
#pragma unroll
            for (int j = 0; j < 16; ++j) bank_slot[lane * 16 + j] = 0xFE;
        }
        __syncwarp();
        int n_runs = 0;
        for (int i0 = 0; i0 < kS8Blk; i0 += kWarp) {
            const int i = i0 + lane;
            const bool is_start = i == 0 || colv[i] != colv[i - 1];
            const unsigned bal = __ballot_sync(0xffffffffu, is_start);
            if (is_start) run_start[n_runs + __popc(bal & ((1u << lane) - 1u))] = (uint16_t)i;
            n_runs += __popc(bal);
        }
        if (lane == 0) run_start[n_runs] = (uint16_t)kS8Blk;
        __syncwarp();
        // lanes 0..15 each keep one step (its fill count); all lanes place the entries of a run piece
        int fill = 0;
        for (int r = 0; r < n_runs; ++r) {
            int s0 = run_start[r];
            int left = (int)run_start[r + 1] - s0;
            const uint32_t rv = colv[s0], rb = rv & 31u;
            while (left > 0) {
                // my step's offer: its room, preferred if the bank is free there (or serves this slot already)
                uint32_t offer = 0;
                if (lane < kS8Steps && fill < 32) {
                    const uint8_t cur = bank_slot[lane * 32 + rb];
                    const bool clean = cur == 0xFE || cur == (uint8_t)rv;
                    offer = ((clean ? 64u : 0u) + (uint32_t)(32 - fill)) << 8 | (uint32_t)(31 - lane);
                }
                const uint32_t top = __reduce_max_sync(0xffffffffu, offer);
                const int who = 31 - (int)(top & 0xFFu);
                const int wroom = (int)((top >> 8) & 63u);     // room is 1..32; the clean flag is bit 6
                const int wfill = 32 - wroom;
                const int take = left < wroom ? left : wroom;
                if (lane < take) dest[s0 + lane] = (uint16_t)((wfill + lane) * 16 + who);
                if (lane == who) { fill += take; bank_slot[who * 32 + rb] = (uint8_t)rv; }
                __syncwarp();
                s0 += take;
                left -= take;
            }
        }
        __syncwarp();
        for (int i = lane; i < kS8Blk; i += kWarp) inv[dest[i]] = (uint16_t)i;
        __syncwarp();
        {   // my vector = bytes [16 lane, 16 lane + 16) of the arranged block; ranks of its entries in list order
            uint32_t sl[16], ps[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) { const int i = inv[lane * 16 + j]; sl[j] = colv[i]; ps[j] = posv[i]; }
            unsigned long long word = 0ull;
            uint32_t w4[4] = {0u, 0u, 0u, 0u};
#pragma unroll
            for (int j = 0; j < 16; ++j) {
                int rank = 0;
#pragma unroll
                for (int k = 0; k < 16; ++k) rank += (ps[k] < ps[j] || (ps[k] == ps[j] && k < j)) ? 1 : 0;   // padding: arbitrary
                word |= (unsigned long long)j << (4 * rank);           // entry index by list-order rank
                w4[j >> 2] |= sl[j] << (8 * (j & 3));
            }
            *reinterpret_cast<uint4 *>(stream + base + lane * 16) = make_uint4(w4[0], w4[1], w4[2], w4[3]);
#pragma unroll
            for (int j = 0; j < 4; ++j)
                *reinterpret_cast<uint4 *>(pos_s + base + lane * 16 + j * 4) =
                    make_uint4(ps[4 * j], ps[4 * j + 1], ps[4 * j + 2], ps[4 * j + 3]);
            vrank[blk * kWarp + lane] = word;
        }
        __syncwarp();
    }
}

// ---- the persistent kernel ----------------------------------------------------------------------------------------

constexpr int kS8TieCap = 24;               // parked ties per thread (local memory, touched on the rare path only)

struct S8Chunk {                     // one per CTA, written by the host's layout
    uint32_t slot0, n_slots;         // its stream slots
    uint32_t n_rows;                 // distinct sub-rows among them (gain rows / count rows it stages)
    uint32_t start;                  // stream offset of its first candidate (its end = next chunk's start)
};

struct MiS8 {
    MiState s;
    uint32_t *n_alt;                 // unused here (the 2-byte loop's second table image)
    uint8_t *stream;
    const uint32_t *pos_s;
    const unsigned long long *vrank; // [stream bytes / 16] list-order ranks inside every vector
    const uint32_t *slot_start;      // [n_slots + 1] stream offsets of the slots (a slot = a sub-row or a piece of one)
    const uint32_t *slot_row;        // [n_slots] sub-row id (c1 * n_sub + sub)
    const uint32_t *slot_u;          // [n_slots] index of that sub-row among the distinct sub-rows of the slot's chunk
    const S8Chunk *chunks;           // [grid + 1]
    MiPub *pub;
    unsigned int *bar;
    int64_t n_picks;
    int64_t *out_pos;
    float *out_gain;
    S8Geom g;
    int32_t rows_smem;               // distinct sub-rows whose gain row + counts fit in shared memory
    int32_t slots_smem;              // slots per chunk the shared-memory tables hold
    int32_t fixed_bytes;             // bytes of the arrays in front of the gain rows
    int32_t prefetch;                // ask what settling will read into L2 when a block is recorded
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
// seen), evict-first in L2 so the 100 MB stream does not push the positions and the log table out
__device__ __forceinline__ uint4 s8_ld_stream(const uint4 *p, uint64_t pol) {
    uint4 v;
    asm volatile("ld.global.cg.L2::cache_hint.v4.u32 {%0, %1, %2, %3}, [%4], %5;"
                 : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p), "l"(pol));
    return v;
}
__device__ __forceinline__ float s8_gather(uint32_t word, int byte, uint32_t row_base) {
    // address = row_base + 4 * byte `byte` of `word`: one PRMT (integer pipe) and one IMAD (FMA pipe), then the LDS.
    // (shift + LOP3 -- both on the integer pipe, which issues a warp instruction every other cycle -- made that pipe the
    // co-limiter of the scan next to the issue slots: ncu math_pipe_throttle 1.3 per issued instruction.)
    uint32_t v, addr;
    asm("prmt.b32 %0, %1, 0, %2;" : "=r"(v) : "r"(word), "n"(0x4440));
    if (byte == 1) asm("prmt.b32 %0, %1, 0, %2;" : "=r"(v) : "r"(word), "n"(0x4441));
    if (byte == 2) asm("prmt.b32 %0, %1, 0, %2;" : "=r"(v) : "r"(word), "n"(0x4442));
    if (byte == 3) asm("prmt.b32 %0, %1, 0, %2;" : "=r"(v) : "r"(word), "n"(0x4443));
    asm("mad.lo.u32 %0, %1, 4, %2;" : "=r"(addr) : "r"(v), "r"(row_base));
    float g;
    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(g) : "r"(addr));
    return g;
}

struct S8Ctx {                                 // what it takes to look a block up again after the scan
    const uint4 *vec;
    const unsigned long long *vrank;           // per vector: list-order rank of each of its 16 entries (4 bits each)
    const uint32_t *pos_s;
    const uint32_t *rs_loc;                    // shared: stream offsets of my slots [ns + 1]
    const uint32_t *su_loc;                    // shared: gain row of every slot
    uint32_t gain_b;
    int32_t ns;
    uint32_t lane;
};

// A recorded block: block index in the low 26 bits, its gain row (sub-row of the CTA, < 64) above.
constexpr uint32_t kS8BlkMask = (1u << 26) - 1u;

// Of the entries of my vectors of the recorded blocks rec[0..n) (n <= 4) that hold gain `bs`, the one with the smallest
// original position.  vrank lists the entries of a vector in list order (16 x 4-bit indices), so per vector only the first
// listed entry holding `bs` matters: the n vector (and rank word) loads go out together, then n position loads -- two memory
// latencies for up to four blocks.
__device__ __forceinline__ void s8_resolve_blocks(const S8Ctx &c, const uint32_t *rec, int n, float bs, uint32_t &bp,
                                                  uint32_t &bi, int32_t &brow, uint32_t &bbyte) {
    uint4 q[4];
    unsigned long long rk[4];
#pragma unroll
    for (int t = 0; t < 4; ++t) {
        q[t] = make_uint4(0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu);
        rk[t] = 0ull;
        if (t < n) {
            const size_t vi = (size_t)(rec[t] & kS8BlkMask) * kWarp + c.lane;
            q[t] = __ldcg(c.vec + vi);
            rk[t] = __ldg(c.vrank + vi);
        }
    }
    int first[4];
    uint32_t p[4];
#pragma unroll
    for (int t = 0; t < 4; ++t) {
        first[t] = -1; p[t] = 0xFFFFFFFFu;
        if (t < n) {
            const uint32_t words[4] = {q[t].x, q[t].y, q[t].z, q[t].w};
            const uint32_t grow_b = c.gain_b + (rec[t] >> 26) * (kS8GainStride * 4u);
            uint32_t eq = 0;
#pragma unroll
            for (int j = 0; j < 16; ++j)
                eq |= (s8_gather(words[j >> 2], j & 3, grow_b) == bs ? 1u : 0u) << j;      // removed / padding: -inf
            if (eq) {
                first[t] = __ffs(eq) - 1;
                if (eq & (eq - 1u)) {                          // several entries hold it: the first one in list order
                    unsigned long long w = rk[t];                 // entry indices by list-order rank, 4 bits each
                    for (int r = 0; r < 16; ++r, w >>= 4) {
                        const int j = (int)(w & 15ull);
                        if ((eq >> j) & 1u) { first[t] = j; break; }
                    }
                }
                p[t] = __ldg(c.pos_s + (size_t)(rec[t] & kS8BlkMask) * kS8Blk + c.lane * 16u + (uint32_t)first[t]);
            }
        }
    }
#pragma unroll
    for (int t = 0; t < 4; ++t) {
        if (p[t] < bp) {
            const uint32_t words[4] = {q[t].x, q[t].y, q[t].z, q[t].w};
            const int j = first[t];
            bp = p[t]; bi = (rec[t] & kS8BlkMask) * kS8Blk + c.lane * 16u + (uint32_t)j; brow = (int32_t)(rec[t] >> 26);
            bbyte = (words[j >> 2] >> (8 * (j & 3))) & 0xFFu;
        }
    }
}

// What settling a recorded block will read -- my vector, its order word, my 16 positions -- is asked into L2 when the
// block is recorded (fire and forget): the stream itself is loaded evict-first, so by the end of the span the vector
// would come from HBM again, and the position after it.
__device__ __forceinline__ void s8_prefetch_settle(const uint4 *vec, const unsigned long long *vrank, const uint32_t *pos_s,
                                                   uint32_t blk, uint32_t lane) {
    const size_t vi = (size_t)blk * kWarp + lane;
    asm volatile("prefetch.global.L2 [%0];" ::"l"(vec + vi));
    asm volatile("prefetch.global.L2 [%0];" ::"l"(vrank + vi));
    asm volatile("prefetch.global.L2 [%0];" ::"l"(pos_s + vi * 16));
}

struct S8Found {
    uint32_t bp, bi;                           // original position and stream index of the candidate
    int32_t brow;                              // its gain row (index among the CTA's distinct sub-rows)
    uint32_t bbyte;                            // its column inside the sub-row
};

// Everything a thread has recorded -- its first block `bx` and the parked ties -- four blocks at a time.  Out of line on
// purpose: it runs once per iteration, does not belong to the scan loop's register budget, and must
// not weigh on the loop's register allocation (for the same reason the loop itself never calls anything: a thread whose
// tie list overflows just stops parking -- ntie = kS8TieCap + 1 -- and every block of its span after the last parked
// one is looked at here instead).
__device__ __noinline__ S8Found s8_resolve_all(const S8Ctx *c, float bs, uint32_t bx, const uint32_t *tie, int ntie,
                                               uint32_t span_hi) {
    S8Found f;
    f.bp = 0xFFFFFFFFu; f.bi = (bx & kS8BlkMask) * kS8Blk; f.brow = 0; f.bbyte = 0;
    const int nt = ntie > kS8TieCap ? kS8TieCap : ntie;
    uint32_t extra = ntie > kS8TieCap ? (tie[kS8TieCap - 1] & kS8BlkMask) + 1u : span_hi;   // overflow: blocks [extra, span_hi) too
    uint32_t rec[4];
    rec[0] = bx; rec[1] = rec[2] = rec[3] = bx;
    int n = 1, t = 0;
    for (;;) {
        while (n < 4 && (t < nt || extra < span_hi)) {
            uint32_t v;
            if (t < nt) v = tie[t++];
            else {                                             // a block nobody recorded: look its gain row up
                const uint32_t ef = extra * kS8Blk;
                int32_t a = 0, b = c->ns;
                while (a < b) { const int32_t m = (a + b) >> 1; if (c->rs_loc[m + 1] > ef) b = m; else a = m + 1; }
                v = extra++ | (c->su_loc[a] << 26);
            }
            if (n == 0) rec[0] = v; else if (n == 1) rec[1] = v; else if (n == 2) rec[2] = v; else rec[3] = v;
            ++n;
        }
        s8_resolve_blocks(*c, rec, n, bs, f.bp, f.bi, f.brow, f.bbyte);
        if (t >= nt && extra >= span_hi) break;
        n = 0;
    }
    return f;
}

// THREADS per CTA, DEPTH 16-byte stream loads in flight per thread, staged in registers.  (A shared-memory cp.async ring
// as in mi_persistent.cu was measured too: 8 more shared-memory wavefronts and 5 more instructions per block, 10 % slower.)
template <int THREADS, int DEPTH>
__global__ void __launch_bounds__(THREADS, 1) mi_stream8_kernel(MiS8 P) {
    extern __shared__ __align__(16) unsigned char s8_smem[];
    const MiState &s = P.s;
    const int32_t k_v = s.k_v, k_a = s.k_a;
    const int32_t n_sub = P.g.n_sub, sub_w = P.g.sub_w;
    float *col_term = reinterpret_cast<float *>(s8_smem);                          // [k_v]
    float *tn_small = col_term + k_v;                                              // [kSmallCounts]
    uint32_t *a_cnt = reinterpret_cast<uint32_t *>(tn_small + kSmallCounts);       // [k_v]
    uint32_t *b_cnt = a_cnt + k_v;                                                 // [k_a]
    uint32_t *rs_loc = b_cnt + k_a;                                                // [slots_smem + 1] stream offsets
    uint32_t *su_loc = rs_loc + P.slots_smem + 1;                                  // [slots_smem] gain row of the slot
    float *rt_local = reinterpret_cast<float *>(su_loc + P.slots_smem);            // [rows_smem]
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
    __shared__ float ps[6];                                   // {NlogN, aloga, blogb, n, fN0, fa0}
    constexpr int kWarps = THREADS / kWarp;

    for (int32_t i = threadIdx.x; i < k_v; i += THREADS) a_cnt[i] = __ldcg(s.a_cols + i);
    for (int32_t i = threadIdx.x; i < k_a; i += THREADS) b_cnt[i] = __ldcg(s.b_rows + i);
    if (threadIdx.x < 6) ps[threadIdx.x] = __ldcg(s.sums + threadIdx.x);
    const S8Chunk ch = P.chunks[blockIdx.x];
    const uint32_t e_lo = ch.start, e_hi = P.chunks[blockIdx.x + 1].start;
    const int32_t ns = min((int32_t)ch.n_slots, P.slots_smem);       // the host's layout keeps both within the tables
    const int32_t nr = min((int32_t)ch.n_rows, P.rows_smem);
    for (int32_t i = threadIdx.x; i <= ns; i += THREADS) {
        rs_loc[i] = __ldg(P.slot_start + ch.slot0 + i);
        if (i < ns) {
            const int32_t grow = (int32_t)__ldg(P.slot_row + ch.slot0 + i), c1 = grow / n_sub;
            const uint32_t u = min(__ldg(P.slot_u + ch.slot0 + i), (uint32_t)(nr > 0 ? nr - 1 : 0));
            su_loc[i] = u;
            row_c1[u] = c1;                                      // pieces of one sub-row write the same values
            row_c2[u] = (grow - c1 * n_sub) * sub_w;
        }
    }
    __syncthreads();
    // table counts of my sub-rows stay here for the whole launch
    for (int32_t idx = threadIdx.x; idx < nr * kS8GainStride; idx += THREADS) {
        const int32_t i = idx >> 8, j = idx & 255;
        const int32_t c2 = row_c2[i] + j;
        cnt[idx] = (j < sub_w && c2 < k_v) ? __ldcg(s.n_cells + (int64_t)row_c1[i] * k_v + c2) : 0u;
    }
    const uint32_t base_pos = (uint32_t)s.pos_base;
    const uint32_t grid = gridDim.x;
    const uint64_t stream_pol = s8_policy_evict_first();
    const uint32_t lane = threadIdx.x % kWarp;
    const uint4 *vec = reinterpret_cast<const uint4 *>(P.stream);
    // my warp's contiguous span of blocks (coalesced 512-byte loads; lane l owns vector l of every block)
    const uint32_t b_lo = e_lo / kS8Blk, b_hi = ns > 0 ? min(e_hi, rs_loc[ns]) / kS8Blk : b_lo;
    const uint32_t span = (b_hi - b_lo + kWarps - 1) / kWarps;
    const uint32_t wb_lo = min(b_hi, b_lo + (threadIdx.x / kWarp) * span);
    const uint32_t wb_hi = min(b_hi, wb_lo + span);
    int32_t crow0 = 0;
    if (wb_lo < wb_hi) {                                       // slot of my first block (uniform per warp)
        const uint32_t ef = wb_lo * kS8Blk;
        int32_t a = 0, b = ns;
        while (a < b) { const int32_t m = (a + b) >> 1; if (rs_loc[m + 1] > ef) b = m; else a = m + 1; }
        crow0 = a;
    }
    S8Ctx ctx;
    ctx.vec = vec; ctx.vrank = P.vrank; ctx.pos_s = P.pos_s; ctx.rs_loc = rs_loc; ctx.su_loc = su_loc; ctx.gain_b = gain_b; ctx.ns = ns;
    ctx.lane = lane;
    int64_t done = 0;
    bool broke = false;
    long long t_learn = 0;
    uint32_t tie[kS8TieCap];
    __syncthreads();

    for (int64_t it = 0; it < P.n_picks; ++it) {
        const int cur = (int)(it & 1);
        const long long t0 = P.dbg ? clock64() : 0;
        long long t_gain = 0, t_pre = 0;
        // ---------------- score my chunk ----------------
        const float NlogN = ps[0], aloga = ps[1], blogb = ps[2], fN0 = ps[4], fa0 = ps[5];
        const float np = __fadd_rn(ps[3], 1.0f);
        const float lognp = __ldg(s.logs + (int64_t)np);
        for (int32_t i = threadIdx.x; i < k_v; i += THREADS)
            col_term[i] = __fdiv_rn(-bump_sum(aloga, a_cnt[i], fa0, s.logs), np);
        for (int32_t i = threadIdx.x; i < kSmallCounts; i += THREADS)
            tn_small[i] = __fdiv_rn(bump_sum(NlogN, (uint32_t)i, fN0, s.logs), np);
        for (int32_t i = threadIdx.x; i < nr; i += THREADS)
            rt_local[i] = __fdiv_rn(-bump_sum(blogb, b_cnt[row_c1[i]], fa0, s.logs), np);
        __syncthreads();
        if (P.dbg) t_pre = clock64() - t0;
        const long long tg0 = P.dbg ? clock64() : 0;
        for (int32_t idx = threadIdx.x; idx < nr * kS8GainStride; idx += THREADS) {
            const int32_t i = idx >> 8, j = idx & 255;
            const int32_t c2 = row_c2[i] + j;
            float gv = -INFINITY;                               // slot 255 and columns beyond the sub-row
            if (j < sub_w && c2 < k_v) {
                const uint32_t x = cnt[idx];
                const float tN = x < (uint32_t)kSmallCounts ? tn_small[x] : __fdiv_rn(bump_sum(NlogN, x, fN0, s.logs), np);
                gv = __fadd_rn(__fadd_rn(__fadd_rn(tN, col_term[c2]), rt_local[i]), lognp);
            }
            gain[idx] = gv;
        }
        __syncthreads();
        if (P.dbg) t_gain = clock64() - tg0;
        // ---- the scan: slots are whole blocks and chunk edges are block aligned, so a block never straddles a slot or
        // a chunk and the gain row is uniform per warp-iteration
        float bs = -INFINITY;
        uint32_t bx = 0, bend = 0;
        int ntie = 0;
        unsigned long long seen = 0ull;                        // gain rows that hold `bs` so far
        {
            int32_t crow = crow0;
            uint32_t seg_end = wb_lo < wb_hi ? rs_loc[crow + 1] / kS8Blk : 0u;      // first block of the next slot
            uint32_t cur_u = su_loc[crow];                                          // gain row of the slot (< 64)
            uint32_t grow_b = gain_b + cur_u * (kS8GainStride * 4u);
            const uint4 *src4 = vec + (size_t)wb_lo * kWarp + lane;
            uint4 stage[DEPTH];
#pragma unroll
            for (int r = 0; r < DEPTH; ++r)
                stage[r] = wb_lo + (uint32_t)r < wb_hi ? s8_ld_stream(src4 + (size_t)r * kWarp, stream_pol)
                                                       : make_uint4(0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu);
            // the load that re-uses the staging registers of block (blk0 + r) is issued AFTER that block is scored: its
            // registers are dead by then (issued before, the compiler copies the four of them aside: 8 moves per block);
            // `check`: the tail of the span, where that load may lie beyond it
            auto refill = [&](int r, uint32_t blk, bool check) {
                if (!check || blk + DEPTH < wb_hi) stage[r] = s8_ld_stream(src4 + (size_t)(DEPTH + r) * kWarp, stream_pol);
            };
            auto score_block = [&](const uint4 q, const uint32_t blk) {
                if (blk >= seg_end) {                            // next slot (uniform per warp; slots are not empty)
                    do { ++crow; seg_end = rs_loc[crow + 1] / kS8Blk; } while (blk >= seg_end);
                    cur_u = su_loc[crow];
                    grow_b = gain_b + cur_u * (kS8GainStride * 4u);
                }
                // two halves of eight gathers (the eight results of a half are folded before the next half's addresses are
                // formed: 16 fewer live registers than sixteen gathers at once, which is what lets DEPTH grow)
                float m;
                {
                    float g[8];
#pragma unroll
                    for (int j = 0; j < 4; ++j) { g[j] = s8_gather(q.x, j, grow_b); g[4 + j] = s8_gather(q.y, j, grow_b); }
                    m = fmaxf(fmaxf(fmaxf(g[0], g[1]), fmaxf(g[2], g[3])), fmaxf(fmaxf(g[4], g[5]), fmaxf(g[6], g[7])));
                }
                {
                    float g[8];
#pragma unroll
                    for (int j = 0; j < 4; ++j) { g[j] = s8_gather(q.z, j, grow_b); g[4 + j] = s8_gather(q.w, j, grow_b); }
                    m = fmaxf(m, fmaxf(fmaxf(fmaxf(g[0], g[1]), fmaxf(g[2], g[3])), fmaxf(fmaxf(g[4], g[5]), fmaxf(g[6], g[7]))));
                }
                if (m >= bs) {                                   // rare once the thread has seen a good candidate
                    if (m > bs) {                                // registers only: which block (and its gain row)
                        bs = m; bx = blk | (cur_u << 26); bend = seg_end; ntie = 0; seen = 1ull << cur_u;
                        if (P.prefetch) s8_prefetch_settle(vec, P.vrank, P.pos_s, blk, lane);
                    } else if (blk >= bend && m > -INFINITY) {   // same gain in a later segment
                        const uint32_t u = cur_u;
                        bend = seg_end;
                        // a later PIECE of a sub-row that already holds this gain cannot win: its candidates all come
                        // after those of the earlier piece in the list.  Another sub-row: park the block.
                        if (!((seen >> u) & 1ull)) {
                            seen |= 1ull << u;
                            if (ntie < kS8TieCap) tie[ntie] = blk | (u << 26);   // (list full: see s8_resolve_all)
                            if (P.prefetch) s8_prefetch_settle(vec, P.vrank, P.pos_s, blk, lane);
                            ntie = min(ntie + 1, kS8TieCap + 1);
                            }
                    }
                }
            };
            uint32_t blk0 = wb_lo;
            // main part: every load of the body is in range, no bounds checks
            for (; blk0 + 2 * DEPTH <= wb_hi; blk0 += DEPTH) {
#pragma unroll
                for (int r = 0; r < DEPTH; ++r) {
                    score_block(stage[r], blk0 + (uint32_t)r);
                    refill(r, blk0 + (uint32_t)r, false);
                }
                src4 += (size_t)DEPTH * kWarp;
            }
            // tail: fewer than 2 * DEPTH blocks left
            for (; blk0 < wb_hi; blk0 += DEPTH) {
#pragma unroll
                for (int r = 0; r < DEPTH; ++r) {
                    const uint32_t blk = blk0 + (uint32_t)r;
                    if (blk < wb_hi) {
                        score_block(stage[r], blk);
                        refill(r, blk, true);
                    }
                }
                src4 += (size_t)DEPTH * kWarp;
            }
        }
        const long long t1 = P.dbg ? clock64() : 0;
        if (P.dbg && lane == 0) P.dbg[8 * (long long)grid + 64 * blockIdx.x + threadIdx.x / kWarp] = t1 - t0;   // my warp's scan end
        // ---------------- warp arg-max, then block arg-max ----------------
        // Every warp settles its own best as soon as its span is done (the two memory latencies of that overlap with
        // the warps still streaming); the block then only compares 32 finished keys.
        const uint32_t my32 = bs > -INFINITY ? orderable(bs) : 0u;
        uint32_t m32 = my32;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) m32 = max(m32, __shfl_xor_sync(0xffffffffu, m32, o));
        unsigned long long key = 0ull, pay = 0ull;
        uint32_t bi = 0xFFFFFFFFu;
        if (my32 != 0u && my32 == m32) {
            // this thread holds the warp's best gain: read its recorded blocks again (the gain rows are still staged)
            // and take the earliest candidate; its count comes from the shared-memory copy
            const S8Found f = s8_resolve_all(&ctx, bs, bx, tie, ntie, wb_hi);
            bi = f.bi;
            const int32_t c1 = row_c1[f.brow];
            const uint32_t c2 = (uint32_t)row_c2[f.brow] + f.bbyte;
            const uint32_t x = cnt[f.brow * kS8GainStride + (int32_t)f.bbyte];
            key = ((unsigned long long)m32 << 32) | (unsigned long long)(0xFFFFFFFFu - (base_pos + f.bp));
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
            if (P.dbg && lane == 0) P.dbg[8 * (long long)grid + 64 * blockIdx.x + 32 + threadIdx.x / kWarp] = clock64() - t0;   // settled
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
            d[0] = t_gain; d[1] = t1 - t0 - t_gain - t_pre; d[2] = t2 - t1; d[3] = clock64() - t2;
            unsigned int smid;
            asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
            d[4] = (long long)(e_hi - e_lo) / kS8Blk; d[5] = nr | ((long long)smid << 16); d[6] = t_pre; d[7] = t_learn;
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
        const uint32_t xw = (uint32_t)(wpayload & 0xFFFFFFFFull);
        {   // my copy of the cell's count, if the cell lies in one of my sub-rows (at most one thread matches)
            const int32_t sub_c2 = (c2w / sub_w) * sub_w;
            for (int32_t i = threadIdx.x; i < nr; i += THREADS)
                if (row_c1[i] == c1 && row_c2[i] == sub_c2) cnt[i * kS8GainStride + (c2w - sub_c2)] = xw + 1;
        }
        if (threadIdx.x == 0) {
            if (sh_best_key == win) {                          // keys are unique: exactly one owner CTA
                P.stream[sh_best_idx] = (uint8_t)kS8Removed;   // remove_idx_all mi.py:104-106 (the -inf slot)
                s.cells[(int64_t)key_pos(win) - s.pos_base] = 0xFFFFFFFFu;     // list-order view stays in sync
            }
            const uint32_t y = a_cnt[c2w], z = b_cnt[c1];
            ps[0] = bump_sum(ps[0], xw, ps[4], s.logs);        // update_cache mi.py:383-389
            ps[1] = bump_sum(ps[1], y, ps[5], s.logs);
            ps[2] = bump_sum(ps[2], z, ps[5], s.logs);
            ps[3] = __fadd_rn(ps[3], 1.0f);                    // update_mats :401-406
            a_cnt[c2w] = y + 1; b_cnt[c1] = z + 1;
            if (blockIdx.x == 0) {
                P.out_pos[it] = (int64_t)key_pos(win);
                P.out_gain[it] = key_score(win);
                // the global table: nobody reads it during the launch (every CTA keeps the counts of its sub-rows in
                // shared memory), so it simply follows pick by pick -- a reduction that nothing waits for
                atomicAdd(s.n_cells + (int64_t)c1 * k_v + c2w, 1u);
            }
        }
        done = it + 1;
        __syncthreads();
        if (P.dbg) t_learn = clock64() - t3;
    }
    // ---------------- write the replicated state back (CTA 0) ----------------
    if (blockIdx.x == 0) {
        if (threadIdx.x == 0) {
            for (int64_t j = done; j < P.n_picks; ++j) { P.out_pos[j] = -1; P.out_gain[j] = nanf(""); }
            for (int i = 0; i < 4; ++i) s.sums[i] = ps[i];
        }
        for (int32_t i = threadIdx.x; i < k_v; i += THREADS) s.a_cols[i] = a_cnt[i];
        for (int32_t i = threadIdx.x; i < k_a; i += THREADS) s.b_rows[i] = b_cnt[i];
    }
}

int s8_slots_for_rows(int32_t rows) { return 2 * rows + 8; }
size_t s8_fixed_bytes(int32_t k_a, int32_t k_v, int32_t rows) {
    const size_t slots = (size_t)s8_slots_for_rows(rows);
    const size_t words = 2 * (size_t)k_v + kSmallCounts + (size_t)k_a + (slots + 1) + slots + 3 * (size_t)rows;
    return words * 4;
}
size_t s8_table_bytes(int32_t k_a, int32_t k_v, int32_t rows) {         // dynamic shared memory of the kernel
    return ((s8_fixed_bytes(k_a, k_v, rows) + 1024 + (size_t)rows * kS8GainStride * 8) + 15) & ~(size_t)15;
}

}  // namespace

int mi_s8_k_rows(int32_t k_a, int32_t k_v) { return s8_geom(k_a, k_v).k_rows; }
int mi_s8_block() { return kS8Blk; }
int mi_s8_tiles(int64_t w) { return (int)ceil_div(w > 0 ? w : 1, kS8Tile); }
int64_t mi_s8_stream_capacity(int64_t w, int32_t k_a, int32_t k_v) {
    return w + (int64_t)kS8Blk * s8_geom(k_a, k_v).k_rows + kS8Blk;
}
int mi_s8_slots_for_rows(int32_t rows) { return s8_slots_for_rows(rows); }

// distinct sub-rows a CTA can stage (gain row + count row each)
int mi_s8_rows_that_fit(int32_t k_a, int32_t k_v) {
    int32_t rows = 0;
    while (rows < 64 && s8_table_bytes(k_a, k_v, rows + 1) <= kS8SmemBudget) ++rows;
    return rows;                                              // <= 64: the scan keeps a 64-bit set of gain rows
}

// shape test (the layout of a concrete list can still fail when its sub-rows do not fit: mi_prepare_stream8)
bool mi_s8_supported(int32_t k_a, int32_t k_v) {
    const S8Geom g = s8_geom(k_a, k_v);
    return g.k_rows <= 24000 && mi_s8_rows_that_fit(k_a, k_v) >= 4 && (int64_t)k_a * k_v < (1ll << 31);   // partition histogram: 96 KB of shared memory
}

// first half of the partition: per-tile histograms, their prefix over tiles and the sub-row totals
int launch_mi_s8_count(const uint32_t *cells, int64_t w, int32_t k_a, int32_t k_v, uint32_t *tilehist,
                       uint32_t *row_total, cudaStream_t st) {
    const S8Geom g = s8_geom(k_a, k_v);
    const int ntiles = mi_s8_tiles(w);
    const size_t smem = (size_t)g.k_rows * sizeof(uint32_t);
    if (smem > 96 * 1024) return ACAV_E_UNSUPPORTED;
    static size_t done_count[kMaxDevices];
    if (smem > 48 * 1024) {
        int rc = ensure_dynamic_smem(s8_count_kernel, (size_t)96 * 1024, done_count);
        if (rc) return rc;
    }
    s8_count_kernel<<<ntiles, kS8PartThreads, smem, st>>>(cells, w, g, tilehist);
    ACAV_LAUNCH_CHECK();
    s8_prefix_kernel<<<(unsigned)ceil_div(g.k_rows, 256), 256, 0, st>>>(tilehist, ntiles, g.k_rows, row_total);
    ACAV_LAUNCH_CHECK();
    return 0;
}

// second half: scatter into the staging layout (row_start: offset of every sub-row there, by sub-row id)
int launch_mi_s8_scatter(const uint32_t *cells, int64_t w, int32_t k_a, int32_t k_v, const uint32_t *tilehist,
                         const uint32_t *row_start, uint8_t *stage_stream, uint32_t *stage_pos, int64_t stream_capacity,
                         cudaStream_t st) {
    const S8Geom g = s8_geom(k_a, k_v);
    const int ntiles = mi_s8_tiles(w);
    const size_t smem = (size_t)g.k_rows * sizeof(uint32_t);
    static size_t done_scatter[kMaxDevices];
    if (smem > 48 * 1024) {
        int rc = ensure_dynamic_smem(s8_scatter_kernel, (size_t)96 * 1024, done_scatter);
        if (rc) return rc;
    }
    // padding entries of every sub-row: the "removed" byte (stream_capacity is a multiple of 16)
    const int64_t n16 = stream_capacity / 16;
    s8_fill_kernel<<<(unsigned)ceil_div(n16, 256), 256, 0, st>>>(reinterpret_cast<uint4 *>(stage_stream), n16, 0xFFFFFFFFu);
    ACAV_LAUNCH_CHECK();
    s8_scatter_kernel<<<ntiles, kS8PartThreads, smem, st>>>(cells, w, g, tilehist, row_start, stage_stream, stage_pos);
    ACAV_LAUNCH_CHECK();
    return 0;
}

// third step: every block of the stream is fetched from its place in the staging layout, sorted, and arranged
int launch_mi_s8_block_arrange(const uint8_t *stage_stream, const uint32_t *stage_pos, const uint32_t *blk_src,
                               uint8_t *stream, uint32_t *pos_s, unsigned long long *vrank, int64_t w_padded,
                               cudaStream_t st) {
    const int64_t n_blocks = w_padded / kS8Blk;
    if (n_blocks == 0) return 0;
    const unsigned grid = (unsigned)(n_blocks < 65535 * 16 ? n_blocks : 65535 * 16);
    s8_block_sort_kernel<<<grid, kS8Blk, 0, st>>>(stage_stream, stage_pos, blk_src, stream, pos_s, n_blocks);
    ACAV_LAUNCH_CHECK();
    const int64_t ctas = ceil_div(n_blocks, (int64_t)kS8ArrangeWarps);
    s8_block_arrange_kernel<<<(unsigned)(ctas < 148 * 40 ? ctas : 148 * 40), kS8ArrangeWarps * kWarp, 0, st>>>(
        stream, pos_s, vrank, n_blocks);
    ACAV_LAUNCH_CHECK();
    return 0;
}

template <int THREADS, int DEPTH>
static int launch_s8_variant(MiS8 &P, size_t smem, int32_t grid, cudaStream_t st) {
    static size_t attr_done[kMaxDevices];
    { int rc = ensure_dynamic_smem(mi_stream8_kernel<THREADS, DEPTH>, smem, attr_done); if (rc) return rc; }
    void *args[] = {&P};
    ACAV_CUDA_TRY(cudaLaunchCooperativeKernel((void *)mi_stream8_kernel<THREADS, DEPTH>, dim3(grid), dim3(THREADS), args,
                                              smem, st));
    return 0;
}

int launch_mi_stream8(const MiState &s, uint32_t *n_alt, uint8_t *stream, const uint32_t *pos_s,
                      const unsigned long long *vrank, const uint32_t *slot_start, const uint32_t *slot_row, const uint32_t *slot_u, const void *chunks,
                      int32_t grid, void *pub, unsigned int *bar, int64_t n_picks, int64_t *out_pos, float *out_gain,
                      int32_t rows_smem, int32_t variant, int32_t world, int32_t rank, unsigned int seq_base,
                      void *mail_local, void *const *mail_peer, long long *dbg, int *status,
                      unsigned long long spin_limit_ns, cudaStream_t st, bool sync_clean) {
    MiS8 P;
    P.s = s; P.n_alt = n_alt; P.stream = stream; P.pos_s = pos_s; P.vrank = vrank; P.slot_start = slot_start; P.slot_row = slot_row;
    P.slot_u = slot_u; P.chunks = reinterpret_cast<const S8Chunk *>(chunks);
    P.pub = reinterpret_cast<MiPub *>(pub); P.bar = bar; P.n_picks = n_picks; P.out_pos = out_pos; P.out_gain = out_gain;
    P.g = s8_geom(s.k_a, s.k_v); P.rows_smem = rows_smem; P.slots_smem = s8_slots_for_rows(rows_smem);
    P.fixed_bytes = (int32_t)s8_fixed_bytes(s.k_a, s.k_v, rows_smem);
    P.prefetch = 1;
    if (const char *e = std::getenv("ACAV_MI_S8_PREFETCH")) P.prefetch = std::atoi(e);
    P.world = world; P.rank = rank; P.seq_base = seq_base;
    P.mail_local = reinterpret_cast<MiMail *>(mail_local);
    for (int r = 0; r < kMaxWorld; ++r)
        P.mail_peer[r] = (world > 1 && r < world) ? reinterpret_cast<MiMail *>(mail_peer[r]) : nullptr;
    P.dbg = dbg; P.status = status; P.spin_limit_ns = spin_limit_ns;
    const size_t table_bytes = s8_table_bytes(s.k_a, s.k_v, rows_smem);
    // the barrier words and the records start at zero: mi_refresh_kernel leaves them so after every run (sync_clean);
    // the status word can only be set by a peer timeout
    if (!sync_clean) {
        ACAV_CUDA_TRY(cudaMemsetAsync(pub, 0, mi_pub_bytes(grid), st));
        ACAV_CUDA_TRY(cudaMemsetAsync(bar, 0, 2 * sizeof(unsigned int), st));
    }
    if (world > 1 || !sync_clean) ACAV_CUDA_TRY(cudaMemsetAsync(status, 0, sizeof(int), st));
    switch (variant) {                             // measured at W = 1e8, K = 1024: 32.7 / 37.9 / 33.9 / 33.0 / 38.2 / 33.0 us
        case 1: return launch_s8_variant<512, 8>(P, table_bytes, grid, st);
        case 2: return launch_s8_variant<768, 4>(P, table_bytes, grid, st);
        case 3: return launch_s8_variant<1024, 2>(P, table_bytes, grid, st);
        case 4: return launch_s8_variant<512, 4>(P, table_bytes, grid, st);
        case 5: return launch_s8_variant<1024, 4>(P, table_bytes, grid, st);
        default: return launch_s8_variant<1024, 3>(P, table_bytes, grid, st);
    }
}

}  // namespace acav
