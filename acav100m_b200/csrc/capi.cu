// extern "C" surface of libacav_b200.so (declared in include/acav_b200.h).
#include <algorithm>
#include <cstdlib>
#include <map>
#include <mutex>
#include <unordered_map>
#include <new>
#include <vector>

#include "common.cuh"
#include "kernels.cuh"

using namespace acav;

namespace acav {
bool pdl_enabled() {
    static int on = -1;
    if (on < 0) {
        const char *e = std::getenv("ACAV_NO_PDL");
        on = (e && e[0] == '1') ? 0 : 1;
    }
    return on == 1;
}
}  // namespace acav

struct acav_kmeans {
    int32_t k, d;
    int64_t max_batch;
    int32_t sm_count;
    int64_t bytes;
    // assignment scratch
    float *xn, *cn, *mind;
    unsigned long long *packed;      // exact kernel: per-row (distance, index) keys merged across centroid groups (idle: ~0)
    unsigned int *tickets;           // exact kernel, single-launch variant: finished centroid groups per 64-row block (idle: 0)
    // tensor-core path: bf16 copies, epilogue parameters, screening results, TMA descriptors
    int32_t dp;
    void *xb, *cb, *cparams, *partial;
    int32_t *cand_rows, *cand_ids, *full_rows, *counters;      // counters = {n_cand, n_full}
    alignas(64) unsigned char tmap_x[128];
    alignas(64) unsigned char tmap_c[128];
    alignas(64) unsigned char tmap_c128[128];      // centroid map with 128-row boxes (CTA-pair kernels)
    bool tensor_ready;
    int32_t tile_variant;            // ACAV_TILE_*
    // partition scratch
    uint32_t *blockhist, *lrank, *total, *seg_start, *sorted_rows;
    float *lr_eff;
    bool partition_valid;
    int64_t partition_rows;
    const float *hist_max_of;        // the counts_b buffer whose maximum lr_eff[2] holds (nullptr: none)
    // side stream for the independent kernels of a step (ACAV_KM_NOFORK=1: everything on the caller's stream)
    KmFork fork;
    bool fork_ok;
};

struct acav_kmeans_comm {
    KmComm c;
    float *counts_global;          // [k] summed batch histogram of the current step
    float *lr_eff;
    bool exported, connected, ptrs_borrowed;
};

struct acav_mi {
    MiState s;
    int64_t max_picks;
    int32_t sm_count;
    float *consts_dev;
    bool loaded, tabled;
    // persistent-loop resources (row-partitioned stream), built on first use
    uint16_t *c2s;
    uint32_t *pos_s, *row_start, *row_total, *tilehist, *chunk_start;
    uint32_t *n_alt;
    unsigned char *pub;
    unsigned int *bar;
    int32_t grid, rows_smem;
    int64_t w_sorted;
    bool sorted_valid;
    // multi-GPU mailbox (persistent loop, world > 1)
    int32_t world, rank;
    unsigned int seq_base;
    unsigned char *mail_local;
    void *mail_peer[kMiMaxWorld];
    bool comm_connected;
    int *run_status;             // device word written by the persistent loops (kMiRun*)
    void *clean_pub; unsigned int *clean_bar;   // the pub / bar buffers mi_refresh_kernel zeroed after the last run
    unsigned long long spin_limit_ns;
    long long *dbg;              // optional per-CTA phase timers of the persistent loop
    // cell-index loop resources (candidates sorted by table cell), built on first use
    uint32_t *cx_sorted_pos, *cx_cell_start, *cx_head, *cx_first_pos;
    bool cells_valid;
    // byte-stream loop resources (sub-row partitioned stream, one byte per candidate), built on first use
    uint8_t *s8_stream;
    uint32_t *s8_pos, *s8_row_total, *s8_tilehist;
    unsigned long long *s8_vrank;
    // layout of the one-byte stream (host-built, see s8_build_layout): slots, chunks, pieces of every sub-row
    uint32_t *s8_slot_start, *s8_slot_row, *s8_slot_u, *s8_chunks, *s8_row_start, *s8_blk_src;
    uint8_t *s8_stage_stream;            // staging layout the scatter writes ((c1, sub) order), read by the block sort
    uint32_t *s8_stage_pos;
    int64_t s8_slot_cap;
    int32_t s8_rows_smem, s8_variant, s8_use_cache;
    bool s8_valid;
};

namespace {

// Device buffers of at least 1 MiB are kept by the library when an engine is destroyed and handed to the next engine
// that asks for exactly that size (per device, at most kCacheMaxBytes in total): cudaMalloc / cudaFree of the ~1.5 GB a
// greedy-MI engine needs at W = 1e8 cost 15-70 ms (page-table work, device synchronisation) -- as much as copying the
// candidate list to the device.  Buffers shared with other processes (CUDA IPC) are small and never come through here.
// ACAV_NO_BUFFER_CACHE=1 switches the cache off.
constexpr size_t kCacheMinBytes = (size_t)1 << 20;
constexpr size_t kCacheMaxBytes = (size_t)24 << 30;
struct BufCache {
    std::mutex mu;
    std::multimap<size_t, void *> idle;                   // size -> buffer
    std::unordered_map<void *, size_t> sizes;             // every live or idle buffer handed out by dev_alloc
    size_t idle_bytes = 0;
};
BufCache g_buf_cache[kMaxDevices];
bool buf_cache_on() {
    static int on = -1;
    if (on < 0) { const char *e = std::getenv("ACAV_NO_BUFFER_CACHE"); on = (e && e[0] == '1') ? 0 : 1; }
    return on == 1;
}
BufCache *buf_cache() {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= kMaxDevices) return nullptr;
    return &g_buf_cache[dev];
}

template <typename T>
int dev_alloc(T **p, size_t n, int64_t *bytes) {
    *p = nullptr;
    size_t sz = sizeof(T) * (n ? n : 1);
    BufCache *c = (sz >= kCacheMinBytes && buf_cache_on()) ? buf_cache() : nullptr;
    if (c) {
        std::lock_guard<std::mutex> lock(c->mu);
        auto it = c->idle.find(sz);
        if (it != c->idle.end()) {
            *p = reinterpret_cast<T *>(it->second);
            c->idle.erase(it);
            c->idle_bytes -= sz;
            if (bytes) *bytes += (int64_t)sz;
            return 0;
        }
    }
    cudaError_t e = cudaMalloc(reinterpret_cast<void **>(p), sz);
    if (e != cudaSuccess && c) {                          // out of memory with idle buffers around: give them back, retry
        {
            std::lock_guard<std::mutex> lock(c->mu);
            for (auto &kv : c->idle) { c->sizes.erase(kv.second); cudaFree(kv.second); }
            c->idle.clear();
            c->idle_bytes = 0;
        }
        cudaGetLastError();
        e = cudaMalloc(reinterpret_cast<void **>(p), sz);
    }
    if (e != cudaSuccess) return (int)e;
    if (c) { std::lock_guard<std::mutex> lock(c->mu); c->sizes[*p] = sz; }
    if (bytes) *bytes += (int64_t)sz;
    return 0;
}

// counterpart of dev_alloc for buffers that may be big; the caller has synchronised with every stream that used it
void dev_free(void *p) {
    if (!p) return;
    BufCache *c = buf_cache_on() ? buf_cache() : nullptr;
    if (c) {
        std::lock_guard<std::mutex> lock(c->mu);
        auto it = c->sizes.find(p);
        if (it != c->sizes.end()) {
            if (c->idle_bytes + it->second <= kCacheMaxBytes) {
                c->idle.emplace(it->second, p);
                c->idle_bytes += it->second;
                return;
            }
            c->sizes.erase(it);
        }
    }
    cudaFree(p);
}

int query_sm_count(int32_t *out) {
    int dev = 0;
    ACAV_CUDA_TRY(cudaGetDevice(&dev));
    int n = 0;
    ACAV_CUDA_TRY(cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev));
    *out = n;
    return 0;
}

// Balanced cut of a row-partitioned stream into `grid` contiguous chunks: cost(e) = e + row_cost * (#non-empty rows that
// start before e); chunk edges are multiples of `blk` (rows are padded to whole blocks).  rs = row offsets [n_rows + 1].
std::vector<uint32_t> cut_chunks(const std::vector<uint32_t> &rs, int32_t n_rows, double row_cost, uint32_t blk, int32_t grid) {
    std::vector<double> cum((size_t)n_rows + 1);
    double run = 0.0;
    for (int32_t r = 0; r < n_rows; ++r) {
        cum[r] = run;
        const uint32_t n = rs[r + 1] - rs[r];
        if (n) run += row_cost + (double)n;
    }
    cum[n_rows] = run;
    std::vector<uint32_t> chunks((size_t)grid + 1);
    int32_t r = 0;
    for (int32_t g = 0; g <= grid; ++g) {
        const double t = run * (double)g / (double)grid;
        while (r < n_rows && cum[r + 1] <= t) ++r;
        uint32_t e;
        if (r >= n_rows) e = rs[n_rows];
        else {
            const uint32_t n = rs[r + 1] - rs[r];
            double off = t - cum[r] - row_cost;
            if (off < 0) off = 0;
            if (off > (double)n) off = (double)n;
            e = rs[r] + ((uint32_t)off / blk) * blk;
        }
        chunks[g] = e;
    }
    chunks[0] = 0;
    chunks[grid] = rs[n_rows];
    for (int32_t g = 1; g <= grid; ++g)
        if (chunks[g] < chunks[g - 1]) chunks[g] = chunks[g - 1];
    return chunks;
}

// The same with a cap on the rows a chunk may touch (so that every CTA's gain rows and table counts fit in its shared
// memory and it never falls back to the multi-batch path).  Costs are in blocks: cost(chunk) = blocks + row_cost *
// rows touched.  The smallest per-chunk budget T that needs <= grid chunks is found by bisection; rows are split at
// block boundaries.  Returns grid + 1 offsets (trailing chunks may be empty).
std::vector<uint32_t> cut_chunks_capped(const std::vector<uint32_t> &rs, int32_t n_rows, double row_cost, uint32_t blk,
                                        int32_t grid, int32_t max_rows) {
    auto fill = [&](double T, std::vector<uint32_t> *out) -> int64_t {
        int64_t n_chunks = 0;
        double cost = 0.0;
        int32_t rows_in = 0;
        uint32_t pos = 0;
        if (out) { out->clear(); out->push_back(0); }
        auto close = [&]() {
            if (out) out->push_back(pos);
            ++n_chunks;
            cost = 0.0;
            rows_in = 0;
        };
        for (int32_t r = 0; r < n_rows; ++r) {
            uint32_t nb = (rs[r + 1] - rs[r]) / blk;
            pos = rs[r];
            while (nb > 0) {
                if (rows_in > 0 && (rows_in >= max_rows || cost + row_cost + 1.0 > T)) close();
                double room = T - cost - row_cost;
                uint32_t take = room >= (double)nb ? nb : (room < 1.0 ? 1u : (uint32_t)room);
                cost += row_cost + (double)take;
                ++rows_in;
                pos += take * blk;
                nb -= take;
                if (nb > 0) close();
            }
        }
        pos = rs[n_rows];
        if (rows_in > 0) close();
        return n_chunks;
    };
    double total = 0.0;
    for (int32_t r = 0; r < n_rows; ++r)
        if (rs[r + 1] > rs[r]) total += row_cost + (double)((rs[r + 1] - rs[r]) / blk);
    double lo = row_cost + 1.0, hi = total + row_cost + 1.0;
    for (int it = 0; it < 48 && hi - lo > 0.5; ++it) {
        const double mid = 0.5 * (lo + hi);
        if (fill(mid, nullptr) <= grid) hi = mid; else lo = mid;
    }
    std::vector<uint32_t> chunks;
    fill(hi, &chunks);
    while ((int32_t)chunks.size() < grid + 1) chunks.push_back(rs[n_rows]);
    chunks[grid] = rs[n_rows];
    chunks.resize((size_t)grid + 1);
    return chunks;
}

// Layout of the one-byte stream (ACAV_MI_LOOP_BYTES).  The order of the sub-rows in the stream is free (a candidate
// only has to lie in list order inside its sub-row), so the host places them such that the CTAs' chunks come out equal:
//   * one BIN per CTA, all bins the same number of blocks;
//   * the small sub-rows (the Zipf tail: thousands of one-block sub-rows) are dealt to the bins in snake order, so
//     every bin gets the same number of them -- in (c1, sub) order the CTAs that get the tail would hit the shared-
//     memory row limit with a quarter of the average blocks;
//   * the big sub-rows fill what is left of the bins, in order, cut at bin edges;
//   * inside a bin the big blocks are cut into as many sub-pieces as there are small sub-rows and the two alternate, so
//     every warp's span of the chunk crosses only a few segment borders (each border can park a tie in the scan).
// A stream SLOT is a sub-row or a piece of one; all slots are whole blocks.  Pieces of one sub-row in one bin share a
// gain row (slot_u).  The scatter kernel writes a plain staging layout (sub-rows in id order); blk_src tells the block
// sort where every block of the stream comes from.
struct S8Layout {
    std::vector<uint32_t> slot_start, slot_row, slot_u;          // [n_slots + 1], [n_slots], [n_slots]
    std::vector<uint32_t> chunks;                                // [grid + 1][4] {slot0, n_slots, n_rows, start}
    std::vector<uint32_t> row_start;                             // [k_rows + 1] staging layout: sub-rows in id order
    std::vector<uint32_t> blk_src;                               // [blocks] stream block -> staging block
    uint32_t total = 0;                                          // padded stream length (candidates)
};

bool s8_build_layout(const std::vector<uint32_t> &nb, uint32_t blk, int32_t grid, int32_t rows_cap, int32_t slots_cap,
                     S8Layout *L) {
    const int32_t k_rows = (int32_t)nb.size();
    std::vector<uint32_t> live;
    uint64_t total_blocks = 0;
    for (int32_t r = 0; r < k_rows; ++r)
        if (nb[r]) { live.push_back((uint32_t)r); total_blocks += nb[r]; }
    std::stable_sort(live.begin(), live.end(), [&](uint32_t a, uint32_t b) { return nb[a] > nb[b]; });
    const uint64_t t0 = (total_blocks + grid - 1) / grid;
    const uint64_t small_max = t0 / 4 > 1 ? t0 / 4 : 1;
    struct Piece { uint32_t row, rank0, n; };                   // in blocks
    std::vector<std::vector<uint32_t>> small_of((size_t)grid);
    std::vector<std::vector<Piece>> big_of((size_t)grid);
    std::vector<uint64_t> small_blocks((size_t)grid, 0);
    std::vector<uint32_t> big;
    {   // snake deal of the small sub-rows (sorted by size)
        int64_t i = 0;
        for (uint32_t r : live) {
            if (nb[r] > small_max) { big.push_back(r); continue; }
            const int64_t pass = i / grid, k = i % grid;
            const int32_t g = (int32_t)((pass & 1) ? grid - 1 - k : k);
            small_of[g].push_back(r);
            small_blocks[g] += nb[r];
            ++i;
        }
    }
    uint64_t big_blocks = 0;
    for (uint32_t r : big) big_blocks += nb[r];
    // bin size: the smallest T with room for all big blocks next to the small ones
    uint64_t T = t0;
    for (int32_t g = 0; g < grid; ++g) T = std::max(T, small_blocks[g]);
    for (;;) {
        uint64_t room = 0;
        for (int32_t g = 0; g < grid; ++g) room += T - small_blocks[g];
        if (room >= big_blocks) break;
        T += (big_blocks - room + grid - 1) / grid;
    }
    {   // big sub-rows fill the bins in order
        int32_t g = 0;
        uint64_t room = T - small_blocks[0];
        for (uint32_t r : big) {
            uint32_t left = nb[r], rank0 = 0;
            while (left) {
                while (room == 0) { ++g; room = T - small_blocks[g]; }
                const uint32_t take = (uint32_t)std::min<uint64_t>(left, room);
                big_of[g].push_back({r, rank0, take});
                rank0 += take; left -= take; room -= take;
            }
        }
    }
    // staging layout: sub-rows in id order
    L->row_start.assign((size_t)k_rows + 1, 0);
    for (int32_t r = 0; r < k_rows; ++r) L->row_start[r + 1] = L->row_start[r] + nb[r] * blk;
    // emit: per bin, big sub-pieces and small sub-rows alternate
    L->slot_start.clear(); L->slot_row.clear(); L->slot_u.clear(); L->blk_src.clear();
    L->chunks.assign((size_t)(grid + 1) * 4, 0);
    uint32_t off = 0;
    for (int32_t g = 0; g < grid; ++g) {
        const uint32_t slot0 = (uint32_t)L->slot_row.size();
        uint32_t n_rows = 0;
        auto emit = [&](uint32_t row, uint32_t rank0, uint32_t n, uint32_t u) {
            L->slot_start.push_back(off);
            L->slot_row.push_back(row);
            L->slot_u.push_back(u);
            const uint32_t src0 = L->row_start[row] / blk + rank0;
            for (uint32_t i = 0; i < n; ++i) L->blk_src.push_back(src0 + i);
            off += n * blk;
        };
        uint64_t bigs = 0;
        for (const Piece &p : big_of[g]) bigs += p.n;
        const size_t n_small = small_of[g].size();
        const uint64_t sub = n_small ? std::max<uint64_t>(1, (bigs + n_small) / (n_small + 1)) : (bigs ? bigs : 1);
        size_t si = 0;
        for (const Piece &p : big_of[g]) {
            const uint32_t u = n_rows++;
            uint32_t done = 0;
            while (done < p.n) {
                const uint32_t take = (uint32_t)std::min<uint64_t>(p.n - done, sub);
                emit(p.row, p.rank0 + done, take, u);
                done += take;
                if (si < n_small) { const uint32_t r = small_of[g][si++]; emit(r, 0, nb[r], n_rows++); }
            }
        }
        while (si < n_small) { const uint32_t r = small_of[g][si++]; emit(r, 0, nb[r], n_rows++); }
        const uint32_t n_slots = (uint32_t)L->slot_row.size() - slot0;
        if ((int32_t)n_rows > rows_cap || (int32_t)n_slots > slots_cap) return false;
        uint32_t *c = &L->chunks[(size_t)g * 4];
        c[0] = slot0; c[1] = n_slots; c[2] = n_rows; c[3] = n_slots ? L->slot_start[slot0] : off;
    }
    L->slot_start.push_back(off);
    L->total = off;
    {   uint32_t *c = &L->chunks[(size_t)grid * 4]; c[0] = (uint32_t)L->slot_row.size(); c[1] = 0; c[2] = 0; c[3] = off; }
    return true;
}

// Build (or rebuild) the sub-row partitioned one-byte stream of ACAV_MI_LOOP_BYTES and its per-CTA chunk table.
int mi_prepare_stream8(acav_mi *h, cudaStream_t st) {
    if (h->s8_valid) return 0;
    MiState &s = h->s;
    if (!mi_s8_supported(s.k_a, s.k_v)) return ACAV_E_UNSUPPORTED;
    h->s8_rows_smem = mi_s8_rows_that_fit(s.k_a, s.k_v);
    const int32_t k_rows = mi_s8_k_rows(s.k_a, s.k_v);
    const int ntiles = mi_s8_tiles(s.w);
    // pieces never add padding (they are whole blocks of a padded sub-row), so the capacity of the plain layout holds
    const int64_t cap = (mi_s8_stream_capacity(s.w, s.k_a, s.k_v) + 15) / 16 * 16;
    const uint32_t blk = (uint32_t)mi_s8_block();
    const int32_t slots_cap = mi_s8_slots_for_rows(h->s8_rows_smem);
    const int64_t slot_cap_total = (int64_t)k_rows + (int64_t)h->sm_count * (slots_cap + 2) + 16;
    int rc = 0;
    if (!h->s8_stream) {
        h->s8_slot_cap = slot_cap_total;
        if (!rc) rc = dev_alloc(&h->s8_stream, (size_t)cap, nullptr);
        if (!rc) rc = dev_alloc(&h->s8_pos, (size_t)cap, nullptr);
        if (!rc) rc = dev_alloc(&h->s8_vrank, (size_t)cap / 16 + 1, nullptr);
        if (!rc) rc = dev_alloc(&h->s8_slot_start, (size_t)slot_cap_total + 1, nullptr);
        if (!rc) rc = dev_alloc(&h->s8_slot_row, (size_t)slot_cap_total, nullptr);
        if (!rc) rc = dev_alloc(&h->s8_slot_u, (size_t)slot_cap_total, nullptr);
        if (!rc) rc = dev_alloc(&h->s8_stage_stream, (size_t)cap, nullptr);
        if (!rc) rc = dev_alloc(&h->s8_stage_pos, (size_t)cap, nullptr);
        if (!rc) rc = dev_alloc(&h->s8_row_start, (size_t)k_rows + 1, nullptr);
        if (!rc) rc = dev_alloc(&h->s8_blk_src, (size_t)cap / blk + 1, nullptr);
        if (!rc) rc = dev_alloc(&h->s8_chunks, ((size_t)h->sm_count + 1) * 4, nullptr);
        if (!rc) rc = dev_alloc(&h->s8_row_total, (size_t)k_rows, nullptr);
        if (!rc) rc = dev_alloc(&h->s8_tilehist, (size_t)ntiles * k_rows, nullptr);
        if (!h->n_alt && !rc) rc = dev_alloc(&h->n_alt, (size_t)s.k_a * s.k_v, nullptr);
        if (!h->pub && !rc) rc = dev_alloc(&h->pub, mi_pub_bytes(h->sm_count), nullptr);
        if (!h->bar && !rc) rc = dev_alloc(&h->bar, 2, nullptr);
        if (rc) return rc;
    }
    rc = launch_mi_s8_count(s.cells, s.w, s.k_a, s.k_v, h->s8_tilehist, h->s8_row_total, st);
    if (rc) return rc;
    std::vector<uint32_t> nb((size_t)k_rows);
    ACAV_CUDA_TRY(cudaMemcpyAsync(nb.data(), h->s8_row_total, sizeof(uint32_t) * nb.size(), cudaMemcpyDeviceToHost, st));
    ACAV_CUDA_TRY(cudaStreamSynchronize(st));
    for (int32_t r = 0; r < k_rows; ++r) nb[r] = (nb[r] + blk - 1) / blk;
    S8Layout L;
    if (!s8_build_layout(nb, blk, h->sm_count, h->s8_rows_smem, slots_cap, &L)) return ACAV_E_UNSUPPORTED;
    if ((int64_t)L.slot_row.size() > h->s8_slot_cap || (int64_t)L.total > cap) return ACAV_E_UNSUPPORTED;
    auto up = [&](uint32_t *dst, const std::vector<uint32_t> &v) {
        return v.empty() ? cudaSuccess
                         : cudaMemcpyAsync(dst, v.data(), sizeof(uint32_t) * v.size(), cudaMemcpyHostToDevice, st);
    };
    ACAV_CUDA_TRY(up(h->s8_slot_start, L.slot_start));
    ACAV_CUDA_TRY(up(h->s8_slot_row, L.slot_row));
    ACAV_CUDA_TRY(up(h->s8_slot_u, L.slot_u));
    ACAV_CUDA_TRY(up(h->s8_chunks, L.chunks));
    ACAV_CUDA_TRY(up(h->s8_row_start, L.row_start));
    ACAV_CUDA_TRY(up(h->s8_blk_src, L.blk_src));
    rc = launch_mi_s8_scatter(s.cells, s.w, s.k_a, s.k_v, h->s8_tilehist, h->s8_row_start, h->s8_stage_stream,
                              h->s8_stage_pos, cap, st);
    if (!rc) rc = launch_mi_s8_block_arrange(h->s8_stage_stream, h->s8_stage_pos, h->s8_blk_src, h->s8_stream, h->s8_pos,
                                             h->s8_vrank, L.total, st);
    if (rc) return rc;
    ACAV_CUDA_TRY(cudaStreamSynchronize(st));      // the layout tables are host temporaries
    h->grid = h->sm_count;
    h->s8_valid = true;
    return 0;
}

// Build (or rebuild) the row-partitioned candidate stream and the per-CTA chunk table.
int mi_prepare_persistent(acav_mi *h, cudaStream_t st) {
    if (h->sorted_valid) return 0;
    MiState &s = h->s;
    h->rows_smem = mi_persistent_rows_that_fit(s.k_a, s.k_v);
    if (h->rows_smem < 1 || s.k_v > 16383 || (int64_t)s.k_a * s.k_v >= (1ll << 31)) return ACAV_E_UNSUPPORTED;   // stream holds 4*c2 in 16 bits
    const int ntiles = mi_partition_scratch_tiles(s.w);
    int rc = 0;
    if (!h->c2s) {
        if (!rc) rc = dev_alloc(&h->c2s, (size_t)mi_stream_capacity(s.w, s.k_a), nullptr);
        if (!rc) rc = dev_alloc(&h->pos_s, (size_t)mi_stream_capacity(s.w, s.k_a), nullptr);
        if (!rc) rc = dev_alloc(&h->row_start, (size_t)s.k_a + 1, nullptr);
        if (!rc) rc = dev_alloc(&h->row_total, (size_t)s.k_a, nullptr);
        if (!rc) rc = dev_alloc(&h->tilehist, (size_t)ntiles * s.k_a, nullptr);
        if (!rc) rc = dev_alloc(&h->chunk_start, (size_t)h->sm_count + 1, nullptr);
        if (!rc) rc = dev_alloc(&h->n_alt, (size_t)s.k_a * s.k_v, nullptr);
        if (!rc) rc = dev_alloc(&h->pub, mi_pub_bytes(h->sm_count), nullptr);
        if (!rc) rc = dev_alloc(&h->bar, 2, nullptr);
        if (rc) return rc;
    }
    rc = launch_mi_partition(s.cells, s.w, s.k_a, h->tilehist, h->row_total, h->row_start, h->c2s, h->pos_s,
                             mi_stream_capacity(s.w, s.k_a), (uint16_t)(s.k_v << 2), st);
    if (rc) return rc;
    std::vector<uint32_t> rs((size_t)s.k_a + 1);
    ACAV_CUDA_TRY(cudaMemcpyAsync(rs.data(), h->row_start, sizeof(uint32_t) * rs.size(), cudaMemcpyDeviceToHost, st));
    ACAV_CUDA_TRY(cudaStreamSynchronize(st));
    h->w_sorted = rs[s.k_a];                       // rows padded to whole blocks
    rc = launch_mi_block_sort(h->c2s, h->pos_s, h->w_sorted, st);
    if (rc) return rc;
    const int32_t grid = h->sm_count;
    double row_cost = 6.0 * s.k_v;                // building one gain row ~ scanning 6*K_v candidates (measured)
    if (const char *e = std::getenv("ACAV_MI_ROWCOST")) row_cost = std::atof(e) * s.k_v;
    const std::vector<uint32_t> chunks = cut_chunks(rs, s.k_a, row_cost, (uint32_t)mi_stream_block(), grid);
    ACAV_CUDA_TRY(cudaMemcpyAsync(h->chunk_start, chunks.data(), sizeof(uint32_t) * chunks.size(),
                                  cudaMemcpyHostToDevice, st));
    ACAV_CUDA_TRY(cudaStreamSynchronize(st));      // `chunks` is a host temporary
    h->grid = grid;
    h->sorted_valid = true;
    return 0;
}

// Build (or rebuild) the cell index from the list-order view.
int mi_prepare_cells(acav_mi *h, cudaStream_t st) {
    if (h->cells_valid) return 0;
    MiState &s = h->s;
    const size_t n_cells = (size_t)s.k_a * s.k_v;
    if (s.k_a > 16384 || s.k_v > 16384) return ACAV_E_UNSUPPORTED;
    int rc = 0;
    if (!h->cx_sorted_pos) {
        if (!rc) rc = dev_alloc(&h->cx_sorted_pos, (size_t)s.w + 4, nullptr);
        if (!rc) rc = dev_alloc(&h->cx_cell_start, n_cells + 1, nullptr);
        if (!rc) rc = dev_alloc(&h->cx_head, n_cells, nullptr);
        if (!rc) rc = dev_alloc(&h->cx_first_pos, n_cells, nullptr);
        if (!h->pub && !rc) rc = dev_alloc(&h->pub, mi_pub_bytes(h->sm_count), nullptr);
        if (!h->bar && !rc) rc = dev_alloc(&h->bar, 2, nullptr);
        if (rc) return rc;
    }
    // scratch of the two sort passes, released again after the build
    const int32_t kmax = s.k_a > s.k_v ? s.k_a : s.k_v;
    uint32_t *tilehist = nullptr, *total = nullptr, *start = nullptr, *tmp_cells = nullptr, *tmp_pos = nullptr,
             *sorted_cells = nullptr;
    if (!rc) rc = dev_alloc(&tilehist, (size_t)mi_cells_tiles(s.w) * kmax, nullptr);
    if (!rc) rc = dev_alloc(&total, (size_t)kmax, nullptr);
    if (!rc) rc = dev_alloc(&start, (size_t)kmax + 1, nullptr);
    if (!rc) rc = dev_alloc(&tmp_cells, (size_t)s.w + 4, nullptr);
    if (!rc) rc = dev_alloc(&tmp_pos, (size_t)s.w + 4, nullptr);
    if (!rc) rc = dev_alloc(&sorted_cells, (size_t)s.w + 4, nullptr);
    int64_t n_live = 0;
    if (!rc) rc = launch_mi_cells_build(s, tilehist, total, start, tmp_cells, tmp_pos, sorted_cells, h->cx_sorted_pos,
                                        h->cx_cell_start, h->cx_head, h->cx_first_pos, &n_live, st);
    if (!rc) rc = (int)cudaStreamSynchronize(st);
    if (rc) cudaStreamSynchronize(st);               // the scratch goes back to the buffer cache: no kernel may still use it
    dev_free(tilehist); dev_free(total); dev_free(start); dev_free(tmp_cells); dev_free(tmp_pos); dev_free(sorted_cells);
    if (rc) return rc;
    h->cells_valid = true;
    return 0;
}

}  // namespace

extern "C" {

int acav_abi_version(void) { return ACAV_B200_ABI_VERSION; }

const char *acav_status_string(int status) {
    switch (status) {
        case ACAV_OK: return "ok";
        case ACAV_E_INVALID: return "acav: invalid argument";
        case ACAV_E_UNSUPPORTED: return "acav: unsupported shape or alignment";
        case ACAV_E_STATE: return "acav: call order violated";
        case ACAV_E_NO_DEVICE: return "acav: no sm_100 device";
        default: break;
    }
    if (status > 0) return cudaGetErrorString((cudaError_t)status);
    return "acav: unknown status";
}

int acav_device_info(int *sm_count, int *cc_major, int *cc_minor) {
    int dev = 0;
    ACAV_CUDA_TRY(cudaGetDevice(&dev));
    int v = 0;
    if (sm_count) { ACAV_CUDA_TRY(cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev)); *sm_count = v; }
    if (cc_major) { ACAV_CUDA_TRY(cudaDeviceGetAttribute(&v, cudaDevAttrComputeCapabilityMajor, dev)); *cc_major = v; }
    if (cc_minor) { ACAV_CUDA_TRY(cudaDeviceGetAttribute(&v, cudaDevAttrComputeCapabilityMinor, dev)); *cc_minor = v; }
    return 0;
}

/* ---------------------------------------------------------------- k-means ------------------- */

int acav_kmeans_destroy(acav_kmeans_t *h) {
    if (!h) return 0;
    cudaFree(h->xn); cudaFree(h->cn); cudaFree(h->mind); cudaFree(h->packed); cudaFree(h->tickets);
    cudaFree(h->blockhist); cudaFree(h->lrank); cudaFree(h->total); cudaFree(h->seg_start);
    cudaFree(h->sorted_rows); cudaFree(h->lr_eff);
    cudaFree(h->xb); cudaFree(h->cb); cudaFree(h->cparams); cudaFree(h->partial);
    cudaFree(h->cand_rows); cudaFree(h->cand_ids); cudaFree(h->full_rows); cudaFree(h->counters);
    if (h->fork.side) cudaStreamDestroy(h->fork.side);
    if (h->fork.ev_fork) cudaEventDestroy(h->fork.ev_fork);
    if (h->fork.ev_join) cudaEventDestroy(h->fork.ev_join);
    delete h;
    return 0;
}

int acav_kmeans_create(acav_kmeans_t **out, int32_t k, int32_t d, int64_t max_batch) {
    if (!out || k <= 0 || d <= 0 || max_batch < 0) return ACAV_E_INVALID;
    if (max_batch >= (int64_t)1 << 31) return ACAV_E_UNSUPPORTED;
    *out = nullptr;
    acav_kmeans *h = new (std::nothrow) acav_kmeans();
    if (!h) return (int)cudaErrorMemoryAllocation;
    h->k = k; h->d = d; h->max_batch = max_batch; h->bytes = 0;
    h->partition_valid = false; h->partition_rows = 0; h->hist_max_of = nullptr;
    h->fork.side = nullptr; h->fork.ev_fork = nullptr; h->fork.ev_join = nullptr; h->fork_ok = false;
    {
        const char *e = std::getenv("ACAV_KM_NOFORK");
        if (!(e && e[0] == '1') && cudaStreamCreateWithFlags(&h->fork.side, cudaStreamNonBlocking) == cudaSuccess &&
            cudaEventCreateWithFlags(&h->fork.ev_fork, cudaEventDisableTiming) == cudaSuccess &&
            cudaEventCreateWithFlags(&h->fork.ev_join, cudaEventDisableTiming) == cudaSuccess)
            h->fork_ok = true;
    }
    int rc = query_sm_count(&h->sm_count);
    const int64_t nblk = ceil_div(max_batch, 512);
    if (!rc) rc = dev_alloc(&h->xn, (size_t)max_batch, &h->bytes);
    if (!rc) rc = dev_alloc(&h->cn, (size_t)k, &h->bytes);
    if (!rc) rc = dev_alloc(&h->mind, (size_t)max_batch, &h->bytes);
    if (!rc) rc = dev_alloc(&h->packed, (size_t)max_batch, &h->bytes);
    if (!rc) rc = dev_alloc(&h->tickets, (size_t)max_batch / 64 + 2, &h->bytes);
    if (!rc && cudaMemset(h->packed, 0xFF, sizeof(unsigned long long) * (size_t)(max_batch ? max_batch : 1)) != cudaSuccess) rc = ACAV_E_STATE;
    if (!rc && cudaMemset(h->tickets, 0, sizeof(unsigned int) * ((size_t)max_batch / 64 + 2)) != cudaSuccess) rc = ACAV_E_STATE;
    if (!rc) rc = dev_alloc(&h->blockhist, (size_t)(nblk * k), &h->bytes);
    if (!rc) rc = dev_alloc(&h->lrank, (size_t)max_batch, &h->bytes);
    if (!rc) rc = dev_alloc(&h->total, (size_t)k, &h->bytes);
    if (!rc) rc = dev_alloc(&h->seg_start, (size_t)2 * k + 2, &h->bytes);     // offsets [k+1] + heavy-centroid list
    if (!rc) rc = dev_alloc(&h->sorted_rows, (size_t)max_batch, &h->bytes);
    if (!rc) rc = dev_alloc(&h->lr_eff, 4, &h->bytes);              // [0] lr, [1] 1 - lr (sequential), [2] max of the batch histogram
    // tensor-core path (bf16 operands padded to a multiple of 64 columns)
    h->dp = (int32_t)ceil_div(d, 64) * 64;
    h->tensor_ready = false;
    unsigned char *raw = nullptr;
    if (!rc) { rc = dev_alloc(&raw, (size_t)(max_batch ? max_batch : 1) * h->dp * 2, &h->bytes); h->xb = raw; }
    if (!rc) { rc = dev_alloc(&raw, (size_t)k * h->dp * 2, &h->bytes); h->cb = raw; }
    if (!rc) { rc = dev_alloc(&raw, (size_t)umma_param_bytes(k), &h->bytes); h->cparams = raw; }
    if (!rc) { rc = dev_alloc(&raw, (size_t)umma_partial_bytes(max_batch ? max_batch : 1), &h->bytes); h->partial = raw; }
    if (!rc) rc = dev_alloc(&h->cand_rows, (size_t)max_batch, &h->bytes);
    if (!rc) rc = dev_alloc(&h->cand_ids, (size_t)max_batch * 16, &h->bytes);       // kMaxCand per row
    if (!rc) rc = dev_alloc(&h->full_rows, (size_t)max_batch, &h->bytes);
    if (!rc) rc = dev_alloc(&h->counters, 2, &h->bytes);
    if (rc) { acav_kmeans_destroy(h); return rc; }
    if (max_batch > 0 &&
        make_bf16_tensor_map(h->tmap_x, h->xb, max_batch, h->dp, 128) == 0 &&
        make_bf16_tensor_map(h->tmap_c, h->cb, k, h->dp, umma_tile_n(k)) == 0 &&
        make_bf16_tensor_map(h->tmap_c128, h->cb, k, h->dp, 128) == 0)
        h->tensor_ready = true;
    h->tile_variant = ACAV_TILE_AUTO;
    if (const char *e = std::getenv("ACAV_TILE_VARIANT")) h->tile_variant = std::atoi(e);
    *out = h;
    return 0;
}

int64_t acav_kmeans_workspace_bytes(const acav_kmeans_t *h) { return h ? h->bytes : 0; }

int acav_kmeans_set_tile_variant(acav_kmeans_t *h, int32_t variant) {
    if (!h || variant < ACAV_TILE_AUTO || variant > ACAV_TILE_PAIR_512) return ACAV_E_INVALID;
    h->tile_variant = variant;
    return 0;
}

// which distance-GEMM kernel runs for this shape
static int32_t resolve_tile_variant(const acav_kmeans *h) {
    if (h->tile_variant != ACAV_TILE_AUTO) return h->tile_variant;
    if (h->k >= 256) return ACAV_TILE_PAIR_256;     // measured fastest at K = 1024, D = 2048 (profiles/README.md)
    return ACAV_TILE_SINGLE;
}

int acav_kmeans_assign(acav_kmeans_t *h, const float *x, int64_t b, int64_t ldx,
                       const float *centers, const float *counts,
                       float underused_threshold, float reinit_r,
                       int64_t *best, float *min_dist, float *mean_dist, int32_t *n_refined,
                       int32_t mode, void *stream) {
    if (!h || !centers || !counts || b < 0 || b > h->max_batch) return ACAV_E_INVALID;
    if (b > 0 && (!x || !best || ldx < h->d)) return ACAV_E_INVALID;
    cudaStream_t st = (cudaStream_t)stream;
    if (mode == ACAV_ASSIGN_EXACT) {
        int rc = launch_row_norm2(x, b, h->d, ldx, nullptr, h->xn, st);
        if (!rc) rc = launch_row_norm2(centers, h->k, h->d, h->d, nullptr, h->cn, st);
        float *mind = min_dist ? min_dist : h->mind;
        if (!rc) rc = launch_assign_exact(x, ldx, nullptr, b, nullptr, centers, h->k, h->d, h->xn, h->cn, counts,
                                          underused_threshold, reinit_r, best, mind, h->packed, h->sm_count, st);
        if (!rc && mean_dist) rc = launch_mean(mind, b, mean_dist, st);
        if (!rc && n_refined) ACAV_CUDA_TRY(cudaMemsetAsync(n_refined, 0, 2 * sizeof(int32_t), st));
        return rc;
    }
    if (mode != ACAV_ASSIGN_TENSOR) return ACAV_E_INVALID;
    // the bf16 copy of the batch and the preparation of the centroids do not depend on each other
    int rc = 0;
    if (h->fork_ok) {
        rc = km_fork(&h->fork, st);
        if (!rc) rc = acav_kmeans_prepare_batch(h, x, b, ldx, h->fork.side);
        if (!rc) rc = acav_kmeans_prepare_centers(h, centers, counts, underused_threshold, reinit_r, stream);
        if (!rc) rc = km_join(&h->fork, st);
    } else {
        rc = acav_kmeans_prepare_centers(h, centers, counts, underused_threshold, reinit_r, stream);
        if (!rc) rc = acav_kmeans_prepare_batch(h, x, b, ldx, stream);
    }
    if (!rc) rc = acav_kmeans_assign_prepared(h, x, b, ldx, centers, counts, underused_threshold, reinit_r, best,
                                              min_dist, mean_dist, n_refined, stream);
    return rc;
}

int acav_kmeans_prepare_centers(acav_kmeans_t *h, const float *centers, const float *counts,
                                float underused_threshold, float reinit_r, void *stream) {
    if (!h || !centers || !counts) return ACAV_E_INVALID;
    if (!h->tensor_ready) return ACAV_E_NO_DEVICE;
    cudaStream_t st = (cudaStream_t)stream;
    int rc = launch_prep_rows(centers, h->k, h->d, h->d, h->dp, h->cb, h->cn, st);
    if (!rc) rc = launch_centroid_params(h->cn, counts, h->k, underused_threshold, reinit_r, h->cparams, st);
    return rc;
}

int acav_kmeans_prepare_batch(acav_kmeans_t *h, const float *x, int64_t b, int64_t ldx, void *stream) {
    if (!h || b < 0 || b > h->max_batch || (b > 0 && (!x || ldx < h->d))) return ACAV_E_INVALID;
    if (!h->tensor_ready) return ACAV_E_NO_DEVICE;
    return launch_prep_rows(x, b, h->d, ldx, h->dp, h->xb, h->xn, (cudaStream_t)stream);
}

int acav_kmeans_assign_prepared(acav_kmeans_t *h, const float *x, int64_t b, int64_t ldx,
                                const float *centers, const float *counts,
                                float underused_threshold, float reinit_r,
                                int64_t *best, float *min_dist, float *mean_dist, int32_t *n_refined,
                                void *stream) {
    if (!h || !centers || !counts || b < 0 || b > h->max_batch) return ACAV_E_INVALID;
    if (b > 0 && (!x || !best || ldx < h->d)) return ACAV_E_INVALID;
    if (!h->tensor_ready) return ACAV_E_NO_DEVICE;
    cudaStream_t st = (cudaStream_t)stream;
    // tcgen05 distance GEMM + top-4 screen; merge / classify; exact re-check of near-ties
    int32_t n_split = 1;                       // partial top-4 lists per row
    const int32_t variant = resolve_tile_variant(h);
    ACAV_CUDA_TRY(cudaMemsetAsync(h->counters, 0, 2 * sizeof(int32_t), st));     // for launch_merge_classify
    int rc = variant == ACAV_TILE_SINGLE
                 ? launch_assign_umma(h->tmap_x, h->tmap_c, h->xn, h->cparams, (int32_t)b, h->k, h->dp, h->sm_count,
                                      h->partial, &n_split, st)
                 : launch_assign_pair(h->tmap_x, h->tmap_c128, h->xn, h->cparams, (int32_t)b, h->k, h->dp,
                                      h->sm_count, variant == ACAV_TILE_PAIR_512 ? 2 : 1, h->partial, &n_split, st);
    float *mind = min_dist ? min_dist : h->mind;
    if (!rc) rc = launch_merge_classify(h->partial, (int32_t)b, n_split, h->xn, h->cparams, h->cn, h->k, best, mind, h->cand_rows,
                                        h->cand_ids, h->full_rows, h->counters, st);
    // rows with a short candidate list and rows that need all K centroids are disjoint: both re-checks side by side
    const bool fk = h->fork_ok && !rc;
    if (fk) rc = km_fork(&h->fork, st);
    if (!rc) rc = launch_candidate_refine(x, ldx, h->d, centers, h->xn, h->cn, counts, underused_threshold, reinit_r,
                                          h->cand_rows, h->cand_ids, h->counters, (int32_t)b, best, mind,
                                          fk ? h->fork.side : st);
    if (!rc) rc = launch_assign_exact(x, ldx, h->full_rows, b, h->counters + 1, centers, h->k, h->d, h->xn, h->cn,
                                      counts, underused_threshold, reinit_r, best, mind, h->packed, h->sm_count, st,
                                      h->tickets);
    if (fk && !rc) rc = km_join(&h->fork, st);
    // exact distance to the assigned centroid, only when the caller wants distances back
    if (!rc && (min_dist || mean_dist))
        rc = launch_exact_min_dist(x, b, h->d, ldx, centers, best, h->xn, h->cn, counts, underused_threshold,
                                   reinit_r, mind, st);
    if (!rc && mean_dist) rc = launch_mean(mind, b, mean_dist, st);
    if (!rc && n_refined)
        ACAV_CUDA_TRY(cudaMemcpyAsync(n_refined, h->counters, 2 * sizeof(int32_t), cudaMemcpyDeviceToDevice, st));
    return rc;
}

int acav_kmeans_assign_noise(const float *noise, int32_t k, int64_t b,
                             int64_t *best, float *min_dist, float *mean_dist, void *stream) {
    if (!noise || !best || k <= 0 || b < 0) return ACAV_E_INVALID;
    if (mean_dist && !min_dist) return ACAV_E_INVALID;
    cudaStream_t st = (cudaStream_t)stream;
    int rc = launch_assign_noise(noise, k, b, best, min_dist, st);
    if (!rc && mean_dist) rc = launch_mean(min_dist, b, mean_dist, st);
    return rc;
}

int acav_kmeans_histogram(acav_kmeans_t *h, const int64_t *best, int64_t b, float *counts_b, void *stream) {
    if (!h || !best || !counts_b || b < 0 || b > h->max_batch) return ACAV_E_INVALID;
    // lr_eff[2] <- the histogram's maximum (small batches: the single-block prefix kernel has it anyway), so that the
    // update kernels of acav_kmeans_update_fused can take the learning-rate decision themselves
    bool max_written = false;
    int rc = launch_partition(best, b, h->k, h->blockhist, h->lrank, h->total, h->seg_start, h->sorted_rows,
                              counts_b, (cudaStream_t)stream, h->lr_eff + 2, &max_written);
    h->hist_max_of = (rc == 0 && max_written && b > 0) ? counts_b : nullptr;
    h->partition_valid = (rc == 0);
    h->partition_rows = b;
    return rc;
}

static int update_common(acav_kmeans_t *h, const float *x, int64_t b, int64_t ldx, const float *counts_b,
                         double lr, float *centers, float *counts, float *deltas, int32_t *fallback,
                         cudaStream_t st) {
    if (!h || !x || !counts_b || !centers || !counts || ldx < h->d) return ACAV_E_INVALID;
    if (!h->partition_valid || h->partition_rows != b) return ACAV_E_STATE;
    // One process (no deltas to exchange) and counts_b is the very buffer acav_kmeans_histogram filled: the update
    // kernels decide the step's lr from the maximum left in lr_eff[2] (same arithmetic, one launch less on the chain).
    // Otherwise (all-reduced counts, another buffer, batches beyond the single-block prefix) km_effective_lr_kernel does.
    const bool fold = !deltas && h->hist_max_of == counts_b && counts_b != nullptr && !km_heavy_ring();
    int rc = fold ? 0 : launch_effective_lr(counts_b, h->k, lr, h->lr_eff, fallback, st);
    if (!rc) rc = launch_update(x, ldx, h->k, h->d, h->seg_start, h->sorted_rows, counts_b, h->lr_eff, centers,
                                counts, deltas, nullptr, false, st, h->fork_ok ? &h->fork : nullptr,
                                fold ? lr : -1.0, fold ? fallback : nullptr);
    h->partition_valid = false;
    h->hist_max_of = nullptr;
    return rc;
}

int acav_kmeans_update_fused(acav_kmeans_t *h, const float *x, int64_t b, int64_t ldx,
                             const float *counts_b, double lr,
                             float *centers, float *counts, int32_t *fallback, void *stream) {
    return update_common(h, x, b, ldx, counts_b, lr, centers, counts, nullptr, fallback, (cudaStream_t)stream);
}

int acav_kmeans_update_sequential(acav_kmeans_t *h, const float *x, int64_t b, int64_t ldx, const float *counts_b,
                                  double lr, float *centers, float *counts, void *stream) {
    if (!h || !x || !counts_b || !centers || !counts || ldx < h->d) return ACAV_E_INVALID;
    if (!h->partition_valid || h->partition_rows != b) return ACAV_E_STATE;
    cudaStream_t st = (cudaStream_t)stream;
    int rc = launch_sequential_lr(lr, h->lr_eff, st);
    if (!rc) rc = launch_update(x, ldx, h->k, h->d, h->seg_start, h->sorted_rows, counts_b, h->lr_eff, centers, counts,
                                nullptr, nullptr, true, st);
    h->partition_valid = false;
    return rc;
}

int acav_kmeans_update_local(acav_kmeans_t *h, const float *x, int64_t b, int64_t ldx,
                             const float *counts_b_global, double lr,
                             float *centers, float *counts, float *deltas, int32_t *fallback, void *stream) {
    if (!deltas) return ACAV_E_INVALID;
    return update_common(h, x, b, ldx, counts_b_global, lr, centers, counts, deltas, fallback, (cudaStream_t)stream);
}

int acav_kmeans_apply_deltas(float *centers, const float *deltas, int64_t n, void *stream) {
    if (!centers || !deltas || n < 0) return ACAV_E_INVALID;
    return launch_apply_deltas(centers, deltas, n, (cudaStream_t)stream);
}

/* ---- multi-GPU step over NVLink peer memory (kmeans_comm.cu) ---- */

int acav_kmeans_comm_destroy(acav_kmeans_comm_t *h) {
    if (!h) return 0;
    if (h->connected && !h->ptrs_borrowed)
        for (int r = 0; r < h->c.world; ++r)
            if (r != h->c.rank && h->c.arena[r]) cudaIpcCloseMemHandle(h->c.arena[r]);
    if (h->exported) cudaFree(h->c.arena[h->c.rank]);
    cudaFree(h->c.seq); cudaFree(h->c.status); cudaFree(h->counts_global); cudaFree(h->lr_eff);
    delete h;
    return 0;
}

int acav_kmeans_comm_create(acav_kmeans_comm_t **out, int32_t k, int32_t d, int32_t world, int32_t rank) {
    if (!out || k <= 0 || d <= 0 || world < 1 || rank < 0 || rank >= world) return ACAV_E_INVALID;
    if (world > kKmMaxWorld || d % 4 != 0) return ACAV_E_UNSUPPORTED;
    *out = nullptr;
    acav_kmeans_comm *h = new (std::nothrow) acav_kmeans_comm();
    if (!h) return (int)cudaErrorMemoryAllocation;
    KmComm &c = h->c;
    for (int r = 0; r < kKmMaxWorld; ++r) c.arena[r] = nullptr;
    c.world = world; c.rank = rank; c.k = k; c.d = d; c.k_own = (k + world - 1) / world;
    c.seq = nullptr; c.status = nullptr;
    c.spin_limit_ns = 20ull * 1000000000ull;                               // 20 s; ACAV_KM_SPIN_TIMEOUT_MS overrides
    if (const char *e = std::getenv("ACAV_KM_SPIN_TIMEOUT_MS")) {
        const double ms = std::atof(e);
        if (ms > 0) c.spin_limit_ns = (unsigned long long)(ms * 1e6);
    }
    h->counts_global = nullptr; h->lr_eff = nullptr; h->exported = false; h->connected = false; h->ptrs_borrowed = false;
    int rc = dev_alloc(&c.seq, 1, nullptr);
    if (!rc) rc = dev_alloc(&c.status, 1, nullptr);
    if (!rc) rc = dev_alloc(&h->counts_global, (size_t)k, nullptr);
    if (!rc) rc = dev_alloc(&h->lr_eff, 1, nullptr);
    if (!rc) rc = (int)cudaMemset(c.seq, 0, sizeof(unsigned int));
    if (!rc) rc = (int)cudaMemset(c.status, 0, sizeof(int));
    if (rc) { acav_kmeans_comm_destroy(h); return rc; }
    *out = h;
    return 0;
}

int acav_kmeans_comm_handle_bytes(void) { return (int)sizeof(cudaIpcMemHandle_t); }

int acav_kmeans_comm_export(acav_kmeans_comm_t *h, void *handle_out) {
    if (!h || !handle_out) return ACAV_E_INVALID;
    if (h->exported) return ACAV_E_STATE;
    KmComm &c = h->c;
    const size_t bytes = km_comm_arena_bytes(c.world, c.k, c.d);
    ACAV_CUDA_TRY(cudaMalloc(reinterpret_cast<void **>(&c.arena[c.rank]), bytes));
    h->exported = true;
    ACAV_CUDA_TRY(cudaMemset(c.arena[c.rank], 0, bytes));
    ACAV_CUDA_TRY(cudaIpcGetMemHandle(reinterpret_cast<cudaIpcMemHandle_t *>(handle_out), c.arena[c.rank]));
    return 0;
}

int acav_kmeans_comm_connect(acav_kmeans_comm_t *h, const void *handles) {
    if (!h || !handles) return ACAV_E_INVALID;
    if (!h->exported || h->connected) return ACAV_E_STATE;
    const cudaIpcMemHandle_t *hs = reinterpret_cast<const cudaIpcMemHandle_t *>(handles);
    for (int r = 0; r < h->c.world; ++r) {
        if (r == h->c.rank) continue;
        void *p = nullptr;
        ACAV_CUDA_TRY(cudaIpcOpenMemHandle(&p, hs[r], cudaIpcMemLazyEnablePeerAccess));
        h->c.arena[r] = reinterpret_cast<unsigned char *>(p);
    }
    h->connected = true;
    return 0;
}

void *acav_kmeans_comm_arena(acav_kmeans_comm_t *h) { return (h && h->exported) ? h->c.arena[h->c.rank] : nullptr; }

int acav_kmeans_comm_connect_ptrs(acav_kmeans_comm_t *h, void *const *arenas) {
    if (!h || !arenas) return ACAV_E_INVALID;
    if (!h->exported || h->connected) return ACAV_E_STATE;
    for (int r = 0; r < h->c.world; ++r) {
        if (r == h->c.rank) continue;
        if (!arenas[r]) return ACAV_E_INVALID;
        h->c.arena[r] = reinterpret_cast<unsigned char *>(arenas[r]);
    }
    h->connected = true;
    h->exported = true;
    h->ptrs_borrowed = true;
    return 0;
}

int acav_kmeans_comm_status(acav_kmeans_comm_t *h, int32_t *status_host, void *stream) {
    if (!h || !status_host) return ACAV_E_INVALID;
    int v = 0;
    ACAV_CUDA_TRY(cudaMemcpyAsync(&v, h->c.status, sizeof(int), cudaMemcpyDeviceToHost, (cudaStream_t)stream));
    ACAV_CUDA_TRY(cudaStreamSynchronize((cudaStream_t)stream));
    *status_host = v;
    return 0;
}

int acav_kmeans_update_p2p(acav_kmeans_t *h, acav_kmeans_comm_t *comm, const float *x, int64_t b, int64_t ldx,
                           const float *counts_b_local, double lr, float *centers, float *counts, int32_t *fallback,
                           void *stream) {
    if (!h || !comm || !x || !counts_b_local || !centers || !counts || ldx < h->d) return ACAV_E_INVALID;
    if (comm->c.k != h->k || comm->c.d != h->d) return ACAV_E_INVALID;
    if (!comm->connected && comm->c.world > 1) return ACAV_E_STATE;
    if (comm->c.world == 1 && !comm->exported) return ACAV_E_STATE;
    if (!h->partition_valid || h->partition_rows != b) return ACAV_E_STATE;
    cudaStream_t st = (cudaStream_t)stream;
    int rc = launch_km_hist_exchange(comm->c, counts_b_local, lr, comm->counts_global, comm->lr_eff, fallback, counts, st);
    const KmPush push = km_comm_push_target(comm->c);
    if (!rc) rc = launch_update(x, ldx, h->k, h->d, h->seg_start, h->sorted_rows, comm->counts_global, comm->lr_eff,
                                centers, counts, nullptr, &push, false, st, h->fork_ok ? &h->fork : nullptr);
    if (!rc) rc = launch_km_reduce_broadcast(comm->c, comm->counts_global, comm->lr_eff, centers, st);
    h->partition_valid = false;
    return rc;
}

int acav_kmeans_underused_flags(const float *counts, int32_t k, const float *threshold_dev, float *flags, void *stream) {
    if (!counts || !threshold_dev || !flags || k <= 0) return ACAV_E_INVALID;
    return launch_underused_flags(counts, k, threshold_dev, flags, (cudaStream_t)stream);
}

/* ---------------------------------------------------------------- greedy MI ----------------- */

int acav_mi_destroy(acav_mi_t *h) {
    if (!h) return 0;
    MiState &s = h->s;
    cudaDeviceSynchronize();                 // the buffers go back to the library's cache, not through cudaFree
    dev_free(s.cells); dev_free(s.n_cells); dev_free(s.a_cols); dev_free(s.b_rows); dev_free(s.gain);
    dev_free(s.col_term); dev_free(s.row_term); dev_free(s.sums); dev_free(s.key); dev_free(h->consts_dev);
    dev_free(h->c2s); dev_free(h->pos_s); dev_free(h->row_start); dev_free(h->row_total); dev_free(h->tilehist);
    dev_free(h->chunk_start); dev_free(h->n_alt); dev_free(h->pub); dev_free(h->bar);
    if (h->comm_connected)
        for (int r = 0; r < h->world; ++r)
            if (r != h->rank && h->mail_peer[r]) cudaIpcCloseMemHandle(h->mail_peer[r]);
    cudaFree(h->mail_local); dev_free(h->run_status);
    dev_free(h->cx_sorted_pos); dev_free(h->cx_cell_start); dev_free(h->cx_head); dev_free(h->cx_first_pos);
    dev_free(h->s8_stream); dev_free(h->s8_pos); dev_free(h->s8_vrank); dev_free(h->s8_row_total); dev_free(h->s8_tilehist);
    dev_free(h->s8_slot_start); dev_free(h->s8_slot_row); dev_free(h->s8_slot_u); dev_free(h->s8_chunks);
    dev_free(h->s8_row_start); dev_free(h->s8_blk_src); dev_free(h->s8_stage_stream); dev_free(h->s8_stage_pos);
    delete h;
    return 0;
}

int acav_mi_create(acav_mi_t **out, int64_t w, int32_t k_a, int32_t k_v, int64_t max_picks, int64_t pos_base) {
    if (!out || w < 0 || k_a <= 0 || k_v <= 0 || max_picks < 0 || pos_base < 0) return ACAV_E_INVALID;
    if (k_a > 65535 || k_v > 65535) return ACAV_E_UNSUPPORTED;            // 2 x uint16 packing, 0xFFFF.. = tombstone
    if (pos_base + w >= 0xFFFFFFFFll) return ACAV_E_UNSUPPORTED;          // positions live in 32 bits of the key
    if (max_picks >= (1ll << 24)) return ACAV_E_UNSUPPORTED;              // fp32 counts must stay exact integers
    *out = nullptr;
    acav_mi *h = new (std::nothrow) acav_mi();
    if (!h) return (int)cudaErrorMemoryAllocation;
    MiState &s = h->s;
    s = MiState();
    s.w = w; s.k_a = k_a; s.k_v = k_v; s.pos_base = pos_base; s.logs = nullptr; s.n_logs = 0;
    h->max_picks = max_picks; h->loaded = false; h->tabled = false; h->consts_dev = nullptr;
    h->c2s = nullptr; h->pos_s = nullptr; h->row_start = nullptr; h->row_total = nullptr; h->tilehist = nullptr;
    h->chunk_start = nullptr; h->n_alt = nullptr; h->pub = nullptr; h->bar = nullptr; h->clean_pub = nullptr; h->clean_bar = nullptr; h->grid = 0; h->rows_smem = 0;
    h->w_sorted = 0; h->world = 1; h->rank = 0; h->seq_base = 0; h->mail_local = nullptr; h->comm_connected = false;
    for (int r = 0; r < kMiMaxWorld; ++r) h->mail_peer[r] = nullptr;
    h->dbg = nullptr; h->run_status = nullptr; h->clean_pub = nullptr; h->clean_bar = nullptr;
    h->spin_limit_ns = 20ull * 1000000000ull;                             // 20 s; ACAV_MI_SPIN_TIMEOUT_MS overrides
    if (const char *e = std::getenv("ACAV_MI_SPIN_TIMEOUT_MS")) {
        const double ms = std::atof(e);
        if (ms > 0) h->spin_limit_ns = (unsigned long long)(ms * 1e6);
    }
    h->sorted_valid = false;
    h->cx_sorted_pos = nullptr; h->cx_cell_start = nullptr; h->cx_head = nullptr; h->cx_first_pos = nullptr;
    h->cells_valid = false;
    h->s8_stream = nullptr; h->s8_pos = nullptr; h->s8_vrank = nullptr; h->s8_row_total = nullptr; h->s8_tilehist = nullptr;
    h->s8_slot_start = nullptr; h->s8_slot_row = nullptr; h->s8_slot_u = nullptr; h->s8_chunks = nullptr;
    h->s8_row_start = nullptr; h->s8_blk_src = nullptr; h->s8_stage_stream = nullptr; h->s8_stage_pos = nullptr;
    h->s8_slot_cap = 0;
    h->s8_rows_smem = 0; h->s8_valid = false;
    h->s8_variant = 0; h->s8_use_cache = 1;
    if (const char *e = std::getenv("ACAV_MI_S8_VARIANT")) h->s8_variant = std::atoi(e);
    if (const char *e = std::getenv("ACAV_MI_S8_CACHE")) h->s8_use_cache = std::atoi(e);
    int rc = query_sm_count(&h->sm_count);
    const size_t cells = (size_t)k_a * k_v;
    if (!rc) rc = dev_alloc(&s.cells, (size_t)w + 4, nullptr);
    if (!rc) rc = dev_alloc(&s.n_cells, cells, nullptr);
    if (!rc) rc = dev_alloc(&s.a_cols, (size_t)k_v, nullptr);
    if (!rc) rc = dev_alloc(&s.b_rows, (size_t)k_a, nullptr);
    if (!rc) rc = dev_alloc(&s.gain, cells, nullptr);
    if (!rc) rc = dev_alloc(&s.col_term, (size_t)k_v, nullptr);
    if (!rc) rc = dev_alloc(&s.row_term, (size_t)k_a, nullptr);
    if (!rc) rc = dev_alloc(&s.sums, 8, nullptr);
    if (!rc) rc = dev_alloc(&s.key, 2, nullptr);
    if (!rc) rc = dev_alloc(&h->consts_dev, 8, nullptr);
    if (!rc) rc = dev_alloc(&h->run_status, 1, nullptr);
    if (!rc) rc = (int)cudaMemset(h->run_status, 0, sizeof(int));
    if (rc) { acav_mi_destroy(h); return rc; }
    *out = h;
    return 0;
}

int acav_mi_load_candidates(acav_mi_t *h, const int64_t *cells, void *stream) {
    if (!h || (!cells && h->s.w > 0)) return ACAV_E_INVALID;
    int rc = launch_mi_pack(cells, h->s.w, h->s.cells, (cudaStream_t)stream);
    h->loaded = (rc == 0);
    h->sorted_valid = false;
    h->cells_valid = false;
    h->s8_valid = false;
    return rc;
}

int acav_mi_set_tables(acav_mi_t *h, const float *logs, int64_t n_logs, const float *consts, void *stream) {
    if (!h || !logs || !consts || n_logs < h->max_picks + 3) return ACAV_E_INVALID;
    cudaStream_t st = (cudaStream_t)stream;
    h->s.logs = logs; h->s.n_logs = n_logs;
    ACAV_CUDA_TRY(cudaMemcpyAsync(h->consts_dev, consts, 6 * sizeof(float), cudaMemcpyHostToDevice, st));
    ACAV_CUDA_TRY(cudaStreamSynchronize(st));      // `consts` is pageable host memory owned by the caller
    int rc = launch_mi_reset(h->s, h->consts_dev, st);
    h->tabled = (rc == 0);
    return rc;
}

int acav_mi_add_sample(acav_mi_t *h, int32_t c1, int32_t c2, void *stream) {
    if (!h || c1 < 0 || c2 < 0 || c1 >= h->s.k_a || c2 >= h->s.k_v) return ACAV_E_INVALID;
    if (!h->tabled) return ACAV_E_STATE;
    return launch_mi_add_sample(h->s, c1, c2, (cudaStream_t)stream);
}

int acav_mi_local_best(acav_mi_t *h, uint64_t *key_cell, void *stream) {
    if (!h || !key_cell) return ACAV_E_INVALID;
    if (!h->tabled || !h->loaded) return ACAV_E_STATE;
    cudaStream_t st = (cudaStream_t)stream;
    int rc = launch_mi_gain_table(h->s, st);
    if (!rc) rc = launch_mi_scan(h->s, h->sm_count, st);
    if (!rc) rc = launch_mi_emit(h->s, reinterpret_cast<unsigned long long *>(key_cell), st);
    return rc;
}

int acav_mi_apply(acav_mi_t *h, const uint64_t *key_cells, int32_t n, int64_t *out_pos, float *out_gain,
                  void *stream) {
    if (!h || !key_cells || n <= 0) return ACAV_E_INVALID;
    if (!h->tabled || !h->loaded) return ACAV_E_STATE;
    h->sorted_valid = false;                 // the row-partitioned stream does not see this removal
    h->cells_valid = false;                  // nor does the cell index
    h->s8_valid = false;
    return launch_mi_apply(h->s, reinterpret_cast<const unsigned long long *>(key_cells), n, out_pos, out_gain,
                           (cudaStream_t)stream);
}

static bool mi_sync_clean(const acav_mi_t *h) {
    return h->pub && h->bar && h->clean_pub == (void *)h->pub && h->clean_bar == h->bar;
}
// row / column terms from the state the loop wrote back; the loop's barrier words and records zeroed for the next run
static int mi_refresh_after_run(acav_mi_t *h, cudaStream_t st) {
    int rc = launch_mi_refresh_terms(h->s, st, h->pub, mi_pub_bytes(h->sm_count), h->bar);
    h->clean_pub = rc ? nullptr : (void *)h->pub;
    h->clean_bar = rc ? nullptr : h->bar;
    return rc;
}

int acav_mi_run(acav_mi_t *h, int64_t n_picks, int64_t *out_pos, float *out_gain, int32_t mode, void *stream) {
    if (!h || n_picks < 0 || (n_picks > 0 && (!out_pos || !out_gain))) return ACAV_E_INVALID;
    if (!h->tabled || !h->loaded) return ACAV_E_STATE;
    cudaStream_t st = (cudaStream_t)stream;
    if (mode == ACAV_MI_LOOP_CELLS) {
        if (n_picks == 0) return 0;
        if (h->world > 1 && !h->comm_connected) return ACAV_E_STATE;
        int rc = mi_prepare_cells(h, st);
        if (rc) return rc;
        const int32_t grid = h->sm_count < h->s.k_a ? h->sm_count : h->s.k_a;
        h->sorted_valid = false;                 // the candidate streams do not see these removals
        h->s8_valid = false;
        rc = launch_mi_cells(h->s, h->cx_cell_start, h->cx_sorted_pos, h->cx_head, h->cx_first_pos, grid, h->pub, h->bar,
                             n_picks, out_pos, out_gain, h->world, h->rank, h->seq_base, h->mail_local, h->mail_peer,
                             h->run_status, h->spin_limit_ns, st, mi_sync_clean(h));
        h->clean_pub = nullptr;
        h->seq_base += (unsigned int)n_picks + 1u;
        if (!rc) rc = mi_refresh_after_run(h, st);
        return rc;
    }
    if (mode == ACAV_MI_LOOP_KERNELS) {
        if (n_picks > 0) { h->sorted_valid = false; h->cells_valid = false; h->s8_valid = false; }
        for (int64_t it = 0; it < n_picks; ++it) {
            int rc = launch_mi_gain_table(h->s, st);
            if (!rc) rc = launch_mi_scan(h->s, h->sm_count, st);
            if (!rc) rc = launch_mi_apply(h->s, nullptr, 0, out_pos + it, out_gain + it, st);
            if (rc) return rc;
        }
        return 0;
    }
    if (mode == ACAV_MI_LOOP_BYTES) {
        if (n_picks == 0) return 0;
        if (h->world > 1 && !h->comm_connected) return ACAV_E_STATE;
        int rc = mi_prepare_stream8(h, st);
        if (rc) return rc;
        h->cells_valid = false;                  // neither the cell index nor the 2-byte stream see these removals
        h->sorted_valid = false;
        rc = launch_mi_stream8(h->s, h->n_alt, h->s8_stream, h->s8_pos, h->s8_vrank, h->s8_slot_start, h->s8_slot_row, h->s8_slot_u,
                               h->s8_chunks, h->grid, h->pub, h->bar, n_picks, out_pos, out_gain, h->s8_rows_smem, h->s8_variant,
                               h->world, h->rank, h->seq_base, h->mail_local, h->mail_peer, h->dbg, h->run_status,
                               h->spin_limit_ns, st, mi_sync_clean(h));
        h->clean_pub = nullptr;
        h->seq_base += (unsigned int)n_picks + 1u;
        if (!rc) rc = mi_refresh_after_run(h, st);
        return rc;
    }
    if (mode != ACAV_MI_LOOP_PERSISTENT) return ACAV_E_INVALID;
    if (n_picks == 0) return 0;
    if (h->world > 1 && !h->comm_connected) return ACAV_E_STATE;
    int rc = mi_prepare_persistent(h, st);
    if (rc) return rc;
    h->cells_valid = false;                      // the cell index does not see these removals
    h->s8_valid = false;
    rc = launch_mi_persistent(h->s, h->n_alt, h->c2s, h->pos_s, h->row_start, h->chunk_start, h->grid, h->pub, h->bar,
                              n_picks, out_pos, out_gain, h->rows_smem, h->world, h->rank, h->seq_base, h->mail_local,
                              h->mail_peer, h->dbg, h->run_status, h->spin_limit_ns, st, mi_sync_clean(h));
    h->clean_pub = nullptr;
    h->seq_base += (unsigned int)n_picks + 1u;          // mailbox tags never repeat across runs
    if (!rc) rc = mi_refresh_after_run(h, st);
    return rc;
}

int acav_mi_loop_supported(int32_t k_a, int32_t k_v, int32_t mode) {
    if (k_a <= 0 || k_v <= 0 || k_a > 65535 || k_v > 65535) return 0;
    if (mode == ACAV_MI_LOOP_KERNELS) return 1;
    if (mode == ACAV_MI_LOOP_PERSISTENT)       // one gain row must fit in shared memory, the stream holds 4*c2 in 16 bits
        return mi_persistent_rows_that_fit(k_a, k_v) >= 1 && k_v <= 16383 && (int64_t)k_a * k_v < (1ll << 31);
    if (mode == ACAV_MI_LOOP_CELLS) return k_a <= 16384 && k_v <= 16384 && mi_cells_smem_fits(k_a, k_v);
    if (mode == ACAV_MI_LOOP_BYTES) return mi_s8_supported(k_a, k_v) ? 1 : 0;
    return 0;
}

int acav_mi_prepare(acav_mi_t *h, int32_t mode, void *stream) {
    if (!h) return ACAV_E_INVALID;
    if (!h->loaded) return ACAV_E_STATE;
    if (!acav_mi_loop_supported(h->s.k_a, h->s.k_v, mode)) return ACAV_E_UNSUPPORTED;
    if (mode == ACAV_MI_LOOP_PERSISTENT) return mi_prepare_persistent(h, (cudaStream_t)stream);
    if (mode == ACAV_MI_LOOP_CELLS) return mi_prepare_cells(h, (cudaStream_t)stream);
    if (mode == ACAV_MI_LOOP_BYTES) return mi_prepare_stream8(h, (cudaStream_t)stream);
    return 0;
}

int acav_mi_status(acav_mi_t *h, int32_t *status_host, void *stream) {
    if (!h || !status_host) return ACAV_E_INVALID;
    int v = 0;
    ACAV_CUDA_TRY(cudaMemcpyAsync(&v, h->run_status, sizeof(int), cudaMemcpyDeviceToHost, (cudaStream_t)stream));
    ACAV_CUDA_TRY(cudaStreamSynchronize((cudaStream_t)stream));
    *status_host = v;
    return 0;
}

int acav_mi_debug_timers(acav_mi_t *h, int64_t *cycles) {
    if (!h) return ACAV_E_INVALID;
    h->dbg = reinterpret_cast<long long *>(cycles);
    return 0;
}

int acav_mi_set_stream_variant(acav_mi_t *h, int32_t variant, int32_t use_cache) {
    if (!h || variant < 0 || variant > 5) return ACAV_E_INVALID;
    h->s8_variant = variant;
    h->s8_use_cache = use_cache ? 1 : 0;
    return 0;
}

int acav_mi_comm_handle_bytes(void) { return (int)sizeof(cudaIpcMemHandle_t); }

int acav_mi_comm_export(acav_mi_t *h, int32_t world, int32_t rank, void *handle_out) {
    if (!h || !handle_out || world < 1 || world > kMiMaxWorld || rank < 0 || rank >= world) return ACAV_E_INVALID;
    if (h->mail_local) return ACAV_E_STATE;
    int rc = dev_alloc(&h->mail_local, mi_mail_bytes(world), nullptr);
    if (rc) return rc;
    ACAV_CUDA_TRY(cudaMemset(h->mail_local, 0, mi_mail_bytes(world)));
    h->world = world; h->rank = rank;
    ACAV_CUDA_TRY(cudaIpcGetMemHandle(reinterpret_cast<cudaIpcMemHandle_t *>(handle_out), h->mail_local));
    return 0;
}

int acav_mi_comm_connect(acav_mi_t *h, const void *handles) {
    if (!h || !handles) return ACAV_E_INVALID;
    if (!h->mail_local || h->comm_connected) return ACAV_E_STATE;
    const cudaIpcMemHandle_t *hs = reinterpret_cast<const cudaIpcMemHandle_t *>(handles);
    for (int r = 0; r < h->world; ++r) {
        if (r == h->rank) { h->mail_peer[r] = h->mail_local; continue; }
        void *p = nullptr;
        ACAV_CUDA_TRY(cudaIpcOpenMemHandle(&p, hs[r], cudaIpcMemLazyEnablePeerAccess));
        h->mail_peer[r] = p;
    }
    h->comm_connected = true;
    return 0;
}

/* ---------------------------------------------------------------- batch_mi scoring -------- */

struct acav_mi_dense {
    MiDense s;
    float *per_pair;
    int64_t per_pair_cap;
    double *ami_scratch;         // 4*P*C + P doubles, allocated on the first acav_mi_dense_score_ami
};

int acav_mi_dense_destroy(acav_mi_dense_t *h) {
    if (!h) return 0;
    cudaFree(h->s.n_cells); cudaFree(h->s.a_cols); cudaFree(h->s.b_rows); cudaFree(h->s.n); cudaFree(h->s.sums);
    cudaFree(h->per_pair);
    cudaFree(h->ami_scratch);
    delete h;
    return 0;
}

int acav_mi_dense_create(acav_mi_dense_t **out, int32_t p, int32_t c, void *stream) {
    if (!out || p <= 0 || c <= 0) return ACAV_E_INVALID;
    *out = nullptr;
    acav_mi_dense *h = new (std::nothrow) acav_mi_dense();
    if (!h) return (int)cudaErrorMemoryAllocation;
    h->s = MiDense(); h->s.p = p; h->s.c = c; h->per_pair = nullptr; h->per_pair_cap = 0; h->ami_scratch = nullptr;
    int rc = dev_alloc(&h->s.n_cells, (size_t)p * c * c, nullptr);
    if (!rc) rc = dev_alloc(&h->s.a_cols, (size_t)p * c, nullptr);
    if (!rc) rc = dev_alloc(&h->s.b_rows, (size_t)p * c, nullptr);
    if (!rc) rc = dev_alloc(&h->s.n, (size_t)p, nullptr);
    if (!rc) rc = dev_alloc(&h->s.sums, (size_t)p * 3, nullptr);
    if (!rc) rc = launch_mi_dense_reset(h->s, (cudaStream_t)stream);
    if (rc) { acav_mi_dense_destroy(h); return rc; }
    *out = h;
    return 0;
}

int acav_mi_dense_add(acav_mi_dense_t *h, const int64_t *cells, int64_t m, void *stream) {
    if (!h || m < 0 || (m > 0 && !cells)) return ACAV_E_INVALID;
    return launch_mi_dense_add(h->s, cells, m, (cudaStream_t)stream);
}

namespace {
// per-(candidate, pair) scratch of the dense scorers when the caller does not ask for the per-pair values
int dense_per_pair(acav_mi_dense *h, int64_t nb, float *per_pair, float **out, cudaStream_t st) {
    *out = per_pair;
    if (per_pair) return 0;
    if (h->per_pair_cap < nb * h->s.p) {
        ACAV_CUDA_TRY(cudaStreamSynchronize(st));
        cudaFree(h->per_pair);
        h->per_pair = nullptr;
        h->per_pair_cap = 0;
        int rc = dev_alloc(&h->per_pair, (size_t)(nb * h->s.p), nullptr);
        if (rc) return rc;
        h->per_pair_cap = nb * h->s.p;
    }
    *out = h->per_pair;
    return 0;
}
}  // namespace

int acav_mi_dense_score(acav_mi_dense_t *h, const int64_t *cells, int64_t nb, float *scores, float *per_pair,
                        void *stream) {
    if (!h || nb < 0 || (nb > 0 && (!cells || !scores))) return ACAV_E_INVALID;
    float *pp = nullptr;
    int rc = dense_per_pair(h, nb, per_pair, &pp, (cudaStream_t)stream);
    if (rc) return rc;
    return launch_mi_dense_score(h->s, cells, nb, pp, scores, (cudaStream_t)stream);
}

int acav_mi_dense_score_exact(acav_mi_dense_t *h, const int64_t *cells, int64_t nb, const float *logs, int64_t n_logs,
                              const float *consts, float *scores, float *per_pair, void *stream) {
    if (!h || nb < 0 || (nb > 0 && (!cells || !scores)) || !logs || !consts) return ACAV_E_INVALID;
    if (nb >= 65536ll * 32768ll) return ACAV_E_UNSUPPORTED;
    (void)n_logs;                                  // the caller guarantees logs[k] for k <= samples in the tables + 1
    float *pp = nullptr;
    int rc = dense_per_pair(h, nb, per_pair, &pp, (cudaStream_t)stream);
    if (rc) return rc;
    return launch_mi_dense_score_exact(h->s, cells, nb, logs, consts, pp, scores, (cudaStream_t)stream);
}

int acav_mi_dense_score_ami(acav_mi_dense_t *h, const int64_t *cells, int64_t nb, int32_t average_method,
                            float *scores, float *per_pair, void *stream) {
    if (!h || nb < 0 || (nb > 0 && (!cells || !scores)) || average_method < 0 || average_method > 2) return ACAV_E_INVALID;
    if (h->s.p > 65535) return ACAV_E_UNSUPPORTED;                     // grid.y of the line sums
    float *pp = nullptr;
    int rc = dense_per_pair(h, nb, per_pair, &pp, (cudaStream_t)stream);
    if (rc) return rc;
    if (!h->ami_scratch) {
        rc = dev_alloc(&h->ami_scratch, (size_t)4 * h->s.p * h->s.c + (size_t)h->s.p, nullptr);
        if (rc) return rc;
    }
    return launch_mi_dense_score_ami(h->s, cells, nb, h->ami_scratch, average_method, pp, scores, (cudaStream_t)stream);
}

int acav_mi_read_state(acav_mi_t *h, uint32_t *n_cells, uint32_t *a_cols, uint32_t *b_rows, float *sums,
                       void *stream) {
    if (!h) return ACAV_E_INVALID;
    cudaStream_t st = (cudaStream_t)stream;
    const MiState &s = h->s;
    if (n_cells) ACAV_CUDA_TRY(cudaMemcpyAsync(n_cells, s.n_cells, sizeof(uint32_t) * (size_t)s.k_a * s.k_v, cudaMemcpyDeviceToDevice, st));
    if (a_cols) ACAV_CUDA_TRY(cudaMemcpyAsync(a_cols, s.a_cols, sizeof(uint32_t) * (size_t)s.k_v, cudaMemcpyDeviceToDevice, st));
    if (b_rows) ACAV_CUDA_TRY(cudaMemcpyAsync(b_rows, s.b_rows, sizeof(uint32_t) * (size_t)s.k_a, cudaMemcpyDeviceToDevice, st));
    if (sums) ACAV_CUDA_TRY(cudaMemcpyAsync(sums, s.sums, sizeof(float) * 4, cudaMemcpyDeviceToDevice, st));
    return 0;
}

}  // extern "C"
