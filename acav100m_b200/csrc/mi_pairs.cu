// Exact greedy mutual-information selection over P > 1 clustering pairs (the reference's `mem_mi` with its default
// `combination` pairing of ten layer clusterings: P = 45).
//
// Replaces EfficientMemMI (subset_selection/code/measures/mi.py:284-412) driven by EfficientMI.run_greedy (:150-192) and
// calc_score (:76-80) for any number of pairs.  Per iteration the reference builds, for every remaining candidate w and
// pair p, the score of adding w to table p (get_last :322-333, calc_MI :368-381 -- through [P, W, C] gathers), takes
// `scores.mean(dim=-1)` and the first arg-max, adopts the winner's running sums and bumps the P tables (:383-406).
//
// Here one iteration is four launches:
//   mip_gain_kernel   P*C*C scores, one per table cell (the score of pair p depends on the candidate only through its
//                     cell of table p) -- same fp32 operation order as the reference, no FMA contraction;
//   mip_scan_kernel   streams the candidates' cluster ids (uint16, one column per clustering, column-major so a warp
//                     reads 64 contiguous bytes per column: 2*D bytes per candidate instead of the reference's
//                     16*P), gathers the P cell scores from the L2-resident score tables, adds them in the order
//                     torch's CPU `mean` adds them (mi_pairs_math.h) and keeps (score, earliest position) in a 64-bit key;
//   mip_emit_kernel   writes the winner's record (key + its D ids) -- also the unit of the multi-GPU exchange;
//   mip_apply_kernel  one CTA per pair: bump table, marginals, running sums, recompute the pair's marginal terms;
//                     CTA 0 tombstones the winner (nothing moves, so list order is preserved).
// The P = 1 engine (mi_scan.cu / mi_persistent.cu / mi_cells.cu) stays the path for one pair.
#include <new>

#include "common.cuh"
#include "mi_pairs_math.h"

namespace acav {

constexpr uint16_t kGone = 0xFFFFu;
constexpr int kScanThreads = 256;

struct MiPairs {
    uint16_t *ids;             // [d][w_pad] cluster ids, column-major; ids[0][w] == 0xFFFF: candidate removed
    uint32_t *n_cells;         // [P][C][C] contingency counts        (cache['N'])
    uint32_t *a_cols;          // [P][C] column marginals, index c2    (cache['a'])
    uint32_t *b_rows;          // [P][C] row marginals, index c1       (cache['b'])
    float *gain;               // [P][C][C] score of adding one sample to a cell, current iteration
    float *col_term;           // [P][C]  (-aloga')/n'
    float *row_term;           // [P][C]  (-blogb')/n'
    float *sums;               // [P][4]  NlogN, aloga, blogb, n
    float *consts;             // [P][6]  fN0, fa0, n0, NlogN0, aloga0, blogb0 of the empty tables
    const float *logs;         // torch-CPU log table (caller-owned)
    int64_t n_logs;
    unsigned long long *key;   // [1] running arg-max key of the iteration
    unsigned long long *rec;   // [rec_words] winner record of this engine: key, then ids packed 4 x 16 bit
    uint8_t *pair_cols;        // [P][2] id column holding c1 / c2 of the pair
    int64_t w, w_pad, pos_base;
    int32_t d, c, p, rec_words;
};

__global__ void mip_pack_kernel(const int64_t *__restrict__ rows, int64_t w, int32_t d, int64_t w_pad,
                                uint16_t *__restrict__ ids) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= w) return;
    for (int32_t j = 0; j < d; ++j) ids[(int64_t)j * w_pad + i] = (uint16_t)rows[i * d + j];
}

// marginal terms of pair p for its CURRENT table
__device__ void mip_terms(const MiPairs &s, int32_t p) {
    const float *sm = s.sums + 4 * p;
    const float n1 = sm[3] + 1.0f;
    const float fa0 = s.consts[6 * p + 1];
    const int64_t o = (int64_t)p * s.c;
    for (int32_t i = threadIdx.x; i < s.c; i += blockDim.x) {
        s.col_term[o + i] = pairs_marginal_term(sm[1], s.a_cols[o + i], fa0, n1, s.logs);
        s.row_term[o + i] = pairs_marginal_term(sm[2], s.b_rows[o + i], fa0, n1, s.logs);
    }
}

__global__ void __launch_bounds__(256) mip_reset_kernel(MiPairs s) {
    const int32_t p = blockIdx.x;
    if (threadIdx.x == 0) {
        const float *c = s.consts + 6 * p;
        float *sm = s.sums + 4 * p;
        sm[0] = c[3]; sm[1] = c[4]; sm[2] = c[5]; sm[3] = c[2];
        if (p == 0) { s.key[0] = 0ull; s.rec[0] = 0ull; }
    }
    __syncthreads();
    mip_terms(s, p);
}

__global__ void __launch_bounds__(256) mip_gain_kernel(MiPairs s) {
    if (blockIdx.x == 0 && threadIdx.x == 0) s.key[0] = 0ull;          // the scan of this iteration starts from "nothing"
    const int64_t cc = (int64_t)s.c * s.c;
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= cc * s.p) return;
    const int32_t p = (int32_t)(idx / cc);
    const int64_t cell = idx - (int64_t)p * cc;
    const int32_t c1 = (int32_t)(cell / s.c), c2 = (int32_t)(cell - (int64_t)c1 * s.c);
    const float *sm = s.sums + 4 * p;
    const float n1 = sm[3] + 1.0f;
    s.gain[idx] = pairs_cell_score(sm[0], s.n_cells[idx], s.consts[6 * p], n1, s.col_term[(int64_t)p * s.c + c2],
                                   s.row_term[(int64_t)p * s.c + c1], s.logs);
}

// dynamic shared memory: uint16 ids[d][256] (thread-private columns: no barrier needed), then uint8 pair_cols[2P]
__global__ void __launch_bounds__(kScanThreads) mip_scan_kernel(MiPairs s) {
    extern __shared__ __align__(16) unsigned char smem[];
    uint16_t *my = reinterpret_cast<uint16_t *>(smem) + threadIdx.x;
    uint8_t *pc = smem + (size_t)s.d * kScanThreads * sizeof(uint16_t);
    __shared__ unsigned long long wbest[kScanThreads / kWarp];
    for (int32_t i = threadIdx.x; i < 2 * s.p; i += blockDim.x) pc[i] = s.pair_cols[i];
    __syncthreads();

    const int32_t P = s.p, C = s.c;
    const float *__restrict__ gain = s.gain;
    const int64_t cc = (int64_t)C * C;
    float bs = 0.f;
    uint32_t bp = 0;
    bool have = false;
    for (int64_t i = (int64_t)blockIdx.x * kScanThreads + threadIdx.x; i < s.w; i += (int64_t)gridDim.x * kScanThreads) {
        const uint16_t first = s.ids[i];
        if (first == kGone) continue;
        my[0] = first;
        for (int32_t j = 1; j < s.d; ++j) my[j * kScanThreads] = s.ids[(int64_t)j * s.w_pad + i];
        const float sc = pairs_mean(P, [&](int p) {
            const uint32_t c1 = my[pc[2 * p] * kScanThreads], c2 = my[pc[2 * p + 1] * kScanThreads];
            return __ldg(gain + (int64_t)p * cc + (int64_t)c1 * C + c2);
        });
        if (!have || sc > bs) { bs = sc; bp = (uint32_t)(s.pos_base + i); have = true; }
    }
    unsigned long long key = have ? make_key(bs, bp) : 0ull;
    key = warp_max_u64(key);
    if (threadIdx.x % kWarp == 0) wbest[threadIdx.x / kWarp] = key;
    __syncthreads();
    if (threadIdx.x < kWarp) {
        key = threadIdx.x < kScanThreads / kWarp ? wbest[threadIdx.x] : 0ull;
        key = warp_max_u64(key);
        if (threadIdx.x == 0 && key) atomicMax(s.key, key);
    }
}

// record of the local winner: out[0] = key (0: no candidate left), out[1 + j/4] holds id j in bits 16*(j%4)..
__global__ void __launch_bounds__(64) mip_emit_kernel(MiPairs s, unsigned long long *__restrict__ out) {
    const unsigned long long key = s.key[0];
    const int32_t t = threadIdx.x;
    if (t == 0) out[0] = key;
    if (t < s.rec_words - 1) {
        unsigned long long word = 0ull;
        if (key) {
            const int64_t i = (int64_t)key_pos(key) - s.pos_base;
            for (int32_t j = 4 * t; j < 4 * t + 4 && j < s.d; ++j)
                word |= (unsigned long long)s.ids[(int64_t)j * s.w_pad + i] << (16 * (j - 4 * t));
        }
        out[1 + t] = word;
    }
}

struct PairsRow { uint16_t v[kPairColsMax]; };

// bump pair p by the sample whose ids are in `row` (update_cache :383-389, update_mats :401-406); thread 0 of the CTA
__device__ void mip_bump_pair(const MiPairs &s, int32_t p, const uint16_t *row) {
    const int32_t c1 = row[s.pair_cols[2 * p]], c2 = row[s.pair_cols[2 * p + 1]];
    const int64_t o = (int64_t)p * s.c;
    uint32_t *xc = s.n_cells + (o + c1) * s.c + c2;
    const uint32_t x = *xc, y = s.a_cols[o + c2], z = s.b_rows[o + c1];
    float *sm = s.sums + 4 * p;
    const float fn0 = s.consts[6 * p], fa0 = s.consts[6 * p + 1];
    sm[0] = pairs_bump(sm[0], x, fn0, s.logs);
    sm[1] = pairs_bump(sm[1], y, fa0, s.logs);
    sm[2] = pairs_bump(sm[2], z, fa0, s.logs);
    sm[3] = sm[3] + 1.0f;
    *xc = x + 1; s.a_cols[o + c2] = y + 1; s.b_rows[o + c1] = z + 1;
}

// every CTA (one per pair) picks the record with the largest key among the n gathered ones -- highest score, earliest
// global position -- and applies it to its pair; CTA 0 removes the candidate if this engine owns it
__global__ void __launch_bounds__(256)
mip_apply_kernel(MiPairs s, const unsigned long long *__restrict__ recs, int32_t n, int64_t *__restrict__ out_pos,
                 float *__restrict__ out_gain) {
    __shared__ uint16_t row[kPairColsMax];
    __shared__ unsigned long long win_key;
    const int32_t p = blockIdx.x;
    if (threadIdx.x == 0) {
        unsigned long long key = 0ull;
        int32_t best = -1;
        for (int32_t i = 0; i < n; ++i) {
            const unsigned long long k = recs[(int64_t)i * s.rec_words];
            if (k > key) { key = k; best = i; }
        }
        win_key = key;
        if (key) {
            const unsigned long long *r = recs + (int64_t)best * s.rec_words + 1;
            for (int32_t j = 0; j < s.d; ++j) row[j] = (uint16_t)(r[j >> 2] >> (16 * (j & 3)));
            mip_bump_pair(s, p, row);
        }
        if (p == 0) {
            if (key) {
                const int64_t pos = (int64_t)key_pos(key);
                if (pos >= s.pos_base && pos < s.pos_base + s.w) s.ids[pos - s.pos_base] = kGone;
                if (out_pos) *out_pos = pos;
                if (out_gain) *out_gain = key_score(key);
            } else {
                if (out_pos) *out_pos = -1;
                if (out_gain) *out_gain = nanf("");
            }
        }
    }
    __syncthreads();
    if (win_key) mip_terms(s, p);
}

// add_samples (mi.py:408-412): count one sample into every table without selecting anything
__global__ void __launch_bounds__(256) mip_add_kernel(MiPairs s, PairsRow sample) {
    const int32_t p = blockIdx.x;
    if (threadIdx.x == 0) mip_bump_pair(s, p, sample.v);
    __syncthreads();
    mip_terms(s, p);
}

}  // namespace acav

using namespace acav;

struct acav_mi_pairs {
    MiPairs s;
    int64_t max_picks;
    int32_t sm_count;
    bool loaded, tabled;
};

namespace {

template <typename T>
int pairs_alloc(T **p, size_t n) {
    *p = nullptr;
    cudaError_t e = cudaMalloc(reinterpret_cast<void **>(p), sizeof(T) * (n ? n : 1));
    return e == cudaSuccess ? 0 : (int)e;
}

int pairs_gain(const MiPairs &s, cudaStream_t st) {
    const int64_t n = (int64_t)s.p * s.c * s.c;
    mip_gain_kernel<<<(unsigned)ceil_div(n, 256), 256, 0, st>>>(s);
    ACAV_LAUNCH_CHECK();
    return 0;
}

int pairs_scan(const acav_mi_pairs *h, cudaStream_t st) {
    const MiPairs &s = h->s;
    if (s.w == 0) return 0;
    const int64_t want = ceil_div(s.w, kScanThreads), cap = (int64_t)h->sm_count * 8;
    const size_t smem = (size_t)s.d * kScanThreads * sizeof(uint16_t) + 2 * (size_t)s.p;
    mip_scan_kernel<<<(unsigned)(want < cap ? want : cap), kScanThreads, smem, st>>>(s);
    ACAV_LAUNCH_CHECK();
    return 0;
}

int pairs_emit(const MiPairs &s, unsigned long long *out, cudaStream_t st) {
    mip_emit_kernel<<<1, 64, 0, st>>>(s, out);
    ACAV_LAUNCH_CHECK();
    return 0;
}

int pairs_apply(const MiPairs &s, const unsigned long long *recs, int32_t n, int64_t *out_pos, float *out_gain,
                cudaStream_t st) {
    mip_apply_kernel<<<(unsigned)s.p, 256, 0, st>>>(s, recs, n, out_pos, out_gain);
    ACAV_LAUNCH_CHECK();
    return 0;
}

}  // namespace

extern "C" {

int acav_mi_pairs_destroy(acav_mi_pairs_t *h) {
    if (!h) return 0;
    MiPairs &s = h->s;
    cudaFree(s.ids); cudaFree(s.n_cells); cudaFree(s.a_cols); cudaFree(s.b_rows); cudaFree(s.gain);
    cudaFree(s.col_term); cudaFree(s.row_term); cudaFree(s.sums); cudaFree(s.consts); cudaFree(s.key);
    cudaFree(s.rec); cudaFree(s.pair_cols);
    delete h;
    return 0;
}

int acav_mi_pairs_create(acav_mi_pairs_t **out, int64_t w, int32_t d, int32_t c, int32_t p, const int32_t *pairs,
                         int64_t max_picks, int64_t pos_base) {
    if (!out || w < 0 || d <= 0 || c <= 0 || p <= 0 || !pairs || max_picks < 0 || pos_base < 0) return ACAV_E_INVALID;
    if (d > kPairColsMax || p > kPairsMax || c > 65535) return ACAV_E_UNSUPPORTED;      // uint16 ids, 0xFFFF = removed
    if (pos_base + w >= 0xFFFFFFFFll) return ACAV_E_UNSUPPORTED;                        // positions live in 32 bits of the key
    if (max_picks >= (1ll << 24)) return ACAV_E_UNSUPPORTED;                            // fp32 counts must stay exact integers
    for (int32_t i = 0; i < 2 * p; ++i)
        if (pairs[i] < 0 || pairs[i] >= d) return ACAV_E_INVALID;
    *out = nullptr;
    acav_mi_pairs *h = new (std::nothrow) acav_mi_pairs();
    if (!h) return (int)cudaErrorMemoryAllocation;
    MiPairs &s = h->s;
    s = MiPairs();
    s.w = w; s.w_pad = (w + 7) / 8 * 8; s.pos_base = pos_base; s.d = d; s.c = c; s.p = p;
    s.rec_words = 1 + (d + 3) / 4;
    s.logs = nullptr; s.n_logs = 0;
    h->max_picks = max_picks; h->loaded = false; h->tabled = false; h->sm_count = 0;
    int dev = 0, rc = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e == cudaSuccess) e = cudaDeviceGetAttribute(&h->sm_count, cudaDevAttrMultiProcessorCount, dev);
    if (e != cudaSuccess) rc = (int)e;
    const size_t cells = (size_t)p * c * c, marg = (size_t)p * c;
    if (!rc) rc = pairs_alloc(&s.ids, (size_t)d * s.w_pad);
    if (!rc) rc = pairs_alloc(&s.n_cells, cells);
    if (!rc) rc = pairs_alloc(&s.a_cols, marg);
    if (!rc) rc = pairs_alloc(&s.b_rows, marg);
    if (!rc) rc = pairs_alloc(&s.gain, cells);
    if (!rc) rc = pairs_alloc(&s.col_term, marg);
    if (!rc) rc = pairs_alloc(&s.row_term, marg);
    if (!rc) rc = pairs_alloc(&s.sums, (size_t)4 * p);
    if (!rc) rc = pairs_alloc(&s.consts, (size_t)6 * p);
    if (!rc) rc = pairs_alloc(&s.key, 1);
    if (!rc) rc = pairs_alloc(&s.rec, (size_t)s.rec_words);
    if (!rc) rc = pairs_alloc(&s.pair_cols, (size_t)2 * p);
    if (!rc) {
        uint8_t cols[2 * kPairsMax];
        for (int32_t i = 0; i < 2 * p; ++i) cols[i] = (uint8_t)pairs[i];
        e = cudaMemcpy(s.pair_cols, cols, (size_t)2 * p, cudaMemcpyHostToDevice);
        if (e != cudaSuccess) rc = (int)e;
    }
    if (rc) { acav_mi_pairs_destroy(h); return rc; }
    *out = h;
    return 0;
}

int acav_mi_pairs_record_words(const acav_mi_pairs_t *h) { return h ? h->s.rec_words : 0; }

int acav_mi_pairs_load_candidates(acav_mi_pairs_t *h, const int64_t *ids, void *stream) {
    if (!h || (!ids && h->s.w > 0)) return ACAV_E_INVALID;
    const MiPairs &s = h->s;
    if (s.w > 0) {
        mip_pack_kernel<<<(unsigned)ceil_div(s.w, 256), 256, 0, (cudaStream_t)stream>>>(ids, s.w, s.d, s.w_pad, s.ids);
        ACAV_LAUNCH_CHECK();
    }
    h->loaded = true;
    return 0;
}

int acav_mi_pairs_set_tables(acav_mi_pairs_t *h, const float *logs, int64_t n_logs, const float *consts, void *stream) {
    if (!h || !logs || !consts || n_logs < h->max_picks + 3) return ACAV_E_INVALID;
    cudaStream_t st = (cudaStream_t)stream;
    MiPairs &s = h->s;
    s.logs = logs; s.n_logs = n_logs;
    ACAV_CUDA_TRY(cudaMemcpyAsync(s.consts, consts, sizeof(float) * 6 * (size_t)s.p, cudaMemcpyHostToDevice, st));
    ACAV_CUDA_TRY(cudaStreamSynchronize(st));      // `consts` is pageable host memory owned by the caller
    const size_t cells = (size_t)s.p * s.c * s.c, marg = (size_t)s.p * s.c;
    ACAV_CUDA_TRY(cudaMemsetAsync(s.n_cells, 0, sizeof(uint32_t) * cells, st));
    ACAV_CUDA_TRY(cudaMemsetAsync(s.a_cols, 0, sizeof(uint32_t) * marg, st));
    ACAV_CUDA_TRY(cudaMemsetAsync(s.b_rows, 0, sizeof(uint32_t) * marg, st));
    mip_reset_kernel<<<(unsigned)s.p, 256, 0, st>>>(s);
    ACAV_LAUNCH_CHECK();
    h->tabled = true;
    return 0;
}

int acav_mi_pairs_add_sample(acav_mi_pairs_t *h, const int64_t *ids, void *stream) {
    if (!h || !ids) return ACAV_E_INVALID;
    if (!h->tabled) return ACAV_E_STATE;
    PairsRow row;
    for (int32_t j = 0; j < kPairColsMax; ++j) row.v[j] = 0;
    for (int32_t j = 0; j < h->s.d; ++j) {
        if (ids[j] < 0 || ids[j] >= h->s.c) return ACAV_E_INVALID;
        row.v[j] = (uint16_t)ids[j];
    }
    mip_add_kernel<<<(unsigned)h->s.p, 256, 0, (cudaStream_t)stream>>>(h->s, row);
    ACAV_LAUNCH_CHECK();
    return 0;
}

int acav_mi_pairs_local_best(acav_mi_pairs_t *h, uint64_t *record, void *stream) {
    if (!h || !record) return ACAV_E_INVALID;
    if (!h->tabled || !h->loaded) return ACAV_E_STATE;
    cudaStream_t st = (cudaStream_t)stream;
    int rc = pairs_gain(h->s, st);
    if (!rc) rc = pairs_scan(h, st);
    if (!rc) rc = pairs_emit(h->s, reinterpret_cast<unsigned long long *>(record), st);
    return rc;
}

int acav_mi_pairs_apply(acav_mi_pairs_t *h, const uint64_t *records, int32_t n, int64_t *out_pos, float *out_gain,
                        void *stream) {
    if (!h || !records || n <= 0) return ACAV_E_INVALID;
    if (!h->tabled || !h->loaded) return ACAV_E_STATE;
    return pairs_apply(h->s, reinterpret_cast<const unsigned long long *>(records), n, out_pos, out_gain,
                       (cudaStream_t)stream);
}

int acav_mi_pairs_run(acav_mi_pairs_t *h, int64_t n_picks, int64_t *out_pos, float *out_gain, void *stream) {
    if (!h || n_picks < 0 || (n_picks > 0 && (!out_pos || !out_gain))) return ACAV_E_INVALID;
    if (!h->tabled || !h->loaded) return ACAV_E_STATE;
    cudaStream_t st = (cudaStream_t)stream;
    for (int64_t it = 0; it < n_picks; ++it) {
        int rc = pairs_gain(h->s, st);
        if (!rc) rc = pairs_scan(h, st);
        if (!rc) rc = pairs_emit(h->s, h->s.rec, st);
        if (!rc) rc = pairs_apply(h->s, h->s.rec, 1, out_pos + it, out_gain + it, st);
        if (rc) return rc;
    }
    return 0;
}

int acav_mi_pairs_read_state(acav_mi_pairs_t *h, uint32_t *n_cells, uint32_t *a_cols, uint32_t *b_rows, float *sums,
                             void *stream) {
    if (!h) return ACAV_E_INVALID;
    cudaStream_t st = (cudaStream_t)stream;
    const MiPairs &s = h->s;
    const size_t cells = (size_t)s.p * s.c * s.c, marg = (size_t)s.p * s.c;
    if (n_cells) ACAV_CUDA_TRY(cudaMemcpyAsync(n_cells, s.n_cells, sizeof(uint32_t) * cells, cudaMemcpyDeviceToDevice, st));
    if (a_cols) ACAV_CUDA_TRY(cudaMemcpyAsync(a_cols, s.a_cols, sizeof(uint32_t) * marg, cudaMemcpyDeviceToDevice, st));
    if (b_rows) ACAV_CUDA_TRY(cudaMemcpyAsync(b_rows, s.b_rows, sizeof(uint32_t) * marg, cudaMemcpyDeviceToDevice, st));
    if (sums) ACAV_CUDA_TRY(cudaMemcpyAsync(sums, s.sums, sizeof(float) * 4 * (size_t)s.p, cudaMemcpyDeviceToDevice, st));
    return 0;
}

}  // extern "C"
