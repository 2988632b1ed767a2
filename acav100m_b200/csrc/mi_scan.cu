// Exact greedy mutual-information selection, one clustering pair (P = 1), list-order layout.
//
// Replaces EfficientMemMI of the reference (subset_selection/code/measures/mi.py:284-412) as driven by
// EfficientMI.run_greedy (:150-192).  Per iteration the reference scores EVERY remaining candidate:
//     x = N[c1,c2], y = a[c2], z = b[c1]                                      get_last  :322-333
//     NlogN' = NlogN - x log x + (x+1) log(x+1)   (same for aloga', blogb')   update_nlogn :339-340
//     score  = ((NlogN'/n' + (-aloga')/n') + (-blogb')/n') + log n',  n' = n+1   calc_MI :368-381
// takes the first maximum (:79), adopts the winner's sums and bumps N, a, b, n (:383-406), deletes it
// from the list preserving order (:104-125).
//
// The score depends on the candidate only through its table cell, so an iteration is
//   (1) gain table: one score per cell, K_a*K_v values (same fp32 op sequence, no FMA contraction);
//   (2) scan: stream the packed (c1,c2) ids with 128-bit loads, gather gain[cell], argmax with
//       "earliest position wins" folded into a 64-bit key;
//   (3) apply: one thread updates table + sums and tombstones the winner (order is preserved because
//       nothing moves).
// log() never runs on the device: `logs` holds torch's CPU fp32 log of every integer the table can
// reach (DESIGN.md "MI exactness"), so scores are bit-identical to the reference's.
#include "common.cuh"
#include "kernels.cuh"

namespace acav {

constexpr uint32_t kTomb = 0xFFFFFFFFu;

__device__ __forceinline__ float xlogx_count(uint32_t k, float f0, const float *__restrict__ logs) {
    return k == 0 ? f0 : __fmul_rn((float)k, logs[k]);
}
// prev - f(k) + f(k+1), evaluated left to right in fp32 (mi.py:339-340)
__device__ __forceinline__ float bump(float prev, uint32_t k, float f0, const float *__restrict__ logs) {
    return __fadd_rn(__fsub_rn(prev, xlogx_count(k, f0, logs)), xlogx_count(k + 1, 0.f, logs));
}

__global__ void mi_pack_kernel(const int64_t *__restrict__ cells, int64_t w, uint32_t *__restrict__ packed) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= w) return;
    uint32_t c1 = (uint32_t)cells[2 * i], c2 = (uint32_t)cells[2 * i + 1];
    packed[i] = (c1 << 16) | (c2 & 0xFFFFu);
}

// per-row / per-column terms of the score for the CURRENT table: (-blogb')/n' and (-aloga')/n'
__device__ void mi_terms(const MiState &s) {
    const float np = __fadd_rn(s.sums[3], 1.0f);
    const float aloga = s.sums[1], blogb = s.sums[2], fa0 = s.sums[5];
    for (int32_t i = threadIdx.x; i < s.k_v; i += blockDim.x)
        s.col_term[i] = __fdiv_rn(-bump(aloga, s.a_cols[i], fa0, s.logs), np);
    for (int32_t i = threadIdx.x; i < s.k_a; i += blockDim.x)
        s.row_term[i] = __fdiv_rn(-bump(blogb, s.b_rows[i], fa0, s.logs), np);
}

__global__ void __launch_bounds__(1024) mi_reset_kernel(MiState s, const float *__restrict__ consts) {
    if (threadIdx.x == 0) {
        s.sums[0] = consts[3]; s.sums[1] = consts[4]; s.sums[2] = consts[5]; s.sums[3] = consts[2];
        s.sums[4] = consts[0]; s.sums[5] = consts[1];
        s.key[0] = 0ull; s.key[1] = 0ull;
    }
    __syncthreads();
    mi_terms(s);
}

__global__ void __launch_bounds__(1024) mi_add_sample_kernel(MiState s, int32_t c1, int32_t c2) {
    if (threadIdx.x == 0) {
        const uint32_t x = s.n_cells[(int64_t)c1 * s.k_v + c2], y = s.a_cols[c2], z = s.b_rows[c1];
        s.sums[0] = bump(s.sums[0], x, s.sums[4], s.logs);
        s.sums[1] = bump(s.sums[1], y, s.sums[5], s.logs);
        s.sums[2] = bump(s.sums[2], z, s.sums[5], s.logs);
        s.sums[3] = __fadd_rn(s.sums[3], 1.0f);
        s.n_cells[(int64_t)c1 * s.k_v + c2] = x + 1; s.a_cols[c2] = y + 1; s.b_rows[c1] = z + 1;
    }
    __syncthreads();
    mi_terms(s);
}

// recompute the row / column terms from the current marginals and sums (after the persistent kernel
// wrote its replicated state back)
__global__ void __launch_bounds__(1024) mi_refresh_kernel(MiState s, uint32_t *pub, uint32_t pub_words, unsigned int *bar) {
    if (threadIdx.x == 0) { s.key[0] = 0ull; s.key[1] = 0ull; }
    // the grid-barrier words and the per-CTA records of the persistent loops: zero for the next run (saves that run
    // two memsets in front of its launch)
    for (uint32_t i = threadIdx.x; i < pub_words; i += blockDim.x) pub[i] = 0u;
    if (bar && threadIdx.x < 2) bar[threadIdx.x] = 0u;
    mi_terms(s);
}

__global__ void mi_gain_kernel(MiState s) {
    const int64_t cell = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (cell >= (int64_t)s.k_a * s.k_v) return;
    const int32_t c1 = (int32_t)(cell / s.k_v), c2 = (int32_t)(cell % s.k_v);
    const float np = __fadd_rn(s.sums[3], 1.0f);
    const float t1 = bump(s.sums[0], s.n_cells[cell], s.sums[4], s.logs);
    const float tN = __fdiv_rn(t1, np);
    s.gain[cell] = __fadd_rn(__fadd_rn(__fadd_rn(tN, s.col_term[c2]), s.row_term[c1]),
                             s.logs[(int64_t)np]);
}

__device__ __forceinline__ void scan_one(uint32_t cell, uint32_t pos, const float *__restrict__ gain,
                                         int32_t k_v, float &bs, uint32_t &bp, bool &have) {
    if (cell == kTomb) return;
    const float g = gain[(cell >> 16) * (uint32_t)k_v + (cell & 0xFFFFu)];
    if (!have || g > bs) { bs = g; bp = pos; have = true; }
}

__global__ void __launch_bounds__(256) mi_scan_kernel(MiState s) {
    __shared__ unsigned long long wbest[8];
    const int64_t nvec = s.w / 4;
    const uint4 *__restrict__ v = reinterpret_cast<const uint4 *>(s.cells);
    float bs = 0.f;
    uint32_t bp = 0;
    bool have = false;
    const uint32_t base = (uint32_t)s.pos_base;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nvec;
         i += (int64_t)gridDim.x * blockDim.x) {
        const uint4 c = __ldg(v + i);
        const uint32_t p = base + (uint32_t)(4 * i);
        scan_one(c.x, p, s.gain, s.k_v, bs, bp, have);
        scan_one(c.y, p + 1, s.gain, s.k_v, bs, bp, have);
        scan_one(c.z, p + 2, s.gain, s.k_v, bs, bp, have);
        scan_one(c.w, p + 3, s.gain, s.k_v, bs, bp, have);
    }
    if (blockIdx.x == 0 && threadIdx.x < (s.w & 3)) {
        const int64_t i = nvec * 4 + threadIdx.x;
        scan_one(s.cells[i], base + (uint32_t)i, s.gain, s.k_v, bs, bp, have);
    }
    unsigned long long key = have ? make_key(bs, bp) : 0ull;
    key = warp_max_u64(key);
    if (threadIdx.x % kWarp == 0) wbest[threadIdx.x / kWarp] = key;
    __syncthreads();
    if (threadIdx.x < kWarp) {
        key = threadIdx.x < 8 ? wbest[threadIdx.x] : 0ull;
        key = warp_max_u64(key);
        if (threadIdx.x == 0 && key) atomicMax(s.key, key);
    }
}

// key[1] = packed cell of the local winner (for the cross-rank exchange)
__global__ void mi_emit_kernel(MiState s, unsigned long long *__restrict__ out) {
    const unsigned long long key = s.key[0];
    out[0] = key;
    out[1] = key ? (unsigned long long)s.cells[(int64_t)key_pos(key) - s.pos_base] : 0ull;
}

// n == 0: use the engine's own running key (single GPU); otherwise the max of n gathered pairs.
__global__ void __launch_bounds__(1024)
mi_apply_kernel(MiState s, const unsigned long long *__restrict__ key_cells, int32_t n,
                int64_t *__restrict__ out_pos, float *__restrict__ out_gain) {
    if (threadIdx.x == 0) {
        unsigned long long key = 0ull, cellw = 0ull;
        if (n == 0) {
            key = s.key[0];
            if (key) cellw = s.cells[(int64_t)key_pos(key) - s.pos_base];
        } else {
            for (int32_t i = 0; i < n; ++i)
                if (key_cells[2 * i] > key) { key = key_cells[2 * i]; cellw = key_cells[2 * i + 1]; }
        }
        s.key[0] = 0ull;
        if (key) {
            const int64_t pos = (int64_t)key_pos(key);
            const int32_t c1 = (int32_t)(cellw >> 16), c2 = (int32_t)(cellw & 0xFFFFu);
            const uint32_t x = s.n_cells[(int64_t)c1 * s.k_v + c2], y = s.a_cols[c2], z = s.b_rows[c1];
            s.sums[0] = bump(s.sums[0], x, s.sums[4], s.logs);         // update_cache :383-389
            s.sums[1] = bump(s.sums[1], y, s.sums[5], s.logs);
            s.sums[2] = bump(s.sums[2], z, s.sums[5], s.logs);
            s.sums[3] = __fadd_rn(s.sums[3], 1.0f);                    // update_mats :401-406
            s.n_cells[(int64_t)c1 * s.k_v + c2] = x + 1; s.a_cols[c2] = y + 1; s.b_rows[c1] = z + 1;
            if (pos >= s.pos_base && pos < s.pos_base + s.w) s.cells[pos - s.pos_base] = kTomb;
            if (out_pos) *out_pos = pos;
            if (out_gain) *out_gain = key_score(key);
        } else {
            if (out_pos) *out_pos = -1;
            if (out_gain) *out_gain = nanf("");
        }
    }
    __syncthreads();
    mi_terms(s);
}

int launch_mi_pack(const int64_t *cells, int64_t w, uint32_t *packed, cudaStream_t st) {
    if (w == 0) return 0;
    mi_pack_kernel<<<(unsigned)ceil_div(w, 256), 256, 0, st>>>(cells, w, packed);
    ACAV_LAUNCH_CHECK();
    return 0;
}

int launch_mi_reset(const MiState &s, const float *consts_dev, cudaStream_t st) {
    ACAV_CUDA_TRY(cudaMemsetAsync(s.n_cells, 0, sizeof(uint32_t) * (size_t)s.k_a * s.k_v, st));
    ACAV_CUDA_TRY(cudaMemsetAsync(s.a_cols, 0, sizeof(uint32_t) * (size_t)s.k_v, st));
    ACAV_CUDA_TRY(cudaMemsetAsync(s.b_rows, 0, sizeof(uint32_t) * (size_t)s.k_a, st));
    mi_reset_kernel<<<1, 1024, 0, st>>>(s, consts_dev);
    ACAV_LAUNCH_CHECK();
    return 0;
}

int launch_mi_add_sample(const MiState &s, int32_t c1, int32_t c2, cudaStream_t st) {
    mi_add_sample_kernel<<<1, 1024, 0, st>>>(s, c1, c2);
    ACAV_LAUNCH_CHECK();
    return 0;
}

int launch_mi_refresh_terms(const MiState &s, cudaStream_t st, void *pub, size_t pub_bytes, unsigned int *bar) {
    mi_refresh_kernel<<<1, 1024, 0, st>>>(s, reinterpret_cast<uint32_t *>(pub), pub ? (uint32_t)(pub_bytes / 4) : 0u, bar);
    ACAV_LAUNCH_CHECK();
    return 0;
}

int launch_mi_gain_table(const MiState &s, cudaStream_t st) {
    const int64_t cells = (int64_t)s.k_a * s.k_v;
    mi_gain_kernel<<<(unsigned)ceil_div(cells, 256), 256, 0, st>>>(s);
    ACAV_LAUNCH_CHECK();
    return 0;
}

int launch_mi_scan(const MiState &s, int sm_count, cudaStream_t st) {
    if (s.w == 0) return 0;
    int64_t want = ceil_div(ceil_div(s.w, 4), 256);
    int64_t cap = (int64_t)sm_count * 8;
    mi_scan_kernel<<<(unsigned)(want < cap ? (want > 0 ? want : 1) : cap), 256, 0, st>>>(s);
    ACAV_LAUNCH_CHECK();
    return 0;
}

int launch_mi_emit(const MiState &s, unsigned long long *out, cudaStream_t st) {
    mi_emit_kernel<<<1, 1, 0, st>>>(s, out);
    ACAV_LAUNCH_CHECK();
    return 0;
}

int launch_mi_apply(const MiState &s, const unsigned long long *key_cells, int32_t n,
                    int64_t *out_pos, float *out_gain, cudaStream_t st) {
    mi_apply_kernel<<<1, 1024, 0, st>>>(s, key_cells, n, out_pos, out_gain);
    ACAV_LAUNCH_CHECK();
    return 0;
}

}  // namespace acav
