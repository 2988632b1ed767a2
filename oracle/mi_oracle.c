/*
 * Plain-C restatement of the reference's exact greedy MI selection (``mem_mi``, one clustering pair).
 *
 * TEST INFRASTRUCTURE ONLY (see oracle/__init__.py) -- never linked into the product library.
 * Build: gcc -O2 -ffp-contract=off -fno-fast-math -fopenmp   (no FMA contraction: every fp32
 * operation below rounds exactly where the reference's torch fp32 tensor ops round).
 *
 * Reference (paths under /root/reference/subset_selection/code):
 *   measures/mi.py:322-333  get_last      x = N[c1,c2], y = a[c2], z = b[c1]
 *   measures/mi.py:339-340  update_nlogn  prev - x*log(x) + (x+1)*log(x+1)   (left to right, fp32)
 *   measures/mi.py:368-381  calc_MI       ((NlogN/n + (-aloga)/n) + (-blogb)/n) + log(n), n = n_old+1
 *   measures/mi.py:76-80    calc_score    mean over P (=1: identity), max over candidates, first index
 *   measures/mi.py:383-406  update_cache / update_mats   adopt winner's scalars, N,a,b,n += 1
 *   measures/mi.py:104-125  remove_idx_all   delete winner, order preserved
 *
 * log() is NOT evaluated here: the reference uses torch's CPU fp32 log, whose bits differ from libm's.
 * The caller passes `logs[k]` = torch.log(float32(k)) for every integer the tables can hold, and
 * `consts` = the x*log(x) values / initial sums for a count of zero (table value 2^-52, marginal
 * C*2^-52), computed with the same torch operators (oracle/mi_oracle.py: log_table,
 * zero_count_constants).  Counts stay below 2^24 so fp32 table entries are exact integers.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#ifdef _OPENMP
#include <omp.h>
#endif

typedef struct {
    float NlogN, aloga, blogb, n;   /* running sums (fp32 like the reference's cache) */
    float fN0, fa0;                 /* x*log(x) for a zero count in the table / in a marginal */
} mi_scalars;

static inline float xlogx(int64_t k, float f0, const float *logs) {
    return k == 0 ? f0 : (float)k * logs[k];
}

/* score of adding one sample to cell (x = N[c1,c2], y = a[c2], z = b[c1]); also returns the three
 * updated sums so the winner's can be adopted verbatim (update_cache). */
static inline float cell_score(const mi_scalars *s, int64_t x, int64_t y, int64_t z, const float *logs,
                               float *oN, float *oa, float *ob) {
    float t1 = (s->NlogN - xlogx(x, s->fN0, logs)) + xlogx(x + 1, 0.f, logs);
    float t2 = (s->aloga - xlogx(y, s->fa0, logs)) + xlogx(y + 1, 0.f, logs);
    float t3 = (s->blogb - xlogx(z, s->fa0, logs)) + xlogx(z + 1, 0.f, logs);
    float np = s->n + 1.0f;
    float tN = t1 / np;
    float ta = (-t2) / np;
    float tb = (-t3) / np;
    *oN = t1; *oa = t2; *ob = t3;
    return ((tN + ta) + tb) + logs[(int64_t)np];
}

static void load_consts(mi_scalars *s, const float *consts) {
    s->fN0 = consts[0]; s->fa0 = consts[1]; s->n = consts[2];
    s->NlogN = consts[3]; s->aloga = consts[4]; s->blogb = consts[5];
}

/* Literal form: every iteration scores every remaining candidate in list order. */
int64_t mi_oracle_greedy_scan(const int32_t *c1, const int32_t *c2, int64_t W, int32_t C,
                              const float *logs, int64_t nlogs, const float *consts,
                              int64_t n_picks, int64_t *out_pos, float *out_gain) {
    (void)nlogs;
    mi_scalars s; load_consts(&s, consts);
    int64_t *N = calloc((size_t)C * C, sizeof(int64_t));
    int64_t *a = calloc((size_t)C, sizeof(int64_t));
    int64_t *b = calloc((size_t)C, sizeof(int64_t));
    uint8_t *gone = calloc((size_t)W, 1);
    int64_t it;
    for (it = 0; it < n_picks; ++it) {
        int64_t best = -1; float bs = 0.f, bN = 0.f, ba = 0.f, bb = 0.f;
        for (int64_t w = 0; w < W; ++w) {
            if (gone[w]) continue;
            float tN, ta, tb;
            float sc = cell_score(&s, N[(int64_t)c1[w] * C + c2[w]], a[c2[w]], b[c1[w]], logs, &tN, &ta, &tb);
            if (best < 0 || sc > bs) { best = w; bs = sc; bN = tN; ba = ta; bb = tb; }
        }
        if (best < 0) break;
        out_pos[it] = best; out_gain[it] = bs;
        s.NlogN = bN; s.aloga = ba; s.blogb = bb; s.n = s.n + 1.0f;
        N[(int64_t)c1[best] * C + c2[best]] += 1; a[c2[best]] += 1; b[c1[best]] += 1;
        gone[best] = 1;
    }
    free(N); free(a); free(b); free(gone);
    return it;
}

/* Same picks, bucketed by cell: the score depends only on the candidate's cell, so the winner is the
 * earliest remaining candidate among the cells holding the maximal score. */
int64_t mi_oracle_greedy_bucketed(const int32_t *c1, const int32_t *c2, int64_t W, int32_t C,
                                  const float *logs, int64_t nlogs, const float *consts,
                                  int64_t n_picks, int64_t *out_pos, float *out_gain) {
    (void)nlogs;
    mi_scalars s; load_consts(&s, consts);
    int64_t cells = (int64_t)C * C;
    int64_t *N = calloc((size_t)cells, sizeof(int64_t));
    int64_t *a = calloc((size_t)C, sizeof(int64_t));
    int64_t *b = calloc((size_t)C, sizeof(int64_t));
    int64_t *start = calloc((size_t)cells + 1, sizeof(int64_t));
    int64_t *head = malloc((size_t)cells * sizeof(int64_t));
    int64_t *order = malloc((size_t)(W > 0 ? W : 1) * sizeof(int64_t));
    for (int64_t w = 0; w < W; ++w) start[(int64_t)c1[w] * C + c2[w] + 1]++;
    for (int64_t c = 0; c < cells; ++c) start[c + 1] += start[c];
    memcpy(head, start, (size_t)cells * sizeof(int64_t));
    for (int64_t w = 0; w < W; ++w) order[head[(int64_t)c1[w] * C + c2[w]]++] = w;   /* stable */
    memcpy(head, start, (size_t)cells * sizeof(int64_t));
    int32_t *live = malloc((size_t)cells * sizeof(int32_t));
    int64_t nlive = 0;
    for (int64_t c = 0; c < cells; ++c) if (start[c + 1] > start[c]) live[nlive++] = (int32_t)c;
    int64_t it;
    for (it = 0; it < n_picks; ++it) {
        int64_t best = -1, bcell = -1, bslot = -1; float bs = 0.f, bN = 0.f, ba = 0.f, bb = 0.f;
        for (int64_t i = 0; i < nlive; ++i) {
            int64_t c = live[i];
            int64_t r = c / C, q = c % C;
            float tN, ta, tb;
            float sc = cell_score(&s, N[c], a[q], b[r], logs, &tN, &ta, &tb);
            int64_t pos = order[head[c]];
            if (best < 0 || sc > bs || (sc == bs && pos < best)) {
                best = pos; bcell = c; bslot = i; bs = sc; bN = tN; ba = ta; bb = tb;
            }
        }
        if (best < 0) break;
        out_pos[it] = best; out_gain[it] = bs;
        s.NlogN = bN; s.aloga = ba; s.blogb = bb; s.n = s.n + 1.0f;
        N[bcell] += 1; a[bcell % C] += 1; b[bcell / C] += 1;
        if (++head[bcell] == start[bcell + 1]) live[bslot] = live[--nlive];
    }
    free(N); free(a); free(b); free(start); free(head); free(order); free(live);
    return it;
}

/* Timing helper for bench.py's cpu_baseline: `repeats` full scans of W candidates (no removal, the
 * table advances by the winner each time), parallel over `threads` contiguous ranges, partial
 * results combined in range order so the pick equals the serial scan's.  Returns seconds. */
double mi_oracle_scan_once(const int32_t *c1, const int32_t *c2, int64_t W, int32_t C,
                           const float *logs, int64_t nlogs, const float *consts,
                           int32_t repeats, int32_t threads) {
    (void)nlogs;
    mi_scalars s; load_consts(&s, consts);
    int64_t *N = calloc((size_t)C * C, sizeof(int64_t));
    int64_t *a = calloc((size_t)C, sizeof(int64_t));
    int64_t *b = calloc((size_t)C, sizeof(int64_t));
    if (threads < 1) threads = 1;
    int64_t *tbest = malloc(sizeof(int64_t) * (size_t)threads);
    float *tvals = malloc(sizeof(float) * 4 * (size_t)threads);
    struct timespec t0, t1;
    clock_gettime(CLOCK_MONOTONIC, &t0);
    for (int32_t rep = 0; rep < repeats; ++rep) {
#ifdef _OPENMP
#pragma omp parallel num_threads(threads)
#endif
        {
#ifdef _OPENMP
            int t = omp_get_thread_num(), nt = omp_get_num_threads();
#else
            int t = 0, nt = 1;
#endif
            for (int tt = t; tt < threads; tt += nt) {
                int64_t lo = W * tt / threads, hi = W * (tt + 1) / threads;
                int64_t best = -1; float bs = 0.f, bN = 0.f, ba = 0.f, bb = 0.f;
                for (int64_t w = lo; w < hi; ++w) {
                    float tN, ta, tb;
                    float sc = cell_score(&s, N[(int64_t)c1[w] * C + c2[w]], a[c2[w]], b[c1[w]], logs, &tN, &ta, &tb);
                    if (best < 0 || sc > bs) { best = w; bs = sc; bN = tN; ba = ta; bb = tb; }
                }
                tbest[tt] = best; tvals[4 * tt] = bs; tvals[4 * tt + 1] = bN; tvals[4 * tt + 2] = ba; tvals[4 * tt + 3] = bb;
            }
        }
        int64_t best = -1; int bt = 0;
        for (int tt = 0; tt < threads; ++tt)
            if (tbest[tt] >= 0 && (best < 0 || tvals[4 * tt] > tvals[4 * bt])) { best = tbest[tt]; bt = tt; }
        if (best >= 0) {
            s.NlogN = tvals[4 * bt + 1]; s.aloga = tvals[4 * bt + 2]; s.blogb = tvals[4 * bt + 3]; s.n += 1.0f;
            N[(int64_t)c1[best] * C + c2[best]] += 1; a[c2[best]] += 1; b[c1[best]] += 1;
        }
    }
    clock_gettime(CLOCK_MONOTONIC, &t1);
    free(N); free(a); free(b); free(tbest); free(tvals);
    return (double)(t1.tv_sec - t0.tv_sec) + 1e-9 * (double)(t1.tv_nsec - t0.tv_nsec);
}

/* ------------------------------------------------------------------------------------------------
 * P > 1 clustering pairs.
 *
 *   measures/mi.py:285-295,310-320  calc_N / gather_pairs   candidate w, pair p -> (c1, c2) = (ids[w][pairs[p][0]],
 *                                                           ids[w][pairs[p][1]])
 *   measures/mi.py:76-80            calc_score              scores.mean(dim=-1) over the P pairs, then first max
 *
 * `mean(dim=-1)` of a contiguous fp32 [W, P] tensor is torch's CPU sum over the inner dimension followed by one
 * division by P.  The sum is NOT a left-to-right loop: it is ATen's cascade_sum (ATen/native/cpu/SumKernel.cpp, an
 * un-vendored dependency of the reference: torch, requirements.txt:2), restated below in the structure of that
 * file -- multi_row_sum / row_sum / the inner-reduction drivers -- for its 8-lane float vectors.  The restatement
 * is pinned against torch itself on random rows for P = 1..600 (tests/test_oracle_golden.py) and, through the
 * greedy loop, against goldens produced by the unmodified reference for P = 3, 10 and 45.
 * ---------------------------------------------------------------------------------------------- */

#define ATEN_LANES 8          /* Vectorized<float>::size() of the sum kernel on this x86 build */
#define ATEN_ILP 4
#define ATEN_LEVELS 4

static int ceil_log2_i64(int64_t x) {
    int r = 0;
    if (x <= 2) return 1;                       /* utils::CeilLog2 */
    --x;
    while (x > 0) { x >>= 1; ++r; }
    return r;
}

/* multi_row_sum: `nrows` interleaved running sums (each `width` lanes wide) over `size` steps with cascade levels.
 * in(i, k, l) = data[(i * row_stride + k * col_stride) + l]; out[k][l]. */
static void multi_row_sum(const float *data, int64_t row_stride, int64_t col_stride, int64_t size, int width,
                          float out[ATEN_ILP][ATEN_LANES]) {
    int level_power = ceil_log2_i64(size) / ATEN_LEVELS;
    if (level_power < 4) level_power = 4;
    const int64_t level_step = (int64_t)1 << level_power, level_mask = level_step - 1;
    float acc[ATEN_LEVELS][ATEN_ILP][ATEN_LANES];
    memset(acc, 0, sizeof(acc));
    int64_t i = 0;
    for (; i + level_step <= size;) {
        for (int64_t j = 0; j < level_step; ++j, ++i)
            for (int k = 0; k < ATEN_ILP; ++k)
                for (int l = 0; l < width; ++l) acc[0][k][l] += data[i * row_stride + k * col_stride + l];
        for (int j = 1; j < ATEN_LEVELS; ++j) {
            for (int k = 0; k < ATEN_ILP; ++k)
                for (int l = 0; l < width; ++l) { acc[j][k][l] += acc[j - 1][k][l]; acc[j - 1][k][l] = 0.f; }
            const int64_t mask = level_mask << (j * level_power);
            if ((i & mask) != 0) break;
        }
    }
    for (; i < size; ++i)
        for (int k = 0; k < ATEN_ILP; ++k)
            for (int l = 0; l < width; ++l) acc[0][k][l] += data[i * row_stride + k * col_stride + l];
    for (int j = 1; j < ATEN_LEVELS; ++j)
        for (int k = 0; k < ATEN_ILP; ++k)
            for (int l = 0; l < width; ++l) acc[0][k][l] += acc[j][k][l];
    memcpy(out, acc[0], sizeof(acc[0]));
}

/* row_sum over `size` elements of `width` lanes, consecutive elements `stride` floats apart */
static void row_sum(const float *data, int64_t stride, int64_t size, int width, float out[ATEN_LANES]) {
    float part[ATEN_ILP][ATEN_LANES];
    const int64_t size_ilp = size / ATEN_ILP;
    multi_row_sum(data, stride * ATEN_ILP, stride, size_ilp, width, part);
    for (int64_t i = size_ilp * ATEN_ILP; i < size; ++i)
        for (int l = 0; l < width; ++l) part[0][l] += data[i * stride + l];
    for (int k = 1; k < ATEN_ILP; ++k)
        for (int l = 0; l < width; ++l) part[0][l] += part[k][l];
    memcpy(out, part[0], sizeof(float) * ATEN_LANES);
}

/* sum of one contiguous row of P floats: vectorized_inner_sum (P >= lanes) or scalar_inner_sum */
float mi_oracle_aten_row_sum(const float *x, int64_t P) {
    float v[ATEN_LANES];
    if (P < ATEN_LANES) {
        row_sum(x, 1, P, 1, v);
        return v[0];
    }
    const int64_t vec_size = P / ATEN_LANES;
    row_sum(x, ATEN_LANES, vec_size, ATEN_LANES, v);
    float acc = 0.f;
    for (int64_t k = vec_size * ATEN_LANES; k < P; ++k) acc += x[k];
    for (int l = 0; l < ATEN_LANES; ++l) acc += v[l];
    return acc;
}

float mi_oracle_aten_row_mean(const float *x, int64_t P) { return mi_oracle_aten_row_sum(x, P) / (float)P; }

/* Literal greedy scan over P pairs.  ids: int32 [W, D] row-major; pairs: int32 [P, 2] columns of ids;
 * consts: [P, 6] as load_consts, per pair.  out_sums (may be NULL): final [P, 4] NlogN, aloga, blogb, n. */
int64_t mi_oracle_greedy_pairs(const int32_t *ids, int64_t W, int32_t D, int32_t C, const int32_t *pairs, int32_t P,
                               const float *logs, int64_t nlogs, const float *consts, int64_t n_picks,
                               int64_t *out_pos, float *out_gain, float *out_sums) {
    (void)nlogs;
    mi_scalars *s = malloc(sizeof(mi_scalars) * (size_t)P);
    for (int p = 0; p < P; ++p) load_consts(&s[p], consts + 6 * p);
    const int64_t cc = (int64_t)C * C;
    int64_t *N = calloc((size_t)P * cc, sizeof(int64_t));
    int64_t *a = calloc((size_t)P * C, sizeof(int64_t));
    int64_t *b = calloc((size_t)P * C, sizeof(int64_t));
    uint8_t *gone = calloc((size_t)(W > 0 ? W : 1), 1);
    float *row = malloc(sizeof(float) * (size_t)P);
    float *tmp = malloc(sizeof(float) * 3 * (size_t)P), *win = malloc(sizeof(float) * 3 * (size_t)P);
    int64_t it;
    for (it = 0; it < n_picks; ++it) {
        int64_t best = -1; float bs = 0.f;
        for (int64_t w = 0; w < W; ++w) {
            if (gone[w]) continue;
            for (int p = 0; p < P; ++p) {
                const int64_t c1 = ids[w * D + pairs[2 * p]], c2 = ids[w * D + pairs[2 * p + 1]];
                row[p] = cell_score(&s[p], N[p * cc + c1 * C + c2], a[(int64_t)p * C + c2], b[(int64_t)p * C + c1], logs,
                                    &tmp[3 * p], &tmp[3 * p + 1], &tmp[3 * p + 2]);
            }
            const float sc = mi_oracle_aten_row_mean(row, P);
            if (best < 0 || sc > bs) { best = w; bs = sc; memcpy(win, tmp, sizeof(float) * 3 * (size_t)P); }
        }
        if (best < 0) break;
        out_pos[it] = best; out_gain[it] = bs;
        for (int p = 0; p < P; ++p) {
            const int64_t c1 = ids[best * D + pairs[2 * p]], c2 = ids[best * D + pairs[2 * p + 1]];
            s[p].NlogN = win[3 * p]; s[p].aloga = win[3 * p + 1]; s[p].blogb = win[3 * p + 2]; s[p].n = s[p].n + 1.0f;
            N[p * cc + c1 * C + c2] += 1; a[(int64_t)p * C + c2] += 1; b[(int64_t)p * C + c1] += 1;
        }
        gone[best] = 1;
    }
    if (out_sums)
        for (int p = 0; p < P; ++p) {
            out_sums[4 * p] = s[p].NlogN; out_sums[4 * p + 1] = s[p].aloga; out_sums[4 * p + 2] = s[p].blogb;
            out_sums[4 * p + 3] = s[p].n;
        }
    free(s); free(N); free(a); free(b); free(gone); free(row); free(tmp); free(win);
    return it;
}
