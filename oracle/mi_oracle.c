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
