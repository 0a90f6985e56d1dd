/* C restatement of the pinned-order fp32 ranking (oracle/search.py) for sizes the
 * NumPy form is too slow for.  TEST INFRASTRUCTURE ONLY: built by oracle/Makefile
 * into oracle/_build/libasr_oracle.so and loaded only by tests/, smoke() and
 * bench.py's cpu_baseline leg.
 *
 * Follows asr/audio_sheet_server.py:530-551 (one query vs whole DB, sort, take
 * n_candidates) and asr/utils/train_dcca_pool.py:46-74 (rank of the correct
 * item), with the deterministic definition stated in oracle/search.py:
 *   every * and + individually rounded to fp32 (compile with -ffp-contract=off),
 *   sequential k = 0..d-1, NaN -> -inf, order = (score desc, index asc).
 */
#include <math.h>
#include <pthread.h>
#include <stdlib.h>
#include <string.h>

/* minimal fork-join helper (no OpenMP runtime in this image) */
static int g_threads = 1;
void asr_oracle_set_threads(int n) { g_threads = n < 1 ? 1 : (n > 256 ? 256 : n); }
typedef void (*range_fn)(long lo, long hi, void *ctx);
typedef struct { range_fn fn; long lo, hi; void *ctx; } job_t;
static void *job_main(void *p) { job_t *j = (job_t *)p; j->fn(j->lo, j->hi, j->ctx); return 0; }
static void parallel_for(long n, range_fn fn, void *ctx) {
    int T = g_threads; if (T > n) T = n > 0 ? (int)n : 1;
    if (T <= 1) { fn(0, n, ctx); return; }
    pthread_t th[256]; job_t jobs[256];
    for (int t = 0; t < T; ++t) {
        jobs[t].fn = fn; jobs[t].ctx = ctx;
        jobs[t].lo = n * t / T; jobs[t].hi = n * (t + 1) / T;
        pthread_create(&th[t], 0, job_main, &jobs[t]);
    }
    for (int t = 0; t < T; ++t) pthread_join(th[t], 0);
}

static void normalise_row(const float *x, int d, float *out) {
    float ss = 0.0f;
    for (int k = 0; k < d; ++k) {
        float p = x[k] * x[k];
        ss = ss + p;
    }
    float inv = 1.0f / sqrtf(ss);
    for (int k = 0; k < d; ++k) out[k] = x[k] * inv;
}

void asr_oracle_normalise(const float *x, long n, int d, float *out) {
    for (long i = 0; i < n; ++i) normalise_row(x + i * d, d, out + i * d);
}

static inline float score_rows(const float *q, const float *r, int d) {
    float acc = 0.0f;
    for (int k = 0; k < d; ++k) {
        float p = q[k] * r[k];
        acc = acc + p;
    }
    if (acc != acc) acc = -INFINITY;
    return acc;
}

/* a beats b ? */
static inline int beats(float sa, long long ia, float sb, long long ib) {
    return (sa > sb) || (sa == sb && ia < ib);
}

typedef struct { const float *qn, *dn; long nd; int d, k; long long idx_base; float *out_s; long long *out_i; } topk_ctx;
static void topk_range(long lo, long hi, void *vp);

/* out_s (nq,k), out_i (nq,k); unfilled slots = (-inf, -1) */
void asr_oracle_topk(const float *q, long nq, const float *db, long nd, int d, int k,
                     int normalise, long long idx_base, float *out_s, long long *out_i) {
    float *qn = (float *)malloc(sizeof(float) * nq * d);
    float *dn = (float *)malloc(sizeof(float) * nd * d);
    if (normalise) {
        asr_oracle_normalise(q, nq, d, qn);
        asr_oracle_normalise(db, nd, d, dn);
    } else {
        memcpy(qn, q, sizeof(float) * nq * d);
        memcpy(dn, db, sizeof(float) * nd * d);
    }
    topk_ctx c = {qn, dn, nd, d, k, idx_base, out_s, out_i};
    parallel_for(nq, topk_range, &c);
    free(qn); free(dn);
}

static void topk_range(long lo, long hi, void *vp) {
    topk_ctx *c = (topk_ctx *)vp;
    const float *qn = c->qn, *dn = c->dn; long nd = c->nd; int d = c->d, k = c->k;
    long long idx_base = c->idx_base; float *out_s = c->out_s; long long *out_i = c->out_i;
    for (long qi = lo; qi < hi; ++qi) {
        float *s = out_s + qi * k;
        long long *ix = out_i + qi * k;
        int cnt = 0;
        for (int j = 0; j < k; ++j) { s[j] = -INFINITY; ix[j] = -1; }
        for (long r = 0; r < nd; ++r) {
            float sc = score_rows(qn + qi * d, dn + r * d, d);
            if (cnt == k && !beats(sc, r, s[k - 1], ix[k - 1] - idx_base)) continue;
            int p = cnt < k ? cnt : k - 1;
            while (p > 0 && beats(sc, r, s[p - 1], ix[p - 1] - idx_base)) {
                s[p] = s[p - 1]; ix[p] = ix[p - 1]; --p;
            }
            s[p] = sc; ix[p] = r + idx_base;
            if (cnt < k) ++cnt;
        }
    }
}

typedef struct { const float *qn, *dn; long nd; int d; long kg, hg; long long *ranks; float *tscore; } rank_ctx;
static void rank_range(long lo, long hi, void *vp);

/* rank of the best correct item, groups as in eval_retrieval: correct(j) <=> j / kg == i / hg */
void asr_oracle_rank(const float *q, long nq, const float *db, long nd, int d, long kg, long hg,
                     int normalise, long long *ranks, float *tscore) {
    float *qn = (float *)malloc(sizeof(float) * nq * d);
    float *dn = (float *)malloc(sizeof(float) * nd * d);
    if (normalise) {
        asr_oracle_normalise(q, nq, d, qn);
        asr_oracle_normalise(db, nd, d, dn);
    } else {
        memcpy(qn, q, sizeof(float) * nq * d);
        memcpy(dn, db, sizeof(float) * nd * d);
    }
    rank_ctx c = {qn, dn, nd, d, kg, hg, ranks, tscore};
    parallel_for(nq, rank_range, &c);
    free(qn); free(dn);
}

static void rank_range(long lo, long hi, void *vp) {
    rank_ctx *c = (rank_ctx *)vp;
    const float *qn = c->qn, *dn = c->dn; long nd = c->nd, kg = c->kg, hg = c->hg; int d = c->d;
    long long *ranks = c->ranks; float *tscore = c->tscore;
    for (long i = lo; i < hi; ++i) {
        long g = i / hg;
        float best = -INFINITY; long bj = -1;
        for (long j = g * kg; j < (g + 1) * kg && j < nd; ++j) {
            float sc = score_rows(qn + i * d, dn + j * d, d);
            if (bj < 0 || sc > best) { best = sc; bj = j; }
        }
        long long cnt = 0;
        for (long j = 0; j < nd; ++j) {
            float sc = score_rows(qn + i * d, dn + j * d, d);
            cnt += beats(sc, j, best, bj);
        }
        ranks[i] = 1 + cnt;
        tscore[i] = best;
    }
}
