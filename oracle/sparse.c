/*
 * oracle/sparse.c — CPU ORACLE (test infrastructure only; see sgs_oracle.h).
 * Restatement of the sparse search, FullState (cpp/state.cpp:552-687; py/state.py:88-147):
 * a binary include/exclude tree over the terms in paulis_list order; every leaf is a closed
 * commuting group carrying the 2^m table of energies over its generators' signs.
 *
 * The reference walks the tree level-synchronously (BFS) and keeps every state alive; this
 * restatement walks the same tree depth-first (include branch first), which visits the leaves in
 * exactly the order of the reference's final `states` vector (each BFS step replaces a state by
 * [include, exclude] in place, cpp/state.cpp:651-660).  Selection = first strict minimum in that
 * order, as py/state.py:142 does; the C++ iterates a std::map keyed by a boost hash instead
 * (cpp/state.cpp:663-679), an order that is not reproducible, so the leaf order is the
 * documented tie-break.  With OpenMP the same winner is chosen by comparing (energy, included-term
 * sequence) — the include-first order is the lexicographic order of those sequences with "end of
 * sequence" sorting last.
 */
#include <assert.h>
#include <float.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#ifdef _OPENMP
#include <omp.h>
#endif
#include "oracle_internal.h"

typedef struct {
    oste st;
    int np, *p;        /* remaining terms: positions in paulis_list order (cpp/state.hpp:127) */
    int nex, *ex;      /* excluded terms */
    int ninc, *inc;    /* included terms so far (tie-break bookkeeping, not in the reference) */
} ofull;

/* per-thread accumulator: leaves are counted and compared without any lock; the threads' records are
 * merged at the end with the same (energy, included-term sequence) order, so the result does not depend on
 * the schedule */
typedef struct {
    uint64_t candidates;
    uint64_t rank_hist[OMAXM + 1];
    double sum_2m;
    int have;
    double bestE;
    oste best;
    int nbest, *bestseq;
    uint64_t since_check;
    char pad[64];
} oacc;

typedef struct {
    const oham *H;
    const orow *P;     /* [T] terms in paulis_list order */
    const double *w;
    oacc *acc;         /* [threads] */
    int par_min;       /* a node with at least this many remaining terms hands its two branches to OpenMP tasks */
    /* bounded sampling (bench.py cpu_baseline): stop after max_leaves leaves or at the deadline */
    uint64_t max_leaves;
    double deadline;
    volatile uint64_t leaves_seen;
    volatile int stop;
    /* uniform sample (bench.py cpu_baseline): of the subtrees rooted at the nodes with exactly two included
     * terms, only every `stride`-th (ordinal = index pair hashed) is walked; leaves with fewer than two
     * included terms are always visited (a few hundred) */
    int stride, offset;
} octx;

static void ofull_free(ofull *s) { oste_free(&s->st); free(s->p); free(s->ex); free(s->inc); }

static void ofull_copy(ofull *d, const ofull *s, int T)
{
    oste_copy(&d->st, &s->st);
    d->np = s->np; d->nex = s->nex; d->ninc = s->ninc;
    d->p = (int *)malloc(sizeof(int) * (T + 1));
    d->ex = (int *)malloc(sizeof(int) * (T + 1));
    d->inc = (int *)malloc(sizeof(int) * (T + 1));
    memcpy(d->p, s->p, sizeof(int) * s->np);
    memcpy(d->ex, s->ex, sizeof(int) * s->nex);
    memcpy(d->inc, s->inc, sizeof(int) * s->ninc);
}

/* FullState::include_pauli (cpp/state.cpp:578-606) */
static void include_pauli(const octx *c, const ofull *self, int t, ofull *out)
{
    int T = c->H->T;
    ofull_copy(out, self, T);
    oste_add(&out->st, &c->P[t], 1, -1);
    out->inc[out->ninc++] = t;
    int np = 0;
    for (int i = 0; i < self->np; i++) {
        int q = self->p[i];
        const orow *Q = &c->P[q];
        if (!ostab_commute(&out->st.s, Q)) continue;
        if (ostab_include(&out->st.s, Q)) {
            oste_add_energy(&out->st, Q, c->w[q]);
        } else {
            /* tmp_stab = stab.add(Q, false); drop Q if it would pull an excluded term into the group */
            ostab tmp = out->st.s;
            ostab_add(&tmp, Q);
            int flag = 0;
            for (int r = 0; r < self->nex; r++)
                if (ostab_include(&tmp, &c->P[self->ex[r]])) { flag = 1; break; }
            if (!flag) out->p[np++] = q;
        }
    }
    out->np = np;
}

/* FullState::exclude_pauli (cpp/state.cpp:608-621) */
static void exclude_pauli(const octx *c, const ofull *self, int t, ofull *out)
{
    int T = c->H->T;
    ofull_copy(out, self, T);
    out->ex[out->nex++] = t;
    int np = 0;
    for (int i = 0; i < self->np; i++) {
        int q = self->p[i];
        ostab tmp = self->st.s;
        ostab_add(&tmp, &c->P[q]);
        if (!ostab_include(&tmp, &c->P[t])) out->p[np++] = q;
    }
    out->np = np;
}

static int seq_before(const int *a, int na, const int *b, int nb)
{
    for (int i = 0;; i++) {
        if (i == na) return 0;            /* a ended: a is a prefix of b (or equal) → b (longer) comes first */
        if (i == nb) return 1;            /* b ended first → a (longer) comes first */
        if (a[i] != b[i]) return a[i] < b[i];
    }
}

static double now_s(void);

static int my_thread(void)
{
#ifdef _OPENMP
    return omp_get_thread_num();
#else
    return 0;
#endif
}

static void leaf(octx *c, const ofull *s)
{
    int m = s->st.s.m;
    size_t sz = (size_t)1 << m;
    double mn = DBL_MAX;
    for (size_t o = 0; o < sz; o++) if (s->st.Es[o] < mn) mn = s->st.Es[o];   /* Es.minCoeff() */
    oacc *a = &c->acc[my_thread()];
    a->candidates++;
    a->rank_hist[m]++;
    a->sum_2m += (double)sz;
    int better = !a->have || mn < a->bestE ||
                 (mn == a->bestE && seq_before(s->inc, s->ninc, a->bestseq, a->nbest));
    if (better) {
        if (a->have) oste_free(&a->best);
        oste_copy(&a->best, &s->st);
        a->bestE = mn; a->have = 1;
        a->nbest = s->ninc;
        memcpy(a->bestseq, s->inc, sizeof(int) * s->ninc);
    }
    if ((c->max_leaves || c->deadline > 0) && ++a->since_check >= 64) {
        uint64_t seen = __atomic_add_fetch(&c->leaves_seen, a->since_check, __ATOMIC_RELAXED);
        a->since_check = 0;
        if ((c->max_leaves && seen >= c->max_leaves) || (c->deadline > 0 && now_s() > c->deadline)) c->stop = 1;
    }
}

static unsigned pair_hash(int a, int b)
{
    unsigned h = (unsigned)a * 2654435761u ^ ((unsigned)b * 40503u + 0x9E3779B9u);
    h ^= h >> 15; h *= 2246822519u; h ^= h >> 13;
    return h;
}

/* FullState::evolve (cpp/state.cpp:636-662) as a recursion; takes ownership of s */
static void walk(octx *c, ofull *s, int depth)
{
    if (c->stop) { ofull_free(s); free(s); return; }
    if (s->np == 0) { leaf(c, s); ofull_free(s); free(s); return; }
    int t = s->p[0];
    ofull *inc = (ofull *)malloc(sizeof(ofull)), *exc = (ofull *)malloc(sizeof(ofull));
    const int big = c->par_min > 0 && s->np >= c->par_min;
    /* uniform sample: the include branch that would create a node with two included terms is taken for
     * every stride-th (first term, second term) pair only */
    int take_inc = 1;
    if (c->stride > 1 && s->ninc == 1 && (int)(pair_hash(s->inc[0], t) % (unsigned)c->stride) != c->offset) take_inc = 0;
    if (take_inc) include_pauli(c, s, t, inc); else { free(inc); inc = NULL; }
    exclude_pauli(c, s, t, exc);
    ofull_free(s); free(s);
    if (big) {
        if (inc) {
#pragma omp task firstprivate(inc, depth)
            walk(c, inc, depth + 1);
        }
#pragma omp task firstprivate(exc, depth)
        walk(c, exc, depth + 1);
    } else {
        if (inc) walk(c, inc, depth + 1);
        walk(c, exc, depth + 1);
    }
}

static double now_s(void)
{
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return ts.tv_sec + 1e-9 * ts.tv_nsec;
}

static ores *oracle_sparse_impl(const oham *H, int threads, uint64_t max_leaves, double seconds, int stride, int offset)
{
    double t0 = now_s();
    int T = H->T;
    octx c;
    memset(&c, 0, sizeof(c));
    c.H = H;
    orow *P = (orow *)malloc(sizeof(orow) * (T > 0 ? T : 1));
    double *w = (double *)malloc(sizeof(double) * (T > 0 ? T : 1));
    for (int i = 0; i < T; i++) { P[i] = H->P[H->order[i]]; w[i] = H->w[H->order[i]]; }
    c.P = P; c.w = w;
    if (threads < 1) threads = 1;
    c.acc = (oacc *)calloc((size_t)threads, sizeof(oacc));
    for (int i = 0; i < threads; i++) c.acc[i].bestseq = (int *)malloc(sizeof(int) * (T + 1));
    c.par_min = threads > 1 ? 6 : 0;
    c.max_leaves = max_leaves;
    c.deadline = seconds > 0 ? t0 + seconds : 0.0;
    c.stride = stride > 1 ? stride : 1;
    c.offset = stride > 1 ? ((offset % stride) + stride) % stride : 0;

    ofull *root = (ofull *)malloc(sizeof(ofull));
    oste_init(&root->st, 0, H->n, 0);          /* FullState ctor, cpp/state.cpp:555-564 */
    root->p = (int *)malloc(sizeof(int) * (T + 1));
    root->ex = (int *)malloc(sizeof(int) * (T + 1));
    root->inc = (int *)malloc(sizeof(int) * (T + 1));
    root->np = T; root->nex = 0; root->ninc = 0;
    for (int i = 0; i < T; i++) root->p[i] = i;
#ifdef _OPENMP
    if (threads > 1) {
#pragma omp parallel num_threads(threads)
#pragma omp single
        walk(&c, root, 0);
    } else
#endif
        walk(&c, root, 0);

    ores *r = (ores *)calloc(1, sizeof(ores));
    r->n = H->n; r->T = T;
    /* merge the per-thread records: counts add up; best = minimum energy, ties by the include-first order */
    int bi = -1;
    for (int i = 0; i < threads; i++) {
        const oacc *a = &c.acc[i];
        r->candidates += a->candidates;
        for (int k = 0; k <= OMAXM; k++) r->rank_hist[k] += a->rank_hist[k];
        r->sum_2m += a->sum_2m;
        if (!a->have) continue;
        if (bi < 0 || a->bestE < c.acc[bi].bestE ||
            (a->bestE == c.acc[bi].bestE && seq_before(a->bestseq, a->nbest, c.acc[bi].bestseq, c.acc[bi].nbest))) bi = i;
    }
    oste cbest = c.acc[bi].best;
    /* cpp/state.cpp:680-686: signs from the first argmin (generator 0 = MSB), then canonicalise */
    int arg = oste_argmin(&cbest);
    int m = cbest.s.m;
    ostab st = cbest.s;
    for (int i = 0; i < m; i++)
        if ((arg >> (m - 1 - i)) & 1) st.g[i].q = (st.g[i].q + 2) & 3;
    ostab_canonicalize(&st, NULL);
    r->energy = cbest.Es[arg];
    r->m = m;
    memcpy(r->gens, st.g, sizeof(orow) * m);
    r->term_signs = (int8_t *)malloc(T > 0 ? T : 1);
    for (int t = 0; t < T; t++) r->term_signs[t] = (int8_t)ostab_energy(&st, &H->P[t]);   /* cpp/sparse.cpp:13-16 */
    for (int i = 0; i < threads; i++) { if (c.acc[i].have) oste_free(&c.acc[i].best); free(c.acc[i].bestseq); }
    free(c.acc); free(P); free(w);
    r->seconds = now_s() - t0;
    return r;
}

ores *oracle_sparse(const oham *H, int threads) { return oracle_sparse_impl(H, threads, 0, 0.0, 1, 0); }

/* bounded sample of the same search: stops after max_leaves leaves or `seconds` of wall time; the
 * result's `candidates` / `seconds` give the CPU rate, its energy is only the best of the sample */
ores *oracle_sparse_sample(const oham *H, int threads, uint64_t max_leaves, double seconds)
{
    return oracle_sparse_impl(H, threads, max_leaves, seconds, 1, 0);
}

/* uniform sample of the same search: only the subtrees below every stride-th (hashed) pair of first two
 * included terms are walked, completely; `candidates` / `seconds` give a CPU rate that is representative of
 * the whole tree (the include-first prefix of oracle_sparse_sample sits in its densest corner) */
ores *oracle_sparse_stride(const oham *H, int threads, int stride, int offset)
{
    return oracle_sparse_impl(H, threads, 0, 0.0, stride, offset);
}
