/*
 * oracle/klocal.c — CPU ORACLE (test infrastructure only; see sgs_oracle.h).
 * Restatement of the 1D k-local dynamic programme:
 *   right-environment table  branch_include/branch_exclude/evolve_stab/get_map_from_stabs/
 *                            generate_Sright_all          (cpp/state.cpp:129-238, py/state.py:469-539)
 *   DP state + transition    State::{add, move_to_next, check_Sright_validity, hash_value, merge_gs,
 *                            get_stab, get_stab_gs}       (cpp/state.cpp:251-442)
 *   per-site batch           StateMachine::{ctor, evolve_single, evolve, generate_gs} (cpp/state.cpp:445-550)
 * Wherever the reference iterates a std::map keyed by a 64-bit boost hash (Sright lists, state
 * dictionaries) this restatement keys on the full canonical content and iterates in sorted key
 * order; the reference's own order depends on the Boost version and, for the OpenMP merge, on
 * thread timing, so only exact-tie selections can differ.
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
#include "klocal_internal.h"

/* ------------------------------------------------------------------ right environments */

typedef struct { ostab s; int nex; int *ex; } obranch;

/* evolve_stab (cpp/state.cpp:147-185): branch include/exclude over the terms whose last site is m,
 * widen one site to the left, project onto the k sites ending at m.  Returns the number of groups. */
static int evolve_stab(const okl *K, const ostab *stab, int m, ostab **out)
{
    const oham *H = K->H;
    int b0 = H->bucket_off[m], b1 = H->bucket_off[m + 1];
    int nb = 1, cap = 4;
    obranch *br = (obranch *)malloc(sizeof(obranch) * cap);
    br[0].s = *stab; br[0].nex = 0; br[0].ex = (int *)malloc(sizeof(int) * (b1 - b0 + 1));
    for (int pos = b0; pos < b1; pos++) {
        const orow *P = &K->P[pos];
        int nn = 0, ncap = nb * 2 + 1;
        obranch *nbr = (obranch *)malloc(sizeof(obranch) * ncap);
        for (int i = 0; i < nb; i++) {
            obranch *B = &br[i];
            if (ostab_commute(&B->s, P) && !ostab_include(&B->s, P)) {
                /* branch_include (cpp/state.cpp:129-138): void if it makes an excluded term a member */
                obranch inc;
                inc.s = B->s; ostab_add(&inc.s, P);
                int ok = 1;
                for (int r = 0; r < B->nex; r++) if (ostab_include(&inc.s, &K->P[B->ex[r]])) { ok = 0; break; }
                if (ok) {
                    inc.nex = B->nex; inc.ex = (int *)malloc(sizeof(int) * (b1 - b0 + 1));
                    memcpy(inc.ex, B->ex, sizeof(int) * B->nex);
                    nbr[nn++] = inc;
                }
                /* branch_exclude (cpp/state.cpp:140-145) */
                B->ex[B->nex++] = pos;
                nbr[nn++] = *B;
            } else {
                nbr[nn++] = *B;
            }
        }
        free(br); br = nbr; nb = nn; cap = ncap;
    }
    int begin_next = stab->begin - 1 > 0 ? stab->begin - 1 : 0;
    ostab *res = (ostab *)malloc(sizeof(ostab) * nb);
    for (int i = 0; i < nb; i++) {
        br[i].s.begin = begin_next; br[i].s.end = stab->end;               /* stab.range(begin_next, end) */
        ostab_project(&br[i].s, begin_next, stab->end - 1, &res[i]);       /* :176-181 */
        free(br[i].ex);
    }
    free(br);
    *out = res;
    return nb;
}

static int cmp_pair(const void *a, const void *b)
{
    const int *x = (const int *)a, *y = (const int *)b;
    if (x[0] != y[0]) return x[0] < y[0] ? -1 : 1;
    return x[1] < y[1] ? -1 : (x[1] > y[1]);
}

/* one backward step: level m from level m+1 (loop body of generate_Sright_all, cpp/state.cpp:222-236,
 * + get_map_from_stabs :187-216).  out->map lists, for each level-m group, the level-(m+1) groups that
 * evolve into it (duplicates removed, as the Python reference's `set` does, py/state.py:523). */
void okl_level_step(const okl *K, const olevel *next, int m, olevel *out)
{
    int total = 0, cap = 16;
    ostab *all = (ostab *)malloc(sizeof(ostab) * cap);
    int *parent = (int *)malloc(sizeof(int) * cap);
    for (int i = 0; i < next->count; i++) {
        ostab *ch;
        int nc = evolve_stab(K, &next->g[i], m, &ch);
        for (int j = 0; j < nc; j++) {
            if (total == cap) { cap *= 2; all = (ostab *)realloc(all, sizeof(ostab) * cap); parent = (int *)realloc(parent, sizeof(int) * cap); }
            all[total] = ch[j]; parent[total] = i; total++;
        }
        free(ch);
    }
    /* distinct children, sorted by content */
    int *idx = (int *)malloc(sizeof(int) * total);
    for (int i = 0; i < total; i++) idx[i] = i;
    for (int i = 1; i < total; i++) {       /* insertion sort by ostab_cmp (lists are small) */
        int o = idx[i], j = i - 1;
        while (j >= 0 && ostab_cmp(&all[idx[j]], &all[o]) > 0) { idx[j + 1] = idx[j]; j--; }
        idx[j + 1] = o;
    }
    out->count = 0;
    out->g = (ostab *)malloc(sizeof(ostab) * (total > 0 ? total : 1));
    int *pairs = (int *)malloc(sizeof(int) * 2 * (total > 0 ? total : 1));
    for (int i = 0; i < total; i++) {
        if (i == 0 || ostab_cmp(&all[idx[i - 1]], &all[idx[i]]) != 0) out->g[out->count++] = all[idx[i]];
        pairs[2 * i] = out->count - 1; pairs[2 * i + 1] = parent[idx[i]];
    }
    qsort(pairs, total, sizeof(int) * 2, cmp_pair);
    out->map_off = (int *)calloc(out->count + 1, sizeof(int));
    out->map_idx = (int *)malloc(sizeof(int) * (total > 0 ? total : 1));
    int nl = 0;
    for (int i = 0; i < total; i++) {
        if (i > 0 && pairs[2 * i] == pairs[2 * i - 2] && pairs[2 * i + 1] == pairs[2 * i - 1]) continue;
        out->map_idx[nl++] = pairs[2 * i + 1];
        out->map_off[pairs[2 * i] + 1] = nl;
    }
    for (int i = 1; i <= out->count; i++) if (out->map_off[i] < out->map_off[i - 1]) out->map_off[i] = out->map_off[i - 1];
    free(all); free(parent); free(idx); free(pairs);
}

void olevel_free(olevel *l) { free(l->g); free(l->map_off); free(l->map_idx); l->g = NULL; l->map_off = l->map_idx = NULL; l->count = 0; }

okl *okl_new(const oham *H, int nmax_hist)
{
    okl *K = (okl *)calloc(1, sizeof(okl));
    K->H = H; K->n = H->n; K->k = H->k; K->T = H->T; K->nmax = nmax_hist;
    K->P = (orow *)malloc(sizeof(orow) * (H->T > 0 ? H->T : 1));
    K->w = (double *)malloc(sizeof(double) * (H->T > 0 ? H->T : 1));
    for (int i = 0; i < H->T; i++) { K->P[i] = H->P[H->order[i]]; K->w[i] = H->w[H->order[i]]; }
    K->lv = (olevel *)calloc(H->n + 2, sizeof(olevel));
    K->vwords = (H->T + 63) / 64 + 1;
    return K;
}

void okl_free(okl *K)
{
    if (!K) return;
    for (int i = 0; i < K->nstates; i++) ostate_free(&K->states[i]);
    free(K->states);
    for (int m = 0; m <= K->n + 1; m++) olevel_free(&K->lv[m]);
    free(K->lv); free(K->P); free(K->w); free(K);
}

/* generate_Sright_all (cpp/state.cpp:218-238) */
void okl_generate_sright_all(okl *K)
{
    int n = K->n, k = K->k;
    olevel *top = &K->lv[n];
    top->count = 1;
    top->g = (ostab *)malloc(sizeof(ostab));
    ostab_init(&top->g[0], n - k + 1, n + 1);         /* Stab::empty(H.k, H.n - H.k + 1) */
    top->map_off = (int *)calloc(2, sizeof(int));
    top->map_idx = (int *)malloc(sizeof(int));
    for (int m = n - 1; m >= 0; m--) okl_level_step(K, &K->lv[m + 1], m, &K->lv[m]);
}

/* ------------------------------------------------------------------ DP states */

void ostate_free(ostate *s) { oste_free(&s->st); free(s->valid); free(s->key); s->valid = NULL; s->key = NULL; }

void ostate_copy(const okl *K, ostate *d, const ostate *s)
{
    d->m = s->m; d->sright = s->sright;
    oste_copy(&d->st, &s->st);
    d->valid = (uint64_t *)malloc(sizeof(uint64_t) * K->vwords);
    memcpy(d->valid, s->valid, sizeof(uint64_t) * K->vwords);
    d->key = NULL; d->keylen = 0;
}

void ostate_init(const okl *K, ostate *s, int sright)   /* State ctor (cpp/state.cpp:251-262) */
{
    s->m = 0; s->sright = sright;
    oste_init(&s->st, 0, 0, K->nmax);
    s->valid = (uint64_t *)malloc(sizeof(uint64_t) * K->vwords);
    memset(s->valid, 0xff, sizeof(uint64_t) * K->vwords);
    s->key = NULL; s->keylen = 0;
}

static int vget(const uint64_t *v, int i) { return (int)((v[i >> 6] >> (i & 63)) & 1); }
static void vclr(uint64_t *v, int i) { v[i >> 6] &= ~(1ULL << (i & 63)); }
static void vset(uint64_t *v, int i, int b) { if (b) v[i >> 6] |= 1ULL << (i & 63); else vclr(v, i); }

/* State::hash_value content (cpp/state.cpp:311-330): m, window group (with position), valid indices for
 * last site in [m+1, min(m+k-1, n)), Sright.  Energies and history are NOT part of the key. */
void ostate_make_key(const okl *K, ostate *s)
{
    int lo = s->m + 1, hi = s->m + K->k - 1 < K->n ? s->m + K->k - 1 : K->n;
    int nbits = 0;
    if (hi > lo) nbits = K->H->bucket_off[hi] - K->H->bucket_off[lo];
    int len = 4 * 4 + s->st.s.m * OW * 8 + (nbits + 7) / 8;
    unsigned char *key = (unsigned char *)calloc(len, 1);
    int32_t hdr[4] = { s->m, s->st.s.m, s->st.s.begin, s->sright };
    memcpy(key, hdr, sizeof(hdr));
    unsigned char *p = key + sizeof(hdr);
    for (int i = 0; i < s->st.s.m; i++) { memcpy(p, s->st.s.g[i].b, OW * 8); p += OW * 8; }
    for (int i = 0; i < nbits; i++)
        if (vget(s->valid, K->H->bucket_off[lo] + i)) p[i >> 3] |= (unsigned char)(1u << (i & 7));
    free(s->key);
    s->key = key; s->keylen = len;
}

static int key_cmp(const ostate *a, const ostate *b)
{
    int l = a->keylen < b->keylen ? a->keylen : b->keylen;
    int c = memcmp(a->key, b->key, l);
    if (c) return c;
    return a->keylen < b->keylen ? -1 : (a->keylen > b->keylen);
}

static int state_cmp(const void *a, const void *b) { return key_cmp((const ostate *)a, (const ostate *)b); }

/* State::add (cpp/state.cpp:367-380) */
static void state_add(const okl *K, ostate *s, int pos)
{
    const oham *H = K->H;
    int file_index = H->order[pos];
    oste_add(&s->st, &K->P[pos], 1, file_index);
    int qlast = H->end[file_index] - 1;
    int qhi = qlast + K->k < K->n ? qlast + K->k : K->n;
    for (int q = qlast; q < qhi; q++)
        for (int t = H->bucket_off[q]; t < H->bucket_off[q + 1]; t++)
            if (vget(s->valid, t) && !ostab_commute(&s->st.s, &K->P[t])) vclr(s->valid, t);
}

/* State::move_to_next (cpp/state.cpp:382-413).  Returns 0 for a dead state. */
static int state_move_to_next(const okl *K, ostate *r)
{
    const oham *H = K->H;
    int m = r->m;
    const ostab *Sright = &K->lv[m].g[r->sright];
    ostab tmp = r->st.s;                                   /* Stab stab_tmp(result.stab.Ps) */
    for (int i = 0; i < Sright->m; i++) {
        assert(ostab_commute(&tmp, &Sright->g[i]));
        if (!ostab_include(&tmp, &Sright->g[i])) ostab_add(&tmp, &Sright->g[i]);
    }
    for (int pos = H->bucket_off[m]; pos < H->bucket_off[m + 1]; pos++) {
        const orow *P = &K->P[pos];
        if (!ostab_include(&tmp, P)) continue;
        if (!ostab_include(Sright, P)) return 0;
        if (!ostab_include(&r->st.s, P)) state_add(K, r, pos);
    }
    for (int pos = H->bucket_off[m]; pos < H->bucket_off[m + 1]; pos++)
        if (ostab_include(&r->st.s, &K->P[pos])) oste_add_energy(&r->st, &K->P[pos], K->w[pos]);
    r->st.s.begin = m - K->k + 1 > 0 ? m - K->k + 1 : 0;   /* stab.range(max(m-k+1,0), m+1) */
    r->st.s.end = m + 1;
    return 1;
}

/* State::check_Sright_validity (cpp/state.cpp:347-365) */
static int state_check_sright(const okl *K, const ostate *s, const ostab *Snext)
{
    const oham *H = K->H;
    int lo = s->m + 1, hi = s->m + K->k + 1 < K->n ? s->m + K->k + 1 : K->n;
    int begin = s->m - K->k + 2 > 0 ? s->m - K->k + 2 : 0, end = s->m + 2;
    int cnt = 0;
    orow *rows = NULL;
    if (hi > lo) {
        int t0 = H->bucket_off[lo], t1 = H->bucket_off[hi];
        rows = (orow *)malloc(sizeof(orow) * (t1 - t0 + 1));
        for (int t = t0; t < t1; t++)
            if (vget(s->valid, t)) { rows[cnt] = K->P[t]; orow_mask_window(&rows[cnt], begin, end); cnt++; }
    }
    int ok = 1;
    for (int i = 0; i < Snext->m && ok; i++) ok = orows_span_contains(rows, cnt, &Snext->g[i]);
    free(rows);
    return ok;
}

/* StateMachine::evolve_single (cpp/state.cpp:460-481).  Appends children to *out. */
int okl_evolve_single(const okl *K, const ostate *raw, ostate **out, int *nout, int *cap)
{
    ostate st;
    ostate_copy(K, &st, raw);
    if (!state_move_to_next(K, &st)) { ostate_free(&st); return 0; }
    oste_project_to(&st.st, K->k - 1);
    const olevel *L = &K->lv[st.m], *Ln = &K->lv[st.m + 1];
    int made = 0;
    for (int j = L->map_off[st.sright]; j < L->map_off[st.sright + 1]; j++) {
        int idx = L->map_idx[j];
        if (!state_check_sright(K, &st, &Ln->g[idx])) continue;
        if (*nout == *cap) { *cap = *cap ? *cap * 2 : 8; *out = (ostate *)realloc(*out, sizeof(ostate) * *cap); }
        ostate *c = &(*out)[(*nout)++];
        ostate_copy(K, c, &st);
        c->sright = idx; c->m = st.m + 1;
        ostate_make_key(K, c);
        made++;
    }
    ostate_free(&st);
    return made;
}

/* keyed merge (loop body of StateMachine::evolve, cpp/state.cpp:501-516): states[] is kept sorted by key */
static uint64_t key_hash(const ostate *s)
{
    uint64_t h = 1469598103934665603ULL;
    for (int i = 0; i < s->keylen; i++) { h ^= s->key[i]; h *= 1099511628211ULL; }
    return h;
}

/* StateMachine::evolve (cpp/state.cpp:483-533): all states of site m → merged states of site m+1 */
void okl_evolve(okl *K, int threads)
{
    int ns = K->nstates;
    ostate **kids = (ostate **)calloc(ns > 0 ? ns : 1, sizeof(ostate *));
    int *nk = (int *)calloc(ns > 0 ? ns : 1, sizeof(int));
#pragma omp parallel for schedule(dynamic, 1) num_threads(threads > 1 ? threads : 1)
    for (int i = 0; i < ns; i++) {
        int cap = 0;
        okl_evolve_single(K, &K->states[i], &kids[i], &nk[i], &cap);
    }
    K->transitions += (uint64_t)ns;
    /* sequential merge in deterministic (parent, child) order; on an exact tie the later one wins,
     * like merge_gs(dict[h], new_state) with its strict '<' (cpp/stab.cpp:832) */
    int total = 0;
    for (int i = 0; i < ns; i++) total += nk[i];
    int hcap = 16;
    while (hcap < total * 2) hcap *= 2;
    int *slot = (int *)malloc(sizeof(int) * hcap);
    for (int i = 0; i < hcap; i++) slot[i] = -1;
    ostate *ns_states = (ostate *)malloc(sizeof(ostate) * (total > 0 ? total : 1));
    int nn = 0;
    for (int i = 0; i < ns; i++) {
        for (int j = 0; j < nk[i]; j++) {
            ostate *c = &kids[i][j];
            uint64_t h = key_hash(c) & (uint64_t)(hcap - 1);
            for (;;) {
                if (slot[h] < 0) { slot[h] = nn; ns_states[nn++] = *c; break; }
                ostate *e = &ns_states[slot[h]];
                if (key_cmp(e, c) == 0) { oste_merge_gs(&e->st, &c->st); ostate_free(c); break; }
                h = (h + 1) & (uint64_t)(hcap - 1);
            }
        }
        free(kids[i]);
    }
    free(kids); free(nk); free(slot);
    for (int i = 0; i < K->nstates; i++) ostate_free(&K->states[i]);
    free(K->states);
    qsort(ns_states, nn, sizeof(ostate), state_cmp);
    K->states = ns_states; K->nstates = nn;
    K->m += 1;
}

/* StateMachine ctor (cpp/state.cpp:445-458): one state per right environment at site 0 */
void okl_init_states(okl *K)
{
    int c = K->lv[0].count;
    K->states = (ostate *)malloc(sizeof(ostate) * (c > 0 ? c : 1));
    K->nstates = c; K->m = 0;
    for (int i = 0; i < c; i++) { ostate_init(K, &K->states[i], i); ostate_make_key(K, &K->states[i]); }
    qsort(K->states, c, sizeof(ostate), state_cmp);
}

static double now_s(void)
{
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return ts.tv_sec + 1e-9 * ts.tv_nsec;
}

ores *oracle_klocal(const oham *H, int threads)
{
    double t0 = now_s();
    okl *K = okl_new(H, H->n);                   /* history capacity n: StabEnergies::empty(0,0,n) */
    okl_generate_sright_all(K);
    okl_init_states(K);
    ores *r = (ores *)calloc(1, sizeof(ores));
    r->n = H->n; r->T = H->T; r->nsites = H->n;
    r->sright_counts = (int *)malloc(sizeof(int) * (H->n + 1));
    r->states_per_site = (int *)malloc(sizeof(int) * (H->n + 1));
    for (int m = 0; m <= H->n; m++) r->sright_counts[m] = K->lv[m].count;
    r->states_per_site[0] = K->nstates;
    while (K->m < H->n) { okl_evolve(K, threads); r->states_per_site[K->m] = K->nstates; }
    r->transitions = K->transitions;
    /* StateMachine::generate_gs (cpp/state.cpp:535-550): first strict minimum */
    int best = -1;
    double bestE = DBL_MAX;
    for (int i = 0; i < K->nstates; i++) {
        const oste *e = &K->states[i].st;
        double mn = e->Es[oste_argmin(e)];
        if (mn < bestE) { bestE = mn; best = i; }
    }
    r->term_signs = (int8_t *)calloc(H->T > 0 ? H->T : 1, 1);
    if (best >= 0) {
        /* State::get_stab_gs / get_stab (cpp/state.cpp:421-442): rebuild the generators from the history */
        const oste *e = &K->states[best].st;
        int arg = oste_argmin(e);
        ostab st;
        ostab_init(&st, 0, H->n);
        for (int i = 0; i < e->nPs[arg]; i++) {
            orow g = H->P[e->Pidx[arg * e->nmax + i]];
            if (e->Psgn[arg * e->nmax + i] < 0) g.q = (g.q + 2) & 3;
            st.g[st.m++] = g;
        }
        ostab_canonicalize(&st, NULL);
        r->energy = e->Es[arg];
        r->m = st.m;
        memcpy(r->gens, st.g, sizeof(orow) * st.m);
        for (int t = 0; t < H->T; t++) r->term_signs[t] = (int8_t)ostab_energy(&st, &H->P[t]);
    }
    okl_free(K);
    r->seconds = now_s() - t0;
    return r;
}

/* helpers for the periodic driver (oracle/periodic.c) */
void okl_vshift(const okl *K, ostate *s, int add)
{
    /* State.shift (py/state.py:309-322): invalid_pauli_indices[qlast + add] = invalid_pauli_indices[qlast];
     * buckets that have no source become "all valid". */
    const oham *H = K->H;
    uint64_t *nv = (uint64_t *)malloc(sizeof(uint64_t) * K->vwords);
    memset(nv, 0xff, sizeof(uint64_t) * K->vwords);
    for (int q = 0; q < K->n; q++) {
        int d = q + add;
        if (d < 0 || d >= K->n) continue;
        int ns = H->bucket_off[q + 1] - H->bucket_off[q], nd = H->bucket_off[d + 1] - H->bucket_off[d];
        for (int i = 0; i < nd; i++) {
            int b = i < ns ? vget(s->valid, H->bucket_off[q] + i) : 1;
            vset(nv, H->bucket_off[d] + i, b);
        }
    }
    free(s->valid);
    s->valid = nv;
}
