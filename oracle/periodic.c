/*
 * oracle/periodic.c — CPU ORACLE (test infrastructure only; see sgs_oracle.h).
 * Restatement of the Python-only infinite-periodic path:
 *   Hamiltonian.from_file_periodic   py/state.py:51-72   (in ham.c)
 *   generate_Sright_periodic         py/state.py:542-567
 *   StateMachinePeriodic             py/state.py:402-462 (same transition as StateMachine)
 *   State.shift / clear_history      py/state.py:309-325
 *   EnergyLevel.clear / get_state    py/stab.py:282-305
 *   the driver script                py/klocal_periodic.py:1-63
 */
#include <assert.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "oracle_internal.h"
#include "klocal_internal.h"

static void row_shift(orow *r, int add)      /* Pauli.shift (py/pauli.py:61): move by `add` sites */
{
    uint64_t o[OW];
    memset(o, 0, sizeof(o));
    int bits = 2 * (add < 0 ? -add : add), ws = bits >> 6, bs = bits & 63;
    for (int i = 0; i < OW; i++) {
        if (add >= 0) {
            int d = i + ws;
            if (d < OW) o[d] |= r->b[i] << bs;
            if (bs && d + 1 < OW) o[d + 1] |= r->b[i] >> (64 - bs);
        } else {
            int d = i - ws;
            if (d >= 0) o[d] |= r->b[i] >> bs;
            if (bs && d - 1 >= 0) o[d - 1] |= r->b[i] << (64 - bs);
        }
    }
    memcpy(r->b, o, sizeof(o));
}

static void stab_shift(ostab *s, int add)
{
    for (int i = 0; i < s->m; i++) row_shift(&s->g[i], add);
    s->begin += add; s->end += add;
}

/* State.shift(add, clear_history) (py/state.py:309-325) */
static int state_shift(const okl *K, ostate *s, int add, int clear_history)
{
    ostab sr = K->lv[s->m].g[s->sright];
    stab_shift(&sr, add);
    s->m += add;
    stab_shift(&s->st.s, add);
    const olevel *L = &K->lv[s->m];
    int found = -1;
    for (int i = 0; i < L->count; i++) if (ostab_equal(&L->g[i], &sr)) { found = i; break; }
    if (found < 0) return -1;          /* the Python would raise KeyError at the next evolve */
    s->sright = found;
    okl_vshift(K, s, add);
    if (clear_history) oste_clear_history(&s->st);
    ostate_make_key(K, s);
    return 0;
}

static int key_eq(const ostate *a, const ostate *b)
{
    return a->keylen == b->keylen && memcmp(a->key, b->key, a->keylen) == 0;
}

static int key_cmpq(const void *a, const void *b)
{
    const ostate *x = (const ostate *)a, *y = (const ostate *)b;
    int l = x->keylen < y->keylen ? x->keylen : y->keylen;
    int c = memcmp(x->key, y->key, l);
    if (c) return c;
    return x->keylen < y->keylen ? -1 : (x->keylen > y->keylen);
}

/* periodic_evolve_states (py/klocal_periodic.py:11-18).  Takes copies of `in`; returns a new array. */
static ostate *periodic_evolve_states(okl *K, const ostate *in, int nin, int clear_history, int c, int *nout)
{
    const oham *H = K->H;
    for (int i = 0; i < K->nstates; i++) ostate_free(&K->states[i]);
    free(K->states);
    /* state_dict = {hash(state): state}: a later state with the same key replaces an earlier one */
    K->states = (ostate *)malloc(sizeof(ostate) * (nin > 0 ? nin : 1));
    K->nstates = 0;
    for (int i = 0; i < nin; i++) {
        int dup = -1;
        for (int j = 0; j < K->nstates; j++) if (key_eq(&K->states[j], &in[i])) { dup = j; break; }
        ostate cp;
        ostate_copy(K, &cp, &in[i]);
        ostate_make_key(K, &cp);
        if (dup >= 0) { ostate_free(&K->states[dup]); K->states[dup] = cp; }
        else K->states[K->nstates++] = cp;
    }
    qsort(K->states, K->nstates, sizeof(ostate), key_cmpq);
    K->m = H->k + H->l - 1;
    for (int i = 0; i < H->l * c; i++) okl_evolve(K, 1);
    ostate *out = (ostate *)malloc(sizeof(ostate) * (K->nstates > 0 ? K->nstates : 1));
    for (int i = 0; i < K->nstates; i++) {
        ostate_copy(K, &out[i], &K->states[i]);
        int rc = state_shift(K, &out[i], -H->l * c, clear_history);
        assert(rc == 0); (void)rc;
    }
    *nout = K->nstates;
    qsort(out, *nout, sizeof(ostate), key_cmpq);
    return out;
}

static int orbit_cmp(const void *a, const void *b)
{
    const operiodic_orbit *x = (const operiodic_orbit *)a, *y = (const operiodic_orbit *)b;
    if (x->energy != y->energy) return x->energy < y->energy ? -1 : 1;
    if (x->begin != y->begin) return x->begin < y->begin ? -1 : 1;
    if (x->end != y->end) return x->end < y->end ? -1 : 1;
    if (x->m != y->m) return x->m < y->m ? -1 : 1;
    for (int i = 0; i < x->m; i++) {
        char sa[OMAXQ + 2], sb[OMAXQ + 2];
        orow_to_string(&x->gens[i], x->begin, x->end, sa);
        orow_to_string(&y->gens[i], y->begin, y->end, sb);
        int c = strcmp(sa, sb);
        if (c) return c;
    }
    return 0;
}

operiodic *oracle_periodic(const char *path, int cmax, int c, char *err, int errlen)
{
    const double max_float = 1e100;                          /* py/klocal_periodic.py:4 */
    oham *H = oham_load(path, 1, cmax, err, errlen);
    if (!H) return NULL;
    int l = H->l, k = H->k, n = H->n;
    operiodic *R = (operiodic *)calloc(1, sizeof(operiodic));
    R->l = l; R->k = k; R->n = n; R->nterms = H->T;
    int logcap = 64;
    R->sright_log = (int *)malloc(sizeof(int) * 2 * logcap);
#define LOG(mm, cc) do { if (R->nlog == logcap) { logcap *= 2; R->sright_log = (int *)realloc(R->sright_log, sizeof(int) * 2 * logcap); } \
        R->sright_log[2 * R->nlog] = (mm); R->sright_log[2 * R->nlog + 1] = (cc); R->nlog++; } while (0)

    okl *K = okl_new(H, n + 3 * (n - k + 1));                /* StateMachinePeriodic: n + self.l * 3 (py/state.py:403-414) */

    /* generate_Sright_periodic (py/state.py:542-567) */
    int end = l * (1 + cmax) + k;
    olevel cur;
    cur.count = 1;
    cur.g = (ostab *)malloc(sizeof(ostab));
    ostab_init(&cur.g[0], end - k + 1, end + 1);
    cur.map_off = (int *)calloc(2, sizeof(int));
    cur.map_idx = (int *)malloc(sizeof(int));
    olevel id;
    memset(&id, 0, sizeof(id));
    int have_id = 0;
    for (;;) {
        olevel lev = cur;                                    /* ownership moves along the sweep */
        for (int m = end - 1; m >= end - l; m--) {
            olevel nxt;
            okl_level_step(K, &lev, m, &nxt);
            olevel_free(&lev);
            lev = nxt;
            LOG(m, lev.count);
        }
        int same = have_id && id.count == lev.count;
        for (int i = 0; same && i < lev.count; i++) same = ostab_equal(&id.g[i], &lev.g[i]);
        if (have_id) free(id.g);
        id.count = lev.count;
        id.g = (ostab *)malloc(sizeof(ostab) * (lev.count > 0 ? lev.count : 1));
        memcpy(id.g, lev.g, sizeof(ostab) * lev.count);
        have_id = 1;
        for (int i = 0; i < lev.count; i++) stab_shift(&lev.g[i], l);   /* value.shift(H.l) */
        cur = lev;
        if (same) break;
    }
    free(id.g);
    K->lv[end] = cur;
    for (int m = end - 1; m >= 0; m--) {
        okl_level_step(K, &K->lv[m + 1], m, &K->lv[m]);
        LOG(m, K->lv[m].count);
    }

    /* driver, py/klocal_periodic.py:21-34 */
    okl_init_states(K);
    for (int i = 0; i < k + 2 * l - 1; i++) okl_evolve(K, 1);
    int ns = K->nstates;
    ostate *states = (ostate *)malloc(sizeof(ostate) * (ns > 0 ? ns : 1));
    for (int i = 0; i < ns; i++) {
        ostate_copy(K, &states[i], &K->states[i]);
        int rc = state_shift(K, &states[i], -l, 1);
        assert(rc == 0); (void)rc;
    }
    qsort(states, ns, sizeof(ostate), key_cmpq);
    ostate *prev = NULL;
    int nprev = -1, fcap = 16;
    R->fixed_sizes = (int *)malloc(sizeof(int) * fcap);
    for (;;) {
        int nn;
        ostate *nst = periodic_evolve_states(K, states, ns, 1, 1, &nn);
        if (R->nfixed == fcap) { fcap *= 2; R->fixed_sizes = (int *)realloc(R->fixed_sizes, sizeof(int) * fcap); }
        R->fixed_sizes[R->nfixed++] = nn;
        int same = (nprev == nn);
        for (int i = 0; same && i < nn; i++) same = key_eq(&prev[i], &nst[i]);
        if (prev) { for (int i = 0; i < nprev; i++) ostate_free(&prev[i]); free(prev); }
        for (int i = 0; i < ns; i++) ostate_free(&states[i]);
        free(states);
        states = nst; ns = nn;
        prev = (ostate *)malloc(sizeof(ostate) * (nn > 0 ? nn : 1));
        for (int i = 0; i < nn; i++) { ostate_copy(K, &prev[i], &nst[i]); ostate_make_key(K, &prev[i]); }
        nprev = nn;
        if (same) break;
    }
    for (int i = 0; i < nprev; i++) ostate_free(&prev[i]);
    free(prev);

    /* orbit search, py/klocal_periodic.py:35-58 */
    int ocap = 16;
    R->orbits = (operiodic_orbit *)malloc(sizeof(operiodic_orbit) * ocap);
    for (int u = 0; u < ns; u++) {
        ostate ns_;
        ostate_copy(K, &ns_, &states[u]);
        ostate_make_key(K, &ns_);
        int nall;
        ostate *all = periodic_evolve_states(K, &ns_, 1, 1, c, &nall);
        int recurs = 0;
        for (int i = 0; i < nall; i++) if (key_eq(&all[i], &ns_)) recurs = 1;
        for (int i = 0; i < nall; i++) ostate_free(&all[i]);
        free(all);
        if (!recurs) { ostate_free(&ns_); continue; }
        size_t sz = (size_t)1 << ns_.st.s.m;
        for (size_t index = 0; index < sz; index++) {
            for (size_t o = 0; o < sz; o++) ns_.st.Es[o] = max_float;
            ns_.st.Es[index] = 0.0;
            all = periodic_evolve_states(K, &ns_, 1, 0, c, &nall);
            for (int i = 0; i < nall; i++) {
                if (!key_eq(&all[i], &ns_)) continue;
                /* EnergyLevel.get_state(index, original_paulis, return_stab=False) (py/stab.py:288-305) */
                const oste *e = &all[i].st;
                if (R->norbits == ocap) { ocap *= 2; R->orbits = (operiodic_orbit *)realloc(R->orbits, sizeof(operiodic_orbit) * ocap); }
                operiodic_orbit *ob = &R->orbits[R->norbits++];
                memset(ob, 0, sizeof(*ob));
                ob->energy = e->Es[index];
                ob->m = e->nPs[index];
                ob->begin = n; ob->end = 0;
                for (int g = 0; g < ob->m; g++) {
                    int ti = e->Pidx[index * e->nmax + g];
                    orow P = H->P[ti];
                    if (e->Psgn[index * e->nmax + g] < 0) P.q = (P.q + 2) & 3;
                    ob->gens[g] = P;
                    if (H->begin[ti] < ob->begin) ob->begin = H->begin[ti];   /* Pauli.stack → common window */
                    if (H->end[ti] > ob->end) ob->end = H->end[ti];
                }
                if (ob->m == 0) { ob->begin = 0; ob->end = 0; }
            }
            for (int i = 0; i < nall; i++) ostate_free(&all[i]);
            free(all);
        }
        ostate_free(&ns_);
    }
    qsort(R->orbits, R->norbits, sizeof(operiodic_orbit), orbit_cmp);
    for (int i = 0; i < ns; i++) ostate_free(&states[i]);
    free(states);
    okl_free(K);
    oham_free(H);
    return R;
}

void operiodic_free(operiodic *r)
{
    if (!r) return;
    free(r->sright_log); free(r->fixed_sizes); free(r->orbits); free(r);
}
