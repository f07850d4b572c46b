/*
 * oracle/energies.c — CPU ORACLE (test infrastructure only; see sgs_oracle.h).
 * Restatement of StabEnergies (cpp/stab.cpp:609-843; py/stab.py:176-429): a canonical
 * all-'+' generator set plus the energy of each of the 2^m sign patterns and, per pattern,
 * the history (Hamiltonian term index, sign) of the generators of the best full state.
 */
#include <assert.h>
#include <stdlib.h>
#include <string.h>
#include "oracle_internal.h"

static size_t tabsize(int m) { return (size_t)1 << m; }

void oste_init(oste *e, int begin, int end, int nmax)
{
    ostab_init(&e->s, begin, end);
    e->nmax = nmax;
    e->Es = (double *)calloc(1, sizeof(double));
    e->nPs = e->Pidx = e->Psgn = NULL;
    if (nmax > 0) {
        e->nPs = (int *)calloc(1, sizeof(int));
        e->Pidx = (int *)malloc(sizeof(int) * nmax);
        e->Psgn = (int *)calloc(nmax, sizeof(int));
        for (int i = 0; i < nmax; i++) e->Pidx[i] = -1;       /* py/stab.py:190 */
    }
}

void oste_free(oste *e)
{
    free(e->Es); free(e->nPs); free(e->Pidx); free(e->Psgn);
    e->Es = NULL; e->nPs = e->Pidx = e->Psgn = NULL;
}

void oste_copy(oste *dst, const oste *src)
{
    size_t sz = tabsize(src->s.m);
    dst->s = src->s;
    dst->nmax = src->nmax;
    dst->Es = (double *)malloc(sizeof(double) * sz);
    memcpy(dst->Es, src->Es, sizeof(double) * sz);
    dst->nPs = dst->Pidx = dst->Psgn = NULL;
    if (src->nmax > 0) {
        dst->nPs = (int *)malloc(sizeof(int) * sz);
        dst->Pidx = (int *)malloc(sizeof(int) * sz * src->nmax);
        dst->Psgn = (int *)malloc(sizeof(int) * sz * src->nmax);
        memcpy(dst->nPs, src->nPs, sizeof(int) * sz);
        memcpy(dst->Pidx, src->Pidx, sizeof(int) * sz * src->nmax);
        memcpy(dst->Psgn, src->Psgn, sizeof(int) * sz * src->nmax);
    }
}

/* apply a permutation of table slots: new[perm[o]] = old[o] */
static void permute_tables(oste *e, const size_t *perm)
{
    size_t sz = tabsize(e->s.m);
    double *Es = (double *)malloc(sizeof(double) * sz);
    for (size_t o = 0; o < sz; o++) Es[perm[o]] = e->Es[o];
    free(e->Es); e->Es = Es;
    if (e->nmax > 0) {
        int nm = e->nmax;
        int *nPs = (int *)malloc(sizeof(int) * sz);
        int *Pi = (int *)malloc(sizeof(int) * sz * nm), *Ps = (int *)malloc(sizeof(int) * sz * nm);
        for (size_t o = 0; o < sz; o++) {
            nPs[perm[o]] = e->nPs[o];
            memcpy(Pi + perm[o] * nm, e->Pidx + o * nm, sizeof(int) * nm);
            memcpy(Ps + perm[o] * nm, e->Psgn + o * nm, sizeof(int) * nm);
        }
        free(e->nPs); free(e->Pidx); free(e->Psgn);
        e->nPs = nPs; e->Pidx = Pi; e->Psgn = Ps;
    }
}

/* StabEnergies::canonicalize (cpp/stab.cpp:693-697) = transform(U) (:699-733) + canonicalize_signs (:685-691) */
static void oste_canonicalize(oste *e)
{
    int m = e->s.m;
    assert(m <= OMAXTAB);
    omask U[OMAXM];
    ostab_canonicalize(&e->s, U);
    size_t sz = tabsize(m);
    size_t *perm = (size_t *)malloc(sizeof(size_t) * sz);
    /* the pattern with sign vector s (s_j = generator j negative) moves to s' = U s (cpp/stab.cpp:721-732) */
    for (size_t o = 0; o < sz; o++) {
        uint64_t smask = 0;
        for (int j = 0; j < m; j++) if ((o >> (m - 1 - j)) & 1) smask |= 1ULL << j;
        size_t no = 0;
        for (int i = 0; i < m; i++)
            if (__builtin_popcountll(U[i].w[0] & smask) & 1) no |= (size_t)1 << (m - 1 - i);
        perm[o] = no;
    }
    permute_tables(e, perm);
    /* a generator whose product came out '-' gets its axis reversed so every stored sign is '+' (:667-691) */
    for (int i = 0; i < m; i++) {
        if (orow_sign(&e->s.g[i]) == -1) {
            e->s.g[i].q = (e->s.g[i].q + 2) & 3;
            size_t bit = (size_t)1 << (m - 1 - i);
            for (size_t o = 0; o < sz; o++) perm[o] = o ^ bit;
            permute_tables(e, perm);
        }
    }
    free(perm);
}

/* StabEnergies::add (cpp/stab.cpp:747-761): the new generator becomes the LAST axis (new = 2*old + b,
 * b = 1 ⇔ '-'); history slot nPs := (index, +1 / -1); then canonicalise. */
void oste_add(oste *e, const orow *P, int canon, int index)
{
    int m = e->s.m;
    assert(m + 1 <= OMAXTAB);
    size_t sz = tabsize(m);
    double *Es = (double *)malloc(sizeof(double) * sz * 2);
    for (size_t o = 0; o < sz; o++) Es[2 * o] = Es[2 * o + 1] = e->Es[o];
    free(e->Es); e->Es = Es;
    if (e->nmax > 0) {
        int nm = e->nmax;
        int *nPs = (int *)malloc(sizeof(int) * sz * 2);
        int *Pi = (int *)malloc(sizeof(int) * sz * 2 * nm), *Ps = (int *)malloc(sizeof(int) * sz * 2 * nm);
        for (size_t o = 0; o < sz; o++) {
            int c = e->nPs[o];
            assert(c < nm);
            for (int b = 0; b < 2; b++) {
                memcpy(Pi + (2 * o + b) * nm, e->Pidx + o * nm, sizeof(int) * nm);
                memcpy(Ps + (2 * o + b) * nm, e->Psgn + o * nm, sizeof(int) * nm);
                Pi[(2 * o + b) * nm + c] = index;
                Ps[(2 * o + b) * nm + c] = b ? -1 : 1;
                nPs[2 * o + b] = c + 1;
            }
        }
        free(e->nPs); free(e->Pidx); free(e->Psgn);
        e->nPs = nPs; e->Pidx = Pi; e->Psgn = Ps;
    }
    e->s.g[e->s.m++] = *P;
    if (canon) oste_canonicalize(e);
}

/* StabEnergies::add_energy (cpp/stab.cpp:819-825): Es[idx] += w * sign * (-1)^{x . s(idx)} */
void oste_add_energy(oste *e, const orow *P, double w)
{
    omask x;
    int sign, m = e->s.m;
    int ok = ostab_decompose(&e->s, P, &x, &sign);
    assert(ok); (void)ok;
    size_t xi = 0;
    for (int i = 0; i < m; i++) if ((x.w[0] >> i) & 1) xi |= (size_t)1 << (m - 1 - i);
    size_t sz = tabsize(m);
    for (size_t o = 0; o < sz; o++)
        e->Es[o] += (__builtin_popcountll(o & xi) & 1) ? -(w * sign) : (w * sign);
}

/* StabEnergies::project_to_next (cpp/stab.cpp:782-801): drop the leftmost qubit of the window.  While a
 * generator acts on it, remove generator 0 and keep per remaining pattern the lower half
 * (keep1 = Es1 < Es2 strictly; a tie keeps the '-' half); history follows. */
static void oste_project_to_next(oste *e)
{
    for (;;) {
        int any = 0;
        for (int i = 0; i < e->s.m; i++) if (orow_site(&e->s.g[i], e->s.begin)) any = 1;
        if (!any) break;
        int m = e->s.m;
        size_t half = tabsize(m - 1);
        for (size_t o = 0; o < half; o++) {
            if (!(e->Es[o] < e->Es[o + half])) {           /* take the second half */
                e->Es[o] = e->Es[o + half];
                if (e->nmax > 0) {
                    e->nPs[o] = e->nPs[o + half];
                    memcpy(e->Pidx + o * e->nmax, e->Pidx + (o + half) * e->nmax, sizeof(int) * e->nmax);
                    memcpy(e->Psgn + o * e->nmax, e->Psgn + (o + half) * e->nmax, sizeof(int) * e->nmax);
                }
            }
        }
        memmove(&e->s.g[0], &e->s.g[1], sizeof(orow) * (m - 1));
        e->s.m = m - 1;
    }
    e->s.begin += 1;      /* range(begin + 1, end, allow_truncation): the dropped column is already empty */
}

void oste_project_to(oste *e, int k1)   /* cpp/stab.cpp:803-809 */
{
    while (e->s.begin < e->s.end - k1) oste_project_to_next(e);
}

/* StabEnergies::merge_gs (cpp/stab.cpp:827-843): element-wise, keep dst where dst < other strictly */
void oste_merge_gs(oste *dst, const oste *other)
{
    assert(dst->s.m == other->s.m);
    size_t sz = tabsize(dst->s.m);
    for (size_t o = 0; o < sz; o++) {
        if (!(dst->Es[o] < other->Es[o])) {
            dst->Es[o] = other->Es[o];
            if (dst->nmax > 0) {
                dst->nPs[o] = other->nPs[o];
                memcpy(dst->Pidx + o * dst->nmax, other->Pidx + o * dst->nmax, sizeof(int) * dst->nmax);
                memcpy(dst->Psgn + o * dst->nmax, other->Psgn + o * dst->nmax, sizeof(int) * dst->nmax);
            }
        }
    }
}

void oste_clear_history(oste *e)   /* EnergyLevel.clear (py/stab.py:282-286): Es, nPs := 0 too */
{
    size_t sz = tabsize(e->s.m);
    for (size_t o = 0; o < sz; o++) e->Es[o] = 0.0;
    if (e->nmax > 0) {
        memset(e->nPs, 0, sizeof(int) * sz);
        for (size_t i = 0; i < sz * e->nmax; i++) { e->Pidx[i] = -1; e->Psgn[i] = 0; }
    }
}

int oste_argmin(const oste *e)     /* first minimum (argmin, cpp/linear.hpp:160) */
{
    size_t sz = tabsize(e->s.m), best = 0;
    for (size_t o = 1; o < sz; o++) if (e->Es[o] < e->Es[best]) best = o;
    return (int)best;
}
