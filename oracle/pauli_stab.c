/*
 * oracle/pauli_stab.c — CPU ORACLE (test infrastructure only; see sgs_oracle.h).
 * Packed-bit restatement of the reference's Pauli and Stab classes
 * (cpp/stab.cpp:80-381 and :388-603; py/pauli.py, py/stab.py:8-173).
 * Rows are absolute-position (site j at bits 2j,2j+1), so the reference's
 * re-windowing helpers range/extend/concatenate (cpp/stab.cpp:238-312) reduce
 * to bookkeeping of [begin,end) plus masking on truncation.
 */
#include <assert.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "oracle_internal.h"

/* ------------------------------------------------------------------ rows */

int orow_is_zero(const orow *r)
{
    uint64_t a = 0;
    for (int i = 0; i < OW; i++) a |= r->b[i];
    return a == 0;
}

/* first non-zero column in array() order [z0,x0,z1,x1,...] (cpp/linear.cpp:50); -1 for identity */
int orow_lowbit(const orow *r)
{
    for (int i = 0; i < OW; i++)
        if (r->b[i]) return i * 64 + __builtin_ctzll(r->b[i]);
    return -1;
}

int orow_bit(const orow *r, int c) { return (int)((r->b[c >> 6] >> (c & 63)) & 1); }

static int orow_nY(const orow *r)   /* cpp/stab.cpp:172-174 */
{
    int c = 0;
    for (int i = 0; i < OW; i++) c += __builtin_popcountll(r->b[i] & (r->b[i] >> 1) & OEVEN);
    return c;
}

/* printed sign: group_phase = (ZX_phase + nY) mod 4 must be 0 (+) or 2 (-)  (cpp/stab.cpp:176-192) */
int orow_sign(const orow *r)
{
    int gp = (r->q + orow_nY(r)) & 3;
    assert((gp & 1) == 0);
    return gp == 0 ? 1 : -1;
}

void orow_set_sign(orow *r, int sign)   /* set_group_phase(1 - sign), cpp/stab.cpp:180-182 */
{
    r->q = ((1 - sign) - orow_nY(r)) & 3;
}

/* symplectic form: sum_j z1_j x2_j + z2_j x1_j even  (cpp/stab.cpp:314-318) */
int orow_commute(const orow *a, const orow *b)
{
    int c = 0;
    for (int i = 0; i < OW; i++) {
        uint64_t sw = ((b->b[i] >> 1) & OEVEN) | ((b->b[i] & OEVEN) << 1);
        c += __builtin_popcountll(a->b[i] & sw);
    }
    return (c & 1) == 0;
}

/* product with phase: q = q1 + q2 + 2 * (x1 . z2)  (cpp/stab.cpp:320-330) */
void orow_dot(orow *out, const orow *a, const orow *b)
{
    int c = 0;
    for (int i = 0; i < OW; i++) c += __builtin_popcountll((a->b[i] >> 1) & b->b[i] & OEVEN);
    int q = (a->q + b->q + 2 * c) & 3;
    for (int i = 0; i < OW; i++) out->b[i] = a->b[i] ^ b->b[i];
    out->q = q;
}

void orow_mask_window(orow *r, int begin, int end)   /* range(..., allow_truncation=true): keep bits in [begin,end) */
{
    int sgn = orow_sign(r);
    for (int j = 0; j < OMAXQ; j++)
        if (j < begin || j >= end) r->b[(2 * j) >> 6] &= ~(3ULL << ((2 * j) & 63));
    orow_set_sign(r, sgn);    /* Pauli(z,x,begin,phase) keeps the printed sign (cpp/stab.cpp:238-254) */
}

int orow_support_outside(const orow *r, int begin, int end)
{
    for (int j = 0; j < OMAXQ; j++) {
        if (j >= begin && j < end) continue;
        if ((r->b[(2 * j) >> 6] >> ((2 * j) & 63)) & 3) return 1;
    }
    return 0;
}

int orow_site(const orow *r, int j) { return (int)((r->b[(2 * j) >> 6] >> ((2 * j) & 63)) & 3); }  /* bit0=z bit1=x */

/* "+XIZY": letters on sites 0..L-1 (cpp/stab.cpp:116-144 letter map) */
int orow_from_string(orow *r, const char *s)
{
    memset(r, 0, sizeof(*r));
    int sign = 1;
    if (*s == '+') s++;
    else if (*s == '-') { sign = -1; s++; }
    int L = (int)strlen(s);
    if (L > OMAXQ) return -1;
    for (int j = 0; j < L; j++) {
        uint64_t v;
        switch (s[j]) {
        case 'I': v = 0; break;
        case 'Z': v = 1; break;
        case 'X': v = 2; break;
        case 'Y': v = 3; break;
        default: return -1;
        }
        r->b[(2 * j) >> 6] |= v << ((2 * j) & 63);
    }
    orow_set_sign(r, sign);
    return L;
}

/* cpp/stab.cpp:347-364: sign then one of IXZY (array4 = 2z + x) per site */
void orow_to_string(const orow *r, int begin, int end, char *out)
{
    *out++ = orow_sign(r) == 1 ? '+' : '-';
    for (int j = begin; j < end; j++) {
        int v = orow_site(r, j);   /* bit0 = z, bit1 = x */
        *out++ = "IZXY"[v];
    }
    *out = 0;
}

/* ------------------------------------------------------------------ groups */

static void mask_set(omask *m, int i) { m->w[i >> 6] |= 1ULL << (i & 63); }
static int mask_get(const omask *m, int i) { return (int)((m->w[i >> 6] >> (i & 63)) & 1); }
static void mask_xor(omask *a, const omask *b) { for (int i = 0; i < OMW; i++) a->w[i] ^= b->w[i]; }

void ostab_init(ostab *s, int begin, int end) { s->m = 0; s->begin = begin; s->end = end; }

/*
 * Canonical form = RREF of the generator matrix in [z0,x0,z1,x1,...] column order, rows sorted by
 * pivot, each new generator the phase-tracked product of the old ones its U row selects
 * (Gauss_elimination_full cpp/linear.cpp:129 → :21-75; Stab::transform cpp/stab.cpp:442-449).
 * The elimination order is the reference's: forward pass k = 0..m-1 (pivot = first non-zero column of
 * row k, clear it below), backward pass k = m-1..0 (clear it above), then the stable sort.
 * U (optional) receives, per output row, the subset of input rows multiplied together.
 */
void ostab_canonicalize(ostab *s, omask *U)
{
    int m = s->m;
    omask u[OMAXM];
    int piv[OMAXM];
    memset(u, 0, sizeof(omask) * (m > 0 ? m : 1));
    for (int i = 0; i < m; i++) mask_set(&u[i], i);
    for (int k = 0; k < m; k++) {
        int p = orow_lowbit(&s->g[k]);
        piv[k] = p;
        if (p < 0) continue;
        for (int i = k + 1; i < m; i++)
            if (orow_bit(&s->g[i], p)) { orow_dot(&s->g[i], &s->g[i], &s->g[k]); mask_xor(&u[i], &u[k]); }
    }
    for (int k = m - 1; k >= 0; k--) {
        int p = piv[k];
        if (p < 0) continue;
        for (int i = 0; i < k; i++)
            if (orow_bit(&s->g[i], p)) { orow_dot(&s->g[i], &s->g[i], &s->g[k]); mask_xor(&u[i], &u[k]); }
    }
    /* get_order (cpp/linear.cpp:4-19): by first non-zero column, zero rows last, stable */
    int order[OMAXM], key[OMAXM];
    for (int i = 0; i < m; i++) { int p = orow_lowbit(&s->g[i]); key[i] = p < 0 ? 2 * OMAXQ : p; order[i] = i; }
    for (int i = 1; i < m; i++) {
        int o = order[i], j = i - 1;
        while (j >= 0 && key[order[j]] > key[o]) { order[j + 1] = order[j]; j--; }
        order[j + 1] = o;
    }
    orow tmp[OMAXM];
    omask tu[OMAXM];
    for (int i = 0; i < m; i++) { tmp[i] = s->g[order[i]]; tu[i] = u[order[i]]; }
    memcpy(s->g, tmp, sizeof(orow) * m);
    if (U) memcpy(U, tu, sizeof(omask) * m);
}

/* Stab::add (cpp/stab.cpp:432-435): append, re-canonicalise; window grows to cover P */
void ostab_add(ostab *s, const orow *P)
{
    assert(s->m < OMAXM);
    s->g[s->m++] = *P;
    ostab_canonicalize(s, NULL);
}

/*
 * Stab::decompose (cpp/stab.cpp:493-505): is ±P in S, and if so which generators (x) and with which
 * relative sign = P.phase * product.phase.  S is canonical, so x_i = P[pivot_i] (what solve_modN,
 * cpp/linear.cpp:77-121, returns for an RREF system) and membership ⇔ the product reproduces P's bits.
 */
int ostab_decompose(const ostab *s, const orow *P, omask *x, int *sign)
{
    orow prod;
    memset(&prod, 0, sizeof(prod));       /* identity, q = 0: Pauli::reduce_dot folds from I (cpp/stab.cpp:340-345) */
    omask xm;
    memset(&xm, 0, sizeof(xm));
    for (int i = 0; i < s->m; i++) {
        int p = orow_lowbit(&s->g[i]);
        if (p >= 0 && orow_bit(P, p)) { orow_dot(&prod, &prod, &s->g[i]); mask_set(&xm, i); }
    }
    for (int i = 0; i < OW; i++) if (prod.b[i] != P->b[i]) return 0;
    if (x) *x = xm;
    if (sign) *sign = orow_sign(P) * orow_sign(&prod);
    return 1;
}

int ostab_include(const ostab *s, const orow *P) { return ostab_decompose(s, P, NULL, NULL); }  /* :507-511 */

/* Stab::energy (cpp/stab.cpp:579-587): 0 if not in the group, else the relative sign */
int ostab_energy(const ostab *s, const orow *P)
{
    int sg;
    return ostab_decompose(s, P, NULL, &sg) ? sg : 0;
}

int ostab_commute(const ostab *s, const orow *P)   /* cpp/stab.cpp:513-522 */
{
    for (int i = 0; i < s->m; i++) if (!orow_commute(&s->g[i], P)) return 0;
    return 1;
}

/*
 * Stab::project(begin,end) (cpp/stab.cpp:524-577): the subgroup acting trivially outside the window.
 * The reference takes the GF(2) kernel of the outside-coordinate matrix (linear::get_kernel), multiplies
 * the selected generators, truncates and canonicalises.  Canonical form is unique, so any basis of the
 * kernel gives the same result; here: eliminate on outside columns only, keep the rows left with no
 * outside support.
 */
void ostab_project(const ostab *s, int begin, int end, ostab *out)
{
    ostab t = *s;
    int m = t.m, r = 0;
    for (int j = 0; j < OMAXQ; j++) {
        if (j >= begin && j < end) continue;
        for (int bit = 0; bit < 2; bit++) {
            int c = 2 * j + bit, pr = -1;
            for (int i = r; i < m; i++) if (orow_bit(&t.g[i], c)) { pr = i; break; }
            if (pr < 0) continue;
            orow sw = t.g[pr]; t.g[pr] = t.g[r]; t.g[r] = sw;
            for (int i = 0; i < m; i++)
                if (i != r && orow_bit(&t.g[i], c)) orow_dot(&t.g[i], &t.g[i], &t.g[r]);
            r++;
        }
    }
    ostab_init(out, begin, end);
    for (int i = r; i < m; i++) {
        assert(!orow_support_outside(&t.g[i], begin, end));
        out->g[out->m++] = t.g[i];
    }
    ostab_canonicalize(out, NULL);
}

/* is_in_group (cpp/stab.cpp:856-862): P in the GF(2) span of rows (signs ignored) */
int orows_span_contains(const orow *rows, int cnt, const orow *P)
{
    orow *basis = (orow *)malloc(sizeof(orow) * (cnt > 0 ? cnt : 1));
    int *piv = (int *)malloc(sizeof(int) * (cnt > 0 ? cnt : 1));
    int r = 0;
    for (int i = 0; i < cnt; i++) {
        orow v = rows[i];
        for (int j = 0; j < r; j++) if (orow_bit(&v, piv[j])) for (int w = 0; w < OW; w++) v.b[w] ^= basis[j].b[w];
        int p = orow_lowbit(&v);
        if (p < 0) continue;
        basis[r] = v; piv[r] = p; r++;
    }
    orow v = *P;
    for (int j = 0; j < r; j++) if (orow_bit(&v, piv[j])) for (int w = 0; w < OW; w++) v.b[w] ^= basis[j].b[w];
    int res = orow_is_zero(&v);
    free(basis); free(piv);
    return res;
}

int ostab_equal(const ostab *a, const ostab *b)
{
    if (a->m != b->m || a->begin != b->begin || a->end != b->end) return 0;
    for (int i = 0; i < a->m; i++) {
        if (memcmp(a->g[i].b, b->g[i].b, sizeof(a->g[i].b))) return 0;
        if (orow_sign(&a->g[i]) != orow_sign(&b->g[i])) return 0;
    }
    return 1;
}

/* total order on canonical groups used wherever the reference iterates a std::map keyed by a
 * boost hash (cpp/state.cpp:196-199); the reference's order is not reproducible. */
int ostab_cmp(const ostab *a, const ostab *b)
{
    if (a->m != b->m) return a->m < b->m ? -1 : 1;
    for (int i = 0; i < a->m; i++) {
        for (int w = OW - 1; w >= 0; w--)
            if (a->g[i].b[w] != b->g[i].b[w]) return a->g[i].b[w] < b->g[i].b[w] ? -1 : 1;
        int sa = orow_sign(&a->g[i]), sb = orow_sign(&b->g[i]);
        if (sa != sb) return sa > sb ? -1 : 1;
    }
    return 0;
}
