/*
 * oracle/ham.c — CPU ORACLE (test infrastructure only; see sgs_oracle.h).
 * Hamiltonian file parser: Hamiltonian::from_file (cpp/state.cpp:58-81, py/state.py:34-49),
 * Pauli(str, qubits) (cpp/stab.cpp:116-144), add_Pauli (cpp/state.cpp:83-86),
 * paulis_list order (cpp/state.cpp:88-96) and from_file_periodic (py/state.py:51-72).
 */
#include <ctype.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "oracle_internal.h"

static void seterr(char *err, int errlen, const char *msg)
{
    if (err && errlen > 0) { strncpy(err, msg, errlen - 1); err[errlen - 1] = 0; }
}

static int push_term(oham *H, int *cap, const orow *P, double w, int begin, int end)
{
    if (H->T == *cap) {
        *cap = *cap ? *cap * 2 : 64;
        H->P = (orow *)realloc(H->P, sizeof(orow) * *cap);
        H->w = (double *)realloc(H->w, sizeof(double) * *cap);
        H->begin = (int *)realloc(H->begin, sizeof(int) * *cap);
        H->end = (int *)realloc(H->end, sizeof(int) * *cap);
    }
    H->P[H->T] = *P; H->w[H->T] = w; H->begin[H->T] = begin; H->end[H->T] = end;
    H->T++;
    return 0;
}

oham *oham_from_text(const char *text, int periodic, int cmax, char *err, int errlen)
{
    oham *H = (oham *)calloc(1, sizeof(oham));
    int cap = 0;
    const char *p = text;
    int a = 0, b = 0, consumed = 0;
    if (sscanf(p, "%d %d%n", &a, &b, &consumed) != 2) { seterr(err, errlen, "bad header"); oham_free(H); return NULL; }
    while (*p && *p != '\n') p++;
    if (*p == '\n') p++;
    if (periodic) { H->l = a; H->k = b; H->n = a * (1 + cmax) + b * 2; }   /* py/state.py:55-56 */
    else { H->n = a; H->k = b; H->l = 0; }
    if (H->n + 1 >= OMAXQ || H->n <= 0) { seterr(err, errlen, "n out of range for the oracle"); oham_free(H); return NULL; }
    while (*p) {
        const char *e = p;
        while (*e && *e != '\n') e++;
        char line[4096];
        size_t len = (size_t)(e - p);
        if (len >= sizeof(line)) { seterr(err, errlen, "line too long"); oham_free(H); return NULL; }
        memcpy(line, p, len); line[len] = 0;
        p = *e ? e + 1 : e;
        char *s = line;
        while (*s && isspace((unsigned char)*s)) s++;
        if (!*s) break;                                    /* Python stops at the first blank line (py/state.py:41-43) */
        char *endp;
        double w = strtod(s, &endp);
        if (endp == s) { seterr(err, errlen, "bad weight"); oham_free(H); return NULL; }
        s = endp;
        while (*s && isspace((unsigned char)*s)) s++;
        char letters[OMAXQ + 1];
        int nl = 0;
        while (*s && !isspace((unsigned char)*s)) { if (nl >= OMAXQ) { seterr(err, errlen, "too many letters"); oham_free(H); return NULL; } letters[nl++] = *s++; }
        letters[nl] = 0;
        int sites[OMAXQ], ns = 0;
        for (;;) {
            while (*s && isspace((unsigned char)*s)) s++;
            if (!*s) break;
            long v = strtol(s, &endp, 10);
            if (endp == s) { seterr(err, errlen, "bad site"); oham_free(H); return NULL; }
            if (ns >= OMAXQ) { seterr(err, errlen, "too many sites"); oham_free(H); return NULL; }
            sites[ns++] = (int)v; s = endp;
        }
        if (ns != nl || ns == 0) { seterr(err, errlen, "letters/sites mismatch"); oham_free(H); return NULL; }
        int qmin = sites[0], qmax = sites[0];
        for (int i = 1; i < ns; i++) { if (sites[i] < qmin) qmin = sites[i]; if (sites[i] > qmax) qmax = sites[i]; }
        int lo = periodic ? -H->n : 0, hi = periodic ? H->n : 0;
        for (int sh = lo; sh <= hi; sh++) {               /* py/state.py:69-71: every shift i*l that fits in [0,n) */
            int off = periodic ? sh * H->l : 0;
            if (qmax + off >= H->n || qmin + off < 0) { if (periodic) continue; seterr(err, errlen, "site out of range"); oham_free(H); return NULL; }
            orow P;
            memset(&P, 0, sizeof(P));
            for (int i = 0; i < ns; i++) {
                uint64_t v;
                switch (letters[i]) {
                case 'I': v = 0; break;
                case 'Z': v = 1; break;
                case 'X': v = 2; break;
                case 'Y': v = 3; break;
                default: seterr(err, errlen, "invalid character"); oham_free(H); return NULL;   /* cpp/stab.cpp:138-140 */
                }
                int j = sites[i] + off;
                P.b[(2 * j) >> 6] &= ~(3ULL << ((2 * j) & 63));
                P.b[(2 * j) >> 6] |= v << ((2 * j) & 63);
            }
            orow_set_sign(&P, 1);
            push_term(H, &cap, &P, w, qmin + off, qmax + off + 1);
        }
    }
    /* paulis_list order: buckets by last site, file order inside a bucket (cpp/state.cpp:83-96) */
    H->order = (int *)malloc(sizeof(int) * (H->T > 0 ? H->T : 1));
    H->bucket_off = (int *)calloc(H->n + 1, sizeof(int));
    int c = 0;
    for (int q = 0; q < H->n; q++) {
        H->bucket_off[q] = c;
        for (int t = 0; t < H->T; t++) if (H->end[t] - 1 == q) H->order[c++] = t;
    }
    H->bucket_off[H->n] = c;
    return H;
}

oham *oham_load(const char *path, int periodic, int cmax, char *err, int errlen)
{
    FILE *f = fopen(path, "rb");
    if (!f) { seterr(err, errlen, "cannot open file"); return NULL; }
    fseek(f, 0, SEEK_END);
    long sz = ftell(f);
    fseek(f, 0, SEEK_SET);
    char *buf = (char *)malloc(sz + 1);
    if (fread(buf, 1, sz, f) != (size_t)sz) { fclose(f); free(buf); seterr(err, errlen, "read error"); return NULL; }
    buf[sz] = 0;
    fclose(f);
    oham *H = oham_from_text(buf, periodic, cmax, err, errlen);
    free(buf);
    return H;
}

void oham_free(oham *H)
{
    if (!H) return;
    free(H->P); free(H->w); free(H->begin); free(H->end); free(H->order); free(H->bucket_off);
    free(H);
}

void ores_free(ores *r)
{
    if (!r) return;
    free(r->term_signs); free(r->sright_counts); free(r->states_per_site);
    free(r);
}
