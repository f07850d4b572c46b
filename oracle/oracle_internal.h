/* oracle/oracle_internal.h — CPU ORACLE internals (test infrastructure only; see sgs_oracle.h). */
#ifndef ORACLE_INTERNAL_H
#define ORACLE_INTERNAL_H
#include "sgs_oracle.h"

#define OEVEN 0x5555555555555555ULL

/* pauli_stab.c */
int  orow_is_zero(const orow *r);
int  orow_lowbit(const orow *r);
int  orow_bit(const orow *r, int c);
void orow_set_sign(orow *r, int sign);
void orow_mask_window(orow *r, int begin, int end);
int  orow_support_outside(const orow *r, int begin, int end);
int  orow_site(const orow *r, int j);
int  ostab_include(const ostab *s, const orow *P);
int  ostab_energy(const ostab *s, const orow *P);
int  orows_span_contains(const orow *rows, int cnt, const orow *P);
int  ostab_equal(const ostab *a, const ostab *b);
int  ostab_cmp(const ostab *a, const ostab *b);

/*
 * StabEnergies (cpp/stab.hpp:154-196): a canonical group whose generators all carry sign '+', the
 * energy of every sign pattern, and per-pattern history of the Hamiltonian terms (index, sign) that
 * generate the best full state reaching that pattern.  Flat index bit (m-1-i) set ⇔ generator i is '-'
 * (cpp/stab.cpp:37-45,59-66).
 */
typedef struct {
    ostab s;
    double *Es;       /* [2^m] */
    int nmax;         /* history capacity per pattern; 0 = no history (sparse search) */
    int *nPs;         /* [2^m] */
    int *Pidx;        /* [2^m * nmax] */
    int *Psgn;        /* [2^m * nmax] */
} oste;

void oste_init(oste *e, int begin, int end, int nmax);             /* StabEnergies::empty cpp/stab.cpp:644-653 */
void oste_copy(oste *dst, const oste *src);
void oste_free(oste *e);
void oste_add(oste *e, const orow *P, int canon, int index);       /* cpp/stab.cpp:747-761 */
void oste_add_energy(oste *e, const orow *P, double w);            /* cpp/stab.cpp:819-825 */
void oste_project_to(oste *e, int k1);                             /* cpp/stab.cpp:782-809 */
void oste_merge_gs(oste *dst, const oste *other);                  /* cpp/stab.cpp:827-843: dst = merge(dst, other) */
void oste_clear_history(oste *e);                                  /* py/stab.py:282-286 */
int  oste_argmin(const oste *e);

#endif
