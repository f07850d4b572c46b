/* oracle/klocal_internal.h — CPU ORACLE internals for the k-local DP (test infrastructure only). */
#ifndef KLOCAL_INTERNAL_H
#define KLOCAL_INTERNAL_H
#include "oracle_internal.h"

/* type_Sright (cpp/state.hpp:49-51): the groups of one level and, per group, the indices of the
 * next level's groups that evolve backward into it */
typedef struct { int count; ostab *g; int *map_off; int *map_idx; } olevel;

/* State (cpp/state.hpp:53-98) */
typedef struct {
    int m;
    oste st;            /* projected left group on the window + energy table + history */
    uint64_t *valid;    /* valid_pauli_indices as a bitset over positions in paulis_list order */
    int sright;         /* index into level m's group list */
    unsigned char *key; int keylen;
} ostate;

/* StateMachine (cpp/state.hpp:100-121) */
typedef struct {
    const oham *H;
    int n, k, T, nmax, vwords;
    orow *P; double *w;          /* terms in paulis_list order */
    olevel *lv;                  /* lv[m], m = 0..n */
    ostate *states; int nstates; int m;
    uint64_t transitions;
} okl;

okl *okl_new(const oham *H, int nmax_hist);
void okl_free(okl *K);
void okl_generate_sright_all(okl *K);
void okl_level_step(const okl *K, const olevel *next, int m, olevel *out);
void olevel_free(olevel *l);
void okl_init_states(okl *K);
void okl_evolve(okl *K, int threads);
int  okl_evolve_single(const okl *K, const ostate *raw, ostate **out, int *nout, int *cap);
void ostate_init(const okl *K, ostate *s, int sright);
void ostate_copy(const okl *K, ostate *d, const ostate *s);
void ostate_free(ostate *s);
void ostate_make_key(const okl *K, ostate *s);
void okl_vshift(const okl *K, ostate *s, int add);

#endif
