/*
 * sgs_oracle.h — CPU ORACLE for the exact stabilizer-ground-state search.
 *
 * TEST INFRASTRUCTURE ONLY.  This is a plain-C restatement of the reference's
 * algorithm (SUSYUSTC/stabilizer_gs: cpp/linear.cpp, cpp/stab.cpp, cpp/state.cpp and the
 * Python-only periodic path in py/state.py + py/klocal_periodic.py) on packed
 * bit rows.  It exists to CHECK the CUDA product path and to be timed as the CPU
 * baseline ("port") by bench.py.  Nothing under stabilizer_gs_b200/ may include,
 * link or call it; only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs do.
 *
 * Pinning: every function family is checked against golden vectors produced by
 * the UNMODIFIED Python reference (tests/golden/make_golden.py → tests/golden/
 * *.json; see tests/test_oracle_*.py).  The reference's C++ cannot be built here
 * (Eigen ≥3.4 and Boost headers are absent), so the Python reference is the pin.
 *
 * Packed layout: bit 2j = z_j, bit 2j+1 = x_j, so "lowest set
 * bit" is the reference's pivot column in `Pauli::array()` order
 * (cpp/stab.cpp:194-199).  Operator = i^q * prod_j Z_j^{z_j} X_j^{x_j},
 * q = ZX_phase in Z_4 (cpp/stab.cpp:176-192).
 */
#ifndef SGS_ORACLE_H
#define SGS_ORACLE_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#ifndef OW
#define OW 4              /* u64 words per row: up to 128 qubit positions (make wide: OW = 12, 384 positions) */
#endif
#define OMAXQ (OW * 32)   /* positions; n (plus the DP's one-site overhang) < OMAXQ   */
#define OMAXM OMAXQ       /* max generators of a group                               */
#define OMAXTAB 26        /* max rank for which a 2^m energy table is materialised   */

typedef struct { uint64_t b[OW]; int q; } orow;              /* one Pauli string       */
#define OMW ((OMAXM + 63) / 64)
typedef struct { uint64_t w[OMW]; } omask;                   /* subset of <= OMAXM rows */
typedef struct { int m, begin, end; orow g[OMAXM]; } ostab;  /* canonical signed group */

/* ---- Hamiltonian (cpp/state.cpp:58-96, py/state.py:34-72) ---- */
typedef struct {
    int n, k, l;          /* l = period for periodic files, else 0                      */
    int T;                /* number of terms                                            */
    orow *P;              /* [T] terms in FILE order (original_paulis)                  */
    double *w;            /* [T]                                                        */
    int *begin, *end;     /* [T] window of each term = [min site, max site + 1)         */
    int *order;           /* [T] paulis_list order: by last site, then file order       */
    int *bucket_off;      /* [n+1] offsets into order[] per last site                   */
} oham;

oham *oham_load(const char *path, int periodic, int cmax, char *err, int errlen);
oham *oham_from_text(const char *text, int periodic, int cmax, char *err, int errlen);
void oham_free(oham *H);

/* ---- result of a search ---- */
typedef struct {
    double energy;
    int n, m;                 /* qubits, generators                                   */
    orow gens[OMAXM];         /* canonical signed generators on [0,n)                 */
    int8_t *term_signs;       /* [T] file order: -1, 0, +1  (cpp/sparse.cpp:13-16)    */
    int T;
    /* sparse statistics */
    uint64_t candidates;      /* leaves of the include/exclude tree                   */
    uint64_t rank_hist[OMAXM + 1];
    double sum_2m;
    /* k-local statistics */
    int nsites;
    int *sright_counts;       /* [n+1]                                                */
    int *states_per_site;     /* [n+1] (index 0 = initial)                            */
    uint64_t transitions;
    double seconds;
} ores;
void ores_free(ores *r);

/* sparse search: FullState::evolve (cpp/state.cpp:636-687).  threads<=1 → serial like the reference. */
ores *oracle_sparse(const oham *H, int threads);
/* same walk, stopped after max_leaves leaves or `seconds` (bench.py's bounded CPU sample) */
ores *oracle_sparse_sample(const oham *H, int threads, uint64_t max_leaves, double seconds);
/* uniform sample: walks, completely, the subtrees below every stride-th (hashed) pair of first two included terms */
ores *oracle_sparse_stride(const oham *H, int threads, int stride, int offset);
/* k-local DP: StateMachine (cpp/state.cpp:445-550) + generate_Sright_all (:218-238) */
ores *oracle_klocal(const oham *H, int threads);

/* periodic: py/klocal_periodic.py:1-63 */
typedef struct { double energy; int begin, end, m; orow gens[OMAXM]; } operiodic_orbit;
typedef struct {
    int l, k, n, nterms;
    int nlog; int *sright_log;      /* pairs (m, count) printed by generate_Sright_periodic */
    int nfixed; int *fixed_sizes;   /* sizes printed while the state set converges         */
    int norbits; operiodic_orbit *orbits;
} operiodic;
operiodic *oracle_periodic(const char *path, int cmax, int c, char *err, int errlen);
void operiodic_free(operiodic *r);

/* ---- small pieces exported for unit pinning ---- */
int  orow_from_string(orow *r, const char *signed_letters);   /* "+XIZY" over sites 0..L-1 */
void orow_to_string(const orow *r, int begin, int end, char *out); /* sign + letters       */
int  orow_commute(const orow *a, const orow *b);               /* cpp/stab.cpp:314-318    */
void orow_dot(orow *out, const orow *a, const orow *b);        /* cpp/stab.cpp:320-330    */
int  orow_sign(const orow *r);                                 /* +1 / -1                 */
void ostab_init(ostab *s, int begin, int end);
void ostab_canonicalize(ostab *s, omask *U);                   /* cpp/stab.cpp:437-449    */
void ostab_add(ostab *s, const orow *P);                       /* cpp/stab.cpp:432-435    */
int  ostab_decompose(const ostab *s, const orow *P, omask *x, int *sign); /* :493-505     */
int  ostab_commute(const ostab *s, const orow *P);             /* cpp/stab.cpp:513-522    */
void ostab_project(const ostab *s, int begin, int end, ostab *out); /* cpp/stab.cpp:524-577 */

/* dense mod-N linear algebra, int entries, row-major (cpp/linear.cpp) — for the test_linear KATs */
void olin_inv_list(int N, int *out);
void olin_get_order(const int *A, int n, int m, int *order);
void olin_gauss(int *A, int *b, int n, int m, int kb, int N, int allow_reorder);
int  olin_solve(const int *A, const int *b, int n, int m, int kb, int N, int *is_over, int *x);
int  olin_kernel(const int *A, int n, int m, int N, int *ker /* m x (m-rank), row-major, capacity m*m */);
void olin_simplify_coset(const int *A2, int n, int m, const int *x, int N, int *out);

#ifdef __cplusplus
}
#endif
#endif
