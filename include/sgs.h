/*
 * sgs.h — C ABI of libsgs.so, the B200-native exact stabilizer-ground-state search.
 *
 * The reference (SUSYUSTC/stabilizer_gs) has NO FFI today: cpp/ and py/ are two independent
 * implementations.  This header is the boundary a maintainer binds instead:
 * every entry point names the reference interface it replaces (file:line under the reference
 * root).  Plain C types only; no C++ / CUDA / torch types cross it.
 *
 * Conventions
 *   - Packed Pauli rows: W = ceil(2n/64) little-endian u64 words per row, bit 2j = z_j,
 *     bit 2j+1 = x_j (the reference's Pauli::array() column order, cpp/stab.cpp:194-199).
 *   - Signs are +1 / -1 (int8_t); a term sign of 0 means "not in the group"
 *     (Stab::energy, cpp/stab.cpp:579-587).
 *   - Every function returns 0 on success or a negative SGS_ERR_* code; the message is
 *     available from sgs_last_error() (thread-local).  No exception crosses the boundary.
 *   - Ownership: the caller owns its inputs; result buffers are allocated by the library and
 *     released with the matching *_free.  Handles are not thread-safe.
 *   - There is NO CPU fallback: a search fails with SGS_ERR_NO_DEVICE when no CUDA device
 *     is usable.
 */
#ifndef SGS_H
#define SGS_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SGS_OK 0
#define SGS_ERR_IO (-1)          /* cannot read the Hamiltonian file                         */
#define SGS_ERR_PARSE (-2)       /* malformed file (bad letter, letters/sites mismatch, ...) */
#define SGS_ERR_ARG (-3)         /* invalid argument                                         */
#define SGS_ERR_NO_DEVICE (-4)   /* no usable CUDA device (the product has no CPU path)      */
#define SGS_ERR_CUDA (-5)        /* CUDA runtime error                                       */
#define SGS_ERR_LIMIT (-6)       /* instance exceeds a compiled limit (n, T, rank)           */
#define SGS_ERR_NCCL (-7)        /* NCCL error                                               */
#define SGS_ERR_OVERFLOW (-8)    /* a device work queue overflowed                           */
#define SGS_ERR_DEAD (-9)        /* the DP ended with no state                               */

#define SGS_MAX_QUBITS 128       /* rows are at most 4 u64 words                             */
#define SGS_MAX_RANK 64          /* rank histogram size - 1                                  */
#define SGS_NCCL_ID_BYTES 128

typedef struct sgs_ham sgs_ham;          /* opaque Hamiltonian handle                        */
typedef struct sgs_klocal sgs_klocal;    /* opaque k-local DP machine (StateMachine)         */

/* ---------------------------------------------------------------- Hamiltonian ---------- */

/* Hamiltonian::from_file (cpp/state.cpp:58-81, py/state.py:34-49); periodic != 0 →
 * Hamiltonian.from_file_periodic(path, cmax) (py/state.py:51-72). */
int sgs_ham_load(const char *path, int periodic, int cmax, sgs_ham **out);
/* same parser on an in-memory text (what from_file reads) */
int sgs_ham_from_text(const char *text, int periodic, int cmax, sgs_ham **out);
/* Hamiltonian(n,k) + add_Pauli (cpp/state.cpp:37-45,83-86) from packed arrays: zx is T x W,
 * begin/end give each term's window [min site, max site + 1) (NULL → derived from the bits). */
int sgs_ham_from_arrays(int n, int k, int T, const uint64_t *zx, const double *w,
                        const int *begin, const int *end, sgs_ham **out);
void sgs_ham_free(sgs_ham *h);
int sgs_ham_n(const sgs_ham *h);
int sgs_ham_k(const sgs_ham *h);
int sgs_ham_l(const sgs_ham *h);         /* period of a periodic file, else 0                */
int sgs_ham_nterms(const sgs_ham *h);
int sgs_ham_words(const sgs_ham *h);     /* W                                                */
/* original_paulis (file order): copies T x W words, T weights, T begins, T ends (any may be NULL) */
int sgs_ham_terms(const sgs_ham *h, uint64_t *zx, double *w, int *begin, int *end);
/* paulis_list order (cpp/state.cpp:88-96): order[i] = file index of the i-th term */
int sgs_ham_order(const sgs_ham *h, int *order);

/* ---------------------------------------------------------------- options / result ----- */

typedef struct {
    int device;              /* CUDA device ordinal of this process (default 0)                */
    int ngpus;               /* single-process multi-GPU: use devices device..device+ngpus-1   */
    int rank, world;         /* multi-process sharding: this process is shard `rank` of `world`
                                (world <= 1 → no sharding).  Combined with ngpus: global shard
                                = rank * ngpus + local gpu.                                    */
    const void *nccl_id;     /* SGS_NCCL_ID_BYTES bytes from sgs_nccl_unique_id(), shared by all
                                ranks, or NULL: with world > 1 and NULL the result is this
                                shard's only (no collective).                                  */
    int verbose;             /* klocal: -v level (cpp/klocal_sparse.cpp:10)                    */
    int expand_depth;        /* sparse: force the prefix depth handed to the DFS kernel; 0=auto */
    int prune;               /* sparse: 1 = branch and bound: subtrees that cannot reach the best
                                energy found so far are skipped.  Same energy, generators and term signs as the
                                exhaustive search (ties included); `candidates` then counts the visited groups only */
    int reserved[7];
} sgs_opts;
void sgs_opts_default(sgs_opts *o);

typedef struct {
    double energy;               /* ground-state energy                                       */
    int n, m, W, T;
    uint64_t *gens;              /* m x W canonical generator rows (RREF, sorted by pivot)     */
    int8_t *gen_signs;           /* m signs                                                    */
    int8_t *term_signs;          /* T, file order (cpp/sparse.cpp:13-16)                       */
    /* sparse search statistics */
    uint64_t candidates;         /* closed commuting groups evaluated (= leaves of FullState's tree) */
    uint64_t rank_hist[SGS_MAX_RANK + 1];
    double sum_2m;               /* sum over candidates of 2^rank (signed candidates)          */
    uint64_t slow_nodes;         /* candidates that needed the general (dependent-term) path   */
    /* k-local statistics */
    int nsites;
    int *sright_counts;          /* n+1: "m count" lines of generate_Sright_all                */
    int *states_per_site;        /* n+1: index 0 = initial states                              */
    uint64_t transitions;        /* evolve_single calls                                        */
    /* timing (seconds) */
    double seconds_total;        /* wall, call entry → result on host                          */
    double seconds_device;       /* CUDA-event time of the search kernels (max over local GPUs) */
    double seconds_hot_kernel;   /* CUDA-event time of the dominant kernel                     */
    uint32_t kernel_launches;    /* launches of this library's kernels during the call         */
    /* sparse search diagnostics (the parity tests assert that every path of the walk was exercised):
     * [0] certified frames walked with 32-bit clique masks, [1] with 64-bit masks, [2] multi-word independence
     * proven on the folded 64-bit image, [3] decided by the full-width elimination, [4] dependent residues
     * certified by the clique test, [5] frames not certified, [6] warp-cooperative child frames (level A),
     * [7] certified frames of more than 64 entries;
     * saturating; summed over shards when a communicator combines them */
    uint32_t walk_paths[8];
} sgs_result;
void sgs_result_free(sgs_result *r);     /* frees the buffers inside *r (not r itself)        */

/* ---------------------------------------------------------------- searches ------------- */

/* FullState(paulis).evolve() (cpp/state.cpp:555-564,636-687; py/state.py:88-147) + the
 * per-term signs of cpp/sparse.cpp:13-16.  Runs entirely on the GPU(s). */
/* With opts.world > 1 and no nccl_id the call returns the result of shard opts.rank alone (counts add up over
 * the ranks, the minimum over the ranks is the global answer); a shard that owns no candidate returns
 * energy = +inf and m = 0. */
int sgs_sparse_search(const sgs_ham *h, const sgs_opts *opts, sgs_result *out);

/* Device-resident variant for repeated searches: upload once, search many times. */
typedef struct sgs_sparse_plan sgs_sparse_plan;
int sgs_sparse_plan_create(const sgs_ham *h, const sgs_opts *opts, sgs_sparse_plan **out);
int sgs_sparse_plan_run(sgs_sparse_plan *p, sgs_result *out);
void sgs_sparse_plan_free(sgs_sparse_plan *p);

/* StateMachine(H).evolve() until m == n, generate_gs(), get_stab_gs()
 * (cpp/klocal_sparse.cpp:26-75, cpp/state.cpp:445-550).
 * opts.ngpus > 1 (or world > 1 with an nccl_id): the multi-GPU DP — the counterpart of the reference's omp-parallel
 * state loop (cpp/state.cpp:494-517).  Every GPU holds the site's states; the move phase of a site and the all-pairs
 * key ranking are shared out, the results exchanged with NCCL (all-gather / all-reduce per site), the merge is
 * replicated: bit-identical to the one-GPU result.  Pays from ~10^4-10^5 states per site on. */
int sgs_klocal_search(const sgs_ham *h, const sgs_opts *opts, sgs_result *out);

/* Finer-grained DP handle for the periodic driver (py/klocal_periodic.py):
 * generate_Sright_all / generate_Sright_periodic + StateMachine / StateMachinePeriodic. */
/* Site steps are enqueued on the device without a host round trip; errors met on the device (a limit, or a batch
 * that outgrew the handle's buffers: SGS_ERR_OVERFLOW) surface at the next call that synchronises — nstates
 * (which then returns -1), ground_state, or the next evolve.  sgs_klocal_search / sgs_klocal_periodic redo the
 * search with larger buffers by themselves; with this stepwise handle the caller sees the error. */
int sgs_klocal_create(const sgs_ham *h, const sgs_opts *opts, int periodic_cmax, sgs_klocal **out);
void sgs_klocal_free(sgs_klocal *k);
int sgs_klocal_m(const sgs_klocal *k);                 /* StateMachine.m                          */
int sgs_klocal_nstates(const sgs_klocal *k);           /* len(state_dict)                         */
int sgs_klocal_sright_count(const sgs_klocal *k, int m);
int sgs_klocal_evolve(sgs_klocal *k);                  /* StateMachine::evolve, one site          */
int sgs_klocal_ground_state(sgs_klocal *k, sgs_result *out);   /* generate_gs + get_stab_gs      */

/* ---- stepwise access to the DP states: the Python classes the BODY of py/klocal_periodic.py:9-63 uses
 * (State.shift/copy/__hash__, StateMachinePeriodic.state_dict/.m, EnergyLevel.Es/get_state; py/state.py:150-325,
 * 402-462, py/stab.py:176-305).  With periodic_cmax >= 0, sgs_klocal_create runs generate_Sright_periodic
 * (py/state.py:542-567) and sets the StateMachinePeriodic's initial states. */
typedef struct sgs_klocal_states sgs_klocal_states;       /* host snapshot of a list of states of one site       */
#define SGS_KLOCAL_KEY_WORDS 32
/* (m, count) pairs in the order generate_Sright_all / generate_Sright_periodic print them; returns their number */
int sgs_klocal_sright_log(const sgs_klocal *k, int *pairs, int cap);
/* the i-th right environment of level m (std::get<0>(Sright_all.at(m)), cpp/state.cpp:218-238; printed by
 * klocal_sparse -v 1, cpp/klocal_sparse.cpp:28-37): signed canonical generators on [begin, end); rows = rank x W */
int sgs_klocal_sright_group(sgs_klocal *k, int m, int i, int *rank, int *begin, int *end, uint64_t *rows, int8_t *signs);
/* State::get_valid_Ps(valid_range()) (cpp/state.cpp:293-309): file indices of the terms still valid for the state */
int sgs_klocal_state_valid_terms(const sgs_klocal *k, const sgs_klocal_states *s, int i, int *terms, int cap, int *count);
/* list(state_dict.values()) of the current site, in key order */
int sgs_klocal_get_states(sgs_klocal *k, sgs_klocal_states **out);
int sgs_klocal_states_count(const sgs_klocal_states *s);
void sgs_klocal_states_free(sgs_klocal_states *s);
/* a new list holding copies of entries idx[0..count) (State.copy, list building) */
int sgs_klocal_states_select(const sgs_klocal_states *s, const int *idx, int count, sgs_klocal_states **out);
/* dst.append(copy of src[i]); every state of a list sits at the same site */
int sgs_klocal_states_append(sgs_klocal_states *dst, const sgs_klocal_states *src, int i);
/* State.m, rank and window [begin, end) of State.stab, index of State.Sright in its level (any pointer may be NULL) */
int sgs_klocal_state_info(const sgs_klocal *k, const sgs_klocal_states *s, int i, int *m, int *rank, int *begin, int *end, int *sright);
/* generators of State.stab: rank x W rows at absolute sites, all '+' (cpp/stab.cpp:626) */
int sgs_klocal_state_rows(const sgs_klocal *k, const sgs_klocal_states *s, int i, uint64_t *rows);
/* content of State.__hash__ (py/state.py:224-225): m, window group, (in)valid term indices of the next k-2 sites,
 * right environment; key has room for SGS_KLOCAL_KEY_WORDS words */
int sgs_klocal_state_key(const sgs_klocal *k, const sgs_klocal_states *s, int i, uint64_t *key, int *nwords);
/* State.stab.energy_level.Es: 2^rank doubles, flat index bit (rank-1-i) set <=> generator i is '-' (cpp/stab.cpp:37-45) */
int sgs_klocal_state_energies(const sgs_klocal_states *s, int i, double *Es);
int sgs_klocal_state_set_energies(sgs_klocal_states *s, int i, const double *Es);
/* State.shift(add, clear_history) applied to every state of the list (py/state.py:309-325) */
int sgs_klocal_states_shift(sgs_klocal *k, sgs_klocal_states *s, int add, int clear_history);
/* state_machine.state_dict = {hash(s): s for s in list}; state_machine.m = m (py/klocal_periodic.py:12-14) */
int sgs_klocal_set_states(sgs_klocal *k, const sgs_klocal_states *s, int m);
/* EnergyLevel.get_state(entry, original_paulis, return_stab=False) (py/stab.py:288-305): file indices and signs of the
 * Hamiltonian terms generating the best full state behind table entry `entry`, in history order.  Walks the device
 * history, which is valid until the machine's next evolve / set_states (SGS_ERR_ARG afterwards). */
int sgs_klocal_state_history(sgs_klocal *k, const sgs_klocal_states *s, int i, int entry, int *terms, int8_t *signs, int cap, int *count);

/* py/klocal_periodic.py:1-63 as one call: orbits of the infinite periodic chain */
typedef struct {
    double energy;          /* energy over c periods                                          */
    int begin, end, m;      /* window and number of generators                                */
    uint64_t *gens;         /* m x W rows (history order, NOT canonicalised: return_stab=False) */
    int8_t *signs;          /* m                                                               */
} sgs_orbit;
typedef struct {
    int l, k, n, W, nterms;
    int nlog; int *sright_log;        /* (m, count) pairs printed by generate_Sright_periodic     */
    int nfixed; int *fixed_sizes;     /* sizes printed while the state set converges              */
    int norbits; sgs_orbit *orbits;   /* sorted by (energy, begin, end, generators)               */
    double seconds_total;
} sgs_periodic_result;
int sgs_klocal_periodic(const char *path, int cmax, int c, const sgs_opts *opts, sgs_periodic_result *out);
void sgs_periodic_result_free(sgs_periodic_result *r);

/* ---------------------------------------------------------------- group utilities ------ */
/* The Python classes the reference scripts touch (py/stab.py StabilizerGroup) are thin shims
 * over these; all of them run as CUDA kernels on `device`. */

/* StabilizerGroup(Ps) canonicalisation (py/stab.py:45-51, cpp/stab.cpp:437-449): rows/signs are
 * overwritten with the canonical generators; *rank receives their number. */
int sgs_group_canonicalize(int device, int n, int m, uint64_t *rows, int8_t *signs, int *rank);
/* StabilizerGroup.get_energy(P) / Stab::energy for a batch of B Paulis (py/stab.py:147-153):
 * out[i] = 0 if P_i is not in ±S, else the relative sign. */
int sgs_group_energy(int device, int n, int m, const uint64_t *rows, const int8_t *signs,
                     int B, const uint64_t *P, const int8_t *Psigns, int8_t *out);
/* Pauli::commute for B pairs (cpp/stab.cpp:314-318): out[i] = 1 if a_i and b_i commute */
int sgs_pauli_commute(int device, int n, int B, const uint64_t *a, const uint64_t *b, uint8_t *out);

/* ---------------------------------------------------------------- multi-GPU ------------ */
int sgs_nccl_unique_id(void *id_out /* SGS_NCCL_ID_BYTES */);
int sgs_device_count(void);

/* ---------------------------------------------------------------- misc ----------------- */
const char *sgs_last_error(void);
const char *sgs_version(void);
/* integer-pipe issue-rate microbenchmarks (roofline denominators):
 * 32-bit LOP3 and POPC ops per second over the whole chip at the running clock */
int sgs_measure_int_peaks(int device, double *lop3_ops_per_s, double *popc_ops_per_s, double *sm_mhz);
/* benchmark hygiene: overwrite a buffer larger than the L2 (evicts the term tables between timed steps) */
int sgs_device_flush_l2(int device);

#ifdef __cplusplus
}
#endif
#endif
