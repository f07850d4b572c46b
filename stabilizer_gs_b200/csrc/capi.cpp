// capi.cpp — the extern "C" boundary declared in include/sgs.h.  No exception crosses it.
#include <chrono>
#include <cstdlib>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <string>

#include "common.hpp"
#include "group_ops.hpp"
#include "hamiltonian.hpp"
#include "klocal_search.hpp"
#include "nccl_glue.hpp"
#include "sparse_search.hpp"

using namespace sgs;

struct sgs_ham { Hamiltonian H; };
struct sgs_sparse_plan { std::unique_ptr<SparsePlan> plan; };
struct sgs_klocal { std::unique_ptr<KlocalMachine> M; };

static int fail(int code, const std::string &msg) { set_last_error(msg); return code; }
static int done(const Status &s) { if (!s.ok()) set_last_error(s.msg); return s.code; }

#define SGS_GUARD_BEGIN try {
#define SGS_GUARD_END                                                                       \
    } catch (const std::bad_alloc &) { return fail(SGS_ERR_LIMIT, "out of host memory"); } \
    catch (const std::exception &e) { return fail(SGS_ERR_ARG, e.what()); }                \
    catch (...) { return fail(SGS_ERR_ARG, "unknown error"); }

extern "C" {

const char *sgs_last_error(void) { return get_last_error(); }
const char *sgs_version(void) { return "stabilizer_gs_b200 0.1 (sm_100a)"; }

void sgs_opts_default(sgs_opts *o)
{
    memset(o, 0, sizeof(*o));
    o->ngpus = 1;
    o->world = 1;
}

int sgs_ham_load(const char *path, int periodic, int cmax, sgs_ham **out)
{
    SGS_GUARD_BEGIN
    if (!path || !out) return fail(SGS_ERR_ARG, "null argument");
    std::unique_ptr<sgs_ham> h(new sgs_ham());
    ParseError err;
    if (!load_hamiltonian_file(path, periodic != 0, cmax, h->H, err)) return fail(err.code, err.msg);
    *out = h.release();
    return SGS_OK;
    SGS_GUARD_END
}

int sgs_ham_from_text(const char *text, int periodic, int cmax, sgs_ham **out)
{
    SGS_GUARD_BEGIN
    if (!text || !out) return fail(SGS_ERR_ARG, "null argument");
    std::unique_ptr<sgs_ham> h(new sgs_ham());
    ParseError err;
    if (!parse_hamiltonian_text(text, periodic != 0, cmax, h->H, err)) return fail(err.code, err.msg);
    *out = h.release();
    return SGS_OK;
    SGS_GUARD_END
}

int sgs_ham_from_arrays(int n, int k, int T, const uint64_t *zx, const double *w, const int *begin, const int *end,
                        sgs_ham **out)
{
    SGS_GUARD_BEGIN
    if (!out || (T > 0 && (!zx || !w))) return fail(SGS_ERR_ARG, "null argument");
    if (n <= 0 || n > (1 << 20) || k <= 0 || T < 0) return fail(SGS_ERR_ARG, "n, k or T out of range");
    std::unique_ptr<sgs_ham> h(new sgs_ham());
    h->H.n = n; h->H.k = k;
    h->H.wide = n + 1 > SGS_MAX_QUBITS;
    const int W = (2 * n + 63) / 64;
    for (int t = 0; t < T; t++) {
        const uint64_t *r = zx + (size_t)t * W;
        std::vector<int> sites, codes;
        for (int wi = 0; wi < W; wi++) {
            if (!r[wi]) continue;
            for (int j = 32 * wi; j < 32 * wi + 32; j++) {
                const int v = (int)((r[wi] >> ((2 * j) & 63)) & 3);
                if (!v) continue;
                if (j >= n) return fail(SGS_ERR_ARG, "term has bits beyond qubit n-1");
                sites.push_back(j); codes.push_back(v);
            }
        }
        int b = sites.empty() ? 0 : sites.front(), e = sites.empty() ? 1 : sites.back() + 1;
        if (begin && end) { b = begin[t]; e = end[t]; }
        if (e <= b) { b = 0; e = 1; }     // identity string: window [0,1)
        if (b < 0 || e > n) return fail(SGS_ERR_ARG, "term window out of range");
        if (!sites.empty() && (b > sites.front() || e < sites.back() + 1)) return fail(SGS_ERR_ARG, "term window does not cover its letters");
        // the window may be wider than the support (Pauli(str, qubits) with explicit I letters): pad with identity sites
        if (sites.empty() || sites.front() != b) { sites.insert(sites.begin(), b); codes.insert(codes.begin(), 0); }
        if (sites.back() != e - 1) { sites.push_back(e - 1); codes.push_back(0); }
        if (!h->H.add_term_sites(sites, codes, w[t])) return fail(SGS_ERR_LIMIT, "with n > 127 a term may span at most 32 sites");
    }
    h->H.finalize();
    *out = h.release();
    return SGS_OK;
    SGS_GUARD_END
}

void sgs_ham_free(sgs_ham *h) { delete h; }
int sgs_ham_n(const sgs_ham *h) { return h ? h->H.n : 0; }
int sgs_ham_k(const sgs_ham *h) { return h ? h->H.k : 0; }
int sgs_ham_l(const sgs_ham *h) { return h ? h->H.l : 0; }
int sgs_ham_nterms(const sgs_ham *h) { return h ? h->H.T() : 0; }
int sgs_ham_words(const sgs_ham *h) { return h ? (2 * h->H.n + 63) / 64 : 0; }

int sgs_ham_terms(const sgs_ham *h, uint64_t *zx, double *w, int *begin, int *end)
{
    if (!h) return fail(SGS_ERR_ARG, "null handle");
    const int W = (2 * h->H.n + 63) / 64;
    for (int t = 0; t < h->H.T(); t++) {
        if (zx) h->H.row_words(t, zx + (size_t)t * W, W);
        if (w) w[t] = h->H.weights[t];
        if (begin) begin[t] = h->H.begin[t];
        if (end) end[t] = h->H.end[t];
    }
    return SGS_OK;
}

int sgs_ham_order(const sgs_ham *h, int *order)
{
    if (!h || !order) return fail(SGS_ERR_ARG, "null argument");
    for (int t = 0; t < h->H.T(); t++) order[t] = h->H.order[t];
    return SGS_OK;
}

void sgs_result_free(sgs_result *r)
{
    if (!r) return;
    free(r->gens); free(r->gen_signs); free(r->term_signs); free(r->sright_counts); free(r->states_per_site);
    r->gens = nullptr; r->gen_signs = r->term_signs = nullptr; r->sright_counts = r->states_per_site = nullptr;
}

int sgs_sparse_plan_create(const sgs_ham *h, const sgs_opts *opts, sgs_sparse_plan **out)
{
    SGS_GUARD_BEGIN
    if (!h || !out) return fail(SGS_ERR_ARG, "null argument");
    sgs_opts o;
    if (opts) o = *opts; else sgs_opts_default(&o);
    std::unique_ptr<sgs_sparse_plan> p(new sgs_sparse_plan());
    Status s = SparsePlan::create(h->H, o, p->plan);
    if (!s.ok()) return done(s);
    *out = p.release();
    return SGS_OK;
    SGS_GUARD_END
}

int sgs_sparse_plan_run(sgs_sparse_plan *p, sgs_result *out)
{
    SGS_GUARD_BEGIN
    if (!p || !out) return fail(SGS_ERR_ARG, "null argument");
    Status s = p->plan->run(out);
    if (!s.ok()) sgs_result_free(out);
    return done(s);
    SGS_GUARD_END
}

void sgs_sparse_plan_free(sgs_sparse_plan *p) { delete p; }

int sgs_sparse_search(const sgs_ham *h, const sgs_opts *opts, sgs_result *out)
{
    sgs_sparse_plan *p = nullptr;
    const auto t0 = std::chrono::steady_clock::now();
    int rc = sgs_sparse_plan_create(h, opts, &p);
    if (rc != SGS_OK) return rc;
    const auto t1 = std::chrono::steady_clock::now();
    rc = sgs_sparse_plan_run(p, out);
    const auto t2 = std::chrono::steady_clock::now();
    sgs_sparse_plan_free(p);
    if (getenv("SGS_TRACE"))
        fprintf(stderr, "sgs_sparse_search: plan %.3f ms, run %.3f ms (device %.3f ms), free %.3f ms\n",
                std::chrono::duration<double>(t1 - t0).count() * 1e3, std::chrono::duration<double>(t2 - t1).count() * 1e3,
                rc == SGS_OK ? out->seconds_device * 1e3 : 0.0,
                std::chrono::duration<double>(std::chrono::steady_clock::now() - t2).count() * 1e3);
    return rc;
}

int sgs_klocal_create(const sgs_ham *h, const sgs_opts *opts, int periodic_cmax, sgs_klocal **out)
{
    SGS_GUARD_BEGIN
    if (!h || !out) return fail(SGS_ERR_ARG, "null argument");
    sgs_opts o;
    if (opts) o = *opts; else sgs_opts_default(&o);
    std::unique_ptr<sgs_klocal> k(new sgs_klocal());
    Status s = KlocalMachine::create(h->H, o, periodic_cmax, k->M);
    if (!s.ok()) return done(s);
    *out = k.release();
    return SGS_OK;
    SGS_GUARD_END
}

void sgs_klocal_free(sgs_klocal *k) { delete k; }
int sgs_klocal_m(const sgs_klocal *k) { return k ? k->M->m() : 0; }
int sgs_klocal_nstates(const sgs_klocal *k) { return k ? k->M->nstates() : 0; }
int sgs_klocal_sright_count(const sgs_klocal *k, int m) { return k ? k->M->sright_count(m) : 0; }

int sgs_klocal_evolve(sgs_klocal *k)
{
    SGS_GUARD_BEGIN
    if (!k) return fail(SGS_ERR_ARG, "null handle");
    return done(k->M->evolve());
    SGS_GUARD_END
}

int sgs_klocal_ground_state(sgs_klocal *k, sgs_result *out)
{
    SGS_GUARD_BEGIN
    if (!k || !out) return fail(SGS_ERR_ARG, "null argument");
    memset(out, 0, sizeof(*out));
    Status s = k->M->ground_state(out);
    if (!s.ok()) sgs_result_free(out);
    return done(s);
    SGS_GUARD_END
}

struct sgs_klocal_states { KlocalStates *S; };

int sgs_klocal_sright_log(const sgs_klocal *k, int *pairs, int cap) { return k ? klocal_sright_log(*k->M, pairs, pairs ? cap : 0) : 0; }

int sgs_klocal_sright_group(sgs_klocal *k, int m, int i, int *rank, int *begin, int *end, uint64_t *rows, int8_t *signs)
{
    SGS_GUARD_BEGIN
    if (!k) return fail(SGS_ERR_ARG, "null handle");
    return done(klocal_sright_group(*k->M, m, i, rank, begin, end, rows, signs));
    SGS_GUARD_END
}

int sgs_klocal_get_states(sgs_klocal *k, sgs_klocal_states **out)
{
    SGS_GUARD_BEGIN
    if (!k || !out) return fail(SGS_ERR_ARG, "null argument");
    KlocalStates *S = nullptr;
    Status s = klocal_get_states(*k->M, &S);
    if (!s.ok()) return done(s);
    *out = new sgs_klocal_states{S};
    return SGS_OK;
    SGS_GUARD_END
}

int sgs_klocal_states_count(const sgs_klocal_states *s) { return s ? klocal_states_count(s->S) : 0; }
void sgs_klocal_states_free(sgs_klocal_states *s) { if (s) { klocal_states_free(s->S); delete s; } }

int sgs_klocal_states_select(const sgs_klocal_states *s, const int *idx, int count, sgs_klocal_states **out)
{
    SGS_GUARD_BEGIN
    if (!s || !out || (count > 0 && !idx)) return fail(SGS_ERR_ARG, "null argument");
    KlocalStates *S = nullptr;
    Status st = klocal_states_select(s->S, idx, count, &S);
    if (!st.ok()) return done(st);
    *out = new sgs_klocal_states{S};
    return SGS_OK;
    SGS_GUARD_END
}

int sgs_klocal_states_append(sgs_klocal_states *dst, const sgs_klocal_states *src, int i)
{
    SGS_GUARD_BEGIN
    if (!dst || !src) return fail(SGS_ERR_ARG, "null argument");
    return done(klocal_states_append(dst->S, src->S, i));
    SGS_GUARD_END
}

int sgs_klocal_state_info(const sgs_klocal *k, const sgs_klocal_states *s, int i, int *m, int *rank, int *begin, int *end, int *sright)
{
    SGS_GUARD_BEGIN
    if (!k || !s) return fail(SGS_ERR_ARG, "null argument");
    return done(klocal_state_info(*k->M, s->S, i, m, rank, begin, end, sright));
    SGS_GUARD_END
}

int sgs_klocal_state_rows(const sgs_klocal *k, const sgs_klocal_states *s, int i, uint64_t *rows)
{
    SGS_GUARD_BEGIN
    if (!k || !s || !rows) return fail(SGS_ERR_ARG, "null argument");
    return done(klocal_state_rows(*k->M, s->S, i, rows));
    SGS_GUARD_END
}

int sgs_klocal_state_key(const sgs_klocal *k, const sgs_klocal_states *s, int i, uint64_t *key, int *nwords)
{
    SGS_GUARD_BEGIN
    if (!k || !s || !key || !nwords) return fail(SGS_ERR_ARG, "null argument");
    return done(klocal_state_key(*k->M, s->S, i, key, SGS_KLOCAL_KEY_WORDS, nwords));
    SGS_GUARD_END
}

int sgs_klocal_state_valid_terms(const sgs_klocal *k, const sgs_klocal_states *s, int i, int *terms, int cap, int *count)
{
    SGS_GUARD_BEGIN
    if (!k || !s || !count || (cap > 0 && !terms)) return fail(SGS_ERR_ARG, "null argument");
    return done(klocal_state_valid_terms(*k->M, s->S, i, terms, cap, count));
    SGS_GUARD_END
}

int sgs_klocal_state_energies(const sgs_klocal_states *s, int i, double *Es)
{
    SGS_GUARD_BEGIN
    if (!s || !Es) return fail(SGS_ERR_ARG, "null argument");
    return done(klocal_state_energies(s->S, i, Es, 0));
    SGS_GUARD_END
}

int sgs_klocal_state_set_energies(sgs_klocal_states *s, int i, const double *Es)
{
    SGS_GUARD_BEGIN
    if (!s || !Es) return fail(SGS_ERR_ARG, "null argument");
    return done(klocal_state_energies(s->S, i, const_cast<double *>(Es), 1));
    SGS_GUARD_END
}

int sgs_klocal_states_shift(sgs_klocal *k, sgs_klocal_states *s, int add, int clear_history)
{
    SGS_GUARD_BEGIN
    if (!k || !s) return fail(SGS_ERR_ARG, "null argument");
    return done(klocal_states_shift(*k->M, s->S, add, clear_history));
    SGS_GUARD_END
}

int sgs_klocal_set_states(sgs_klocal *k, const sgs_klocal_states *s, int m)
{
    SGS_GUARD_BEGIN
    if (!k || !s) return fail(SGS_ERR_ARG, "null argument");
    return done(klocal_set_states(*k->M, s->S, m));
    SGS_GUARD_END
}

int sgs_klocal_state_history(sgs_klocal *k, const sgs_klocal_states *s, int i, int entry, int *terms, int8_t *signs, int cap, int *count)
{
    SGS_GUARD_BEGIN
    if (!k || !s || !count || (cap > 0 && (!terms || !signs))) return fail(SGS_ERR_ARG, "null argument");
    return done(klocal_state_history(*k->M, s->S, i, entry, terms, signs, cap, count));
    SGS_GUARD_END
}

int sgs_klocal_search(const sgs_ham *h, const sgs_opts *opts, sgs_result *out)
{
    SGS_GUARD_BEGIN
    if (!h || !out) return fail(SGS_ERR_ARG, "null argument");
    sgs_opts o;
    if (opts) o = *opts; else sgs_opts_default(&o);
    memset(out, 0, sizeof(*out));
    Status s = klocal_search(h->H, o, out);
    if (!s.ok()) sgs_result_free(out);
    return done(s);
    SGS_GUARD_END
}

int sgs_klocal_periodic(const char *path, int cmax, int c, const sgs_opts *opts, sgs_periodic_result *out)
{
    SGS_GUARD_BEGIN
    if (!path || !out) return fail(SGS_ERR_ARG, "null argument");
    sgs_opts o;
    if (opts) o = *opts; else sgs_opts_default(&o);
    memset(out, 0, sizeof(*out));
    Hamiltonian H;
    ParseError err;
    if (!load_hamiltonian_file(path, true, cmax, H, err)) return fail(err.code, err.msg);
    Status s = klocal_periodic(H, cmax, c, o, out);
    if (!s.ok()) sgs_periodic_result_free(out);
    return done(s);
    SGS_GUARD_END
}

void sgs_periodic_result_free(sgs_periodic_result *r)
{
    if (!r) return;
    for (int i = 0; i < r->norbits; i++) { free(r->orbits[i].gens); free(r->orbits[i].signs); }
    free(r->orbits); free(r->sright_log); free(r->fixed_sizes);
    r->orbits = nullptr; r->sright_log = r->fixed_sizes = nullptr; r->norbits = r->nlog = r->nfixed = 0;
}

int sgs_group_canonicalize(int device, int n, int m, uint64_t *rows, int8_t *signs, int *rank)
{
    SGS_GUARD_BEGIN
    if (!rows || !signs || !rank || n <= 0 || n > SGS_MAX_QUBITS || m < 0) return fail(SGS_ERR_ARG, "invalid argument");
    return done(group_canonicalize(device, n, m, rows, signs, rank));
    SGS_GUARD_END
}

int sgs_group_energy(int device, int n, int m, const uint64_t *rows, const int8_t *signs, int B, const uint64_t *P,
                     const int8_t *Psigns, int8_t *out)
{
    SGS_GUARD_BEGIN
    if ((m > 0 && (!rows || !signs)) || (B > 0 && (!P || !out)) || n <= 0 || n > SGS_MAX_QUBITS) return fail(SGS_ERR_ARG, "invalid argument");
    return done(group_energy(device, n, m, rows, signs, B, P, Psigns, out));
    SGS_GUARD_END
}

int sgs_pauli_commute(int device, int n, int B, const uint64_t *a, const uint64_t *b, uint8_t *out)
{
    SGS_GUARD_BEGIN
    if (B > 0 && (!a || !b || !out)) return fail(SGS_ERR_ARG, "invalid argument");
    return done(pauli_commute_batch(device, n, B, a, b, out));
    SGS_GUARD_END
}

int sgs_nccl_unique_id(void *id_out)
{
    SGS_GUARD_BEGIN
    if (!id_out) return fail(SGS_ERR_ARG, "null argument");
    return done(NcclWorld::unique_id(id_out));
    SGS_GUARD_END
}

int sgs_device_count(void) { return device_count(); }

int sgs_device_flush_l2(int device)
{
    SGS_GUARD_BEGIN
    return done(device_flush_l2(device));
    SGS_GUARD_END
}

int sgs_measure_int_peaks(int device, double *lop3, double *popc, double *sm_mhz)
{
    SGS_GUARD_BEGIN
    return done(measure_int_peaks(device, lop3, popc, sm_mhz));
    SGS_GUARD_END
}

}  // extern "C"
