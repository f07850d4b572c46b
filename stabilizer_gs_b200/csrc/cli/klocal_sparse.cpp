// cli/klocal_sparse.cpp — drop-in for `./klocal_sparse [-v N] [-n N] <file>` (cpp/klocal_sparse.cpp:7-76).
// Same stdout lines; timing values differ by nature.  -n/--next is parsed and ignored, as the reference
// stores it in StateMachine::nstates and never reads it (cpp/state.cpp:449).
//
// -v 1 dumps what cpp/klocal_sparse.cpp:28-37,47-60 dumps — every right environment of every level, then after
// each site every state (State::to_string, cpp/state.cpp:332-345: m, window group, right environment, valid
// terms, energy table, per-entry history as the Ps_indices / Ps_signs matrices in Eigen's print layout) — with
// two documented differences: states come in this library's key order (the reference iterates a std::map keyed by
// a boost hash, an order that changes with the Boost version), and -v 2 prints no hash values (they are boost
// hash_combine values of Eigen matrices; nothing here computes them).  The stderr progress line of
// cpp/state.cpp:514,518 is written once per state as well.
#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <iomanip>
#include <iostream>
#include <sstream>
#include <string>
#include <vector>

#include "sgs.h"

static int g_n = 0;       // sites >= n (the one-site overhang of the last right environment) print as I
static std::string letters(const uint64_t *row, int begin, int end)
{
    std::string s;
    for (int j = begin; j < end; j++) s += j < g_n ? "IZXY"[(row[(2 * j) >> 6] >> ((2 * j) & 63)) & 3] : 'I';
    return s;
}

static std::string row_string(const uint64_t *row, int sign, int n) { return std::string(1, sign >= 0 ? '+' : '-') + letters(row, 0, n); }

// Pauli::to_string(true) (cpp/stab.cpp:347-364): "qubit b-e size m " then "+XZ " per string
static std::string pauli_block(const std::vector<uint64_t> &rows, const std::vector<int8_t> &signs, int m, int W, int begin, int end)
{
    std::string s = "qubit " + std::to_string(begin) + "-" + std::to_string(end - 1) + " size " + std::to_string(m) + " ";
    for (int i = 0; i < m; i++) s += std::string(1, signs[i] >= 0 ? '+' : '-') + letters(rows.data() + (size_t)i * W, begin, end) + " ";
    return s;
}

// Eigen's default operator<< for a matrix: every coefficient right-aligned to the widest one, columns separated by
// one space, rows by newlines
template <typename T>
static std::string eigen_matrix(const std::vector<std::vector<T>> &M, int precision = 6)
{
    std::vector<std::vector<std::string>> cells;
    size_t width = 0;
    for (auto &r : M) {
        cells.emplace_back();
        for (auto &v : r) {
            std::ostringstream o;
            o << std::setprecision(precision) << v;
            cells.back().push_back(o.str());
            width = std::max(width, o.str().size());
        }
    }
    std::string out;
    for (size_t i = 0; i < cells.size(); i++) {
        for (size_t j = 0; j < cells[i].size(); j++) {
            if (j) out += " ";
            out += std::string(width - cells[i][j].size(), ' ') + cells[i][j];
        }
        if (i + 1 < cells.size()) out += "\n";
    }
    return out;
}

#define CHECK(call) do { if ((call) != SGS_OK) { fprintf(stderr, "error: %s\n", sgs_last_error()); return 1; } } while (0)

int main(int argc, char **argv)
{
    std::string path;
    sgs_opts opts;
    sgs_opts_default(&opts);
    // boost::program_options forms (cpp/klocal_sparse.cpp:8-19): -v N, -vN, --verbose N, --verbose=N (same for -n / --next),
    // --filename F, --filename=F, or the file as the positional argument
    for (int i = 1; i < argc; i++) {
        std::string a = argv[i];
        auto value = [&](size_t skip, bool attached) -> std::string { return attached ? a.substr(skip) : (i + 1 < argc ? std::string(argv[++i]) : std::string()); };
        if (a == "-v" || a == "--verbose") opts.verbose = atoi(value(0, false).c_str());
        else if (a.rfind("--verbose=", 0) == 0) opts.verbose = atoi(value(10, true).c_str());
        else if (a.rfind("-v", 0) == 0 && a.size() > 2 && a[1] == 'v') opts.verbose = atoi(value(2, true).c_str());
        else if (a == "-n" || a == "--next") value(0, false);
        else if (a.rfind("--next=", 0) == 0) {}
        else if (a.rfind("-n", 0) == 0 && a.size() > 2 && a[1] == 'n') {}
        else if (a == "--device" && i + 1 < argc) opts.device = atoi(argv[++i]);
        else if (a == "--filename") path = value(0, false);
        else if (a.rfind("--filename=", 0) == 0) path = value(11, true);
        else path = a;
    }
    if (path.empty()) { fprintf(stderr, "the option '--filename' is required but missing\n"); return 1; }
    const int verbose = opts.verbose;
    std::cout << "verbose level " << verbose << std::endl;
    sgs_ham *H = nullptr;
    CHECK(sgs_ham_load(path.c_str(), 0, 0, &H));
    const int n = sgs_ham_n(H), T = sgs_ham_nterms(H), W = sgs_ham_words(H);
    g_n = n;
    std::vector<uint64_t> trow((size_t)std::max(T, 1) * W);
    std::vector<int> tbeg(std::max(T, 1)), tend(std::max(T, 1)), order(std::max(T, 1));
    CHECK(sgs_ham_terms(H, trow.data(), nullptr, tbeg.data(), tend.data()));
    CHECK(sgs_ham_order(H, order.data()));
    // history index of a term as the reference encodes it (cpp/state.cpp:393): i * n + m, i = position in the bucket of last site m
    std::vector<int> hist_code(std::max(T, 1), 0);
    {
        std::vector<int> seen(n + 1, 0);
        for (int p = 0; p < T; p++) { const int f = order[p], m = tend[f] - 1; hist_code[f] = seen[m]++ * n + m; }
    }
    sgs_klocal *K = nullptr;
    CHECK(sgs_klocal_create(H, &opts, -1, &K));
    for (int m = n - 1; m >= 0; m--) std::cout << m << ' ' << sgs_klocal_sright_count(K, m) << std::endl;   // cpp/state.cpp:235
    std::vector<uint64_t> rows((size_t)64 * W);
    std::vector<int8_t> signs(64);
    if (verbose > 0) {                                          // cpp/klocal_sparse.cpp:28-37
        for (int m = 0; m < n; m++) {
            std::cout << m << std::endl;
            for (int i = 0; i < sgs_klocal_sright_count(K, m); i++) {
                int r, b, e;
                CHECK(sgs_klocal_sright_group(K, m, i, &r, &b, &e, rows.data(), signs.data()));
                std::cout << pauli_block(rows, signs, r, W, b, e) << std::endl;
            }
            std::cout << std::endl;
        }
    }
    int nprev = sgs_klocal_nstates(K);
    while (true) {
        auto t0 = std::chrono::high_resolution_clock::now();
        CHECK(sgs_klocal_evolve(K));
        const int ns = sgs_klocal_nstates(K);                   // (synchronises: the site step is done)
        auto t1 = std::chrono::high_resolution_clock::now();
        for (int i = 1; i <= nprev; i++) std::cerr << "progress: " << i << " in " << nprev << "\r" << std::flush;   // cpp/state.cpp:514
        std::cerr << std::endl;
        const double sec = std::chrono::duration_cast<std::chrono::microseconds>(t1 - t0).count() * 1e-6;
        std::cout << "evolve time " << sec << std::endl;      // cpp/state.cpp:519-520 (device time of the transition batch)
        std::cout << "update time " << 0 << std::endl;
        std::cout << "m " << sgs_klocal_m(K) << std::endl;
        std::cout << "number of states " << ns << std::endl;
        std::cout << "time " << sec << std::endl;
        std::cout << std::endl;
        nprev = ns;
        if (verbose > 0) {                                      // cpp/klocal_sparse.cpp:47-60 + State::to_string cpp/state.cpp:332-345
            sgs_klocal_states *S = nullptr;
            CHECK(sgs_klocal_get_states(K, &S));
            for (int s = 0; s < sgs_klocal_states_count(S); s++) {
                int m, r, b, e, sr;
                CHECK(sgs_klocal_state_info(K, S, s, &m, &r, &b, &e, &sr));
                std::cout << s << std::endl;
                std::cout << "m = " << m << "\n";
                CHECK(sgs_klocal_state_rows(K, S, s, rows.data()));
                std::fill(signs.begin(), signs.end(), (int8_t)1);
                std::cout << "stab = " << pauli_block(rows, signs, r, W, b, e) << "\n";
                int rr, rb, re;
                CHECK(sgs_klocal_sright_group(K, m, sr, &rr, &rb, &re, rows.data(), signs.data()));
                std::cout << "Sright = " << pauli_block(rows, signs, rr, W, rb, re) << "\n";
                std::vector<int> vt(std::max(T, 1));
                int nv = 0;
                CHECK(sgs_klocal_state_valid_terms(K, S, s, vt.data(), T, &nv));
                std::cout << "valid_pauli_indices = ";
                if (nv) {                                       // Pauli::concatenate: all strings on the union of their windows
                    int cb = n, ce = 0;
                    for (int i = 0; i < nv; i++) { cb = std::min(cb, tbeg[vt[i]]); ce = std::max(ce, tend[vt[i]]); }
                    std::cout << "qubit " << cb << "-" << ce - 1 << " size " << nv << " ";
                    for (int i = 0; i < nv; i++) std::cout << "+" << letters(trow.data() + (size_t)vt[i] * W, cb, ce) << " ";
                }
                std::cout << "\n";
                std::vector<double> Es((size_t)1 << r);
                CHECK(sgs_klocal_state_energies(S, s, Es.data()));
                std::cout << "energies = " << eigen_matrix<double>({Es}) << std::endl;
                // Ps_indices / Ps_signs: one row per table entry, n columns (nmax = n, cpp/state.cpp:251-262), zero padded
                std::vector<std::vector<int>> Pi, Ps;
                std::vector<int> ht(std::max(T, 1));
                std::vector<int8_t> hs(std::max(T, 1));
                for (int en = 0; en < (1 << r); en++) {
                    int cnt = 0;
                    CHECK(sgs_klocal_state_history(K, S, s, en, ht.data(), hs.data(), T, &cnt));
                    std::vector<int> a(n, 0), c(n, 0);
                    for (int i = 0; i < cnt && i < n; i++) { a[i] = hist_code[ht[i]]; c[i] = hs[i]; }
                    Pi.push_back(a); Ps.push_back(c);
                }
                std::cout << "Ps indices \n" << eigen_matrix<int>(Pi) << std::endl;
                std::cout << "Ps signs \n" << eigen_matrix<int>(Ps) << std::endl;
                std::cout << '\n' << std::endl;                 // to_string ends with '\n', then the caller's endl
                std::cout << std::endl;                         // (the hash line of -v 2 would precede this endl: not reproduced)
            }
            sgs_klocal_states_free(S);
        }
        if (sgs_klocal_m(K) == n) break;
    }
    sgs_result r;
    CHECK(sgs_klocal_ground_state(K, &r));
    std::cout << "ground state stabilizers" << std::endl;
    std::cout << "qubit 0-" << n - 1 << " size " << r.m << "\n";
    for (int i = 0; i < r.m; i++) std::cout << row_string(r.gens + (size_t)i * r.W, r.gen_signs[i], n) << " \n";
    std::cout << std::endl;
    std::cout << std::setprecision(16) << "ground state energy " << r.energy << std::endl;
    std::cout << "terms signs" << std::endl;
    for (int t = 0; t < r.T; t++) std::cout << (int)r.term_signs[t] << " ";
    std::cout << std::endl;
    sgs_result_free(&r);
    sgs_klocal_free(K);
    sgs_ham_free(H);
    return 0;
}
