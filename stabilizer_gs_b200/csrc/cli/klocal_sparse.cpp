// cli/klocal_sparse.cpp — drop-in for `./klocal_sparse [-v N] [-n N] <file>` (cpp/klocal_sparse.cpp:7-76).
// Same stdout lines; timing values differ by nature.  -n/--next is parsed and ignored, as the reference
// stores it in StateMachine::nstates and never reads it (cpp/state.cpp:449).
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <iomanip>
#include <iostream>
#include <string>

#include "sgs.h"

static std::string row_string(const uint64_t *row, int sign, int n)
{
    std::string s(1, sign >= 0 ? '+' : '-');
    for (int j = 0; j < n; j++) s += "IZXY"[(row[(2 * j) >> 6] >> ((2 * j) & 63)) & 3];
    return s;
}

int main(int argc, char **argv)
{
    std::string path;
    sgs_opts opts;
    sgs_opts_default(&opts);
    for (int i = 1; i < argc; i++) {
        std::string a = argv[i];
        if ((a == "-v" || a == "--verbose") && i + 1 < argc) opts.verbose = atoi(argv[++i]);
        else if (a.rfind("--verbose=", 0) == 0) opts.verbose = atoi(a.c_str() + 10);
        else if ((a == "-n" || a == "--next") && i + 1 < argc) ++i;
        else if (a.rfind("--next=", 0) == 0) {}
        else if (a == "--device" && i + 1 < argc) opts.device = atoi(argv[++i]);
        else if (a == "--filename" && i + 1 < argc) path = argv[++i];
        else path = a;
    }
    if (path.empty()) { fprintf(stderr, "the option '--filename' is required but missing\n"); return 1; }
    std::cout << "verbose level " << opts.verbose << std::endl;
    sgs_ham *H = nullptr;
    if (sgs_ham_load(path.c_str(), 0, 0, &H) != SGS_OK) { fprintf(stderr, "error: %s\n", sgs_last_error()); return 1; }
    const int n = sgs_ham_n(H);
    sgs_klocal *K = nullptr;
    if (sgs_klocal_create(H, &opts, -1, &K) != SGS_OK) { fprintf(stderr, "error: %s\n", sgs_last_error()); sgs_ham_free(H); return 1; }
    for (int m = n - 1; m >= 0; m--) std::cout << m << ' ' << sgs_klocal_sright_count(K, m) << std::endl;   // cpp/state.cpp:235
    while (true) {
        auto t0 = std::chrono::high_resolution_clock::now();
        if (sgs_klocal_evolve(K) != SGS_OK) { fprintf(stderr, "error: %s\n", sgs_last_error()); return 1; }
        auto t1 = std::chrono::high_resolution_clock::now();
        const double sec = std::chrono::duration_cast<std::chrono::microseconds>(t1 - t0).count() * 1e-6;
        std::cout << "evolve time " << sec << std::endl;      // cpp/state.cpp:519-520 (device time of the transition batch)
        std::cout << "update time " << 0 << std::endl;
        std::cout << "m " << sgs_klocal_m(K) << std::endl;
        std::cout << "number of states " << sgs_klocal_nstates(K) << std::endl;
        std::cout << "time " << sec << std::endl;
        std::cout << std::endl;
        if (sgs_klocal_m(K) == n) break;
    }
    sgs_result r;
    if (sgs_klocal_ground_state(K, &r) != SGS_OK) { fprintf(stderr, "error: %s\n", sgs_last_error()); return 1; }
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
