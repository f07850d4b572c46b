// cli/sparse.cpp — drop-in for the reference's `./sparse <file>` (cpp/sparse.cpp:4-18): same stdout.
// Extra flags (ours): --gpus N, --device D, --expand-depth K, --stats (one JSON line on stderr), --prune (branch and
// bound: same printed result, only the groups that could still beat the incumbent are visited).
#include <cstdio>
#include <cstdlib>
#include <cstring>
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
    bool stats = false;
    for (int i = 1; i < argc; i++) {
        std::string a = argv[i];
        if (a == "--gpus" && i + 1 < argc) opts.ngpus = atoi(argv[++i]);
        else if (a == "--device" && i + 1 < argc) opts.device = atoi(argv[++i]);
        else if (a == "--expand-depth" && i + 1 < argc) opts.expand_depth = atoi(argv[++i]);
        else if (a == "--stats") stats = true;
        else if (a == "--prune") opts.prune = 1;
        else path = a;
    }
    if (path.empty()) { fprintf(stderr, "usage: %s [--gpus N] [--device D] [--stats] [--prune] <hamiltonian file>\n", argv[0]); return 2; }
    sgs_ham *H = nullptr;
    if (sgs_ham_load(path.c_str(), 0, 0, &H) != SGS_OK) { fprintf(stderr, "error: %s\n", sgs_last_error()); return 1; }
    sgs_result r;
    if (sgs_sparse_search(H, &opts, &r) != SGS_OK) { fprintf(stderr, "error: %s\n", sgs_last_error()); sgs_ham_free(H); return 1; }
    const int n = sgs_ham_n(H);
    std::cout << "ground state stabilizers" << std::endl;
    // Stab::to_string(true), cpp/stab.cpp:589-599: header line, one generator per line, then a blank line from endl
    std::cout << "qubit 0-" << n - 1 << " size " << r.m << "\n";
    for (int i = 0; i < r.m; i++) std::cout << row_string(r.gens + (size_t)i * r.W, r.gen_signs[i], n) << " \n";
    std::cout << std::endl;
    std::cout << std::setprecision(16) << "ground state energy " << r.energy << std::endl;
    std::cout << "terms signs" << std::endl;
    for (int t = 0; t < r.T; t++) std::cout << (int)r.term_signs[t] << " ";
    std::cout << std::endl;
    if (stats)
        fprintf(stderr, "{\"candidates\": %llu, \"sum_2m\": %.0f, \"slow_nodes\": %llu, \"seconds_total\": %.6f, \"seconds_device\": %.6f, "
                        "\"seconds_hot_kernel\": %.6f, \"candidates_per_s\": %.4g, \"kernel_launches\": %u}\n",
                (unsigned long long)r.candidates, r.sum_2m, (unsigned long long)r.slow_nodes, r.seconds_total, r.seconds_device,
                r.seconds_hot_kernel, r.seconds_device > 0 ? r.candidates / r.seconds_device : 0.0, r.kernel_launches);
    sgs_result_free(&r);
    sgs_ham_free(H);
    return 0;
}
