// cli/klocal_periodic.cpp — native counterpart of `python klocal_periodic.py` (py/klocal_periodic.py:1-63;
// the reference has no C++ binary for it).  usage: klocal_periodic [--cmax N] [--c N] <file>
// Prints the "m count" lines of generate_Sright_periodic, the fixed-point sizes, then every orbit as
// "energy E" / "qubit b-e size (m,) <generators>" / blank, sorted (the Python order depends on PYTHONHASHSEED).
#include <cstdio>
#include <cstdlib>
#include <iostream>
#include <string>

#include "sgs.h"

int main(int argc, char **argv)
{
    std::string path;
    int cmax = 10, c = 3;                 // py/klocal_periodic.py:5,7
    sgs_opts opts;
    sgs_opts_default(&opts);
    for (int i = 1; i < argc; i++) {
        std::string a = argv[i];
        if (a == "--cmax" && i + 1 < argc) cmax = atoi(argv[++i]);
        else if (a == "--c" && i + 1 < argc) c = atoi(argv[++i]);
        else if (a == "--device" && i + 1 < argc) opts.device = atoi(argv[++i]);
        else path = a;
    }
    if (path.empty()) { fprintf(stderr, "usage: %s [--cmax N] [--c N] <periodic hamiltonian file>\n", argv[0]); return 2; }
    sgs_periodic_result r;
    if (sgs_klocal_periodic(path.c_str(), cmax, c, &opts, &r) != SGS_OK) { fprintf(stderr, "error: %s\n", sgs_last_error()); return 1; }
    for (int i = 0; i < r.nlog; i++) std::cout << r.sright_log[2 * i] << ' ' << r.sright_log[2 * i + 1] << std::endl;
    for (int i = 0; i < r.nfixed; i++) std::cout << r.fixed_sizes[i] << std::endl;
    for (int i = 0; i < r.norbits; i++) {
        const sgs_orbit &o = r.orbits[i];
        std::cout << "energy " << o.energy << std::endl;
        std::cout << "qubit " << o.begin << "-" << o.end - 1 << " size (" << o.m << ",) ";
        for (int g = 0; g < o.m; g++) {
            std::string s(1, o.signs[g] >= 0 ? '+' : '-');
            for (int j = o.begin; j < o.end; j++) s += "IZXY"[(o.gens[(size_t)g * r.W + ((2 * j) >> 6)] >> ((2 * j) & 63)) & 3];
            std::cout << s << (g + 1 < o.m ? " " : "");
        }
        std::cout << std::endl << std::endl;
    }
    sgs_periodic_result_free(&r);
    return 0;
}
