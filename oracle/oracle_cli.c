/*
 * oracle/oracle_cli.c — CPU ORACLE driver (test infrastructure only; see sgs_oracle.h).
 * usage: oracle_cli sparse|klocal <file> [threads]      (stdout in the cpp/sparse.cpp:10-17 layout)
 *        oracle_cli periodic <file> [cmax] [c]          (py/klocal_periodic.py:60-63 layout)
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "sgs_oracle.h"

int main(int argc, char **argv)
{
    if (argc < 3) { fprintf(stderr, "usage: %s sparse|klocal|periodic <file> [threads | cmax c]\n", argv[0]); return 2; }
    char err[256] = "";
    if (!strcmp(argv[1], "periodic")) {
        int cmax = argc > 3 ? atoi(argv[3]) : 10, c = argc > 4 ? atoi(argv[4]) : 3;
        operiodic *R = oracle_periodic(argv[2], cmax, c, err, sizeof(err));
        if (!R) { fprintf(stderr, "error: %s\n", err); return 1; }
        for (int i = 0; i < R->nlog; i++) printf("%d %d\n", R->sright_log[2 * i], R->sright_log[2 * i + 1]);
        for (int i = 0; i < R->nfixed; i++) printf("%d\n", R->fixed_sizes[i]);
        for (int i = 0; i < R->norbits; i++) {
            operiodic_orbit *o = &R->orbits[i];
            printf("energy %.10g\nqubit %d-%d size (%d,)", o->energy, o->begin, o->end - 1, o->m);
            for (int g = 0; g < o->m; g++) { char s[OMAXQ + 2]; orow_to_string(&o->gens[g], o->begin, o->end, s); printf(" %s", s); }
            printf("\n\n");
        }
        operiodic_free(R);
        return 0;
    }
    oham *H = oham_load(argv[2], 0, 0, err, sizeof(err));
    if (!H) { fprintf(stderr, "error: %s\n", err); return 1; }
    int threads = argc > 3 ? atoi(argv[3]) : 1;
    ores *r = !strcmp(argv[1], "sparse") ? oracle_sparse(H, threads) : oracle_klocal(H, threads);
    if (!strcmp(argv[1], "klocal")) {
        for (int m = H->n - 1; m >= 0; m--) printf("%d %d\n", m, r->sright_counts[m]);
        for (int m = 1; m <= H->n; m++) printf("m %d\nnumber of states %d\n\n", m, r->states_per_site[m]);
    }
    printf("ground state stabilizers\nqubit 0-%d size %d\n", H->n - 1, r->m);
    for (int i = 0; i < r->m; i++) { char s[OMAXQ + 2]; orow_to_string(&r->gens[i], 0, H->n, s); printf("%s \n", s); }
    printf("\nground state energy %.16g\nterms signs\n", r->energy);
    for (int t = 0; t < r->T; t++) printf("%d ", r->term_signs[t]);
    printf("\n");
    if (!strcmp(argv[1], "sparse")) fprintf(stderr, "candidates %llu sum_2m %.0f seconds %.3f\n", (unsigned long long)r->candidates, r->sum_2m, r->seconds);
    else fprintf(stderr, "transitions %llu seconds %.3f\n", (unsigned long long)r->transitions, r->seconds);
    ores_free(r);
    oham_free(H);
    return 0;
}
