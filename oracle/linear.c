/*
 * oracle/linear.c — CPU ORACLE (test infrastructure only; see sgs_oracle.h).
 * Dense mod-N linear algebra on int matrices, restating cpp/linear.cpp (and
 * py/linear.py) step for step so the cpp/test_linear.cpp inputs can be pinned.
 * The hot path only ever uses N = 2; the packed GF(2) versions used by the
 * searches live in pauli_stab.c and are cross-checked against these.
 */
#include <stdlib.h>
#include <string.h>
#include "sgs_oracle.h"

static int imod(int a, int N) { int r = a % N; return r < 0 ? r + N : r; }   /* cpp/linear.hpp:104 */

/* cpp/linear.hpp:134-146 get_inv_list: inv[x] = first y with x*y = 1 (mod N), 0 if none */
void olin_inv_list(int N, int *out)
{
    for (int x = 0; x < N; x++) {
        out[x] = 0;
        for (int y = 0; y < N; y++)
            if ((x * y) % N == 1) { out[x] = y; break; }
    }
}

static int first_nonzero(const int *row, int m)   /* argmax(get_binary(row)); 0 for a zero row */
{
    for (int j = 0; j < m; j++) if (row[j] != 0) return j;
    return 0;
}

/* cpp/linear.cpp:4-19 get_order: rows sorted by first non-zero column, zero rows last (stable) */
void olin_get_order(const int *A, int n, int m, int *order)
{
    int *key = (int *)malloc(sizeof(int) * (n > 0 ? n : 1));
    for (int i = 0; i < n; i++) {
        int s = 0;
        for (int j = 0; j < m; j++) s += A[i * m + j];
        key[i] = (s == 0) ? m : first_nonzero(A + i * m, m);
        order[i] = i;
    }
    for (int i = 1; i < n; i++) {            /* stable insertion sort */
        int o = order[i], j = i - 1;
        while (j >= 0 && key[order[j]] > key[o]) { order[j + 1] = order[j]; j--; }
        order[j + 1] = o;
    }
    free(key);
}

/* cpp/linear.cpp:21-75 Gauss_elimination: in place on A (n x m) and b (n x kb) */
void olin_gauss(int *A, int *b, int n, int m, int kb, int N, int allow_reorder)
{
    if (n == 0 || m == 0) return;
    int *inv = (int *)malloc(sizeof(int) * N);
    int *piv = (int *)malloc(sizeof(int) * n);
    olin_inv_list(N, inv);
    for (int pass = 0; pass < 2; pass++) {
        for (int kk = 0; kk < n; kk++) {
            int k = pass == 0 ? kk : n - 1 - kk;
            int pivot;
            int lo, hi;
            if (pass == 0) { pivot = first_nonzero(A + k * m, m); piv[k] = pivot; lo = k + 1; hi = n; }
            else { pivot = piv[k]; lo = 0; hi = k; }
            for (int i = lo; i < hi; i++) {                     /* update(): rows i -= f * row k */
                int f = (A[i * m + pivot] * inv[A[k * m + pivot]]) % N;
                if (f == 0) continue;
                for (int j = 0; j < m; j++) A[i * m + j] = imod(A[i * m + j] - A[k * m + j] * f, N);
                for (int j = 0; j < kb; j++) b[i * kb + j] = imod(b[i * kb + j] - b[k * kb + j] * f, N);
            }
            if (pass == 1) {                                     /* scale the pivot row */
                int iv = inv[A[k * m + pivot]];
                if (iv != 0) {
                    for (int j = 0; j < m; j++) A[k * m + j] = imod(A[k * m + j] * iv, N);
                    for (int j = 0; j < kb; j++) b[k * kb + j] = imod(b[k * kb + j] * iv, N);
                }
            }
        }
    }
    if (allow_reorder) {
        int *order = (int *)malloc(sizeof(int) * n);
        int *A2 = (int *)malloc(sizeof(int) * n * m);
        int *b2 = (int *)malloc(sizeof(int) * n * (kb > 0 ? kb : 1));
        olin_get_order(A, n, m, order);
        for (int i = 0; i < n; i++) {
            memcpy(A2 + i * m, A + order[i] * m, sizeof(int) * m);
            if (kb) memcpy(b2 + i * kb, b + order[i] * kb, sizeof(int) * kb);
        }
        memcpy(A, A2, sizeof(int) * n * m);
        if (kb) memcpy(b, b2, sizeof(int) * n * kb);
        free(order); free(A2); free(b2);
    }
    free(inv); free(piv);
}

/* cpp/linear.cpp:77-121 solve_modN: A (n x m) x = b (n x kb).  Returns is_under; x is m x kb. */
int olin_solve(const int *A, const int *b, int n, int m, int kb, int N, int *is_over, int *x)
{
    memset(x, 0, sizeof(int) * m * kb);
    if (m == 0) {
        for (int c = 0; c < kb; c++) {
            is_over[c] = 0;
            for (int i = 0; i < n; i++) if (b[i * kb + c] > 0) is_over[c] = 1;
        }
        return 0;
    }
    int *A2 = (int *)malloc(sizeof(int) * n * m), *b2 = (int *)malloc(sizeof(int) * n * kb);
    memcpy(A2, A, sizeof(int) * n * m);
    memcpy(b2, b, sizeof(int) * n * kb);
    olin_gauss(A2, b2, n, m, kb, N, 1);
    int is_under = 0;
    for (int c = 0; c < kb; c++) is_over[c] = 0;
    for (int i = 0; i < n; i++) {
        int cond = 0;
        for (int j = 0; j < m; j++) cond += A2[i * m + j];
        if (cond == 0) {
            for (int c = 0; c < kb; c++) if (b2[i * kb + c] > 0) is_over[c] = 1;
        } else {
            if (cond > 1) is_under = 1;
            int p = first_nonzero(A2 + i * m, m);
            for (int c = 0; c < kb; c++) x[p * kb + c] = b2[i * kb + c];
        }
    }
    free(A2); free(b2);
    return is_under;
}

/* cpp/linear.cpp:135-168 get_kernel: basis of {x : A x = 0}; one vector per free column, ascending */
int olin_kernel(const int *A, int n, int m, int N, int *ker)
{
    if (m == 0) return 0;
    int *A2 = (int *)malloc(sizeof(int) * (n > 0 ? n : 1) * m);
    memcpy(A2, A, sizeof(int) * n * m);
    olin_gauss(A2, NULL, n, m, 0, N, 1);
    int r = 0;
    int *main_col = (int *)malloc(sizeof(int) * (n > 0 ? n : 1));
    char *is_main = (char *)calloc(m, 1);
    for (int i = 0; i < n; i++) {
        int s = 0;
        for (int j = 0; j < m; j++) s += A2[i * m + j];
        if (s == 0) continue;
        main_col[r] = first_nonzero(A2 + i * m, m);
        is_main[main_col[r]] = 1;
        if (r != i) memcpy(A2 + r * m, A2 + i * m, sizeof(int) * m);
        r++;
    }
    int nf = m - r, f = 0;
    memset(ker, 0, sizeof(int) * m * (nf > 0 ? nf : 1));
    for (int c = 0; c < m; c++) {
        if (is_main[c]) continue;
        for (int i = 0; i < r; i++) ker[main_col[i] * nf + f] = imod(-A2[i * m + c], N);
        ker[c * nf + f] = 1;
        f++;
    }
    free(A2); free(main_col); free(is_main);
    return nf;
}

/* cpp/linear.cpp:170-178 simplify_coset_representative */
void olin_simplify_coset(const int *A2, int n, int m, const int *x, int N, int *out)
{
    for (int j = 0; j < m; j++) out[j] = x[j];
    if (n == 0) { for (int j = 0; j < m; j++) out[j] = imod(out[j], N); return; }
    for (int j = 0; j < m; j++) {
        int acc = x[j];
        for (int i = 0; i < n; i++) acc -= A2[i * m + j] * x[first_nonzero(A2 + i * m, m)];
        out[j] = imod(acc, N);
    }
}
