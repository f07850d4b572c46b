"""Unit-level pins of the oracle: cpp/test_linear.cpp inputs (KAT-5) and Pauli/Stab known answers
generated with the Python reference (tests/golden/linear_kats.json, pauli_kats.json).  CPU only."""
import ctypes as C
import json
import os

import numpy as np

import oracle_py as O
from conftest import GOLDEN

L = O.lib()


def ints(a):
    a = np.ascontiguousarray(np.asarray(a, dtype=np.int32))
    return a, a.ctypes.data_as(C.POINTER(C.c_int))


def test_linear_kats():
    k = json.load(open(os.path.join(GOLDEN, 'linear_kats.json')))
    out = (C.c_int * 7)()
    L.olin_inv_list(7, out)
    assert list(out) == k['inv_list_7'] == [0, 1, 4, 5, 2, 3, 6]
    A, pA = ints(k['get_order']['A'])
    order = (C.c_int * 3)()
    L.olin_get_order(pA, 3, 3, order)
    assert list(order) == k['get_order']['out'] == [2, 0, 1]
    # Gauss_elimination mod 3 with RHS = I (cpp/test_linear.cpp:52-66)
    A, pA = ints(k['gauss_mod3']['A'])
    b, pb = ints(np.eye(3))
    L.olin_gauss(pA, pb, 3, 3, 3, 3, 0)
    assert A.tolist() == k['gauss_mod3']['A2'] and b.tolist() == k['gauss_mod3']['b2']
    # solve_modN mod 2 (cpp/test_linear.cpp:79-92)
    A, pA = ints(k['solve_mod2']['A'])
    b, pb = ints(np.eye(3))
    over = (C.c_int * 3)()
    x, px = ints(np.zeros((3, 3)))
    under = L.olin_solve(pA, pb, 3, 3, 3, 2, over, px)
    assert [bool(v) for v in over] == k['solve_mod2']['is_over'] and bool(under) == k['solve_mod2']['is_under']
    assert x.tolist() == k['solve_mod2']['x']
    # kernel (cpp/test_linear.cpp:94-103)
    A, pA = ints(k['kernel_mod2']['A'])
    ker, pk = ints(np.zeros((3, 3)))
    nf = L.olin_kernel(pA, 2, 3, 2, pk)
    assert nf == 1 and ker.reshape(-1)[:3].tolist() == [r[0] for r in k['kernel_mod2']['kernel']]
    # simplify_coset_representative (cpp/test_linear.cpp:105-114)
    A2, pA2 = ints(k['coset']['A2'])
    x, px = ints(k['coset']['x'])
    o, po = ints(np.zeros(3))
    L.olin_simplify_coset(pA2, 2, 3, px, 2, po)
    assert o.tolist() == k['coset']['out'] == [0, 0, 1]
    # rank-deficient RREF with transform U
    A, pA = ints(k['rank_deficient']['A'])
    U, pU = ints(np.eye(4))
    L.olin_gauss(pA, pU, 4, 4, 4, 2, 1)
    assert A.tolist() == k['rank_deficient']['rref'] and U.tolist() == k['rank_deficient']['U']


def test_linear_random_mod2_and_packed_agree():
    k = json.load(open(os.path.join(GOLDEN, 'linear_kats.json')))
    for rec in k['random_mod2']:
        A0 = np.array(rec['A'], dtype=np.int32)
        n, m = A0.shape
        A, pA = ints(A0)
        U, pU = ints(np.eye(n))
        L.olin_gauss(pA, pU, n, m, n, 2, 1)
        assert A.tolist() == rec['rref'] and U.tolist() == rec['U']
        ker, pk = ints(np.zeros((m, max(m, 1))))
        A, pA = ints(A0)
        nf = L.olin_kernel(pA, n, m, 2, pk)
        got = ker.reshape(-1)[:m * nf].reshape(m, nf) if nf else np.zeros((m, 0), dtype=int)
        exp = np.array(rec['kernel']).reshape(m, -1) if rec['kernel'] and len(rec['kernel'][0]) else np.zeros((m, 0), dtype=int)
        assert got.tolist() == exp.tolist()
        # packed GF(2) canonicalisation of the same rows (interpreted as [z0,x0,z1,x1,...] columns, padded to even)
        if m % 2 == 0:
            st = O.OStab()
            L.ostab_init(C.byref(st), 0, m // 2)
            rank_rows = [r for r in rec['rref'] if any(r)]
            # only independent inputs are valid generator sets
            if len(rank_rows) == n:
                for i in range(n):
                    letters = ''.join('IZXY'[int(A0[i, 2 * j]) + 2 * int(A0[i, 2 * j + 1])] for j in range(m // 2))
                    st.g[i] = O.row_from('+' + letters)
                st.m = n
                L.ostab_canonicalize(C.byref(st), None)
                got_rows = [[(st.g[i].b[c >> 6] >> (c & 63)) & 1 for c in range(m)] for i in range(n)]
                assert got_rows == rec['rref']


def test_pauli_products_and_commutation():
    k = json.load(open(os.path.join(GOLDEN, 'pauli_kats.json')))
    for rec in k['products']:
        def mk(z, x, p):
            s = ''.join('IXZY'[2 * zz + xx] for zz, xx in zip(z, x))
            return O.row_from(('+' if p == 1 else '-') + s)
        a, b = mk(rec['z1'], rec['x1'], rec['p1']), mk(rec['z2'], rec['x2'], rec['p2'])
        assert bool(L.orow_commute(C.byref(a), C.byref(b))) == rec['commute']
        if rec['commute']:
            out = O.ORow()
            L.orow_dot(C.byref(out), C.byref(a), C.byref(b))
            exp = ('+' if rec['prod_phase'] == 1 else '-') + ''.join('IXZY'[2 * zz + xx] for zz, xx in zip(rec['prod_z'], rec['prod_x']))
            assert O.row_str(out, 0, rec['L']) == exp


def test_group_canonical_form_membership_projection():
    k = json.load(open(os.path.join(GOLDEN, 'pauli_kats.json')))
    for rec in k['groups']:
        n = rec['n']
        st = O.OStab()
        L.ostab_init(C.byref(st), 0, n)
        for s in rec['input']:
            r = O.row_from(s)
            L.ostab_add(C.byref(st), C.byref(r))
        assert [O.row_str(st.g[i], 0, n) for i in range(st.m)] == rec['canonical']
        probe = O.row_from(rec['probe'])
        assert bool(L.ostab_commute(C.byref(st), C.byref(probe))) == rec['probe_commutes']
        if rec['probe_commutes']:
            sg = C.c_int()
            x = O.OMask()
            assert bool(L.ostab_decompose(C.byref(st), C.byref(probe), C.byref(x), C.byref(sg))) == rec['probe_in_group']
            if rec['probe_in_group']:
                assert sg.value == rec['probe_sign']
        if rec['member']:
            mem = O.row_from(rec['member']['line'])
            sg = C.c_int()
            x = O.OMask()
            assert L.ostab_decompose(C.byref(st), C.byref(mem), C.byref(x), C.byref(sg))
            assert [(x.w[0] >> i) & 1 for i in range(st.m)] == rec['member']['x'] and sg.value == rec['member']['sign']
        pr = O.OStab()
        p = rec['project']
        L.ostab_project(C.byref(st), p['begin'], p['end'], C.byref(pr))
        assert [O.row_str(pr.g[i], p['begin'], p['end']) for i in range(pr.m)] == p['generators']
