"""tests/golden/pauli_kats.json (40 Pauli products / commutation checks, 30 groups with canonical form, membership
sign and projection; written by tests/golden/make_golden.py from the unmodified Python reference) through the CUDA
group utilities of the C ABI: sgs_pauli_commute (Pauli::commute, cpp/stab.cpp:314-318), sgs_group_canonicalize
(Stab::canonicalize, cpp/stab.cpp:437-449), sgs_group_energy (Stab::decompose/include/energy, cpp/stab.cpp:493-511,
579-587).  Bit-exact."""
import ctypes as C
import json
import os

import numpy as np
import pytest

from conftest import GOLDEN

pytestmark = pytest.mark.gpu
KATS = json.load(open(os.path.join(GOLDEN, 'pauli_kats.json')))
CODE = {'I': 0, 'Z': 1, 'X': 2, 'Y': 3}


def pack(letters, W):
    words = [0] * W
    for j, ch in enumerate(letters):
        words[(2 * j) >> 6] |= CODE[ch] << ((2 * j) & 63)
    return words


def unpack(words, n):
    return ''.join('IZXY'[(words[(2 * j) >> 6] >> ((2 * j) & 63)) & 3] for j in range(n))


def capi():
    import stabilizer_gs_b200 as S
    return S._capi


def test_pauli_commute_kats():
    L = capi().lib()
    recs = KATS['products']
    # one batch per row length (the C ABI takes B pairs of the same n)
    for n in sorted({r['L'] for r in recs}):
        sub = [r for r in recs if r['L'] == n]
        W = (2 * n + 63) // 64
        a = (C.c_uint64 * (len(sub) * W))()
        b = (C.c_uint64 * (len(sub) * W))()
        for i, r in enumerate(sub):
            for k, w in enumerate(pack(['IXZY'[2 * z + x] for z, x in zip(r['z1'], r['x1'])], W)):
                a[i * W + k] = w
            for k, w in enumerate(pack(['IXZY'[2 * z + x] for z, x in zip(r['z2'], r['x2'])], W)):
                b[i * W + k] = w
        out = (C.c_uint8 * len(sub))()
        capi().check(L.sgs_pauli_commute(0, n, len(sub), a, b, out))
        assert [bool(v) for v in out] == [r['commute'] for r in sub]


def test_pauli_products_through_membership():
    """prod = P1*P2 (with its phase) for commuting pairs: the group <P1, P2> must contain ±prod with exactly the
    recorded relative sign (Stab::energy), which checks the device's phase tracking of Pauli::dot"""
    L = capi().lib()
    done = 0
    for r in KATS['products']:
        if not r['commute']:
            continue
        n = r['L']
        W = (2 * n + 63) // 64
        p1 = ['IXZY'[2 * z + x] for z, x in zip(r['z1'], r['x1'])]
        p2 = ['IXZY'[2 * z + x] for z, x in zip(r['z2'], r['x2'])]
        pr = ['IXZY'[2 * z + x] for z, x in zip(r['prod_z'], r['prod_x'])]
        if all(c == 'I' for c in pr) or p1 == p2:
            continue
        rows = (C.c_uint64 * (2 * W))(*(pack(p1, W) + pack(p2, W)))
        sg = (C.c_int8 * 2)(r['p1'], r['p2'])
        rank = C.c_int()
        capi().check(L.sgs_group_canonicalize(0, n, 2, rows, sg, C.byref(rank)))
        if rank.value != 2:
            continue
        P = (C.c_uint64 * W)(*pack(pr, W))
        ps = (C.c_int8 * 1)(1)
        out = (C.c_int8 * 1)()
        capi().check(L.sgs_group_energy(0, n, 2, rows, sg, 1, P, ps, out))
        assert out[0] == r['prod_phase']
        done += 1
    assert done >= 10


def test_group_canonical_form_and_membership_kats():
    L = capi().lib()
    for rec in KATS['groups']:
        n = rec['n']
        W = (2 * n + 63) // 64
        m = len(rec['input'])
        rows = (C.c_uint64 * (m * W))(*[w for s in rec['input'] for w in pack(s[1:], W)])
        sg = (C.c_int8 * m)(*[1 if s[0] == '+' else -1 for s in rec['input']])
        rank = C.c_int()
        capi().check(L.sgs_group_canonicalize(0, n, m, rows, sg, C.byref(rank)))
        got = [('+' if sg[i] > 0 else '-') + unpack([rows[i * W + k] for k in range(W)], n) for i in range(rank.value)]
        assert got == rec['canonical']
        # membership / relative sign of the probe and of the recorded member
        probes = [(rec['probe'], rec['probe_sign'] if (rec['probe_commutes'] and rec['probe_in_group']) else 0)]
        if rec['member']:
            probes.append((rec['member']['line'], rec['member']['sign']))
        B = len(probes)
        P = (C.c_uint64 * (B * W))(*[w for s, _ in probes for w in pack(s[1:], W)])
        ps = (C.c_int8 * B)(*[1 if s[0] == '+' else -1 for s, _ in probes])
        out = (C.c_int8 * B)()
        capi().check(L.sgs_group_energy(0, n, rank.value, rows, sg, B, P, ps, out))
        # Stab::energy = sign of P relative to the group element (cpp/stab.cpp:579-587): recorded sign is for the
        # signed probe line, the C ABI folds the probe's own sign in the same way
        assert [int(v) for v in out] == [e for _, e in probes], rec
