"""Chains longer than 127 qubits (VERDICT r1 missing #4): the reference's DP carries only a k-site
window (cpp/state.cpp:411, cpp/stab.cpp:782-809) and runs at any n; so does k_dp_run — DP terms are (first site, one
u64) — and the ground state of a wide chain is finished by the generic-width kernels of group_ops.cu.
Parity: the 12-word build of the CPU oracle on gen_klocal(256, 4, 3, 0) (a prefix of the 1024-site instance) and on a
periodic chain with a 12-site unit cell (n = 138); invariants at n = 1024."""
import json
import os
import subprocess
import sys

import pytest

import gen_hamiltonians as G
from conftest import ROOT

pytestmark = pytest.mark.gpu


def sgs():
    import stabilizer_gs_b200 as S
    return S


def wide_oracle(*args):
    p = subprocess.run([sys.executable, os.path.join(ROOT, 'tests', 'oracle_wide_cli.py'), *map(str, args)], capture_output=True, text=True, timeout=900)
    assert p.returncode == 0, p.stderr[-2000:]
    return json.loads(p.stdout.strip().splitlines()[-1])


def test_256_sites_vs_wide_oracle(tmp_path):
    S = sgs()
    text = G.klocal_text(256, 4, 3, 0)
    f = tmp_path / 'k256.txt'
    f.write_text(text)
    ref = wide_oracle('klocal', f)
    r = S.klocal_search(S.HamHandle.from_text(text))
    assert r['energy'] == pytest.approx(ref['energy'], rel=1e-10)
    assert r['generators'] == ref['generators'] and len(r['generators']) == 196
    assert r['term_signs'] == ref['term_signs']
    assert r['states_per_site'] == ref['states_per_site'] and r['sright_counts'] == ref['sright_counts']
    assert r['transitions'] == ref['transitions']


def _invariants(S, text, r, n):
    H = S.HamHandle.from_text(text)
    rows, w, b, e = H.terms()
    assert r['energy'] == pytest.approx(sum(wi * si for wi, si in zip(w, r['term_signs'])), rel=1e-10)
    code = {'I': 0, 'Z': 1, 'X': 2, 'Y': 3}
    gb = [sum(code[ch] << (2 * j) for j, ch in enumerate(s[1:])) for s in r['generators']]
    ev = int('01' * (n + 1), 2)
    sw = lambda x: ((x >> 1) & ev) | ((x & ev) << 1)
    piv = [(x & -x).bit_length() - 1 for x in gb]
    assert piv == sorted(piv) and len(set(piv)) == len(piv)                       # RREF order
    pmask = sum(1 << p for p in piv)
    for i, x in enumerate(gb):
        assert x & pmask == 1 << piv[i]                                           # fully reduced
    for i in range(0, len(gb), 7):                                                # generators commute (sampled: O(m^2) big-int ops)
        for j in range(i):
            assert bin(gb[i] & sw(gb[j])).count('1') % 2 == 0
    # every term sign is what the canonical generators say (Stab::energy): reduce the term by pivots, track nothing but
    # membership here — the sign itself is pinned by E = sum w * sign above and by the 256-site oracle comparison
    by_piv = dict(zip(piv, gb))
    for t, row in enumerate(rows):
        v = sum(x << (64 * k) for k, x in enumerate(row))
        while v:
            p = (v & -v).bit_length() - 1
            if p not in by_piv:
                break
            v ^= by_piv[p]
        assert (v == 0) == (r['term_signs'][t] != 0)


def test_1024_sites():
    S = sgs()
    text = G.klocal_text(1024, 4, 3, 0)
    r = S.klocal_search(S.HamHandle.from_text(text))
    assert r['n'] == 1024 and len(r['states_per_site']) == 1025 and r['states_per_site'][-1] >= 1
    _invariants(S, text, r, 1024)
    # the 256-site instance is a prefix of this one: far from both ends the per-site optimum does not know the length
    # (checked on the energy density only: -0.519 per site at n = 256)
    assert -0.60 < r['energy'] / 1024 < -0.45


def test_wide_hamiltonian_is_refused_by_the_sparse_search_only():
    S = sgs()
    H = S.HamHandle.from_text(G.klocal_text(200, 3, 1, 2))
    assert H.n == 200 and H.words == 7
    with pytest.raises(S.SgsError) as e:
        S.sparse_search(H)
    assert e.value.code == -6
    rows, w, b, e2 = H.terms()
    assert all(sum(bin(x).count('1') for x in r) > 0 for r in rows)               # absolute rows are rebuilt from the compact ones
    r = S.klocal_search(H)
    assert r['energy'] == pytest.approx(sum(wi * si for wi, si in zip(w, r['term_signs'])), rel=1e-10)
    # a term that spans more than 32 sites cannot be stored compactly
    with pytest.raises(S.SgsError):
        S.HamHandle.from_text("200 40\n 1.0 ZZ 0 39\n")


def test_periodic_chain_with_a_12_site_unit_cell(tmp_path):
    """cmax = 10, k = 3, l = 12: the unrolled chain has 138 sites (the r1 parser refused l >= 12)"""
    S = sgs()
    lines = ["12 3"]
    for i in range(12):
        lines.append(f" {-1.0 - 0.05 * (i % 3):.4f} ZZ {i} {i + 1}")
        lines.append(f" {-0.6 + 0.1 * (i % 2):.4f} X {i}")
    f = tmp_path / 'cell12.txt'
    f.write_text('\n'.join(lines) + '\n')
    ref = wide_oracle('periodic', f, 10, 1)
    r = S.klocal_periodic(str(f), cmax=10, c=1)
    key = lambda o: (round(o['energy'], 9), o['begin'], o['end'], tuple(o['generators']))
    assert sorted(map(key, r['results'])) == sorted(map(key, ref['results']))
    assert r['fixed_point_sizes'] == ref['fixed_point_sizes']
