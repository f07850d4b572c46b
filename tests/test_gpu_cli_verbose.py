"""`klocal_sparse -v 1` (cpp/klocal_sparse.cpp:28-37,47-60; State::to_string cpp/state.cpp:332-345): the dumped right
environments and DP states equal, site by site, what the unmodified Python reference holds
(tests/golden/klocal_dump_*.json, written by tests/golden/make_golden_dump.py); the Ps_indices / Ps_signs matrices have
the reference's shape and decode to the ground state; boost's option spellings are accepted; the progress line goes
to stderr."""
import json
import os
import re
import subprocess

import pytest

from conftest import GOLDEN, ROOT

pytestmark = pytest.mark.gpu
BIN = os.path.join(ROOT, 'stabilizer_gs_b200', 'bin', 'klocal_sparse')


def parse_block(s):
    """'qubit b-e size m +XZ -IY ' -> [b, e + 1, [strings]]"""
    m = re.match(r'qubit (-?\d+)-(-?\d+) size (\d+) ?(.*)$', s.strip())
    assert m, s
    strings = m.group(4).split()
    assert len(strings) == int(m.group(3))
    return [int(m.group(1)), int(m.group(2)) + 1, strings]


def parse_dump(out, n):
    lines = out.split('\n')
    i = 0
    assert lines[i].startswith('verbose level')
    i += 1
    counts = {}
    while re.fullmatch(r'\d+ \d+', lines[i]):
        a, b = map(int, lines[i].split()); counts[a] = b; i += 1
    sright = {}
    for m in range(n):                                         # "m\n" then one block per group, then a blank line
        assert lines[i] == str(m), (i, lines[i])
        i += 1
        sright[m] = []
        while lines[i] != '':
            sright[m].append(parse_block(lines[i])); i += 1
        i += 1
    sites = []
    while not lines[i].startswith('ground state stabilizers'):
        assert lines[i].startswith('evolve time'), (i, lines[i])
        m = int(lines[i + 2].split()[1]); ns = int(lines[i + 3].split()[3])
        i += 6
        states = []
        for s in range(ns):
            assert lines[i] == str(s) and lines[i + 1] == f'm = {m}', (i, lines[i:i + 2])
            stab = parse_block(lines[i + 2][len('stab = '):])
            sr = parse_block(lines[i + 3][len('Sright = '):])
            assert lines[i + 4].startswith('valid_pauli_indices = ')
            Es = [float(x) for x in lines[i + 5][len('energies = '):].split()]
            assert lines[i + 6] == 'Ps indices '
            rows = len(Es)
            Pi = [list(map(int, lines[i + 7 + r].split())) for r in range(rows)]
            assert lines[i + 7 + rows] == 'Ps signs '
            Ps = [list(map(int, lines[i + 8 + rows + r].split())) for r in range(rows)]
            assert all(len(r) == n for r in Pi + Ps) and len(Es) == 1 << len(stab[2])
            i += 8 + 2 * rows
            assert lines[i:i + 3] == ['', '', ''], lines[i:i + 4]
            i += 3
            states.append(dict(stab=stab, sright=sr, Es=Es, Pi=Pi, Ps=Ps))
        sites.append(states)
    return counts, sright, sites, lines[i:]


@pytest.mark.parametrize('name', ['klocal_dump_finite_example', 'klocal_dump_8_3_2_5'])
def test_verbose_dump_matches_the_python_reference(name):
    g = json.load(open(os.path.join(GOLDEN, name + '.json')))
    path = os.path.join(GOLDEN, g['file'])
    p = subprocess.run([BIN, '-v', '1', path], capture_output=True, text=True, timeout=300)
    assert p.returncode == 0, p.stderr
    n = g['n']
    counts, sright, sites, tail = parse_dump(p.stdout, n)
    for m in range(n):
        assert sorted(sright[m]) == [list(x) for x in g['sright'][str(m)]], m
        assert counts[m] == len(sright[m])
    assert len(sites) == n
    for m in range(n):
        got = sorted([st['stab'], st['sright'], [round(e, 5) for e in st['Es']]] for st in sites[m])
        exp = sorted([x[0], x[1], [round(e, 5) for e in x[2]]] for x in g['sites'][m])      # (Eigen prints 6 significant digits)
        assert got == exp, m
    # the history of the best entry of the best final state decodes (index = i * n + m, cpp/state.cpp:393) to terms of
    # the Hamiltonian; that many generators are printed below
    best = min(((e, si, ei) for si, st in enumerate(sites[-1]) for ei, e in enumerate(st['Es'])))
    st = sites[-1][best[1]]
    used = [(c, s) for c, s in zip(st['Pi'][best[2]], st['Ps'][best[2]]) if s != 0]
    assert all(s in (1, -1) and c % n < n for c, s in used)
    assert tail[1] == f'qubit 0-{n - 1} size {len(used)}'
    # stderr: one progress line per site
    assert p.stderr.count('progress: 1 in ') == n


def test_boost_option_spellings():
    path = os.path.join(ROOT, 'examples', 'hamiltonian_finite.txt')
    base = subprocess.run([BIN, path], capture_output=True, text=True, check=True).stdout
    strip = lambda s: re.sub(r'(time) [0-9.e+-]+', r'\1 T', s)
    for argv in (['-v0', path], ['--verbose=0', path], ['--verbose', '0', '-n1', path], ['-n', '2', '--filename', path],
                 ['--filename=' + path, '--next=1']):
        out = subprocess.run([BIN] + argv, capture_output=True, text=True, check=True).stdout
        assert strip(out) == strip(base), argv
    v1 = subprocess.run([BIN, '-v1', path], capture_output=True, text=True, check=True).stdout
    assert v1.startswith('verbose level 1\n') and 'Ps indices \n' in v1 and v1.endswith(base[base.index('ground state stabilizers'):])
    v2 = subprocess.run([BIN, '--verbose=2', path], capture_output=True, text=True, check=True).stdout
    assert v2.startswith('verbose level 2\n') and v2.count('Sright = ') == v1.count('Sright = ')
