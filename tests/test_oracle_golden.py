"""Pin the CPU oracle (oracle/*.c) against golden vectors produced by the UNMODIFIED Python reference
(tests/golden/make_golden.py).  CPU only."""
import glob
import json
import os

import pytest

import oracle_py as O
from conftest import GOLDEN

# (oracle_sparse_*.json were written BY the oracle, tests/golden/make_golden_oracle.py: nothing to pin there)
SPARSE = sorted(f for f in glob.glob(os.path.join(GOLDEN, '*.json'))
                if not os.path.basename(f).startswith('oracle_') and 'sparse' in json.load(open(f)))
KLOCAL = sorted(f for f in glob.glob(os.path.join(GOLDEN, '*.json')) if 'klocal' in json.load(open(f)))


def _name(p):
    return os.path.basename(p)[:-5]


@pytest.mark.parametrize('path', SPARSE, ids=_name)
def test_sparse_matches_python_reference(path):
    g = json.load(open(path))
    ref = g['sparse']
    r = O.sparse(path=os.path.join(GOLDEN, g['file']))
    assert r['energy'] == pytest.approx(ref['energy'], rel=1e-10, abs=1e-12)
    assert r['generators'] == ref['generators']
    assert r['term_signs'] == ref['term_signs']
    assert r['candidates'] == ref['candidates'] == ref['distinct_groups']
    assert {str(k): v for k, v in r['rank_hist'].items()} == ref['ranks']
    assert r['sum_2m'] == ref['sum_2m']
    # file order (what py/sparse.py uses) gives the same state
    assert g['sparse_file_order']['generators'] == ref['generators']
    assert g['sparse_file_order']['candidates'] == ref['candidates']


@pytest.mark.parametrize('path', SPARSE[:4], ids=_name)
def test_sparse_openmp_equals_serial(path):
    g = json.load(open(path))
    a = O.sparse(path=os.path.join(GOLDEN, g['file']), threads=1)
    b = O.sparse(path=os.path.join(GOLDEN, g['file']), threads=4)
    for k in ('energy', 'generators', 'term_signs', 'candidates', 'rank_hist'):
        assert a[k] == b[k]


@pytest.mark.parametrize('path', KLOCAL, ids=_name)
def test_klocal_matches_python_reference(path):
    g = json.load(open(path))
    ref = g['klocal']
    r = O.klocal(path=os.path.join(GOLDEN, g['file']), threads=2)
    n = g['n']
    assert r['energy'] == pytest.approx(ref['energy'], rel=1e-10, abs=1e-12)
    assert r['energy'] == pytest.approx(ref['energy_from_signs'], rel=1e-10, abs=1e-12)
    assert r['generators'] == ref['generators']
    assert r['term_signs'] == ref['term_signs']
    assert r['sright_counts'] == [ref['sright_counts'][str(m)] for m in range(n + 1)]
    assert r['states_per_site'] == ref['states_per_site']
    assert r['transitions'] == ref['transitions']


def test_klocal_equals_sparse_on_klocal_inputs():
    for name in ('finite_example', 'kat4_klocal_10_3_2_1', 'klocal_9_2_2_4'):
        g = json.load(open(os.path.join(GOLDEN, name + '.json')))
        a = O.sparse(path=os.path.join(GOLDEN, g['file']))
        b = O.klocal(path=os.path.join(GOLDEN, g['file']))
        assert a['energy'] == pytest.approx(b['energy'], rel=1e-10)
        assert a['generators'] == b['generators']


def test_config3_golden_md5():
    """BASELINE config 3: gen_klocal(64,4,3,0) — E, 47 generators, md5 of the printed group."""
    import hashlib
    g = json.load(open(os.path.join(GOLDEN, 'klocal_64_4_3_0.json')))
    r = O.klocal(path=os.path.join(GOLDEN, g['file']), threads=4)
    assert r['energy'] == pytest.approx(-30.9918743342, abs=1e-9)
    assert len(r['generators']) == 47
    line = 'qubit 0-63 size (47,) ' + ' '.join(r['generators']) + '\n'   # print() adds the newline
    assert hashlib.md5(line.encode()).hexdigest() == 'a715783786513531b60e5877f72a9f30'
    assert max(r['states_per_site']) == 113 and r['transitions'] == 2850


PERIODIC = sorted(f for f in glob.glob(os.path.join(GOLDEN, '*.json')) if 'periodic' in json.load(open(f)))


@pytest.mark.parametrize('path', PERIODIC, ids=lambda p: os.path.basename(p)[:-5])
def test_periodic_matches_python_reference(path):
    """py/klocal_periodic.py on the reference's example and on three more unit cells (fixtures written by the
    unmodified Python reference, tests/golden/make_golden.py)"""
    rec = json.load(open(path))
    g = rec['periodic']
    r = O.periodic(os.path.join(GOLDEN, rec['file']), cmax=g['cmax'], c=g['c'])
    assert (r['l'], r['k'], r['n'], r['nterms']) == (g['l'], g['k'], g['n'], g['nterms'])
    assert r['sright_log'] == g['sright_log']
    assert r['fixed_point_sizes'] == g['fixed_point_sizes']
    key = lambda o: (round(o['energy'], 9), o['begin'], o['end'], o['generators'])
    assert sorted(map(key, r['results'])) == sorted(map(key, g['results']))
    assert min(o['energy'] for o in r['results']) / g['c'] == pytest.approx(g['min_energy_per_period'])
