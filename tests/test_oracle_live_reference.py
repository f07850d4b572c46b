"""Live differential test: the CPU oracle against the UNMODIFIED Python reference on fresh random instances.

Runs only where the reference is present (/root/reference, i.e. the build container); on the GPU box it is skipped
— there the committed fixtures under tests/golden/ (written by the same functions) are the pin.  The instances are
tiny because the reference evaluates ~25 sparse leaves or ~50 DP transitions per second."""
import importlib.util
import os
import tempfile

import pytest

import gen_hamiltonians as G
import oracle_py as O
from conftest import GOLDEN

REF = '/root/reference/py'
pytestmark = pytest.mark.skipif(not os.path.isdir(REF), reason='the Python reference is not on this machine')


@pytest.fixture(scope='module')
def mg():
    spec = importlib.util.spec_from_file_location('make_golden', os.path.join(GOLDEN, 'make_golden.py'))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def _ref_ham(mg, text):
    with tempfile.NamedTemporaryFile('w', suffix='.txt', delete=False) as f:
        f.write(text)
    try:
        return mg.Hamiltonian.from_file(f.name)
    finally:
        os.unlink(f.name)


@pytest.mark.parametrize('n,T,seed', [(3, 7, 101), (4, 9, 102), (5, 10, 103), (4, 12, 104), (6, 9, 105)])
def test_sparse_fresh_instances(mg, n, T, seed):
    text = G.sparse_text(n, T, seed)
    ref = mg.run_sparse(_ref_ham(mg, text), 'cpp')
    r = O.sparse(text=text)
    assert r['energy'] == pytest.approx(ref['energy'], rel=1e-10, abs=1e-12)
    assert r['generators'] == ref['generators'] and r['term_signs'] == ref['term_signs']
    assert r['candidates'] == ref['candidates'] == ref['distinct_groups']
    assert {str(k): v for k, v in r['rank_hist'].items()} == ref['ranks']


@pytest.mark.parametrize('n,k,per,seed', [(6, 2, 2, 111), (7, 3, 2, 112), (6, 3, 3, 113), (8, 2, 3, 114)])
def test_klocal_fresh_instances(mg, n, k, per, seed):
    text = G.klocal_text(n, k, per, seed)
    ref = mg.run_klocal(_ref_ham(mg, text))
    r = O.klocal(text=text, threads=2)
    assert r['energy'] == pytest.approx(ref['energy'], rel=1e-10, abs=1e-12)
    assert r['generators'] == ref['generators'] and r['term_signs'] == ref['term_signs']
    assert r['sright_counts'] == [ref['sright_counts'][str(m)] for m in range(n + 1)]
    assert r['states_per_site'] == ref['states_per_site'] and r['transitions'] == ref['transitions']


@pytest.mark.parametrize('cell', ["1 2\n 0.7 XY 0 1\n -0.4 Z 0\n 0.2 X 0\n", "2 2\n -1.0 ZZ 0 1\n 0.6 XX 1 2\n -0.25 Y 1\n",
                                  "1 3\n -0.9 ZXZ 0 1 2\n 0.35 XX 0 1\n"])
def test_periodic_fresh_cells(mg, cell, tmp_path):
    p = tmp_path / 'cell.txt'
    p.write_text(cell)
    ref = mg.run_periodic(str(p), 10, 2)
    r = O.periodic(str(p), cmax=10, c=2)
    assert (r['l'], r['k'], r['n'], r['nterms']) == (ref['l'], ref['k'], ref['n'], ref['nterms'])
    assert r['sright_log'] == ref['sright_log'] and r['fixed_point_sizes'] == ref['fixed_point_sizes']
    key = lambda o: (round(o['energy'], 9), o['begin'], o['end'], o['generators'])
    assert sorted(map(key, r['results'])) == sorted(map(key, ref['results']))


@pytest.mark.parametrize('n,k,per,seed', [(7, 3, 2, 121), (9, 2, 3, 122), (6, 4, 4, 123)])
def test_parser_order_matches_reference(mg, n, k, per, seed):
    """Hamiltonian::from_file / paulis_list (cpp/state.cpp:58-96, py/state.py:34-49): term windows and the
    (last site, file order) ordering of the product's parser against the reference's own objects."""
    import stabilizer_gs_b200 as S
    text = G.klocal_text(n, k, per, seed)
    Hr = _ref_ham(mg, text)
    H = S.HamHandle.from_text(text)
    assert (H.n, H.k, H.nterms) == (Hr.n, Hr.k, len(Hr.original_paulis))
    ws = [float(w) for _, w in Hr.original_paulis]
    assert len(set(ws)) == len(ws)
    buckets = Hr.paulis.values() if isinstance(Hr.paulis, dict) else Hr.paulis
    expect = [ws.index(float(w)) for bucket in buckets for _, w in bucket]
    assert H.order() == expect
    rows, w, b, e = H.terms()
    for i, (P, wi) in enumerate(Hr.original_paulis):
        assert w[i] == pytest.approx(float(wi)) and (b[i], e[i]) == (P.begin, P.end)
