"""GPU parity tests of the k-local DP and the periodic chain search (through the C ABI) against the CPU
oracle and the golden vectors of the Python reference."""
import glob
import hashlib
import json
import os

import pytest

import gen_hamiltonians as G
import oracle_py as O
from conftest import GOLDEN, ROOT

pytestmark = pytest.mark.gpu

KLOCAL = sorted(f for f in glob.glob(os.path.join(GOLDEN, '*.json')) if 'klocal' in json.load(open(f)))


def _name(p):
    return os.path.basename(p)[:-5]


def sgs():
    import stabilizer_gs_b200 as S
    return S


@pytest.mark.parametrize('path', KLOCAL, ids=_name)
def test_golden_python_reference(path):
    S = sgs()
    g = json.load(open(path))
    ref = g['klocal']
    n = g['n']
    r = S.klocal_search(S.HamHandle.load(os.path.join(GOLDEN, g['file'])))
    assert r['energy'] == pytest.approx(ref['energy'], rel=1e-10, abs=1e-12)
    assert r['generators'] == ref['generators']
    assert r['term_signs'] == ref['term_signs']
    assert r['sright_counts'] == [ref['sright_counts'][str(m)] for m in range(n + 1)]
    assert r['states_per_site'] == ref['states_per_site']
    assert r['transitions'] == ref['transitions']


def test_config3_golden():
    """BASELINE config 3: gen_klocal(64,4,3,0): E = -30.9918743342, 47 generators, md5 of the printed group"""
    S = sgs()
    r = S.klocal_search(S.HamHandle.from_text(G.klocal_text(64, 4, 3, 0)))
    assert r['energy'] == pytest.approx(-30.9918743342, abs=1e-9)
    assert len(r['generators']) == 47
    line = 'qubit 0-63 size (47,) ' + ' '.join(r['generators']) + '\n'
    assert hashlib.md5(line.encode()).hexdigest() == 'a715783786513531b60e5877f72a9f30'
    assert max(r['states_per_site']) == 113 and r['transitions'] == 2850


@pytest.mark.parametrize('n,k,per,seed', [(8, 2, 2, 0), (10, 3, 2, 1), (10, 3, 3, 2), (12, 4, 2, 3), (9, 4, 4, 4), (14, 3, 3, 5),
                                          (8, 4, 6, 3), (16, 5, 2, 6), (20, 4, 3, 7), (6, 6, 5, 8), (12, 1, 2, 9),
                                          (10, 7, 2, 1), (12, 8, 3, 3), (14, 10, 2, 4), (9, 9, 4, 5), (24, 6, 3, 6), (12, 8, 6, 7),
                                          (16, 7, 5, 8), (18, 12, 3, 9), (20, 15, 2, 10)])
def test_random_vs_oracle(n, k, per, seed):
    S = sgs()
    per = min(per, 4 ** k - 1)
    text = G.klocal_text(n, k, per, seed)
    ref = O.klocal(text=text, threads=4)
    r = S.klocal_search(S.HamHandle.from_text(text))
    assert r['energy'] == pytest.approx(ref['energy'], rel=1e-10, abs=1e-12)
    assert r['generators'] == ref['generators']
    assert r['term_signs'] == ref['term_signs']
    assert r['sright_counts'] == ref['sright_counts']
    assert r['states_per_site'] == ref['states_per_site']
    assert r['transitions'] == ref['transitions']


@pytest.mark.parametrize('n,k,per,seed', [(8, 3, 2, 0), (10, 3, 2, 1), (9, 2, 2, 4)])
def test_klocal_equals_sparse_on_gpu(n, k, per, seed):
    S = sgs()
    text = G.klocal_text(n, k, per, seed)
    a = S.sparse_search(S.HamHandle.from_text(text))
    b = S.klocal_search(S.HamHandle.from_text(text))
    assert a['energy'] == pytest.approx(b['energy'], rel=1e-10)
    assert a['generators'] == b['generators'] and a['term_signs'] == b['term_signs']


def test_full_size_properties_n64_k4_dense_windows():
    """n = 64, k = 4 with 6 terms per window: no oracle golden; E must equal sum w * sign and the signs must be
    those of a commuting group built from Hamiltonian terms"""
    S = sgs()
    text = G.klocal_text(64, 4, 6, 1)
    H = S.HamHandle.from_text(text)
    r = S.klocal_search(H)
    rows, w, b, e = H.terms()
    assert r['energy'] == pytest.approx(sum(wi * si for wi, si in zip(w, r['term_signs'])), rel=1e-10)
    ref = O.klocal(text=text, threads=8)
    assert r['energy'] == pytest.approx(ref['energy'], rel=1e-10)
    assert r['generators'] == ref['generators'] and r['states_per_site'] == ref['states_per_site']


PERIODIC = sorted(f for f in glob.glob(os.path.join(GOLDEN, '*.json')) if 'periodic' in json.load(open(f)))


@pytest.mark.parametrize('path', PERIODIC, ids=_name)
def test_periodic_matches_python_reference(path):
    S = sgs()
    rec = json.load(open(path))
    g = rec['periodic']
    r = S.klocal_periodic(os.path.join(GOLDEN, rec['file']), cmax=g['cmax'], c=g['c'])
    assert (r['l'], r['k'], r['n'], r['nterms']) == (g['l'], g['k'], g['n'], g['nterms'])
    assert r['sright_log'] == g['sright_log']
    assert r['fixed_point_sizes'] == g['fixed_point_sizes']
    key = lambda o: (round(o['energy'], 9), o['begin'], o['end'], o['generators'])
    assert sorted(map(key, r['results'])) == sorted(map(key, g['results']))
    assert min(o['energy'] for o in r['results']) / g['c'] == pytest.approx(g['min_energy_per_period'])


def test_periodic_vs_oracle_other_cells(tmp_path):
    S = sgs()
    cases = ["1 2\n -1.0 ZZ 0 1\n -0.5 X 0\n", "2 3\n -1.0 XZX 0 1 2\n 0.7 YY 1 2\n -0.3 ZZ 0 1\n", "1 3\n 1.0 XZX 0 1 2\n -1.0 YY 1 2\n",
             "1 2\n 0.7 XY 0 1\n -0.4 Z 0\n 0.2 X 0\n", "2 2\n -1.0 ZZ 0 1\n 0.6 XX 1 2\n -0.25 Y 1\n", "1 3\n -0.9 ZXZ 0 1 2\n 0.35 XX 0 1\n"]
    for i, text in enumerate(cases):
        p = tmp_path / f'per{i}.txt'
        p.write_text(text)
        ref = O.periodic(str(p), cmax=10, c=2)
        r = S.klocal_periodic(str(p), cmax=10, c=2)
        assert r['sright_log'] == ref['sright_log'] and r['fixed_point_sizes'] == ref['fixed_point_sizes']
        key = lambda o: (round(o['energy'], 9), o['begin'], o['end'], o['generators'])
        assert sorted(map(key, r['results'])) == sorted(map(key, ref['results']))


@pytest.mark.parametrize('env', [{'SGS_DP_CAPS': '16'}, {'SGS_SR_CAP': '64'}, {'SGS_SR_POOL_KB': '16'},
                                 {'SGS_DP_CAPS': '8', 'SGS_SR_CAP': '32', 'SGS_SR_POOL_KB': '8'}])
def test_capacity_overflow_is_redone_with_larger_buffers(monkeypatch, env):
    """Device-side capacity checks (states, children, right-environment branches, level pool) end the persistent
    kernels with an error code; the host grows the buffers and redoes the search: same result as the default."""
    S = sgs()
    for k, v in env.items():
        monkeypatch.setenv(k, v)
    r = S.klocal_search(S.HamHandle.from_text(G.klocal_text(64, 4, 3, 0)))
    assert r['energy'] == pytest.approx(-30.9918743342, abs=1e-9)
    assert len(r['generators']) == 47
    assert max(r['states_per_site']) == 113 and r['transitions'] == 2850
    g = json.load(open(os.path.join(GOLDEN, 'periodic_example.json')))['periodic']
    p = S.klocal_periodic(os.path.join(ROOT, 'examples', 'hamiltonian_periodic.txt'), cmax=g['cmax'], c=g['c'])
    key = lambda o: (o['energy'], o['begin'], o['end'], o['generators'])
    assert sorted(map(key, p['results'])) == sorted(map(key, g['results']))


@pytest.mark.parametrize('n,k,per,seed', [(10, 3, 3, 2), (8, 4, 6, 3), (24, 6, 3, 6), (12, 8, 6, 7), (16, 7, 5, 8), (12, 1, 2, 9)])
def test_right_environment_waves_equal_paths(monkeypatch, n, k, per, seed):
    """k_sr_run builds a level's branches either with one thread per root-to-leaf path of the include/exclude
    tree (default when the site has few terms) or as breadth-first waves (SGS_SR_WAVES=1, and automatically for
    many terms per site): same levels, same search; also with a scratch capacity that overflows first."""
    S = sgs()
    text = G.klocal_text(n, k, min(per, 4 ** k - 1), seed)
    ref = O.klocal(text=text, threads=4)
    for env in ({}, {'SGS_SR_WAVES': '1'}, {'SGS_SR_CAP': '32'}, {'SGS_SR_WAVES': '1', 'SGS_SR_CAP': '32'}):
        for name in ('SGS_SR_WAVES', 'SGS_SR_CAP'):
            monkeypatch.delenv(name, raising=False)
        for name, v in env.items():
            monkeypatch.setenv(name, v)
        r = S.klocal_search(S.HamHandle.from_text(text))
        assert r['energy'] == pytest.approx(ref['energy'], rel=1e-10, abs=1e-12), env
        assert r['generators'] == ref['generators'] and r['term_signs'] == ref['term_signs'], env
        assert r['sright_counts'] == ref['sright_counts'] and r['states_per_site'] == ref['states_per_site'], env
        assert r['transitions'] == ref['transitions'], env


@pytest.mark.parametrize('n,k,per,seed', [(10, 3, 3, 2), (8, 4, 6, 3), (24, 6, 3, 6), (16, 7, 5, 8), (40, 4, 6, 1)])
def test_small_batch_modes_equal_grid_mode(monkeypatch, n, k, per, seed):
    """The persistent kernels run as one CTA (__syncthreads between the phases), then as one thread-block cluster
    (hardware cluster barrier) while the batches are small, and hand over to the cooperative grid otherwise
    (SGS_DP_BLOCK=0 / SGS_DP_CLUSTER=0 turn a mode off, SGS_DP_CLUSTER=8 is the portable cluster size).  Same
    levels, states and result in every combination; (40, 4, 6) outgrows both small modes in mid-run."""
    S = sgs()
    text = G.klocal_text(n, k, min(per, 4 ** k - 1), seed)
    ref = O.klocal(text=text, threads=4)
    for env in ({}, {'SGS_DP_BLOCK': '0'}, {'SGS_DP_CLUSTER': '0'}, {'SGS_DP_BLOCK': '0', 'SGS_DP_CLUSTER': '0'},
                {'SGS_DP_BLOCK': '0', 'SGS_DP_CLUSTER': '8'}):
        for name in ('SGS_DP_BLOCK', 'SGS_DP_CLUSTER'):
            monkeypatch.delenv(name, raising=False)
        for name, v in env.items():
            monkeypatch.setenv(name, v)
        r = S.klocal_search(S.HamHandle.from_text(text))
        assert r['energy'] == pytest.approx(ref['energy'], rel=1e-10, abs=1e-12), env
        assert r['generators'] == ref['generators'] and r['term_signs'] == ref['term_signs'], env
        assert r['sright_counts'] == ref['sright_counts'] and r['states_per_site'] == ref['states_per_site'], env
        assert r['transitions'] == ref['transitions'], env


@pytest.mark.parametrize('n,k,per,seed', [(8, 3, 18, 1), (9, 3, 20, 3), (7, 2, 12, 2)])
def test_many_terms_per_site_vs_oracle(n, k, per, seed):
    """18-20 terms share a last site: more than the thread-per-path mode of k_sr_run takes (16), so the levels
    are built by breadth-first waves without being asked to, next to path-mode levels at the chain's ends."""
    S = sgs()
    text = G.klocal_text(n, k, min(per, 4 ** k - 1), seed)
    ref = O.klocal(text=text, threads=4)
    r = S.klocal_search(S.HamHandle.from_text(text))
    assert r['energy'] == pytest.approx(ref['energy'], rel=1e-10, abs=1e-12)
    assert r['generators'] == ref['generators'] and r['term_signs'] == ref['term_signs']
    assert r['sright_counts'] == ref['sright_counts'] and r['states_per_site'] == ref['states_per_site']
    assert r['transitions'] == ref['transitions']


def test_per_site_api_matches_one_shot():
    """sgs_klocal_create / evolve / nstates / ground_state (the loop of cpp/klocal_sparse.cpp) = sgs_klocal_search"""
    S = sgs()
    H = S.HamHandle.from_text(G.klocal_text(20, 4, 3, 7))
    one = S.klocal_search(H)
    from stabilizer_gs_b200.state import Hamiltonian, StateMachine, generate_Sright_all
    import tempfile
    with tempfile.NamedTemporaryFile('w', suffix='.txt', delete=False) as f:
        f.write(G.klocal_text(20, 4, 3, 7))
    Hm = Hamiltonian.from_file(f.name)
    SM = StateMachine(Hm, generate_Sright_all(Hm))
    sizes = [len(SM.state_dict)]
    while SM.m < Hm.n:
        SM.evolve()
        sizes.append(len(SM.state_dict))
    SM.generate_gs()
    assert sizes == one['states_per_site']
    assert SM.energy_gs == pytest.approx(one['energy'], rel=1e-12)
    os.unlink(f.name)


@pytest.mark.parametrize('kind,k', [('tfim8', 2), ('heisenberg6', 2), ('cluster7', 3), ('tfim8', 4)])
def test_structured_chains_vs_oracle(kind, k):
    """degenerate, commuting-heavy chains through the DP: exact ties must be broken like the oracle does"""
    import test_gpu_sparse as TS
    S = sgs()
    text = TS._structured(kind)
    n = int(text.split()[0])
    text = f"{n} {k}\n" + text.split('\n', 1)[1]
    ref = O.klocal(text=text, threads=2)
    r = S.klocal_search(S.HamHandle.from_text(text))
    assert r['energy'] == pytest.approx(ref['energy'], rel=1e-10, abs=1e-12)
    assert r['generators'] == ref['generators']
    assert r['term_signs'] == ref['term_signs']
    assert r['sright_counts'] == ref['sright_counts'] and r['states_per_site'] == ref['states_per_site']
    sp = S.sparse_search(S.HamHandle.from_text(text))
    assert sp['energy'] == pytest.approx(r['energy'], rel=1e-10, abs=1e-12)
