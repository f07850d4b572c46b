"""Edge cases and limits of the GPU paths: wide chains (4-word rows), k = 1, tiny inputs, explicit errors."""
import pytest

import gen_hamiltonians as G
import oracle_py as O

pytestmark = pytest.mark.gpu


def sgs():
    import stabilizer_gs_b200 as S
    return S


def _same_klocal(S, text):
    ref = O.klocal(text=text, threads=4)
    r = S.klocal_search(S.HamHandle.from_text(text))
    assert r['energy'] == pytest.approx(ref['energy'], rel=1e-10, abs=1e-12)
    assert r['generators'] == ref['generators'] and r['term_signs'] == ref['term_signs']
    assert r['states_per_site'] == ref['states_per_site'] and r['sright_counts'] == ref['sright_counts']


@pytest.mark.parametrize('n,k,per,seed', [(100, 3, 2, 0), (120, 2, 2, 1), (96, 4, 2, 2), (70, 5, 2, 3)])
def test_long_chains_multiword_rows(n, k, per, seed):
    """n > 64: absolute rows span 3-4 words; the DP works on origin-relative single words"""
    _same_klocal(sgs(), G.klocal_text(n, k, per, seed))


def test_tiny_and_degenerate_inputs():
    S = sgs()
    for text in ["1 1\n 0.5 Z 0\n", "2 1\n 0.5 Z 0\n -0.25 X 1\n", "3 2\n 1.0 ZZ 0 1\n 1.0 ZZ 1 2\n -0.5 X 1\n",
                 "4 2\n -1.0 XX 0 1\n -1.0 XX 1 2\n -1.0 XX 2 3\n -1.0 ZZ 0 1\n -1.0 ZZ 2 3\n"]:
        _same_klocal(S, text)
        ref = O.sparse(text=text)
        r = S.sparse_search(S.HamHandle.from_text(text))
        assert r['energy'] == pytest.approx(ref['energy'], abs=1e-12) and r['generators'] == ref['generators']
        assert r['candidates'] == ref['candidates']


def test_dense_window_many_terms_per_site():
    """up to 12 terms share a last site (2^12 right-environment branches per group)"""
    _same_klocal(sgs(), G.klocal_text(7, 3, 12, 5))


def test_locality_and_limit_errors():
    S = sgs()
    with pytest.raises(S.SgsError):          # a term wider than k (Hamiltonian::check_locality)
        S.klocal_search(S.HamHandle.from_text("4 2\n 1.0 ZIZ 0 1 2\n"))
    with pytest.raises(S.SgsError):          # k beyond the single-word window limit
        S.klocal_search(S.HamHandle.from_text("40 20\n 1.0 ZZ 0 19\n"))


def test_sparse_wide_rows_four_words():
    """n > 64 in the sparse search exercises the W = 4 kernels"""
    S = sgs()
    text = G.sparse_text(70, 14, 3)
    ref = O.sparse(text=text, threads=4)
    for depth in (0, 2):
        r = S.sparse_search(S.HamHandle.from_text(text), expand_depth=depth)
        assert r['energy'] == pytest.approx(ref['energy'], rel=1e-10)
        assert r['generators'] == ref['generators'] and r['candidates'] == ref['candidates']


def test_sparse_results_are_run_to_run_deterministic():
    S = sgs()
    plan = S.SparsePlan(S.HamHandle.from_text(G.sparse_text(30, 200, 4)))
    runs = [plan.run() for _ in range(4)]
    plan.close()
    for r in runs[1:]:
        assert r['candidates'] == runs[0]['candidates'] and r['rank_hist'] == runs[0]['rank_hist']
        assert r['energy'] == runs[0]['energy'] and r['generators'] == runs[0]['generators']
