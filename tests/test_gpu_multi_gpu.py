"""Tests that need more than one GPU (skipped on a one-GPU box; `gpurun --gpus 2 -- python -m pytest tests/test_gpu_multi_gpu.py -m gpu`).

* multi-GPU k-local DP (reference: the omp-parallel state loop of StateMachine::evolve,
  cpp/state.cpp:494-517): sgs_klocal_search with opts.ngpus = N moves each site's states in N slices, exchanges the
  results with NCCL and must reproduce the one-GPU run bit for bit — energy, generators, term signs, states per site,
  transitions — on dense-window instances and against the CPU oracle;
* sparse search with opts.ngpus = N (ncclCommInitAll): same answer and counts as one GPU.
"""
import pytest

import gen_hamiltonians as G
import oracle_py as O

pytestmark = pytest.mark.gpu


def sgs():
    import stabilizer_gs_b200 as S
    return S


def ngpu():
    return sgs()._capi.lib().sgs_device_count()


needs2 = pytest.mark.skipif(ngpu() < 2, reason='needs at least two GPUs')
KEYS = ('energy', 'generators', 'term_signs', 'states_per_site', 'sright_counts', 'transitions')


@needs2
@pytest.mark.parametrize('n,k,per,seed', [(16, 4, 3, 1), (20, 5, 5, 2), (18, 6, 8, 3), (64, 4, 6, 0)])
def test_multi_gpu_dp_equals_one_gpu(n, k, per, seed):
    S = sgs()
    H = S.HamHandle.from_text(G.klocal_text(n, k, per, seed))
    one = S.klocal_search(H)
    counts = [c for c in (2, 3, 4, 8) if c <= ngpu()]
    for c in counts:
        r = S.klocal_search(H, ngpus=c)
        for key in KEYS:
            assert r[key] == one[key], (c, key)
    assert max(one['states_per_site']) >= 16


@needs2
def test_multi_gpu_dp_vs_oracle():
    S = sgs()
    text = G.klocal_text(14, 4, 4, 5)
    ref = O.klocal(text=text, threads=4)
    r = S.klocal_search(S.HamHandle.from_text(text), ngpus=2)
    assert r['energy'] == pytest.approx(ref['energy'], rel=1e-10)
    assert r['generators'] == ref['generators'] and r['term_signs'] == ref['term_signs']
    assert r['states_per_site'] == ref['states_per_site'] and r['transitions'] == ref['transitions']


@needs2
def test_multi_gpu_sparse_single_process():
    S = sgs()
    H = S.HamHandle.from_text(G.sparse_text(20, 100, 0))
    one = S.sparse_search(H)
    for c in [c for c in (2, 4, 8) if c <= ngpu()]:
        plan = S.SparsePlan(H, ngpus=c)
        a, b = plan.run(), plan.run()
        plan.close()
        for r in (a, b):
            assert r['energy'] == one['energy'] and r['generators'] == one['generators'] and r['term_signs'] == one['term_signs']
            assert r['candidates'] == one['candidates'] and r['rank_hist'] == one['rank_hist']
