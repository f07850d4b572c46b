"""GPU parity tests of the sparse search: CUDA path (through the C ABI) vs the CPU oracle and the
golden vectors of the Python reference.  Bit-exact for generators / signs / counts; energy to 1e-10."""
import glob
import json
import os

import numpy as np
import pytest

import gen_hamiltonians as G
import oracle_py as O
from conftest import GOLDEN

pytestmark = pytest.mark.gpu

# (oracle_sparse_*.json are the oracle-generated mid-size goldens of test_gpu_sparse_midsize.py)
SPARSE = sorted(f for f in glob.glob(os.path.join(GOLDEN, '*.json'))
                if not os.path.basename(f).startswith('oracle_') and 'sparse' in json.load(open(f)))


def _name(p):
    return os.path.basename(p)[:-5]


def sgs():
    import stabilizer_gs_b200 as S
    return S


def check_same(gpu, ref, counts=True):
    assert gpu['energy'] == pytest.approx(ref['energy'], rel=1e-10, abs=1e-12)
    assert gpu['generators'] == ref['generators']
    assert gpu['term_signs'] == ref['term_signs']
    if counts:
        assert gpu['candidates'] == ref['candidates']
        assert {int(k): v for k, v in gpu['rank_hist'].items()} == {int(k): v for k, v in ref['rank_hist'].items()}


@pytest.mark.parametrize('depth', [0, 1, 2, 3])
@pytest.mark.parametrize('path', SPARSE, ids=_name)
def test_golden_python_reference(path, depth):
    S = sgs()
    g = json.load(open(path))
    ref = dict(g['sparse'])
    ref['rank_hist'] = ref['ranks']
    H = S.HamHandle.load(os.path.join(GOLDEN, g['file']))
    r = S.sparse_search(H, expand_depth=depth)
    check_same(r, ref)
    assert r['sum_2m'] == ref['sum_2m']


@pytest.mark.parametrize('depth', [0, 1, 2])
@pytest.mark.parametrize('n,T,seed', [(5, 12, 0), (6, 18, 3), (7, 22, 4), (8, 26, 5), (9, 30, 6), (10, 34, 7), (12, 36, 8),
                                      (12, 30, 9), (3, 12, 10), (4, 20, 11), (2, 9, 12), (16, 40, 13)])
def test_random_vs_oracle(n, T, seed, depth):
    S = sgs()
    T = min(T, 4 ** n - 1)
    text = G.sparse_text(n, T, seed)
    ref = O.sparse(text=text, threads=4)
    r = S.sparse_search(S.HamHandle.from_text(text), expand_depth=depth)
    check_same(r, ref)


@pytest.mark.parametrize('n,k,per,seed', [(8, 3, 2, 0), (10, 3, 3, 1), (9, 4, 2, 2), (12, 2, 2, 3)])
def test_klocal_inputs_vs_oracle(n, k, per, seed):
    """k-local (structured, dependency-heavy) Hamiltonians through the sparse path"""
    S = sgs()
    text = G.klocal_text(n, k, per, seed)
    ref = O.sparse(text=text, threads=4)
    for depth in (0, 1, 3):
        r = S.sparse_search(S.HamHandle.from_text(text), expand_depth=depth)
        check_same(r, ref)


def _tie_text(n, T, seed):
    """weights from {±1, ±0.5}: sums are exact in fp64, so ties are exact and the tie-break is tested"""
    rng = np.random.default_rng(seed)
    lines = G.sparse_text(n, T, seed).splitlines()
    out = [lines[0]]
    for ln in lines[1:]:
        parts = ln.split()
        w = rng.choice([-1.0, 1.0, -0.5, 0.5])
        out.append(f"{w: .10f} " + ' '.join(parts[1:]))
    return '\n'.join(out) + '\n'


@pytest.mark.parametrize('n,T,seed', [(4, 10, 0), (5, 14, 1), (6, 16, 2), (8, 24, 3), (10, 28, 4)])
def test_exact_ties_follow_documented_order(n, T, seed):
    S = sgs()
    text = _tie_text(n, T, seed)
    ref = O.sparse(text=text)
    for depth in (0, 1, 2):
        r = S.sparse_search(S.HamHandle.from_text(text), expand_depth=depth)
        check_same(r, ref)


def test_edge_cases():
    S = sgs()
    # a single term; a single qubit; duplicate strings; an identity-letter term; zero weights
    cases = [
        "1 1\n 0.5 Z 0\n",
        "1 1\n 0.5 Z 0\n -0.25 X 0\n 0.75 Y 0\n",
        "3 3\n 0.5 ZZ 0 1\n 0.25 ZZ 0 1\n -1.0 XX 0 1\n 0.3 IZ 1 2\n",
        "3 3\n 0.0 ZZ 0 1\n 0.0 XX 0 1\n 1.0 YY 0 1\n 0.0 Z 2\n",
    ]
    for text in cases:
        ref = O.sparse(text=text)
        for depth in (0, 1, 2):
            r = S.sparse_search(S.HamHandle.from_text(text), expand_depth=depth)
            check_same(r, ref)
    # an all-identity term: the reference makes the identity a (zero-row) "generator" (its canonical form
    # then carries a '+II' row, a degenerate input it never meant to support); here it is the constant it
    # is.  Energies and the signs of the real terms agree.
    text = "2 2\n 1.0 II 0 1\n -1.0 ZZ 0 1\n 0.5 XX 0 1\n"
    ref = O.sparse(text=text)
    for depth in (0, 1, 2):
        r = S.sparse_search(S.HamHandle.from_text(text), expand_depth=depth)
        assert r['energy'] == pytest.approx(ref['energy'], abs=1e-12)
        assert r['generators'] == [g for g in ref['generators'] if g != '+II']
        assert r['term_signs'][1:] == ref['term_signs'][1:]


def test_two_word_rows_vs_oracle():
    """n > 32 exercises the W = 2 kernels: few terms on many qubits"""
    S = sgs()
    for n, T, seed in [(40, 16, 0), (48, 18, 1), (33, 20, 2)]:
        text = G.sparse_text(n, T, seed)
        ref = O.sparse(text=text, threads=4)
        for depth in (0, 2):
            r = S.sparse_search(S.HamHandle.from_text(text), expand_depth=depth)
            check_same(r, ref)


def test_virtual_shards_partition_the_search():
    """rank r of world w without a communicator returns that shard only: counts add up, the minimum over
    shards is the global answer (what the NCCL argmin combines)."""
    S = sgs()
    text = G.sparse_text(10, 34, 21)
    full = S.sparse_search(S.HamHandle.from_text(text))
    for world in (2, 3, 8):
        parts = [S.sparse_search(S.HamHandle.from_text(text), rank=r, world=world) for r in range(world)]
        assert sum(p['candidates'] for p in parts) == full['candidates']
        best = min(parts, key=lambda p: p['energy'])
        assert best['energy'] == full['energy'] and best['generators'] == full['generators']


def _structured(kind):
    """small structured (non-random) Hamiltonians: many commuting terms, dependent products, degenerate weights"""
    lines = []
    if kind == 'tfim8':                       # transverse-field Ising chain
        n = 8
        lines += [f"-1.0 ZZ {i} {i + 1}" for i in range(n - 1)] + [f"-0.7 X {i}" for i in range(n)]
    elif kind == 'heisenberg6':               # XXX chain: XX, YY, ZZ on every bond (XX*YY = -ZZ: dependent members)
        n = 6
        for i in range(n - 1):
            lines += [f"{w} {p}{p} {i} {i + 1}" for p, w in (('X', 1.0), ('Y', 1.0), ('Z', 1.0))]
    elif kind == 'cluster7':                  # cluster-state stabilizers plus fields
        n = 7
        lines += [f"-1.0 ZXZ {i} {i + 1} {i + 2}" for i in range(n - 2)] + [f"0.3 Z {i}" for i in range(n)]
    elif kind == 'zonly9':                    # all terms commute: one clique, ranks up to n, dependent Z strings
        n = 9
        import random
        rnd = random.Random(5)
        seen = set()
        while len(lines) < 16:
            sites = sorted(rnd.sample(range(n), rnd.randint(1, 4)))
            if tuple(sites) in seen:
                continue
            seen.add(tuple(sites))
            lines.append(f"{rnd.uniform(-1, 1):.6f} {'Z' * len(sites)} {' '.join(map(str, sites))}")
    elif kind == 'toric2x2':                  # plaquette / star operators of a small torus (commuting, dependent) + fields
        n = 8
        stars = [(0, 1, 4, 6), (0, 1, 5, 7), (2, 3, 4, 6), (2, 3, 5, 7)]
        plaqs = [(0, 2, 4, 5), (1, 3, 4, 5), (0, 2, 6, 7), (1, 3, 6, 7)]
        lines += [f"-1.0 XXXX {' '.join(map(str, q))}" for q in stars] + [f"-1.0 ZZZZ {' '.join(map(str, q))}" for q in plaqs]
        lines += [f"0.25 Z {i}" for i in range(0, n, 2)]
    return f"{n} {n}\n" + "".join(" " + l + "\n" for l in lines)


@pytest.mark.parametrize('kind', ['tfim8', 'heisenberg6', 'cluster7', 'zonly9', 'toric2x2'])
@pytest.mark.parametrize('depth', [0, 1])
def test_structured_hamiltonians_vs_oracle(kind, depth):
    """SURVEY §8(f)4 territory: commuting-heavy instances where most groups contain dependent terms (exact kernel,
    entangled sign variables, degenerate energies)."""
    S = sgs()
    text = _structured(kind)
    ref = O.sparse(text=text)
    r = S.sparse_search(S.HamHandle.from_text(text), expand_depth=depth)
    assert r['energy'] == pytest.approx(ref['energy'], rel=1e-10, abs=1e-12)
    assert r['candidates'] == ref['candidates']
    assert r['generators'] == ref['generators']
    assert r['term_signs'] == ref['term_signs']


@pytest.mark.parametrize('case', ['r10', 'r16', 'r20', 'w2', 'ties', 'tfim8', 'heisenberg6', 'zonly9', 'toric2x2'])
def test_branch_and_bound_returns_the_exhaustive_answer(case):
    """opts.prune: subtrees that cannot reach the incumbent are skipped — energy, generators and term signs are those
    of the exhaustive search (exact ties included), on one shard and on virtual shards; fewer groups are visited."""
    S = sgs()
    if case.startswith('r'):
        n = int(case[1:])
        text = G.sparse_text(n, {10: 34, 16: 70, 20: 130}[n], 3)
    elif case == 'w2':
        text = G.sparse_text(36, 90, 2)
    elif case == 'ties':
        import random
        rnd = random.Random(11)
        lines = set()
        while len(lines) < 40:
            lines.add(''.join(rnd.choice('IXYZ') for _ in range(8)))
        lines.discard('I' * 8)
        text = "8 8\n" + "".join(f" {rnd.choice([-1.0, -0.5, 0.5, 1.0])} {p} {' '.join(map(str, range(8)))}\n" for p in sorted(lines))
    else:
        text = _structured(case)
    H = S.HamHandle.from_text(text)
    full = S.sparse_search(H)
    for depth in (0, 1):
        r = S.sparse_search(H, prune=1, expand_depth=depth)
        assert r['energy'] == full['energy'] and r['generators'] == full['generators'] and r['term_signs'] == full['term_signs']
        assert r['candidates'] <= full['candidates']
    parts = [S.sparse_search(H, prune=1, rank=k, world=3) for k in range(3)]
    best = min(parts, key=lambda p: p['energy'])
    assert best['energy'] == full['energy']
    plan = S.SparsePlan(H, prune=1)
    a, b = plan.run(), plan.run()
    plan.close()
    assert (a['energy'], a['generators']) == (b['energy'], b['generators']) == (full['energy'], full['generators'])


@pytest.mark.parametrize('retry_cap', [None, '8'])
@pytest.mark.parametrize('n,T,seed,worlds', [(18, 110, 5, (2, 5)), (24, 300, 0, (8,))])
def test_multi_round_walk_hand_off(monkeypatch, retry_cap, n, T, seed, worlds):
    """With several shards, roots that cannot be certified at once (or are very heavy) hand their children to a
    second / third k_dfs round; a full hand-off queue (SGS_RETRY_CAP=8) falls back to walking locally.  Counts per
    rank add up to the single-shard run and the best shard holds the global answer, for the adaptive first run of a
    plan and for its recorded launch sequence (second run)."""
    S = sgs()
    H = S.HamHandle.from_text(G.sparse_text(n, T, seed))
    full = S.sparse_search(H)
    monkeypatch.setenv('SGS_SPLIT', '1')          # (the rounds are switched on from 5 shards by default)
    if retry_cap:
        monkeypatch.setenv('SGS_RETRY_CAP', retry_cap)
    for world in worlds:
        first, second = [], []
        for r in range(world):
            plan = S.SparsePlan(H, rank=r, world=world)
            first.append(plan.run())       # adaptive, host-synchronised run: roots owned by hash(root)
            second.append(plan.run())      # recorded launch sequence: last expansion level owned by hash(parent)
            plan.close()
        for parts in (first, second):      # two different, equally valid partitions of the same tree
            assert sum(p['candidates'] for p in parts) == full['candidates']
            merged = {}
            for p in parts:
                for k, v in p['rank_hist'].items():
                    merged[k] = merged.get(k, 0) + v
            assert {k: v for k, v in merged.items() if v} == {k: v for k, v in full['rank_hist'].items() if v}
            best = min(parts, key=lambda p: p['energy'])
            assert best['energy'] == full['energy'] and best['generators'] == full['generators']


def test_shard_without_candidates_is_empty_not_an_error():
    """a tiny instance split over more virtual shards than it has groups: the empty shards return energy = +inf"""
    S = sgs()
    text = "2 2\n 1.0 ZZ 0 1\n -0.5 XX 0 1\n"
    H = S.HamHandle.from_text(text)
    full = S.sparse_search(H)
    parts = [S.sparse_search(H, rank=r, world=8) for r in range(8)]
    assert sum(p['candidates'] for p in parts) == full['candidates']
    assert any(p['candidates'] == 0 and p['energy'] == float('inf') and p['m'] == 0 for p in parts)
    best = min(parts, key=lambda p: p['energy'])
    assert best['energy'] == full['energy'] and best['generators'] == full['generators']


def _check_invariants(S, text, r):
    H = S.HamHandle.from_text(text)
    rows, w, b, e = H.terms()
    assert r['energy'] == pytest.approx(sum(wi * si for wi, si in zip(w, r['term_signs'])), rel=1e-10)
    # generators pairwise commute and are in RREF order (strictly increasing pivots)
    gens = r['generators']
    def bits(s):
        v = 0
        for j, ch in enumerate(s[1:]):
            v |= {'I': 0, 'Z': 1, 'X': 2, 'Y': 3}[ch] << (2 * j)
        return v
    gb = [bits(s) for s in gens]
    ev = int('01' * 128, 2)
    sw = lambda x: ((x >> 1) & ev) | ((x & ev) << 1)
    for i in range(len(gb)):
        for j in range(i):
            assert bin(gb[i] & sw(gb[j])).count('1') % 2 == 0
    piv = [(x & -x).bit_length() - 1 for x in gb]
    assert piv == sorted(piv) and len(set(piv)) == len(piv)
    for i, p in enumerate(piv):
        for j, x in enumerate(gb):
            if i != j:
                assert not (x >> p) & 1


def test_full_size_config4_properties():
    """BASELINE config 4 (n=24, T=300): size-independent properties + invariance under the work split"""
    S = sgs()
    text = G.sparse_text(24, 300, 0)
    r = S.sparse_search(S.HamHandle.from_text(text))
    _check_invariants(S, text, r)
    assert r['slow_nodes'] <= 600000          # the exact kernel only handles the top of the tree here
    r2 = S.sparse_search(S.HamHandle.from_text(text), expand_depth=2)
    assert r2['candidates'] == r['candidates'] and r2['rank_hist'] == r['rank_hist']
    assert r2['energy'] == r['energy'] and r2['generators'] == r['generators']
    parts = [S.sparse_search(S.HamHandle.from_text(text), rank=k, world=4) for k in range(4)]
    assert sum(p['candidates'] for p in parts) == r['candidates']
    assert min(p['energy'] for p in parts) == r['energy']
    # the expected size of the search space (~8.0e7)
    assert 4e7 < r['candidates'] < 2e8


def test_plan_reuse_is_deterministic():
    S = sgs()
    text = G.sparse_text(14, 60, 5)
    plan = S.SparsePlan(S.HamHandle.from_text(text))
    a = plan.run()
    b = plan.run()
    plan.close()
    for k in ('energy', 'generators', 'term_signs', 'candidates', 'rank_hist'):
        assert a[k] == b[k]
    _check_invariants(S, text, a)


def test_errors():
    S = sgs()
    with pytest.raises(S.SgsError):
        S.HamHandle.load('/nonexistent/file.txt')
    with pytest.raises(S.SgsError):
        S.HamHandle.from_text("2 2\n 1.0 QZ 0 1\n")
    with pytest.raises(S.SgsError):
        S.HamHandle.from_text("2 2\n 1.0 ZZ 0\n")
