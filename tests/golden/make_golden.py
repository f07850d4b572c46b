#!/usr/bin/env python
"""Generate golden vectors by running the UNMODIFIED Python reference.

Run in the build container only (it needs /root/reference/py, which does not
exist on the GPU box):

    python tests/golden/make_golden.py [--only NAME ...]

It imports the reference's own classes (py/state.py, py/stab.py, py/pauli.py,
py/linear.py) and records, for every instance listed below, what the three
reference scripts compute:

* sparse        (py/sparse.py:1-15, FullStateEnergies  py/state.py:88-147)
* klocal        (py/klocal_sparse.py:1-19, StateMachine py/state.py:328-399)
* periodic      (py/klocal_periodic.py:1-63)
* linear KATs   (inputs of cpp/test_linear.cpp, evaluated with py/linear.py)

Output: tests/golden/<name>.json (+ the instance text under tests/golden/<name>.txt).
Nothing here is imported by the product; tests read the JSON only.
"""
import argparse
import contextlib
import io
import json
import os
import sys
import time
from collections import Counter

HERE = os.path.dirname(os.path.abspath(__file__))
REF = '/root/reference'
sys.path.insert(0, os.path.join(REF, 'py'))
sys.path.insert(0, os.path.dirname(HERE))

import warnings
warnings.simplefilter('ignore')
import numpy as np  # noqa: E402
import linear  # noqa: E402
from pauli import Pauli  # noqa: E402
from stab import StabilizerGroup  # noqa: E402
import state as refstate  # noqa: E402
from state import (Hamiltonian, FullStateEnergies, StateMachine,  # noqa: E402
                   StateMachinePeriodic, generate_Sright_all, generate_Sright_periodic)
import gen_hamiltonians as G  # noqa: E402

# tqdm progress bars of the reference are noise here
refstate.tqdm.tqdm = lambda x, *a, **k: x


def stab_lines(stab_or_pauli, n):
    """signed Pauli strings over [0,n) in the C++ `to_string(true)` style."""
    Ps = stab_or_pauli.Ps if hasattr(stab_or_pauli, 'Ps') else stab_or_pauli
    Ps = Ps.range(0, n)
    out = []
    a4 = Ps.array4
    ph = Ps.phase
    for i in range(a4.shape[0]):
        out.append(('+' if ph[i] == 1 else '-') + ''.join('IXZY'[v] for v in a4[i]))
    return out


def cpp_order(H):
    """C++ `Hamiltonian::paulis_list` order (cpp/state.cpp:88-96)."""
    return [(P.range(0, H.n), w) for q in range(H.n) for P, w in H.paulis[q]]


def run_sparse(H, order):
    paulis = H.original_paulis if order == 'file' else cpp_order(H)
    full = FullStateEnergies(paulis)
    # same loop as FullStateEnergies.evolve (py/state.py:122-147) but keeping the leaves
    states = [full.copy()]
    levels = []
    while not all(len(s.paulis) == 0 for s in states):
        new_states = []
        for s in states:
            if len(s.paulis) == 0:
                new_states.append(s)
                continue
            P, _ = s.paulis[0]
            new_states.append(s.include_pauli(P))
            new_states.append(s.exclude_pauli(P))
        states = new_states
        levels.append(len(states))
    ranks = Counter(int(s.stab.m) for s in states)
    sum2m = int(sum(2 ** s.stab.m for s in states))
    keys = set(s.stab.Ps.hash(add_position=False) for s in states)
    # the reference's own result through its public call
    with contextlib.redirect_stdout(io.StringIO()):
        full_Ps, E = FullStateEnergies(paulis).evolve()
    stab = StabilizerGroup(full_Ps)
    signs = [int(stab.get_energy(P)) for P, w in H.original_paulis]
    echeck = float(sum(stab.get_energy(P) * w for P, w in H.original_paulis))
    return dict(order=order, energy=float(E), generators=stab_lines(stab, H.n), term_signs=signs,
                energy_from_signs=echeck, candidates=len(states), distinct_groups=len(keys),
                levels=levels, ranks={str(k): v for k, v in sorted(ranks.items())}, sum_2m=sum2m)


def run_klocal(H):
    with contextlib.redirect_stdout(io.StringIO()) as buf:
        Sright_all = generate_Sright_all(H)
    sright_counts = {str(m): len(Sright_all[m][0]) for m in range(H.n + 1)}
    links = {str(m): int(sum(len(v) for v in Sright_all[m][1].values())) for m in range(H.n)}
    SM = StateMachine(H, Sright_all)
    nstates = [len(SM.state_dict)]
    ntrans = 0
    while True:
        ntrans += len(SM.state_dict)
        SM.evolve()
        nstates.append(len(SM.state_dict))
        if SM.m == H.n:
            break
    SM.generate_gs()
    eref = float(sum(SM.stab_gs.get_energy(P) * w for P, w in H.original_paulis))
    signs = [int(SM.stab_gs.get_energy(P)) for P, w in H.original_paulis]
    return dict(energy=float(SM.energy_gs), energy_from_signs=eref, generators=stab_lines(SM.stab_gs, H.n),
                term_signs=signs, sright_counts=sright_counts, sright_links=links,
                states_per_site=nstates, transitions=ntrans,
                terms_per_last_site={str(q): len(H.paulis[q]) for q in range(H.n)})


def run_periodic(path, cmax=10, c=3):
    """py/klocal_periodic.py:1-63 with the path/cmax/c made arguments (the script hard-codes them)."""
    max_float = 1e100
    Hp = Hamiltonian.from_file_periodic(path, cmax)
    log = io.StringIO()

    def periodic_evolve_states(states, clear_history=True, c=1):
        sp.state_dict = {hash(s): s for s in states}
        sp.m = Hp.k + Hp.l - 1
        for _ in range(Hp.l * c):
            sp.evolve()
        return [s.shift(-Hp.l * c, clear_history=clear_history) for s in sp.state_dict.values()]

    with contextlib.redirect_stdout(log):
        Sr = generate_Sright_periodic(Hp, cmax)
    sp = StateMachinePeriodic(Hp, Sr)
    for _ in range(Hp.k + Hp.l * 2 - 1):
        sp.evolve()
    states_hash = None
    states = [s.shift(-Hp.l, clear_history=True) for s in sp.state_dict.values()]
    fixed_sizes = []
    while True:
        states = periodic_evolve_states(states)
        new_hash = tuple(sorted(map(hash, states)))
        fixed_sizes.append(len(new_hash))
        if new_hash == states_hash:
            break
        states_hash = new_hash
    results = []
    for us in states.copy():
        ns = us.copy()
        sa = {hash(s): s for s in periodic_evolve_states([ns.copy()], c=c)}
        if hash(ns) not in sa:
            continue
        shape = ns.stab.energy_level.Es.shape
        indices = np.indices(shape).reshape((len(shape), 2 ** len(shape))).T
        for index in indices:
            Es = np.full(shape, max_float)
            Es[tuple(index)] = 0.0
            ns.stab.energy_level.Es = Es
            sa = {hash(s): s for s in periodic_evolve_states([ns], clear_history=False, c=c)}
            if hash(ns) in sa:
                fs = sa[hash(ns)]
                res = fs.stab.energy_level.get_state(tuple(index), Hp.original_paulis, return_stab=False)
                results.append(res)
    out = []
    for P, e in results:
        a4 = P.array4
        ph = P.phase
        gens = [('+' if ph[i] == 1 else '-') + ''.join('IXZY'[v] for v in a4[i]) for i in range(a4.shape[0])]
        out.append(dict(energy=float(e), begin=int(P.begin), end=int(P.end), generators=gens))
    out.sort(key=lambda r: (r['energy'], r['begin'], r['end'], r['generators']))
    sright_log = [list(map(int, ln.split())) for ln in log.getvalue().strip().splitlines()]
    return dict(l=Hp.l, k=Hp.k, n=Hp.n, cmax=cmax, c=c, nterms=len(Hp.original_paulis), sright_log=sright_log,
                fixed_point_sizes=fixed_sizes, results=out,
                min_energy_per_period=min(r['energy'] for r in out) / c)


def linear_kats():
    """Inputs from cpp/test_linear.cpp, expected values from py/linear.py (SURVEY KAT-5)."""
    A1 = np.array([[1, 0, 1], [0, 1, 1], [1, 1, 0]])
    A2, b2 = linear.Gauss_elimination(A1, np.eye(3, dtype=int), 3)
    A3 = np.array([[1, 0, 1], [0, 1, 1], [1, 1, 1]])
    over, under, x = linear.solve_modN(A3, np.eye(3, dtype=int), 2)
    ker = linear.get_kernel(np.array([[1, 0, 1], [0, 1, 1]]), 2)
    Ad = np.array([[1, 1, 0, 0], [1, 1, 0, 0], [0, 0, 1, 1], [1, 1, 1, 1]])
    R, U = linear.Gauss_elimination_full(Ad, 2, allow_reorder=True)
    rng = np.random.default_rng(1234)
    rand = []
    for _ in range(40):
        r, c = int(rng.integers(1, 9)), int(rng.integers(1, 13))
        A = rng.integers(0, 2, size=(r, c))
        Rr, Ur = linear.Gauss_elimination_full(A, 2, allow_reorder=True)
        K = linear.get_kernel(A, 2)
        rand.append(dict(A=A.tolist(), rref=Rr.tolist(), U=Ur.tolist(), kernel=K.tolist()))
    return dict(
        mod=dict(x=[-2, -1, 0, 1], N=2, out=[int(v) % 2 for v in [-2, -1, 0, 1]]),
        inv_list_7=linear.get_inv_list(7).tolist(),
        get_order=dict(A=[[0, 1, 0], [0, 0, 0], [2, 0, 0]],
                       out=linear.get_order(np.array([[0, 1, 0], [0, 0, 0], [2, 0, 0]])).tolist()),
        gauss_mod3=dict(A=A1.tolist(), A2=A2.tolist(), b2=b2.tolist()),
        solve_mod2=dict(A=A3.tolist(), is_over=[bool(v) for v in over], is_under=bool(under), x=x.tolist()),
        kernel_mod2=dict(A=[[1, 0, 1], [0, 1, 1]], kernel=ker.tolist()),
        coset=dict(A2=[[1, 0, 0], [0, 1, 1]], x=[0, 1, 0],
                   out=linear.simplify_coset_representative(np.array([[1, 0, 0], [0, 1, 1]]), np.array([0, 1, 0]), 2).tolist()),
        rank_deficient=dict(A=Ad.tolist(), rref=R.tolist(), U=U.tolist()),
        random_mod2=rand,
    )


def pauli_kats():
    """Random products / commutation / canonical forms from py/pauli.py + py/stab.py."""
    rng = np.random.default_rng(99)
    np.random.seed(99)
    out = []
    for _ in range(40):
        L = int(rng.integers(1, 9))
        z1, x1, z2, x2 = (rng.integers(0, 2, size=L) for _ in range(4))
        P = Pauli(z1, x1, phase=np.array(1 - 2 * int(rng.integers(0, 2))))
        Q = Pauli(z2, x2, phase=np.array(1 - 2 * int(rng.integers(0, 2))))
        com = bool(P.commute(Q))
        rec = dict(L=L, z1=z1.tolist(), x1=x1.tolist(), p1=int(P.phase), z2=z2.tolist(), x2=x2.tolist(),
                   p2=int(Q.phase), commute=com)
        if com:
            R = P.dot(Q)
            rec.update(prod_z=R.z.tolist(), prod_x=R.x.tolist(), prod_phase=int(R.phase))
        out.append(rec)
    groups = []
    for _ in range(30):
        nq = int(rng.integers(2, 7))
        m = int(rng.integers(1, nq + 1))
        st = StabilizerGroup.random_stab(nq, m)
        raw = st.Ps
        # scramble the generating set: multiply random subsets, keep independence by unit-triangular mixing
        U = np.triu(rng.integers(0, 2, size=(m, m)), 1) + np.eye(m, dtype=int)
        perm = rng.permutation(m)
        U = U[perm]
        mixed = [raw[U[i] > 0].reduce_dot() for i in range(m)]
        mixedP = Pauli.stack(mixed)
        canon = StabilizerGroup(mixedP)
        test = Pauli.random(nq)
        dec = canon.decompose(test) if canon.commute(test) else None
        member = raw[rng.integers(0, 2, size=m) > 0]
        memrec = None
        if member.batch_size[0] > 0:
            mp = member.reduce_dot()
            x, ph = canon.decompose(mp)
            memrec = dict(line=stab_lines(mp[None], nq)[0], x=[int(v) for v in x], sign=int(ph))
        lo = int(rng.integers(0, nq))
        hi = int(rng.integers(lo + 1, nq + 1))
        proj = canon.project(lo, hi)
        groups.append(dict(n=nq, input=stab_lines(mixedP, nq), canonical=stab_lines(canon, nq),
                           probe=stab_lines(test[None], nq)[0], probe_commutes=bool(canon.commute(test)),
                           probe_in_group=dec is not None,
                           probe_sign=(int(dec[1]) if dec is not None else 0), member=memrec,
                           project=dict(begin=lo, end=hi,
                                        generators=[s[0] + s[1 + lo:1 + hi] for s in stab_lines(proj, nq)])))
    return dict(products=out, groups=groups)


INSTANCES = {
    # name: (kind, text-or-path, what to run)
    'finite_example': ('file', os.path.join(REF, 'example/hamiltonian_finite.txt'), ('sparse', 'klocal')),
    'kat3_sparse_6_12_1': ('text', lambda: G.sparse_text(6, 12, 1), ('sparse',)),
    'kat4_klocal_10_3_2_1': ('text', lambda: G.klocal_text(10, 3, 2, 1), ('sparse', 'klocal')),
    'sparse_6_20_1': ('text', lambda: G.sparse_text(6, 20, 1), ('sparse',)),
    'sparse_8_20_1': ('text', lambda: G.sparse_text(8, 20, 1), ('sparse',)),
    'sparse_8_30_1': ('text', lambda: G.sparse_text(8, 30, 1), ('sparse',)),
    'sparse_10_30_1': ('text', lambda: G.sparse_text(10, 30, 1), ('sparse',)),
    'sparse_3_14_5': ('text', lambda: G.sparse_text(3, 14, 5), ('sparse',)),
    'sparse_4_16_7': ('text', lambda: G.sparse_text(4, 16, 7), ('sparse',)),
    'klocal_10_4_3_2': ('text', lambda: G.klocal_text(10, 4, 3, 2), ('klocal',)),
    'klocal_8_4_6_3': ('text', lambda: G.klocal_text(8, 4, 6, 3), ('klocal',)),
    'klocal_12_3_3_5': ('text', lambda: G.klocal_text(12, 3, 3, 5), ('klocal',)),
    'klocal_9_2_2_4': ('text', lambda: G.klocal_text(9, 2, 2, 4), ('sparse', 'klocal')),
    'klocal_64_4_3_0': ('text', lambda: G.klocal_text(64, 4, 3, 0), ('klocal',)),
    'periodic_example': ('file', os.path.join(REF, 'example/hamiltonian_periodic.txt'), ('periodic',)),
    # more unit cells for the periodic path (the reference ships one example only); run with cmax=10, c=2
    'periodic_tfim_1_2': ('text', lambda: "1 2\n -1.0 ZZ 0 1\n -0.5 X 0\n", ('periodic', 10, 2)),
    'periodic_cell2_2_3': ('text', lambda: "2 3\n -1.0 XZX 0 1 2\n 0.7 YY 1 2\n -0.3 ZZ 0 1\n", ('periodic', 10, 2)),
    'periodic_frustrated_1_3': ('text', lambda: "1 3\n 1.0 XZX 0 1 2\n -1.0 YY 1 2\n", ('periodic', 10, 2)),
    # second batch of random instances (other seeds, k = 5, deeper sparse trees)
    'sparse_7_24_3': ('text', lambda: G.sparse_text(7, 24, 3), ('sparse',)),
    'sparse_5_20_9': ('text', lambda: G.sparse_text(5, 20, 9), ('sparse',)),
    'sparse_9_26_2': ('text', lambda: G.sparse_text(9, 26, 2), ('sparse',)),
    'klocal_12_5_2_7': ('text', lambda: G.klocal_text(12, 5, 2, 7), ('klocal',)),
    'klocal_14_3_4_9': ('text', lambda: G.klocal_text(14, 3, 4, 9), ('klocal',)),
    'klocal_11_2_3_6': ('text', lambda: G.klocal_text(11, 2, 3, 6), ('klocal',)),   # 233 240 sparse leaves: too slow for the Python sparse search
}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--only', nargs='*')
    args = ap.parse_args()
    names = args.only or (list(INSTANCES) + ['linear_kats', 'pauli_kats'])
    for name in names:
        t0 = time.time()
        if name == 'linear_kats':
            rec = linear_kats()
        elif name == 'pauli_kats':
            rec = pauli_kats()
        else:
            kind, src, runs = INSTANCES[name]
            text = open(src).read() if kind == 'file' else src()
            path = os.path.join(HERE, name + '.txt')
            with open(path, 'w') as f:
                f.write(text)
            rec = dict(name=name, file=name + '.txt')
            if 'periodic' in runs:
                rec['periodic'] = run_periodic(path, *runs[1:])
            else:
                H = Hamiltonian.from_file(path)
                rec.update(n=H.n, k=H.k, nterms=len(H.original_paulis))
                if 'sparse' in runs:
                    rec['sparse'] = run_sparse(H, 'cpp')
                    rec['sparse_file_order'] = run_sparse(H, 'file')
                if 'klocal' in runs:
                    rec['klocal'] = run_klocal(H)
        rec['generated_by'] = 'tests/golden/make_golden.py (unmodified /root/reference/py)'
        rec['seconds'] = round(time.time() - t0, 2)
        with open(os.path.join(HERE, name + '.json'), 'w') as f:
            json.dump(rec, f, indent=1, sort_keys=True)
        print(name, 'done in', rec['seconds'], 's', flush=True)


if __name__ == '__main__':
    main()
