#!/usr/bin/env python
"""Golden content of the `klocal_sparse -v 1` dumps, from the UNMODIFIED Python reference (/root/reference/py;
build container only): every right environment of every level (generate_Sright_all, py/state.py:526-539) and,
after every site, every DP state (StateMachine.evolve, py/state.py:328-399) as
(window group, right environment, energy table).  The reference's C++ prints the same objects
(cpp/klocal_sparse.cpp:28-37,47-60; State::to_string cpp/state.cpp:332-345) in a std::map order of boost hashes, so
the comparison is per site, as sorted lists.

    python tests/golden/make_golden_dump.py      →  tests/golden/klocal_dump_*.json
"""
import contextlib
import io
import json
import os
import sys
import warnings

HERE = os.path.dirname(os.path.abspath(__file__))
REF = '/root/reference'
sys.path.insert(0, os.path.join(REF, 'py'))
sys.path.insert(0, os.path.dirname(HERE))
warnings.simplefilter('ignore')
import numpy as np  # noqa: E402
import state as refstate  # noqa: E402
from state import Hamiltonian, StateMachine, generate_Sright_all  # noqa: E402
import gen_hamiltonians as G  # noqa: E402

refstate.tqdm.tqdm = lambda x, *a, **k: x


def group(P):
    """signed Pauli batch -> [begin, end, ['+XZ', ...]]"""
    a4 = np.asarray(P.array4).reshape(-1, P.length)
    ph = np.asarray(P.phase).reshape(-1)
    return [int(P.begin), int(P.end), [('+' if ph[i] == 1 else '-') + ''.join('IXZY'[v] for v in a4[i]) for i in range(a4.shape[0])]]


def dump(path):
    H = Hamiltonian.from_file(path)
    with contextlib.redirect_stdout(io.StringIO()):
        S = generate_Sright_all(H)
        SM = StateMachine(H, S)
        sites = []
        while SM.m < H.n:
            SM.evolve()
            sites.append(sorted([group(st.stab.Ps), group(st.Sright.Ps), [round(float(x), 9) for x in np.asarray(st.stab.energy_level.Es).reshape(-1)]]
                                for st in SM.state_dict.values()))
    sright = {str(m): sorted(group(s.Ps) for s in S[m][0].values()) for m in range(H.n)}
    return dict(n=H.n, k=H.k, sright=sright, sites=sites)


def main():
    cases = {'klocal_dump_finite_example': os.path.join(REF, 'example/hamiltonian_finite.txt')}
    tmp = os.path.join(HERE, 'klocal_dump_8_3_2_5.txt')
    with open(tmp, 'w') as f:
        f.write(G.klocal_text(8, 3, 2, 5))
    cases['klocal_dump_8_3_2_5'] = tmp
    for name, path in cases.items():
        d = dump(path)
        d.update(name=name, generated_by='tests/golden/make_golden_dump.py (unmodified /root/reference/py)',
                 file=os.path.basename(path) if path.startswith(HERE) else 'finite_example.txt')
        with open(os.path.join(HERE, name + '.json'), 'w') as f:
            json.dump(d, f)
        print(name, [len(s) for s in d['sites']])


if __name__ == '__main__':
    main()
