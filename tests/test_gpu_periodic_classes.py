"""The Python class surface of the periodic path: py/klocal_periodic.py here is the reference's
driver statement for statement — Hamiltonian.from_file_periodic, generate_Sright_periodic, StateMachinePeriodic
(.evolve, .m, assignable .state_dict), State.shift / copy / hash, EnergyLevel.Es / get_state
(/root/reference/py/klocal_periodic.py:9-63, py/state.py:309-325,402-462,542-567, py/stab.py:282-305) — and must print
the golden orbits of the unmodified Python reference (tests/golden/periodic_*.json, KAT-2 among them)."""
import glob
import json
import os
import re
import subprocess
import sys

import pytest

from conftest import GOLDEN, ROOT

pytestmark = pytest.mark.gpu
PERIODIC = sorted(f for f in glob.glob(os.path.join(GOLDEN, 'periodic_*.json')))


def _run(path, c, cmax):
    p = subprocess.run([sys.executable, os.path.join(ROOT, 'py', 'klocal_periodic.py'), path, str(c), str(cmax)], capture_output=True, text=True,
                       timeout=600)
    assert p.returncode == 0, p.stdout[-2000:] + p.stderr[-2000:]
    return p.stdout


def _parse(out):
    """→ (lines before the results, sorted [(energy, 'qubit b-e size (m,) +.. +..')])"""
    head, res, lines = [], [], out.splitlines()
    i = 0
    while i < len(lines) and not lines[i].startswith('energy'):
        head.append(lines[i]); i += 1
    while i < len(lines):
        if lines[i].startswith('energy'):
            res.append((round(float(lines[i].split()[1]), 9), lines[i + 1].strip())); i += 2
        else:
            i += 1
    return head, sorted(res)


@pytest.mark.parametrize('path', PERIODIC, ids=lambda p: os.path.basename(p)[:-5])
def test_script_on_the_class_api_prints_the_reference_orbits(path):
    g = json.load(open(path))
    P = g['periodic']
    head, res = _parse(_run(os.path.join(GOLDEN, g['file']), P['c'], P['cmax']))
    exp = sorted((round(o['energy'], 9), f"qubit {o['begin']}-{o['end'] - 1} size ({len(o['generators'])},) " + ' '.join(o['generators']))
                 for o in P['results'])
    assert res == exp
    # the "m count" lines of generate_Sright_periodic and the fixed-point sizes, in the reference's order
    nums = [tuple(map(int, ln.split())) for ln in head if re.fullmatch(r'\d+( \d+)?', ln.strip())]
    assert [x for x in nums if len(x) == 2] == [tuple(x) for x in P['sright_log']]
    assert [x[0] for x in nums if len(x) == 1] == P['fixed_point_sizes']


def test_state_handles_behave_like_the_reference_objects():
    from stabilizer_gs_b200.state import Hamiltonian, StateMachine, StateMachinePeriodic, generate_Sright_all, generate_Sright_periodic
    import numpy as np
    H = Hamiltonian.from_file_periodic(os.path.join(ROOT, 'examples', 'hamiltonian_periodic.txt'), 10)
    SM = StateMachinePeriodic(H, generate_Sright_periodic(H, 10))
    assert SM.m == 0 and len(SM.state_dict) >= 1
    for _ in range(H.k + 2 * H.l - 1):
        SM.evolve()
    sd = SM.state_dict
    assert all(hash(st) == key for key, st in sd.items()) and all(st.m == SM.m for st in sd.values())
    st = next(iter(sd.values()))
    cp = st.copy()
    assert cp == st and hash(cp) == hash(st) and cp is not st
    assert st.stab.energy_level.Es.shape == (2,) * st.rank
    assert st.stab.begin == max(st.m - H.k + 1, 0) and st.stab.end == st.m          # py/state.py:411-412 after m += 1
    sh = st.shift(-H.l, clear_history=True)
    assert sh.m == st.m - H.l and np.all(sh.stab.energy_level.Es == 0.0)
    assert str(sh.stab.Ps).split('size')[1] == str(st.stab.Ps).split('size')[1]       # same letters, window moved
    Es = np.full((2,) * cp.rank, 7.0)
    cp.stab.energy_level.Es = Es
    assert np.all(cp.stab.energy_level.Es == 7.0) and not np.all(st.stab.energy_level.Es == 7.0)
    # finite chain: state_dict is a real mapping (py/klocal_sparse.py:11 iterates .values())
    Hf = Hamiltonian.from_file(os.path.join(ROOT, 'examples', 'hamiltonian_finite.txt'))
    SF = StateMachine(Hf, generate_Sright_all(Hf))
    sizes = []
    while SF.m < Hf.n:
        SF.evolve()
        vals = list(SF.state_dict.values())
        sizes.append(len(vals))
        assert all(v.m == SF.m for v in vals) and len({hash(v) for v in vals}) == len(vals)
    assert sizes == [7, 23, 25, 9]
