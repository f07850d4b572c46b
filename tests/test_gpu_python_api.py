"""The Python drop-in surface (stabilizer_gs_b200.{pauli,stab,state} and py/*.py) on the GPU."""
import os
import subprocess
import sys

import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu
EX = os.path.join(ROOT, 'examples', 'hamiltonian_finite.txt')


def test_classes_like_the_reference_scripts():
    from stabilizer_gs_b200.stab import StabilizerGroup
    from stabilizer_gs_b200.state import FullStateEnergies, Hamiltonian, StateMachine, generate_Sright_all
    H = Hamiltonian.from_file(EX)
    assert (H.n, H.k, len(H.original_paulis)) == (4, 3, 9)
    Ps, E = FullStateEnergies(H.original_paulis).evolve()
    stab = StabilizerGroup(Ps)
    assert E == pytest.approx(-2.5937784332, abs=1e-10)
    assert str(stab.Ps) == 'qubit 0-3 size (4,) +YIII +IYII +IIXI +IIIY'
    assert [stab.get_energy(P) for P, w in H.original_paulis] == [1, 1, 0, 1, 0, 0, 0, 0, 1]
    SM = StateMachine(H, generate_Sright_all(H))
    sizes = []
    while SM.m < H.n:
        SM.evolve()
        sizes.append(len(SM.state_dict))
    assert sizes == [7, 23, 25, 9]
    SM.generate_gs()
    assert SM.energy_gs == pytest.approx(E, abs=1e-10)
    assert sum(SM.stab_gs.get_energy(P) * w for P, w in H.original_paulis) == pytest.approx(E, abs=1e-10)
    assert str(SM.stab_gs) == str(stab.Ps)


def test_scripts_and_clis_print_the_reference_results():
    out = subprocess.run([sys.executable, os.path.join(ROOT, 'py', 'sparse.py'), EX], capture_output=True, text=True, check=True).stdout
    assert 'gs energy -2.5937784332' in out and 'qubit 0-3 size (4,) +YIII +IYII +IIXI +IIIY' in out
    out = subprocess.run([sys.executable, os.path.join(ROOT, 'py', 'klocal_sparse.py'), EX], capture_output=True, text=True, check=True).stdout
    assert '3 5\n2 7\n1 4\n0 2\n1 7\n2 23\n3 25\n4 9\n' in out
    out = subprocess.run([sys.executable, os.path.join(ROOT, 'py', 'klocal_periodic.py'), os.path.join(ROOT, 'examples', 'hamiltonian_periodic.txt')],
                         capture_output=True, text=True, check=True).stdout
    assert out.count('energy -3.0') == 8 and out.count('energy -1.0') == 4
    # cpp/sparse.cpp:10-17 layout, byte for byte
    exp = ("ground state stabilizers\nqubit 0-3 size 4\n+YIII \n+IYII \n+IIXI \n+IIIY \n\n"
           "ground state energy -2.5937784332\nterms signs\n1 1 0 1 0 0 0 0 1 \n")
    out = subprocess.run([os.path.join(ROOT, 'stabilizer_gs_b200', 'bin', 'sparse'), EX], capture_output=True, text=True, check=True).stdout
    assert out == exp
    out = subprocess.run([os.path.join(ROOT, 'stabilizer_gs_b200', 'bin', 'klocal_sparse'), EX], capture_output=True, text=True, check=True).stdout
    assert out.startswith('verbose level 0\n3 5\n2 7\n1 4\n0 2\n') and out.endswith(exp)
    assert 'm 3\nnumber of states 25\n' in out
    out = subprocess.run([os.path.join(ROOT, 'stabilizer_gs_b200', 'bin', 'klocal_periodic'), os.path.join(ROOT, 'examples', 'hamiltonian_periodic.txt')],
                         capture_output=True, text=True, check=True).stdout
    assert out.count('energy -3') == 8 and '8\n8\n' in out
