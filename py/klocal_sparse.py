"""Drop-in for the reference's py/klocal_sparse.py (which ignores argv and hard-codes ../hamiltonian_finite.txt;
here the README's documented path argument works, with the reference's default as fallback)."""
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
from stabilizer_gs_b200.state import Hamiltonian, StateMachine, generate_Sright_all

H = Hamiltonian.from_file(sys.argv[1] if len(sys.argv) > 1 else "../hamiltonian_finite.txt")
Sright_all = generate_Sright_all(H)
lists = []
SM2 = StateMachine(H, Sright_all)
while True:
    SM2.evolve()
    print(SM2.m, len(SM2.state_dict))
    lists.append(list(SM2.state_dict.values()))
    if SM2.m == H.n:
        break
SM2.generate_gs()
energy_ref = sum([SM2.stab_gs.get_energy(P) * w for P, w in H.original_paulis])

print(SM2.energy_gs)
print(energy_ref)
print(SM2.stab_gs)
