"""python sparse.py <hamiltonian file> — prints what the reference's py/sparse.py prints (every term with its weight,
`gs energy E`, the canonical stabilizer group), using the same class calls a user of the reference makes:
Hamiltonian.from_file -> FullStateEnergies(paulis).evolve() -> StabilizerGroup(Ps).  The search runs on the GPU."""
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
from stabilizer_gs_b200.stab import StabilizerGroup
from stabilizer_gs_b200.state import FullStateEnergies, Hamiltonian


def main(path):
    ham = Hamiltonian.from_file(path)
    for term in ham.original_paulis:                 # (Pauli, weight) pairs in file order
        print(term)
    generators, energy = FullStateEnergies(ham.original_paulis).evolve()
    print('gs energy', energy)
    print(StabilizerGroup(generators).Ps)


if __name__ == '__main__':
    if len(sys.argv) != 2:
        sys.exit('usage: python sparse.py <hamiltonian file>')
    main(sys.argv[1])
