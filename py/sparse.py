"""Drop-in for the reference's py/sparse.py: python sparse.py <hamiltonian file>  (search runs on the GPU)."""
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
from stabilizer_gs_b200.stab import StabilizerGroup
from stabilizer_gs_b200.state import Hamiltonian, FullStateEnergies

H = Hamiltonian.from_file(sys.argv[1])
print(*H.original_paulis, sep='\n')

# use sparse (and nonlocal) algorithm
fullstateE = FullStateEnergies(H.original_paulis)
full_Ps, full_energy_gs = fullstateE.evolve()
full_stab = StabilizerGroup(full_Ps)
print('gs energy', full_energy_gs)
# stabilizer group S
print(full_stab.Ps)
