"""Drop-in for the reference's py/klocal_periodic.py: the same driver, statement for statement, on this package's
classes (Hamiltonian.from_file_periodic, generate_Sright_periodic, StateMachinePeriodic, State.shift / copy / hash,
EnergyLevel.Es / get_state) — every one of them a handle on data held by libsgs, every step a CUDA kernel.
cmax = 10 and c = 3 as hard-coded in the reference; here the path, c and cmax may also be given as argv[1..3].
(The one-call native equivalent is sgs_klocal_periodic / bin/klocal_periodic.)"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
from stabilizer_gs_b200.state import Hamiltonian, StateMachinePeriodic, generate_Sright_periodic

max_float = 1e100
cmax = int(sys.argv[3]) if len(sys.argv) > 3 else 10  # unrolled periods the handle is built for, should be >= c
path = sys.argv[1] if len(sys.argv) > 1 else "../hamiltonian_periodic.txt"
H_periodic = Hamiltonian.from_file_periodic(path, cmax)
c = int(sys.argv[2]) if len(sys.argv) > 2 else 3  # supercell size


def periodic_evolve_states(states, clear_history=True, c=1):
    """put `states` at site k + l - 1, evolve c periods, bring the results back by -c periods"""
    state_periodic.state_dict = {hash(state): state for state in states}
    state_periodic.m = H_periodic.k + H_periodic.l - 1
    for m in range(H_periodic.l * c):
        state_periodic.evolve()
    return [state.shift(-H_periodic.l * c, clear_history=clear_history) for state in state_periodic.state_dict.values()]


Sright_periodic = generate_Sright_periodic(H_periodic, cmax)
state_periodic = StateMachinePeriodic(H_periodic, Sright_periodic)
for m in range(H_periodic.k + H_periodic.l * 2 - 1):
    state_periodic.evolve()
assert state_periodic.m == H_periodic.k + H_periodic.l * 2 - 1
states = [state.shift(-H_periodic.l, clear_history=True) for state in state_periodic.state_dict.values()]
states_hash = None
while True:                                   # the set of states one period later, until it is a fixed point
    states = periodic_evolve_states(states)
    new_states_hash = tuple(sorted(map(hash, states)))
    print(len(new_states_hash))
    if new_states_hash == states_hash:
        break
    states_hash = new_states_hash

results = []
for unique_state in list(states):
    new_state = unique_state.copy()
    states_all = {hash(state): state for state in periodic_evolve_states([new_state.copy()], c=c)}
    if hash(new_state) not in states_all:
        continue                              # does not come back after c periods
    shape = new_state.stab.energy_level.Es.shape
    for index in np.indices(shape).reshape((len(shape), 2 ** len(shape))).T:
        Es = np.full(shape, max_float)        # start from this sign pattern only
        Es[tuple(index)] = 0.0
        new_state.stab.energy_level.Es = Es
        states_all = {hash(state): state for state in periodic_evolve_states([new_state], clear_history=False, c=c)}
        if hash(new_state) in states_all:
            final_state = states_all[hash(new_state)]
            results.append(final_state.stab.energy_level.get_state(tuple(index), H_periodic.original_paulis, return_stab=False))

# (the reference prints in set-iteration order, which depends on PYTHONHASHSEED: sorted here)
for stab, energy in sorted(results, key=lambda r: (r[1], str(r[0]))):
    print('energy', energy)
    print(stab)
    print()
