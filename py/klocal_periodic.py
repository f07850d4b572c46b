"""Drop-in for the reference's py/klocal_periodic.py (cmax = 10, c = 3 as hard-coded there; the path may be
given as argv[1]).  The whole search — periodic right environments, state fixed point, orbit search — runs on
the GPU behind sgs_klocal_periodic; results are printed in the reference's layout, sorted (the reference's
order depends on PYTHONHASHSEED)."""
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
import stabilizer_gs_b200 as S

cmax = 10  # memory amount to malloc, should >= c
c = 3      # supercell size
path = sys.argv[1] if len(sys.argv) > 1 else "../hamiltonian_periodic.txt"
r = S.klocal_periodic(path, cmax=cmax, c=c)
for m, count in r['sright_log']:
    print(m, count)
for size in r['fixed_point_sizes']:
    print(size)
for o in r['results']:
    print('energy', o['energy'])
    print(f"qubit {o['begin']}-{o['end'] - 1} size ({len(o['generators'])},) " + ' '.join(o['generators']))
    print()
