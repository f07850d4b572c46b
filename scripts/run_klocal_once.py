"""One k-local DP search (config 3 family) on the GPU; used under ncu for the launch list."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import stabilizer_gs_b200 as S
import gen_hamiltonians as G
per = int(sys.argv[1]) if len(sys.argv) > 1 else 3
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 1
H = S.HamHandle.from_text(G.klocal_text(64, 4, per, 0))
for _ in range(reps):
    r = S.klocal_search(H)
print(r['energy'], r['transitions'], max(r['states_per_site']), r['seconds_device'], r['seconds_total'])
