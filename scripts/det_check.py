import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import stabilizer_gs_b200 as S
import gen_hamiltonians as G
n, T, seed, reps = map(int, sys.argv[1:5])
plan = S.SparsePlan(S.HamHandle.from_text(G.sparse_text(n, T, seed)))
ref = None
for i in range(reps):
    r = plan.run()
    h = r['rank_hist']
    if ref is None: ref = h
    diff = {k: h.get(k, 0) - ref.get(k, 0) for k in set(h) | set(ref) if h.get(k, 0) != ref.get(k, 0)}
    print(i, r['candidates'], r['slow_nodes'], 'diff vs run0:', diff, flush=True)
