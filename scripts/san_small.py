"""Small end-to-end pass over every kernel family (used under compute-sanitizer memcheck)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import stabilizer_gs_b200 as S
import gen_hamiltonians as G
H = S.HamHandle.from_text(G.sparse_text(10, 34, 7))
for kw in ({}, {'expand_depth': 1}, {'expand_depth': 2}, {'prune': 1}, {'rank': 1, 'world': 8}):
    r = S.sparse_search(H, **kw)
    print('sparse', kw, r['energy'], r['candidates'])
os.environ['SGS_SPLIT'] = '1'
plan = S.SparsePlan(S.HamHandle.from_text(G.sparse_text(16, 70, 3)), rank=0, world=2)
print('sparse split', plan.run()['candidates'], plan.run()['candidates'])
plan.close()
r = S.sparse_search(S.HamHandle.from_text(G.sparse_text(36, 60, 2)))
print('sparse W=2', r['energy'], r['candidates'])
r = S.klocal_search(S.HamHandle.from_text(G.klocal_text(16, 4, 3, 1)))
print('klocal', r['energy'], r['transitions'])
r = S.klocal_search(S.HamHandle.from_text(G.klocal_text(12, 8, 3, 3)))
print('klocal k=8', r['energy'], r['transitions'])
r = S.klocal_periodic(os.path.join(ROOT, 'examples', 'hamiltonian_periodic.txt'), cmax=10, c=3)
print('periodic', len(r['results']))
